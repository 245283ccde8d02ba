// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A tiny host-side stand-in for the CUDA execution model
// so that the *same kernel source* of aztotmd_b200/csrc can be stepped through without a GPU.
//
// It exists because this repository is developed in a container without a GPU: every indexing
// bug would otherwise cost a multi-minute round trip to a B200 box.  tests/emu/build_emu.sh
// compiles csrc/az_engine.cu with g++ and this header into tests/emu/_build/libaztot_emu.so,
// which only tests/ load (marked "not gpu").  It is NOT a CPU fallback: the product library
// (libaztot_b200.so) is compiled by nvcc only, never includes this file, and the Python
// package cannot load the emulator build.
//
// Model: a kernel launch runs its blocks on a small pool of OS threads; the threads of one block
// are ucontext fibers on one OS thread.  __syncthreads and the warp shuffles are counting barriers
// between fibers, `__shared__` is `static thread_local` (one copy per OS worker = per resident
// block), atomics are real atomics.  Floating-point intrinsics map to libm (so values may differ
// from the GPU's approximate units in the last ulps; tolerances in the tests cover that).
#pragma once
#ifndef AZ_EMU
#error "cuda_emu.h is only for the emulator build (-DAZ_EMU)"
#endif

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#include <algorithm>
#include <ucontext.h>

// ----------------------------------------------------------------------------- vector types
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
static inline int2 make_int2(int x, int y) { return int2{ x, y }; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{ x, y, z, w }; }
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{ x, y }; }

// ----------------------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __shared__ static thread_local
#define __constant__ static

// ----------------------------------------------------------------------------- fibers
struct AzEmuBlock;
struct AzEmuFiber {
    ucontext_t ctx;
    uint3 tid;
    int linear;
    bool done;
    AzEmuBlock* blk;
};
struct AzEmuWarp {
    int expected;     // live lanes
    int arrived[2];
    unsigned gen;
    uint64_t buf[2][32];
    int pred[2][32];
};
struct AzEmuBlock {
    uint3 bid;
    dim3 bdim, gdim;
    int nthreads, live;
    int bar_arrived;
    unsigned bar_gen;
    std::vector<AzEmuFiber> fibers;
    std::vector<AzEmuWarp> warps;
    ucontext_t sched;
    const std::function<void()>* body;
    void* dyn_smem;
};
extern thread_local AzEmuFiber* az_emu_cur;

#define threadIdx (az_emu_cur->tid)
#define blockIdx (az_emu_cur->blk->bid)
#define blockDim (az_emu_cur->blk->bdim)
#define gridDim (az_emu_cur->blk->gdim)

void az_emu_yield();
void az_emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
static inline void* az_emu_dyn_smem() { return az_emu_cur->blk->dyn_smem; }

static inline void __syncthreads()
{
    AzEmuBlock* b = az_emu_cur->blk;
    unsigned g = b->bar_gen;
    if (++b->bar_arrived >= b->live) { b->bar_arrived = 0; b->bar_gen++; return; }
    while (b->bar_gen == g) az_emu_yield();
}
static inline void __syncwarp(unsigned = 0xffffffffu)
{
    AzEmuBlock* b = az_emu_cur->blk;
    AzEmuWarp& w = b->warps[az_emu_cur->linear >> 5];
    unsigned g = w.gen; int p = g & 1;
    if (++w.arrived[p] >= w.expected) { w.arrived[p] = 0; w.gen++; return; }
    while (w.gen == g) az_emu_yield();
}
// exchange one 64-bit payload between the lanes of the calling warp; returns the slot array
static inline const uint64_t* az_emu_warp_exchange(uint64_t v, int pred, const int** preds)
{
    AzEmuBlock* b = az_emu_cur->blk;
    AzEmuWarp& w = b->warps[az_emu_cur->linear >> 5];
    unsigned g = w.gen; int p = g & 1;
    int lane = az_emu_cur->linear & 31;
    w.buf[p][lane] = v; w.pred[p][lane] = pred;
    if (++w.arrived[p] >= w.expected) { w.arrived[p] = 0; w.gen++; }
    else while (w.gen == g) az_emu_yield();
    if (preds) *preds = w.pred[p];
    return w.buf[p];
}
template <class T> static inline uint64_t az_emu_pack(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T az_emu_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
static inline int az_emu_lane() { return az_emu_cur->linear & 31; }
static inline int az_emu_warp_lanes()
{
    AzEmuBlock* b = az_emu_cur->blk; int w = az_emu_cur->linear >> 5;
    return std::min(32, b->nthreads - 32 * w);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    const uint64_t* s = az_emu_warp_exchange(az_emu_pack(v), 0, nullptr);
    int lane = az_emu_lane(); int base = lane & ~(width - 1);
    int from = base + (src & (width - 1));
    return from < az_emu_warp_lanes() ? az_emu_unpack<T>(s[from]) : v;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
    const uint64_t* s = az_emu_warp_exchange(az_emu_pack(v), 0, nullptr);
    int lane = az_emu_lane(); int from = lane + (int)d;
    if ((from & ~(width - 1)) != (lane & ~(width - 1)) || from >= az_emu_warp_lanes()) return v;
    return az_emu_unpack<T>(s[from]);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
    const uint64_t* s = az_emu_warp_exchange(az_emu_pack(v), 0, nullptr);
    int lane = az_emu_lane(); int from = lane - (int)d;
    if (from < (lane & ~(width - 1))) return v;
    return az_emu_unpack<T>(s[from]);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
    const uint64_t* s = az_emu_warp_exchange(az_emu_pack(v), 0, nullptr);
    int lane = az_emu_lane(); int from = lane ^ m;
    if (from >= az_emu_warp_lanes()) return v;
    (void)width;
    return az_emu_unpack<T>(s[from]);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
    const int* p; az_emu_warp_exchange(0, pred, &p);
    unsigned m = 0; int n = az_emu_warp_lanes();
    for (int l = 0; l < n; l++) if (p[l]) m |= 1u << l;
    return m;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred)
{
    unsigned b = __ballot_sync(m, pred); int n = az_emu_warp_lanes();
    return b == (n == 32 ? 0xffffffffu : ((1u << n) - 1));
}
static inline unsigned __activemask() { int n = az_emu_warp_lanes(); return n == 32 ? 0xffffffffu : ((1u << n) - 1); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

// ----------------------------------------------------------------------------- atomics
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int* p, int v)
{
    int o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
static inline unsigned atomicMax(unsigned* p, unsigned v)
{
    unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float* p, float v)
{
    uint32_t* u = (uint32_t*)p; uint32_t o = __atomic_load_n(u, __ATOMIC_RELAXED), n;
    float f;
    do { memcpy(&f, &o, 4); f += v; memcpy(&n, &f, 4); } while (!__atomic_compare_exchange_n(u, &o, n, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    memcpy(&f, &o, 4); return f;
}
static inline double atomicAdd(double* p, double v)
{
    uint64_t* u = (uint64_t*)p; uint64_t o = __atomic_load_n(u, __ATOMIC_RELAXED), n;
    double f;
    do { memcpy(&f, &o, 8); f += v; memcpy(&n, &f, 8); } while (!__atomic_compare_exchange_n(u, &o, n, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    memcpy(&f, &o, 8); return f;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __sync_synchronize(); } // ranks are host threads: a full barrier

// ----------------------------------------------------------------------------- math / bit casts
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __ldg(const float* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline float2 __ldg(const float2* p) { return *p; }
static inline int4 __ldg(const int4* p) { return *p; }
static inline int2 __ldg(const int2* p) { return *p; }
using std::min; using std::max;

// ----------------------------------------------------------------------------- host runtime
typedef int cudaError_t;
typedef struct AzEmuStream* cudaStream_t;
struct AzEmuEvent { std::chrono::steady_clock::time_point t; };
typedef AzEmuEvent* cudaEvent_t;
#define cudaSuccess 0
#define cudaMemcpyHostToDevice 1
#define cudaMemcpyDeviceToHost 2
#define cudaMemcpyDeviceToDevice 3
#define cudaMemcpyDefault 4
#define cudaStreamNonBlocking 1
#define cudaEventDefault 0
#define cudaEventDisableTiming 2
struct cudaDeviceProp { char name[256]; int multiProcessorCount; int major, minor; size_t sharedMemPerBlockOptin; size_t totalGlobalMem; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    memset(p, 0, sizeof(*p)); strcpy(p->name, "cuda-emu (host fibers)");
    p->multiProcessorCount = 148; p->major = 10; p->minor = 0; p->sharedMemPerBlockOptin = 227 * 1024; p->totalGlobalMem = (size_t)8 << 30;
    return 0;
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)aligned_alloc(256, ((n ? n : 1) + 255) / 256 * 256); return *p ? 0 : 2; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new AzEmuEvent(); return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new AzEmuEvent(); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{ *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
