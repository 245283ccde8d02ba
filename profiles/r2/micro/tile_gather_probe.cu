// tile_gather_probe.cu -- microbenchmark behind the tile-list design of round 2.
// Question: what does one list row cost when the j positions are gathered (a) from global memory through L1 (the
// round-1 list pass, bound by L1 wavefronts: ~2 cycles per distinct 128-byte line and warp request) or (b) from a
// shared-memory copy of the CTA's neighbourhood (LDS.128, bank conflicts only), with M = 1, 2 or 4 i-atoms per thread
// sharing every gathered j (packed f32x2 evaluation of two i-atoms per instruction)?
// Synthetic data with the statistics of BASELINE config 3: 256-atom tiles, 2 560-atom footprints, union lists of
// 57 / 81 / 113 rows for M = 1 / 2 / 4.   build: nvcc -arch=sm_100a -O3 tile_gather_probe.cu -o tile_gather_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

#define TPB 128
// one packed evaluation: i-atoms (a, b) against one j
__device__ __forceinline__ void eval2(f2 xi, f2 yi, f2 zi, float4 pj, float gate, f2& fx, f2& fy, f2& fz)
{
    const f2 dx = add2(xi, pk(-pj.x, -pj.x)), dy = add2(yi, pk(-pj.y, -pj.y)), dz = add2(zi, pk(-pj.z, -pj.z));
    const f2 r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    const float ra = lo(r2), rb = hi(r2);
    const f2 r2i = pk(ra <= gate ? frcp(ra) : 0.f, rb <= gate ? frcp(rb) : 0.f);
    const f2 sr2 = mul2(pk(11.5f, 11.5f), r2i);
    const f2 sr6 = mul2(mul2(sr2, sr2), sr2);
    const f2 f = mul2(mul2(mul2(pk(0.24f, 0.24f), r2i), sr6), fma2(pk(2.f, 2.f), sr6, pk(-1.f, -1.f)));
    fx = fma2(f, dx, fx); fy = fma2(f, dy, fy); fz = fma2(f, dz, fz);
}

// MODE 0: global gathers, 1 atom per thread (round-1 shape).  MODE 1/2/4: shared-memory footprint, M atoms per thread.
template <int M, int SMEM>
__global__ void __launch_bounds__(TPB) probe(const float4* __restrict__ pos, const unsigned short* __restrict__ rows16, const int* __restrict__ rows32,
                                              int nrows, int foot, int natoms, float4* __restrict__ out)
{
    extern __shared__ float4 s_pos[];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int base = tile * TPB * M;           // first i-atom of the tile
    const int fbase = max(0, min(natoms - foot, base - foot / 2)); // footprint = window of `foot` atoms around the tile
    float* s_x = (float*)s_pos; float* s_y = s_x + foot; float* s_z = s_y + foot;
    if (SMEM == 1) {
        for (int k = tid; k < foot; k += TPB) s_pos[k] = pos[fbase + k];
        __syncthreads();
    } else if (SMEM == 2) { // structure of arrays: a random gather is three LDS.32 (32 banks) instead of one LDS.128 (8 bank quads)
        for (int k = tid; k < foot; k += TPB) { const float4 p = pos[fbase + k]; s_x[k] = p.x; s_y[k] = p.y; s_z[k] = p.z; }
        __syncthreads();
    }
    float gate = 72.f;
    if (M == 1) {
        const float4 pi = pos[base + tid];
        float fx = 0, fy = 0, fz = 0;
        const size_t rb = ((size_t)(tile * (TPB / 32) + (tid >> 5)) * nrows) * 32 + (tid & 31);
#pragma unroll 4
        for (int k = 0; k < nrows; k++) {
            float4 pj;
            if (SMEM == 1) pj = s_pos[rows16[rb + (size_t)k * 32]];
            else if (SMEM == 2) { const int j = rows16[rb + (size_t)k * 32]; pj = make_float4(s_x[j], s_y[j], s_z[j], 0.f); }
            else pj = __ldg(pos + rows32[rb + (size_t)k * 32]);
            const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            const float r2i = r2 <= gate ? frcp(r2) : 0.f;
            const float sr2 = 11.5f * r2i, sr6 = sr2 * sr2 * sr2;
            const float f = 0.24f * r2i * sr6 * (2.f * sr6 - 1.f);
            fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
        }
        out[base + tid] = make_float4(fx, fy, fz, 0.f);
    } else {
        constexpr int G = M / 2 > 0 ? M / 2 : 1;
        f2 xi[G], yi[G], zi[G], fx[G], fy[G], fz[G];
#pragma unroll
        for (int g = 0; g < M / 2; g++) {
            const float4 pa = pos[base + tid * M + 2 * g], pb = pos[base + tid * M + 2 * g + 1];
            xi[g] = pk(pa.x, pb.x); yi[g] = pk(pa.y, pb.y); zi[g] = pk(pa.z, pb.z);
            fx[g] = fy[g] = fz[g] = pk(0.f, 0.f);
        }
        const size_t rb = ((size_t)(tile * (TPB / 32) + (tid >> 5)) * nrows) * 32 + (tid & 31);
#pragma unroll 2
        for (int k = 0; k < nrows; k++) {
            float4 pj;
            if (SMEM == 1) pj = s_pos[rows16[rb + (size_t)k * 32]];
            else if (SMEM == 2) { const int j = rows16[rb + (size_t)k * 32]; pj = make_float4(s_x[j], s_y[j], s_z[j], 0.f); }
            else pj = __ldg(pos + rows32[rb + (size_t)k * 32]);
#pragma unroll
            for (int g = 0; g < M / 2; g++) eval2(xi[g], yi[g], zi[g], pj, gate, fx[g], fy[g], fz[g]);
        }
#pragma unroll
        for (int g = 0; g < M / 2; g++) {
            out[base + tid * M + 2 * g] = make_float4(lo(fx[g]), lo(fy[g]), lo(fz[g]), 0.f);
            out[base + tid * M + 2 * g + 1] = make_float4(hi(fx[g]), hi(fy[g]), hi(fz[g]), 0.f);
        }
    }
}

template <int M, int SMEM>
static void run(const char* name, int natoms, int nrows, int foot, const float4* d_pos, float4* d_out, int span)
{
    const int ntiles = natoms / (TPB * M);
    const size_t nr = (size_t)ntiles * (TPB / 32) * nrows * 32;
    std::vector<unsigned short> h16(nr); std::vector<int> h32(nr);
    // rows: random picks inside a window of `span` atoms that slides with the row number (neighbouring lanes pick near each other,
    // like the k-th neighbours of atoms of one cell), expressed as footprint-local (16 bit) and global (32 bit) indices
    for (int t = 0; t < ntiles; t++) {
        const int base = t * TPB * M;
        const int fbase = std::max(0, std::min(natoms - foot, base - foot / 2));
        for (int w = 0; w < TPB / 32; w++)
            for (int k = 0; k < nrows; k++)
                for (int l = 0; l < 32; l++) {
                    // lane l's cluster sits l*M atoms further along the pencil, and so does the window its k-th row falls into
                    const int centre = (int)((long long)k * (foot - span - 32 * M) / nrows) + (M > 1 ? l * M : 0);
                    int li = centre + rand() % span;
                    li = std::max(0, std::min(foot - 1, li));
                    const size_t idx = (((size_t)(t * (TPB / 32) + w) * nrows) + k) * 32 + l;
                    h16[idx] = (unsigned short)li; h32[idx] = fbase + li;
                }
    }
    unsigned short* d16; int* d32;
    cudaMalloc(&d16, nr * 2); cudaMalloc(&d32, nr * 4);
    cudaMemcpy(d16, h16.data(), nr * 2, cudaMemcpyHostToDevice); cudaMemcpy(d32, h32.data(), nr * 4, cudaMemcpyHostToDevice);
    const size_t smem = SMEM == 1 ? (size_t)foot * 16 : (SMEM == 2 ? (size_t)foot * 12 : 0);
    cudaFuncSetAttribute(probe<M, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        probe<M, SMEM><<<ntiles, TPB, smem>>>(d_pos, d16, d32, nrows, foot, natoms, d_out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-34s atoms %d rows/thread %d  %.4f ms  -> %.3f ns/atom  (%.1f Matom/ms) %s\n", name, natoms, nrows, best, best * 1e6 / natoms, natoms / best / 1e6,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d16); cudaFree(d32);
}

int main()
{
    const int natoms = 4 * 1024 * 1024, foot = 2560;
    std::vector<float4> h(natoms);
    for (int i = 0; i < natoms; i++) h[i] = make_float4((rand() % 10000) * 0.057f, (rand() % 10000) * 0.057f, (rand() % 10000) * 0.057f, 0.f);
    float4 *d_pos, *d_out;
    cudaMalloc(&d_pos, natoms * 16); cudaMalloc(&d_out, natoms * 16);
    cudaMemcpy(d_pos, h.data(), natoms * 16, cudaMemcpyHostToDevice);
    run<1, 0>("global gather M=1 (round-1 shape)", natoms, 57, foot, d_pos, d_out, 46);
    run<1, 1>("smem AoS LDS.128 M=1 foot 1280", natoms, 57, 1280, d_pos, d_out, 46);
    run<2, 1>("smem AoS LDS.128 M=2 packed", natoms, 81, foot, d_pos, d_out, 46);
    run<4, 1>("smem AoS LDS.128 M=4 packed foot 4900", natoms, 113, 4900, d_pos, d_out, 46);
    run<1, 2>("smem SoA 3xLDS.32 M=1 foot 1280", natoms, 57, 1280, d_pos, d_out, 46);
    run<2, 2>("smem SoA 3xLDS.32 M=2 packed", natoms, 81, foot, d_pos, d_out, 46);
    run<4, 2>("smem SoA 3xLDS.32 M=4 packed foot 4900", natoms, 113, 4900, d_pos, d_out, 46);
    run<2, 2>("smem SoA M=2 packed, rows only 40", natoms, 40, foot, d_pos, d_out, 46);
    run<2, 2>("smem SoA M=2 packed, foot 320", natoms, 81, 320, d_pos, d_out, 46);
    run<2, 0>("global gather M=2 packed", natoms, 81, foot, d_pos, d_out, 46);
    run<4, 0>("global gather M=4 packed", natoms, 113, foot, d_pos, d_out, 46);
    return 0;
}
