// tex_gather_probe.cu -- can the TEXTURE pipe carry the position gathers of the list pass?  ncu says az_k_pair_list is bound by
// the LSU data pipe of L1 (l1tex__data_pipe_lsu_wavefronts 88 %: an LDG.128 with 32 unrelated addresses replays ~2 cycles per
// distinct line).  A tex1Dfetch<float4> goes through the TEX pipe of the same cache.  Variants over the same synthetic lists
// (57 rows, k-th neighbours of a warp inside a 46-atom window): LDG.128 gathers (the shape of az_k_pair_list), texture
// fetches, and both pipes at once (even rows LDG, odd rows TEX).
// build: nvcc -arch=sm_100a -O3 tex_gather_probe.cu -o tex_gather_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#define TPB 128
template <int MODE, int MINB>
__global__ void __launch_bounds__(TPB, MINB) probe(const float4* __restrict__ pos, cudaTextureObject_t tex, const int* __restrict__ rows, int nrows, int natoms, float4* __restrict__ out)
{
    const int i = blockIdx.x * TPB + threadIdx.x;
    if (i >= natoms) return;
    const float gate = 72.f, mx = 1.0e4f;
    const float4 pi = pos[i];
    float fx = 0, fy = 0, fz = 0;
    const int* __restrict__ row = rows + ((size_t)(i >> 5) * nrows) * 32 + (i & 31);
    auto one = [&](const float4 pj) {
        const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
        const float r2i = frcp(r2);
        const float sr2 = 11.5f * r2i, sr6 = sr2 * sr2 * sr2;
        float f = 0.24f * r2i * sr6 * (2.f * sr6 - 1.f);
        f = fminf(fmaxf(f, -mx), mx);
        f = r2 <= gate ? f : 0.f;
        fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
    };
#pragma unroll 4
    for (int k = 0; k < nrows; k++) {
        const int c = row[(size_t)k * 32];
        float4 pj;
        if (MODE == 0) pj = __ldg(pos + c);
        else if (MODE == 1) pj = tex1Dfetch<float4>(tex, c);
        else pj = (k & 1) ? tex1Dfetch<float4>(tex, c) : __ldg(pos + c);
        one(pj);
    }
    out[i] = make_float4(fx, fy, fz, 0.f);
}

template <int MODE, int MINB>
static void run(const char* name, int natoms, int nrows, const float4* d_pos, cudaTextureObject_t tex, const int* d_rows, float4* d_out)
{
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, probe<MODE, MINB>);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        probe<MODE, MINB><<<natoms / TPB, TPB>>>(d_pos, tex, d_rows, nrows, natoms, d_out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-40s regs %2d  %.4f ms %s\n", name, fa.numRegs, best, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    const int natoms = 4 * 1024 * 1024, nrows = 58, span = 46, foot = 1280;
    std::vector<float4> h(natoms);
    for (int i = 0; i < natoms; i++) h[i] = make_float4((rand() % 10000) * 0.057f, (rand() % 10000) * 0.057f, (rand() % 10000) * 0.057f, 0.f);
    float4 *d_pos, *d_out; int* d_rows;
    cudaMalloc(&d_pos, (size_t)natoms * 16); cudaMalloc(&d_out, (size_t)natoms * 16);
    cudaMemcpy(d_pos, h.data(), (size_t)natoms * 16, cudaMemcpyHostToDevice);
    const size_t nr = (size_t)(natoms / 32) * nrows * 32;
    std::vector<int> hr(nr);
    for (int w = 0; w < natoms / 32; w++) {
        const int base = w * 32;
        const int fbase = std::max(0, std::min(natoms - foot, base - foot / 2));
        for (int k = 0; k < nrows; k++)
            for (int l = 0; l < 32; l++) {
                const int centre = (int)((long long)k * (foot - span - 32) / nrows);
                const int li = std::max(0, std::min(foot - 1, centre + rand() % span));
                hr[((size_t)w * nrows + k) * 32 + l] = fbase + li;
            }
    }
    cudaMalloc(&d_rows, nr * 4);
    cudaMemcpy(d_rows, hr.data(), nr * 4, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = d_pos; rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = (size_t)natoms * 16;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType; td.filterMode = cudaFilterModePoint; td.addressMode[0] = cudaAddressModeClamp;
    cudaTextureObject_t tex = 0;
    cudaError_t ce = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    printf("texture object: %s\n", cudaGetErrorString(ce));
    run<0, 16>("LDG.128 gathers, 32 regs", natoms, nrows, d_pos, tex, d_rows, d_out);
    run<1, 16>("tex1Dfetch<float4>, 32 regs", natoms, nrows, d_pos, tex, d_rows, d_out);
    run<2, 16>("even rows LDG, odd rows TEX, 32 regs", natoms, nrows, d_pos, tex, d_rows, d_out);
    run<0, 12>("LDG.128 gathers, 40 regs", natoms, nrows, d_pos, tex, d_rows, d_out);
    run<1, 12>("tex1Dfetch<float4>, 40 regs", natoms, nrows, d_pos, tex, d_rows, d_out);
    run<2, 12>("even rows LDG, odd rows TEX, 40 regs", natoms, nrows, d_pos, tex, d_rows, d_out);
    run<1, 8>("tex1Dfetch<float4>, 64 regs", natoms, nrows, d_pos, tex, d_rows, d_out);
    run<2, 8>("even rows LDG, odd rows TEX, 64 regs", natoms, nrows, d_pos, tex, d_rows, d_out);
    return 0;
}
