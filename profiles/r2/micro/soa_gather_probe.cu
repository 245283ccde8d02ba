// soa_gather_probe.cu -- does the list pass gather fewer L1 wavefronts from three coordinate streams (3 x LDG.32) than from the
// float4 position stream (1 x LDG.128)?  An LDG.128 is served per quarter-warp (8 lanes x 16 B = one line): ncu on the real pass
// shows ~11.7 wavefronts per gather for ~6.4 distinct lines.  An LDG.32 is served per warp: the same 51-atom window is 205 bytes
// = 2-3 lines.  Same synthetic lists as tile_gather_probe.cu (57 rows per atom, k-th neighbours of a warp inside a window of 46
// atoms that slides along the sorted stream).   build: nvcc -arch=sm_100a -O3 soa_gather_probe.cu -o soa_gather_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ int ld_stream(const int* p) { int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
#define TPB 128
#ifndef MINB
#define MINB 16
#endif
// MODE 0: float4 gather.  1: x[], y[], z[] (3 x LDG.32).  2: xy[] (LDG.64) + z[] (LDG.32).  3: like 1, and the kernel also writes three
// coordinate streams (what the fused integrator would add).
template <int MODE>
__global__ void __launch_bounds__(TPB, MINB) probe(const float4* __restrict__ pos, const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ zs,
                                                    const float2* __restrict__ xy, const int* __restrict__ rows, int nrows, float4* __restrict__ out, float* __restrict__ ox, float* __restrict__ oy, float* __restrict__ oz)
{
    const int i = blockIdx.x * TPB + threadIdx.x;
    const float4 pi = pos[i];
    float fx = 0, fy = 0, fz = 0;
    const int* row = rows + ((size_t)(i >> 5) * nrows) * 32 + (i & 31);
    const float gate = 72.f;
#pragma unroll 4
    for (int k = 0; k < nrows; k++) {
        const int j = ld_stream(row + (size_t)k * 32);
        float xj, yj, zj;
        if (MODE == 0) { const float4 pj = __ldg(pos + j); xj = pj.x; yj = pj.y; zj = pj.z; }
        else if (MODE == 2) { const float2 a = __ldg(xy + j); xj = a.x; yj = a.y; zj = __ldg(zs + j); }
        else { xj = __ldg(xs + j); yj = __ldg(ys + j); zj = __ldg(zs + j); }
        const float dx = pi.x - xj, dy = pi.y - yj, dz = pi.z - zj;
        const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
        const float r2i = r2 <= gate ? frcp(r2) : 0.f;
        const float sr2 = 11.5f * r2i, sr6 = sr2 * sr2 * sr2;
        float f = 0.24f * r2i * sr6 * (2.f * sr6 - 1.f);
        f = fminf(fmaxf(f, -1e4f), 1e4f);
        fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
    }
    out[i] = make_float4(fx, fy, fz, 0.f);
    if (MODE == 3) { ox[i] = pi.x + 1e-3f * fx; oy[i] = pi.y + 1e-3f * fy; oz[i] = pi.z + 1e-3f * fz; }
}

int main()
{
    const int natoms = 4 * 1024 * 1024, nrows = 57, span = 46, foot = 2560;
    std::vector<float4> h(natoms);
    std::vector<float> hx(natoms), hy(natoms), hz(natoms); std::vector<float2> hxy(natoms);
    for (int i = 0; i < natoms; i++) {
        h[i] = make_float4((rand() % 10000) * 0.057f, (rand() % 10000) * 0.057f, (rand() % 10000) * 0.057f, 0.f);
        hx[i] = h[i].x; hy[i] = h[i].y; hz[i] = h[i].z; hxy[i] = make_float2(h[i].x, h[i].y);
    }
    const size_t nr = (size_t)(natoms / 32) * nrows * 32;
    std::vector<int> hr(nr);
    for (int t = 0; t < natoms / TPB; t++) {
        const int base = t * TPB, fbase = std::max(0, std::min(natoms - foot, base - foot / 2));
        for (int w = 0; w < TPB / 32; w++)
            for (int k = 0; k < nrows; k++)
                for (int l = 0; l < 32; l++) {
                    const int centre = (int)((long long)k * (foot - span - 32) / nrows);
                    int li = std::max(0, std::min(foot - 1, centre + rand() % span));
                    hr[(((size_t)(t * (TPB / 32) + w) * nrows) + k) * 32 + l] = fbase + li;
                }
    }
    float4 *d_pos, *d_out; float *dx, *dy, *dz, *ox, *oy, *oz; float2* dxy; int* d_rows;
    cudaMalloc(&d_pos, natoms * 16); cudaMalloc(&d_out, natoms * 16); cudaMalloc(&dx, natoms * 4); cudaMalloc(&dy, natoms * 4); cudaMalloc(&dz, natoms * 4);
    cudaMalloc(&ox, natoms * 4); cudaMalloc(&oy, natoms * 4); cudaMalloc(&oz, natoms * 4); cudaMalloc(&dxy, natoms * 8); cudaMalloc(&d_rows, nr * 4);
    cudaMemcpy(d_pos, h.data(), natoms * 16, cudaMemcpyHostToDevice); cudaMemcpy(dx, hx.data(), natoms * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, hy.data(), natoms * 4, cudaMemcpyHostToDevice); cudaMemcpy(dz, hz.data(), natoms * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dxy, hxy.data(), natoms * 8, cudaMemcpyHostToDevice); cudaMemcpy(d_rows, hr.data(), nr * 4, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[4] = { "float4 gather (1 x LDG.128)", "x[] y[] z[] (3 x LDG.32)", "xy[] + z[] (LDG.64 + LDG.32)", "3 x LDG.32 + three streams written" };
    for (int pass = 0; pass < 2; pass++)
    for (int m = 0; m < 4; m++) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; rep++) {
            cudaEventRecord(a);
            if (m == 0) probe<0><<<natoms / TPB, TPB>>>(d_pos, dx, dy, dz, dxy, d_rows, nrows, d_out, ox, oy, oz);
            if (m == 1) probe<1><<<natoms / TPB, TPB>>>(d_pos, dx, dy, dz, dxy, d_rows, nrows, d_out, ox, oy, oz);
            if (m == 2) probe<2><<<natoms / TPB, TPB>>>(d_pos, dx, dy, dz, dxy, d_rows, nrows, d_out, ox, oy, oz);
            if (m == 3) probe<3><<<natoms / TPB, TPB>>>(d_pos, dx, dy, dz, dxy, d_rows, nrows, d_out, ox, oy, oz);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        printf("%-40s MINB %d  %.4f ms %s\n", names[m], MINB, best, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
