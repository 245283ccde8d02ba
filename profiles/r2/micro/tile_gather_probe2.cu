// tile_gather_probe2.cu -- second microbenchmark behind the tile-list design: parameters of the
// shared-memory footprint variant that won tile_gather_probe.cu (structure-of-arrays copy of the CTA's neighbourhood,
// three LDS.32 per list row, 16-bit footprint-local indices).  Questions:
//   * tile size (64 / 128 / 256 atoms per CTA) and footprint size (i.e. resident warps per SM)?
//   * two rows per 32-bit index load (ushort2)?
//   * packed f32x2 evaluation of two rows (FFMA2), now that the gathers are cheaper?
// Synthetic lists with the statistics of BASELINE config 3 (57 rows, k-th neighbours of a warp inside a 46-atom window).
// build: nvcc -arch=sm_100a -O3 tile_gather_probe2.cu -o tile_gather_probe2
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float frcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// VAR 0: one 16-bit index load per row, scalar math.  VAR 1: two rows per 32-bit load, scalar math.
// VAR 2: two rows per 32-bit load, packed math.
template <int TPB, int VAR, int MINB>
__global__ void __launch_bounds__(TPB, MINB) probe(const float4* __restrict__ pos, const unsigned short* __restrict__ rows16, int nrows, int foot, int natoms,
                                                    float4* __restrict__ out)
{
    extern __shared__ float s_x[];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int base = tile * TPB;
    const int fbase = max(0, min(natoms - foot, base - foot / 2));
    float* s_y = s_x + foot; float* s_z = s_y + foot;
    for (int k = tid; k < foot; k += TPB) { const float4 p = pos[fbase + k]; s_x[k] = p.x; s_y[k] = p.y; s_z[k] = p.z; }
    __syncthreads();
    const float gate = 72.f, mx = 1.0e4f;
    const float4 pi = pos[base + tid];
    float fx = 0, fy = 0, fz = 0;
    const size_t rb = ((size_t)(tile * (TPB / 32) + (tid >> 5)) * nrows) * 32 + (tid & 31);
    auto one = [&](int j) {
        const float dx = pi.x - s_x[j], dy = pi.y - s_y[j], dz = pi.z - s_z[j];
        const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
        const float r2i = frcp(r2);
        const float sr2 = 11.5f * r2i, sr6 = sr2 * sr2 * sr2;
        float f = 0.24f * r2i * sr6 * (2.f * sr6 - 1.f);
        f = fminf(fmaxf(f, -mx), mx);
        f = r2 <= gate ? f : 0.f;
        fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
    };
    if (VAR == 0) {
#pragma unroll 4
        for (int k = 0; k < nrows; k++) one(rows16[rb + (size_t)k * 32]);
    } else if (VAR == 1) {
        const unsigned* rows32 = (const unsigned*)rows16 + ((size_t)(tile * (TPB / 32) + (tid >> 5)) * (nrows / 2)) * 32 + (tid & 31);
#pragma unroll 4
        for (int k = 0; k < nrows / 2; k++) { const unsigned v = rows32[(size_t)k * 32]; one(v & 0xffff); one(v >> 16); }
    } else {
        const unsigned* rows32 = (const unsigned*)rows16 + ((size_t)(tile * (TPB / 32) + (tid >> 5)) * (nrows / 2)) * 32 + (tid & 31);
        f2 fx2 = pk(0.f, 0.f), fy2 = fx2, fz2 = fx2;
#pragma unroll 2
        for (int k = 0; k < nrows / 2; k++) {
            const unsigned v = rows32[(size_t)k * 32];
            const int a = v & 0xffff, b = v >> 16;
            const f2 dx = pk(pi.x - s_x[a], pi.x - s_x[b]), dy = pk(pi.y - s_y[a], pi.y - s_y[b]), dz = pk(pi.z - s_z[a], pi.z - s_z[b]);
            const f2 r2 = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
            const float ra = lo(r2), rbb = hi(r2);
            const f2 r2i = pk(ra <= gate ? frcp(ra) : 0.f, rbb <= gate ? frcp(rbb) : 0.f);
            const f2 sr2 = mul2(pk(11.5f, 11.5f), r2i);
            const f2 sr6 = mul2(mul2(sr2, sr2), sr2);
            const f2 f = mul2(mul2(mul2(pk(0.24f, 0.24f), r2i), sr6), fma2(pk(2.f, 2.f), sr6, pk(-1.f, -1.f)));
            fx2 = fma2(f, dx, fx2); fy2 = fma2(f, dy, fy2); fz2 = fma2(f, dz, fz2);
        }
        fx = lo(fx2) + hi(fx2); fy = lo(fy2) + hi(fy2); fz = lo(fz2) + hi(fz2);
    }
    out[base + tid] = make_float4(fx, fy, fz, 0.f);
}

template <int TPB, int VAR, int MINB>
static void run(const char* name, int natoms, int nrows, int foot, const float4* d_pos, float4* d_out, int span)
{
    const int ntiles = natoms / TPB;
    const size_t nr = (size_t)ntiles * (TPB / 32) * nrows * 32;
    std::vector<unsigned short> h16(nr);
    for (int t = 0; t < ntiles; t++)
        for (int w = 0; w < TPB / 32; w++)
            for (int k = 0; k < nrows; k++)
                for (int l = 0; l < 32; l++) {
                    // the window of row k slides through the footprint; the warps of a tile sit w*32 atoms apart
                    const int centre = (int)((long long)k * (foot - span - TPB) / nrows) + w * 32;
                    int li = std::max(0, std::min(foot - 1, centre + rand() % span));
                    size_t idx;
                    if (VAR == 0) idx = (((size_t)(t * (TPB / 32) + w) * nrows) + k) * 32 + l;
                    else idx = ((((size_t)(t * (TPB / 32) + w) * (nrows / 2)) + k / 2) * 32 + l) * 2 + (k & 1);
                    h16[idx] = (unsigned short)li;
                }
    unsigned short* d16;
    cudaMalloc(&d16, nr * 2);
    cudaMemcpy(d16, h16.data(), nr * 2, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)foot * 12;
    cudaFuncSetAttribute(probe<TPB, VAR, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, probe<TPB, VAR, MINB>, TPB, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, probe<TPB, VAR, MINB>);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        probe<TPB, VAR, MINB><<<ntiles, TPB, smem>>>(d_pos, d16, nrows, foot, natoms, d_out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-44s tpb %3d foot %4d regs %2d warps/SM %2d  %.4f ms %s\n", name, TPB, foot, fa.numRegs, nb * TPB / 32, best, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d16);
}

int main()
{
    const int natoms = 4 * 1024 * 1024;
    std::vector<float4> h(natoms);
    for (int i = 0; i < natoms; i++) h[i] = make_float4((rand() % 10000) * 0.057f, (rand() % 10000) * 0.057f, (rand() % 10000) * 0.057f, 0.f);
    float4 *d_pos, *d_out;
    cudaMalloc(&d_pos, natoms * 16); cudaMalloc(&d_out, natoms * 16);
    cudaMemcpy(d_pos, h.data(), natoms * 16, cudaMemcpyHostToDevice);
    const int R = 58;
    run<128, 0, 1>("u16 rows, scalar", natoms, R, 1280, d_pos, d_out, 46);
    run<128, 0, 1>("u16 rows, scalar", natoms, R, 1536, d_pos, d_out, 46);
    run<128, 0, 1>("u16 rows, scalar", natoms, R, 1792, d_pos, d_out, 46);
    run<128, 0, 1>("u16 rows, scalar", natoms, R, 2048, d_pos, d_out, 46);
    run<128, 0, 1>("u16 rows, scalar", natoms, R, 2560, d_pos, d_out, 46);
    run<128, 0, 1>("u16 rows, scalar", natoms, R, 3072, d_pos, d_out, 46);
    run<128, 1, 1>("2 rows per u32, scalar", natoms, R, 1280, d_pos, d_out, 46);
    run<128, 1, 1>("2 rows per u32, scalar", natoms, R, 1536, d_pos, d_out, 46);
    run<128, 1, 1>("2 rows per u32, scalar", natoms, R, 2048, d_pos, d_out, 46);
    run<128, 1, 16>("2 rows per u32, scalar, 32 regs", natoms, R, 1280, d_pos, d_out, 46);
    run<128, 2, 1>("2 rows per u32, packed", natoms, R, 1280, d_pos, d_out, 46);
    run<128, 2, 1>("2 rows per u32, packed", natoms, R, 1536, d_pos, d_out, 46);
    run<128, 2, 1>("2 rows per u32, packed", natoms, R, 2048, d_pos, d_out, 46);
    run<128, 2, 12>("2 rows per u32, packed, 40 regs", natoms, R, 1536, d_pos, d_out, 46);
    run<128, 2, 10>("2 rows per u32, packed, 48 regs", natoms, R, 1536, d_pos, d_out, 46);
    run<128, 1, 12>("2 rows per u32, scalar, 40 regs", natoms, R, 1536, d_pos, d_out, 46);
    run<128, 1, 10>("2 rows per u32, scalar, 48 regs", natoms, R, 1536, d_pos, d_out, 46);
    run<128, 1, 16>("2 rows per u32, scalar, 32 regs", natoms, R, 1536, d_pos, d_out, 46);
    run<128, 1, 8>("2 rows per u32, scalar, 64 regs", natoms, R, 2048, d_pos, d_out, 46);
    run<128, 2, 8>("2 rows per u32, packed, 64 regs", natoms, R, 2048, d_pos, d_out, 46);
    run<64, 1, 1>("2 rows per u32, scalar", natoms, R, 768, d_pos, d_out, 46);
    run<64, 1, 1>("2 rows per u32, scalar", natoms, R, 1024, d_pos, d_out, 46);
    run<64, 2, 1>("2 rows per u32, packed", natoms, R, 768, d_pos, d_out, 46);
    run<64, 2, 1>("2 rows per u32, packed", natoms, R, 1024, d_pos, d_out, 46);
    run<256, 1, 1>("2 rows per u32, scalar", natoms, R, 2560, d_pos, d_out, 46);
    run<256, 1, 1>("2 rows per u32, scalar", natoms, R, 3072, d_pos, d_out, 46);
    run<256, 2, 1>("2 rows per u32, packed", natoms, R, 2560, d_pos, d_out, 46);
    run<256, 2, 1>("2 rows per u32, packed", natoms, R, 3072, d_pos, d_out, 46);
    run<512, 1, 1>("2 rows per u32, scalar", natoms, R, 4608, d_pos, d_out, 46);
    run<512, 2, 1>("2 rows per u32, packed", natoms, R, 4608, d_pos, d_out, 46);
    return 0;
}
