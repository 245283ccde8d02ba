// az_ewald.cuh -- reciprocal-space part of the Ewald sum.
//
// What it computes is ewald_rec of the reference's serial twin (src/elec.cpp:167-336; the CUDA twin recip_ewald /
// ewald_force, src/cuElec.cu:151-381, is compiled out by defines.h:18, keeps N x nK complex numbers in global memory --
// 8 GB for 10^6 ions and 10^3 k-vectors -- and halves both scale factors):
//     S(k)    = sum_i q_i exp(i k r_i)                                   structure factor, k over a half space
//     E_rec   = scale * sum_k a_k |S(k)|^2,   a_k = exp(-k^2 / 4 alpha^2) / k^2,  scale = 4 pi Fcoul / (V eps)
//     F_i     = 2 scale * sum_k a_k (Im c_ik Re S - Re c_ik Im S) k,   c_ik = q_i exp(i k r_i)
// The list of k-vectors is built on the host in double in the reference's own enumeration order and cut-off
// (ewald_setup, az_engine.cu), so the set of vectors is the reference's exactly.
//
// B200 design: nothing of size N x nK is ever stored.
//   az_k_ewald_sfac   one thread per pair of k-vectors (l, m, +n) / (l, m, -n); the atoms of a 128-atom tile staged in
//                     shared memory as the three tables exp(i 2 pi l x / a), l < kx (likewise y, z) built with the
//                     reference's recurrence; every thread walks the tile with its own (l, m, n) -- three shared loads
//                     and 12 FMAs give both members of the pair, since exp(-i n z) is the conjugate of exp(i n z) --
//                     and keeps its partial S(k) in registers over all the tiles of its block: no reduction inside
//                     the tile loop, no atomics.  One partial per block and k.
//   az_k_ewald_reduce one block: partials summed in double in block order (reproducible), a_k S(k) stored for the force
//                     pass, E_rec reduced in a fixed order.
//   az_k_ewald_force  one thread per atom, its tables in shared memory (one column per thread), a_k S(k) and the list
//                     of (l, m) runs staged in shared memory; within a run (n = -nmax .. nmax, contiguous in the
//                     reference's order) exp(i (l x + m y)) is formed once, +n and -n share the four products with
//                     exp(i n z), and the x / y components of the force are applied once per run.
#pragma once
#include "az_device.cuh"

#define AZ_EW_TILE 128      // atoms per shared-memory tile of the structure-factor pass
#define AZ_EW_TPB 256       // threads of the structure-factor pass
#define AZ_EW_KPT 5         // k-vector pairs per thread of the structure-factor pass
#define AZ_EW_MAX_PAIRS (AZ_EW_TPB * AZ_EW_KPT)
#define AZ_EW_MAX_K 2048    // k-vectors (the reference's NTOTKVEC is 1200, src/defines.h:32)
#define AZ_EW_FTPB 128      // threads (atoms) per block of the force pass

struct AzEwald {
    int nk;              // k-vectors kept by the cut-off
    int kx, ky, kz;      // table lengths: entries 0..k-1 per axis (>= 2)
    float tpl[3];        // 2 pi / box length
    float fscale;        // 2 * scale
    double escale;       // scale
    int npairs, nruns;
    const float* ak;     // [nk] a_k, the reference's enumeration order (src/elec.cpp:234-324)
    const int4* pairs;   // [npairs] x = l | |m| << 8 | (m < 0) << 16 | n << 20 (n >= 0), y = index of (l, m, n), z = index of (l, m, -n) or -1
    const int4* runs;    // [nruns] x = l | |m| << 8 | (m < 0) << 16, y = lowest n, z = highest n, w = index of (l, m, lowest n)
};

#ifdef AZ_EWALD_KERNELS // az_engine.cu only: the other units need the argument struct, not the kernels
// exp(i t), exp(i 2 t), ... by repeated complex multiplication (src/elec.cpp:219-231); e[0] = 1
#define AZ_EW_FILL(tab, stride, len, ang)                                              \
    do {                                                                               \
        float s1_, c1_;                                                                \
        sincosf((ang), &s1_, &c1_);                                                    \
        (tab)[0] = make_float2(1.f, 0.f);                                              \
        float2 p_ = make_float2(c1_, s1_);                                             \
        (tab)[(stride)] = p_;                                                          \
        for (int l_ = 2; l_ < (len); l_++) {                                           \
            p_ = make_float2(p_.x * c1_ - p_.y * s1_, p_.y * c1_ + p_.x * s1_);        \
            (tab)[l_ * (stride)] = p_;                                                 \
        }                                                                              \
    } while (0)

__device__ __forceinline__ float2 az_cmul(const float2 a, const float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y); }

// part[block * nk + k] = sum over the atoms of this block's tiles of q_i exp(i k r_i)
__global__ void __launch_bounds__(AZ_EW_TPB)
az_k_ewald_sfac(const __grid_constant__ AzParams P, const __grid_constant__ AzEwald E, const float4* __restrict__ pos, float2* __restrict__ part)
{
    AZ_DYN_SMEM(float2, s_ew);
    float2* tx = s_ew;                          // [atom][l], q folded in
    float2* ty = tx + AZ_EW_TILE * E.kx;        // [atom][m]
    float2* tz = ty + AZ_EW_TILE * E.ky;        // [atom][n]
    const int tid = threadIdx.x;
    const int n = az_n_owned(P);
    const int ntiles = (n + AZ_EW_TILE - 1) / AZ_EW_TILE;
    // this thread's k-vector pairs: table offsets and the conjugation sign of m, packed
    int koff[AZ_EW_KPT];
    float2 accp[AZ_EW_KPT], accm[AZ_EW_KPT];
#pragma unroll
    for (int j = 0; j < AZ_EW_KPT; j++) {
        const int k = tid + j * AZ_EW_TPB;
        accp[j] = accm[j] = make_float2(0.f, 0.f);
        koff[j] = k < E.npairs ? E.pairs[k].x : -1;
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int a0 = tile * AZ_EW_TILE;
        const int na = min(AZ_EW_TILE, n - a0);
        __syncthreads(); // the previous tile has been consumed
        {
            const int a = tid & (AZ_EW_TILE - 1);
            if (a < na) {
                const float4 p = pos[a0 + a];
                if (tid < AZ_EW_TILE) {
                    const float q = P.charge[__float_as_int(p.w)];
                    float2* t = tx + a * E.kx;
                    AZ_EW_FILL(t, 1, E.kx, E.tpl[0] * p.x);
                    for (int l = 0; l < E.kx; l++) { t[l].x *= q; t[l].y *= q; }
                    t = ty + a * E.ky;
                    AZ_EW_FILL(t, 1, E.ky, E.tpl[1] * p.y);
                } else {
                    float2* t = tz + a * E.kz;
                    AZ_EW_FILL(t, 1, E.kz, E.tpl[2] * p.z);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < AZ_EW_KPT; j++) {
            const int ko = koff[j];
            if (ko < 0) continue;
            const float sm = (ko & (1 << 16)) ? -1.f : 1.f; // negative m: conjugate
            const float2* px = tx + (ko & 255);
            const float2* py = ty + ((ko >> 8) & 255);
            const float2* pz = tz + (ko >> 20);
            float2 sp = accp[j], sn = accm[j];
#pragma unroll 4
            for (int a = 0; a < na; a++) {
                const float2 ex = px[a * E.kx];
                float2 ey = py[a * E.ky];
                const float2 ez = pz[a * E.kz];
                ey.y *= sm;
                const float2 lm = az_cmul(ex, ey);
                // lm * ez and lm * conj(ez) share their four products
                sp.x = fmaf(lm.x, ez.x, sp.x); sp.x = fmaf(-lm.y, ez.y, sp.x);
                sp.y = fmaf(lm.y, ez.x, sp.y); sp.y = fmaf(lm.x, ez.y, sp.y);
                sn.x = fmaf(lm.x, ez.x, sn.x); sn.x = fmaf(lm.y, ez.y, sn.x);
                sn.y = fmaf(lm.y, ez.x, sn.y); sn.y = fmaf(-lm.x, ez.y, sn.y);
            }
            accp[j] = sp; accm[j] = sn;
        }
    }
#pragma unroll
    for (int j = 0; j < AZ_EW_KPT; j++) {
        const int k = tid + j * AZ_EW_TPB;
        if (k < E.npairs) {
            const int4 v = E.pairs[k];
            part[(size_t)blockIdx.x * E.nk + v.y] = accp[j];
            if (v.z >= 0) part[(size_t)blockIdx.x * E.nk + v.z] = accm[j];
        }
    }
}

// S(k) = sum of the block partials (double, block order); T(k) = a_k S(k) for the force pass; engCoul2 = scale sum a_k |S|^2.
// Multi-GPU: `Sd` (double2 per k) is what the ranks all-reduce; finish = 0 stops after writing it, finish = 2 starts from it.
__global__ void __launch_bounds__(1024)
az_k_ewald_reduce(const __grid_constant__ AzEwald E, const float2* __restrict__ part, int nparts, double2* __restrict__ Sd, float2* __restrict__ T,
                  double* __restrict__ eng_out, int finish)
{
    __shared__ double s_e[32];
    double e = 0.0;
    for (int k = threadIdx.x; k < E.nk; k += 1024) {
        double sx = 0.0, sy = 0.0;
        if (finish != 2) {
            // eight independent partial sums (eight loads in flight per thread), combined in a fixed order
            double ax[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, ay[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
            int b = 0;
            for (; b + 8 <= nparts; b += 8) {
                float2 p[8];
#pragma unroll
                for (int u = 0; u < 8; u++) p[u] = part[(size_t)(b + u) * E.nk + k];
#pragma unroll
                for (int u = 0; u < 8; u++) { ax[u] += (double)p[u].x; ay[u] += (double)p[u].y; }
            }
            for (; b < nparts; b++) { const float2 p = part[(size_t)b * E.nk + k]; ax[0] += (double)p.x; ay[0] += (double)p.y; }
#pragma unroll
            for (int u = 0; u < 8; u++) { sx += ax[u]; sy += ay[u]; }
            Sd[k] = make_double2(sx, sy);
        } else { const double2 v = Sd[k]; sx = v.x; sy = v.y; }
        if (finish) {
            const float ak = E.ak[k];
            T[k] = make_float2(ak * (float)sx, ak * (float)sy);
            e += (double)ak * (sx * sx + sy * sy);
        }
    }
    if (!finish) return;
    // fixed-order block sum
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) s_e[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; w++) t += s_e[w];
        *eng_out = E.escale * t;
    }
}

// frc[i] += q_i * fscale * sum_k (Im e_ik T_k.x - Re e_ik T_k.y) k,   e_ik = exp(i k r_i),  T_k = a_k S(k)
__global__ void __launch_bounds__(AZ_EW_FTPB)
az_k_ewald_force(const __grid_constant__ AzParams P, const __grid_constant__ AzEwald E, const float4* __restrict__ pos, const float2* __restrict__ T,
                 float4* __restrict__ frc)
{
    AZ_DYN_SMEM(float2, s_ew);
    const int tid = threadIdx.x;
    const int i = blockIdx.x * AZ_EW_FTPB + tid;
    float2* tx = s_ew + tid;                         // [l][thread]
    float2* ty = tx + AZ_EW_FTPB * E.kx;
    float2* tz = ty + AZ_EW_FTPB * E.ky;
    float2* sT = s_ew + AZ_EW_FTPB * (E.kx + E.ky + E.kz);   // [nk]
    int4* sR = (int4*)(sT + ((E.nk + 1) & ~1));               // [nruns], 16-byte aligned
    for (int k = tid; k < E.nk; k += AZ_EW_FTPB) sT[k] = T[k];
    for (int r = tid; r < E.nruns; r += AZ_EW_FTPB) sR[r] = E.runs[r];
    const bool act = i < az_n_owned(P);
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) {
        p = pos[i];
        AZ_EW_FILL(tx, AZ_EW_FTPB, E.kx, E.tpl[0] * p.x);
        AZ_EW_FILL(ty, AZ_EW_FTPB, E.ky, E.tpl[1] * p.y);
        AZ_EW_FILL(tz, AZ_EW_FTPB, E.kz, E.tpl[2] * p.z);
    }
    __syncthreads();
    if (!act) return;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    for (int r = 0; r < E.nruns; r++) {
        const int4 run = sR[r];
        const int l = run.x & 255, m = (run.x >> 8) & 255;
        float2 ey = ty[m * AZ_EW_FTPB];
        const float sgn = (run.x & (1 << 16)) ? -1.f : 1.f;
        ey.y *= sgn;
        const float2 lm = az_cmul(tx[l * AZ_EW_FTPB], ey);
        const float2* t0 = sT + (run.w - run.y); // t0[n] = T of (l, m, n)
        float wsum = 0.f, wz = 0.f;
        if (run.y <= 0) { const float2 t = t0[0]; wsum = lm.y * t.x - lm.x * t.y; } // n = 0
        const bool both = run.y < 0; // the run is symmetric: n = -nmax .. nmax (only l = m = 0 starts at n = 1)
        for (int n = 1; n <= run.z; n++) {
            const float2 ez = tz[n * AZ_EW_FTPB];
            const float a = lm.x * ez.x, b = lm.y * ez.y, c = lm.y * ez.x, d = lm.x * ez.y;
            const float2 tp = t0[n];
            const float wp = (c + d) * tp.x - (a - b) * tp.y;
            float wm = 0.f;
            if (both) { const float2 tm = t0[-n]; wm = (c - d) * tm.x - (a + b) * tm.y; }
            wsum += wp + wm;
            wz = fmaf((float)n, wp - wm, wz);
        }
        fx = fmaf((float)l * E.tpl[0], wsum, fx);
        fy = fmaf(sgn * (float)m * E.tpl[1], wsum, fy);
        fz = fmaf(E.tpl[2], wz, fz);
    }
    const float q = P.charge[__float_as_int(p.w)] * E.fscale;
    float4 f = frc[i];
    f.x = fmaf(q, fx, f.x); f.y = fmaf(q, fy, f.y); f.z = fmaf(q, fz, f.z);
    frc[i] = f;
}
#endif // AZ_EWALD_KERNELS
