// az_device.cuh -- device-side parameter blocks and pair-interaction functions (sm_100a).
//
// No device function pointers and no god-struct: the reference makes two indirect calls per
// pair through md->funcCoul / vdw->feng (src/cuPairs.cu:131,141-146); here the (vdW kind,
// Coulomb kind) combination is a template parameter and all scalars travel by value in a
// __grid_constant__ parameter block.
#pragma once
#ifdef AZ_EMU
#include "cuda_emu.h" // tests/emu: host stand-in for the CUDA execution model (test builds only)
#define AZ_DYN_SMEM(T, name) T* name = (T*)az_emu_dyn_smem()
#define AZ_LAUNCH(kern, grid, block, smem, stream, ...) \
    az_emu_launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); })
#define AZ_LAUNCH_PDL(on, kern, grid, block, smem, stream, ...) AZ_LAUNCH(kern, grid, block, smem, stream, __VA_ARGS__)
static inline void az_pdl_trigger() {}
static inline void az_pdl_wait() {}
#else
#include <cuda_runtime.h>
#define AZ_DYN_SMEM(T, name) \
    extern __shared__ __align__(16) unsigned char az_dyn_smem_raw[]; \
    T* name = reinterpret_cast<T*>(az_dyn_smem_raw)
#define AZ_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// Programmatic dependent launch (sm_90+): the kernel may be brought onto the SMs while its predecessor on the stream is still
// running -- every block of the predecessor has called az_pdl_trigger() -- and parks in az_pdl_wait() until the predecessor has
// finished and its writes are visible.  What it removes is the launch latency between the short kernels of the small decks
// (also as an edge of a captured graph).  `on` = false is the plain launch.
#define AZ_LAUNCH_PDL(on, kern, grid, block, smem, strm_, ...) do { \
        if (on) { \
            cudaLaunchConfig_t cfg_ = {}; \
            cfg_.gridDim = dim3(grid); cfg_.blockDim = dim3(block); cfg_.dynamicSmemBytes = (smem); cfg_.stream = (strm_); \
            cudaLaunchAttribute at_[1]; \
            at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at_[0].val.programmaticStreamSerializationAllowed = 1; \
            cfg_.attrs = at_; cfg_.numAttrs = 1; \
            cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__); \
        } else kern<<<(grid), (block), (smem), (strm_)>>>(__VA_ARGS__); \
    } while (0)
__device__ __forceinline__ void az_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void az_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
#include <stdint.h>
#include "../../include/aztot_b200.h"

#define AZ_VK_MIXED 8 /* template value: per-pair switch over the table's kinds */

struct AzVdw {     // one entry of the nSpec x nSpec table (cudaVdW, src/cuStruct.h:50-63)
    float p0, p1, p2, p3;
    float p4, r2cut, chprod; // chprod = q_i*q_j (md->chProd, src/cuStruct.h:381)
    int kind;                // az_vdw_kind, 0 = no entry
};

struct AzGrid {
    int n[3];      // global cell counts (cNumber)
    int nl[3];     // local grid dims (== n on one GPU; owned + ghost layers when decomposed)
    int goff[3];   // global cell coordinate of local index 0 (0 when not decomposed)
    int dec[3];    // 1 = axis is decomposed across ranks (no local wrap; ghosts carry images)
    int olo[3], ohi[3]; // owned local cell range [olo, ohi)
    int sten[3];   // stencil half widths (cell_dist_int, src/cuCellList.cu:98-132)
    float csize[3], crev[3];
    int ncell_local;
    // storage numbering of local cells: owned cells first (z fastest inside the owned block), ghost cells
    // after them.  NULL on one GPU (identity: storage id == x*cnYZ + y*cNumber.z + z).
    const int* cellmap;
};

// linear local coordinate (including ghost layers) -> storage cell id
__device__ __forceinline__ int az_cell_id(const AzGrid& g, int lin) { return g.cellmap ? g.cellmap[lin] : lin; }

// the step a kernel works on: its launch argument, plus the device-resident base when the launch is part of a replayed graph
#define AZ_STEP(P, k) ((k) + ((P).step_base ? *(P).step_base : 0))
struct AzParams {
    int n_local;        // atoms resident on this GPU that take part as i-atoms (owned); launch bound when n_dyn is set
    const int* n_dyn;   // domain decomposition: device-resident owned-atom count (AzDDCounts::n_owned), else NULL
    const int* step_base; // CUDA-graph replay: kernels receive the step RELATIVE to the first step of the call and add *step_base; else NULL
    int nspec;
    float L[3], revL[3];
    float tSt;
    float r2max;        // pair gate (md->r2Max)
    float r2list;       // cell-pair existence radius^2 = (sqrt(r2max)+skin)^2; == r2max when skin == 0
    float skin;
    float efield[3];
    int has_efield;
    // Coulomb scalars (md->alpha, daipi2, elC1, elC2, rElec; src/cuInit.cu:951-957)
    float alpha, daipi2, elC1, elC2, rElec;
    // per-species tables
    float mass[AZ_MAX_SPECIES], charge[AZ_MAX_SPECIES], rMasshdT[AZ_MAX_SPECIES];
    float radA[AZ_MAX_SPECIES], radB[AZ_MAX_SPECIES], mxEng[AZ_MAX_SPECIES];
    int frozen[AZ_MAX_SPECIES];
    // thermostat
    int tstat;
    float teKin;
    float rQmass, qMassTau2; // Nose-Hoover: 1 / (2 tKin tau^2), 2 tKin (src/sys_init.cpp:1107-1111)
    uint32_t seed_lo, seed_hi;
    AzGrid g;
};

__device__ __forceinline__ int az_n_owned(const AzParams& P) { return P.n_dyn ? *P.n_dyn : P.n_local; }

#define AZ_FCOUL_SCALE 14.3996f // d_Fcoul_scale, src/main.cu:40
#define AZ_SQRTPI 1.772453f     // d_sqrtpi, src/main.cu:41
#define AZ_REV_LIGHT 3.33567e-5f // revLight, src/cuTemp.cu:225
#define AZ_NUM_PI 3.14159f       // numPi, src/cuTemp.cu:228

// ---------------------------------------------------------------------------------------
// Wait until an arrival flag written by a neighbour GPU over peer memory (az_k_ghost_put) has reached `seq`.  Flags only grow;
// a neighbour may already be one step ahead.  Bounded: if the neighbour died the sticky halo flag is raised (bit 8 =
// AZ_FLAG_HALO) and the kernel goes on with stale ghosts instead of hanging the GPU.  In the emulator build the ranks are
// host threads: yield the core while waiting and give them a minute.
#ifdef AZ_EMU
#include <chrono>
#include <thread>
static inline void az_wait_flag(const int* flag, int seq, int* err_flags)
{
    const auto t0 = std::chrono::steady_clock::now();
    while (__atomic_load_n(flag, __ATOMIC_ACQUIRE) - seq < 0) {
        std::this_thread::yield();
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) { __atomic_fetch_or(err_flags, 8, __ATOMIC_RELAXED); break; }
    }
}
#else
__device__ __forceinline__ void az_wait_flag(const int* flag, int seq, int* err_flags)
{
    long long spins = 0;
    while (*(const volatile int*)flag - seq < 0)
        if (++spins > (1ll << 24)) { atomicOr(err_flags, 8); break; } // ~10 s
    __threadfence();
}
#endif

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011); the CPU checker restates the same rounds
__device__ __forceinline__ void az_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                          uint32_t k0, uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---------------------------------------------------------------------------------------
// Pair functions.  Each returns F/r and adds the pair energy.  Arithmetic follows the
// reference forms; 1/x and x/y use the hardware reciprocal (MUFU.RCP, <= 1 ulp) instead of
// the IEEE division sequence -- inside the 1e-5 relative force tolerance of the parity tests.
__device__ __forceinline__ float az_rcp(float x) { return __frcp_rn(x); }
#ifdef AZ_EMU
__device__ __forceinline__ float az_fast_rcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float az_fast_exp(float x) { return expf(x); }
#else
__device__ __forceinline__ float az_fast_exp(float x) { return __expf(x); }
__device__ __forceinline__ float az_fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif

// ---------------------------------------------------------------------------------------
// Packed FP32 pairs (sm_100a: add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2, two IEEE-rn results per issue slot).
// Each half is rounded exactly like the scalar instruction, so a packed evaluation of two list rows gives bitwise the
// terms of two scalar evaluations.  The emulator build carries the two halves in a struct and uses the scalar ops.
#ifdef AZ_EMU
struct az_f2 { float lo, hi; };
__device__ __forceinline__ az_f2 az_pk(float a, float b) { az_f2 r; r.lo = a; r.hi = b; return r; }
__device__ __forceinline__ float az_lo(az_f2 v) { return v.lo; }
__device__ __forceinline__ float az_hi(az_f2 v) { return v.hi; }
__device__ __forceinline__ az_f2 az_fma2(az_f2 a, az_f2 b, az_f2 c) { return az_pk(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
__device__ __forceinline__ az_f2 az_mul2(az_f2 a, az_f2 b) { return az_pk(__fmul_rn(a.lo, b.lo), __fmul_rn(a.hi, b.hi)); }
__device__ __forceinline__ az_f2 az_add2(az_f2 a, az_f2 b) { return az_pk(__fadd_rn(a.lo, b.lo), __fadd_rn(a.hi, b.hi)); }
#else
typedef unsigned long long az_f2;
__device__ __forceinline__ az_f2 az_pk(float a, float b) { az_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float az_lo(az_f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float az_hi(az_f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ az_f2 az_fma2(az_f2 a, az_f2 b, az_f2 c) { az_f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ az_f2 az_mul2(az_f2 a, az_f2 b) { az_f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ az_f2 az_add2(az_f2 a, az_f2 b) { az_f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#endif
__device__ __forceinline__ az_f2 az_pk1(float a) { return az_pk(a, a); }

template <int VK>
__device__ __forceinline__ float az_vdw_feng(const AzVdw& v, float r2, float& r, bool have_r,
                                             float rad1, float rad2, float& eng)
{
    if (VK == AZ_VDW_LJ) { // cu_fe_lj, src/cuVdW.cu:37-47
        float r2i = az_fast_rcp(r2);
        float sr2 = v.p1 * r2i;
        float sr6 = sr2 * sr2 * sr2;
        eng += v.p0 * sr6 * (sr6 - 1.f);
        return v.p2 * r2i * sr6 * (2.f * sr6 - 1.f);
    } else if (VK == AZ_VDW_BUCK) { // cu_fe(r)_buck, src/cuVdW.cu:69-91
        float r2i = az_fast_rcp(r2);
        float r4i = r2i * r2i;
        if (!have_r) r = sqrtf(r2);
        float ex = v.p0 * az_fast_exp(-r * az_fast_rcp(v.p1));
        eng += ex - v.p2 * r4i * r2i;
        return ex * az_fast_rcp(r) * az_fast_rcp(v.p1) - 6.f * v.p2 * r4i * r4i;
    } else if (VK == AZ_VDW_BMH) { // cu_fe(r)_bmh, src/cuVdW.cu:115-136
        float r2i = az_fast_rcp(r2);
        float r4i = r2i * r2i;
        if (!have_r) r = sqrtf(r2);
        float ex = v.p0 * az_fast_exp(v.p1 * (v.p2 - r));
        eng += ex - v.p3 * r4i * r2i - v.p4 * r4i * r4i;
        return v.p1 * ex * az_fast_rcp(r) - 6.f * v.p3 * r4i * r4i - 8.f * v.p4 * r4i * r4i * r2i;
    } else if (VK == AZ_VDW_ELIN) { // cu_fe(r)_elin, src/cuVdW.cu:162-179
        if (!have_r) r = sqrtf(r2);
        float ex = v.p0 * az_fast_exp(-r * az_fast_rcp(v.p1));
        float ir = az_fast_rcp(r);
        eng += ex + v.p2 * r;
        return ex * ir * az_fast_rcp(v.p1) - v.p2 * ir;
    } else if (VK == AZ_VDW_EINV) { // cu_fe(r)_einv, src/cuVdW.cu:200-217
        if (!have_r) r = sqrtf(r2);
        float ex = v.p0 * az_fast_exp(-r * az_fast_rcp(v.p1));
        float ir = az_fast_rcp(r);
        eng += ex - v.p2 * ir;
        return ex * ir * az_fast_rcp(v.p1) - v.p2 * ir * az_fast_rcp(r2);
    } else if (VK == AZ_VDW_SURK) { // surk_pot, src/cuVdW.cu:236-257
        // The bracket 7*C1*a^2b^2/r - 6*C2/(k a + l b) cancels strongly near the potential minimum, so one-ulp
        // differences (FMA contraction, rcp.approx) are amplified past the 1e-5 force tolerance: IEEE
        // divisions and explicitly un-contracted products keep the pair term bit-identical to the oracle.
        float c2ir = v.p1 / __fadd_rn(__fmul_rn(v.p2, rad1), __fmul_rn(v.p3, rad2));
        float r_prod = rad1 * rad2;
        float c1ab2 = __fmul_rn(__fmul_rn(r_prod, r_prod), v.p0);
        float r6 = __fmul_rn(__fmul_rn(r2, r2), r2);
        float rr = sqrtf(r2);
        float ir6 = 1.f / r6;
        float ir = 1.f / rr;
        eng += __fmul_rn(__fmul_rn(r_prod, ir6), __fsub_rn(__fmul_rn(c1ab2, ir), c2ir));
        return __fmul_rn(__fmul_rn(r_prod, ir6) / r2, __fsub_rn(__fmul_rn(__fmul_rn(7.f, c1ab2), ir), __fmul_rn(6.f, c2ir)));
    }
    return 0.f;
}

__device__ __forceinline__ float az_vdw_feng_any(const AzVdw& v, float r2, float& r, bool have_r,
                                                 float rad1, float rad2, float& eng)
{
    switch (v.kind) {
    case AZ_VDW_LJ: return az_vdw_feng<AZ_VDW_LJ>(v, r2, r, have_r, rad1, rad2, eng);
    case AZ_VDW_BUCK: return az_vdw_feng<AZ_VDW_BUCK>(v, r2, r, have_r, rad1, rad2, eng);
    case AZ_VDW_BMH: return az_vdw_feng<AZ_VDW_BMH>(v, r2, r, have_r, rad1, rad2, eng);
    case AZ_VDW_ELIN: return az_vdw_feng<AZ_VDW_ELIN>(v, r2, r, have_r, rad1, rad2, eng);
    case AZ_VDW_EINV: return az_vdw_feng<AZ_VDW_EINV>(v, r2, r, have_r, rad1, rad2, eng);
    case AZ_VDW_SURK: return az_vdw_feng<AZ_VDW_SURK>(v, r2, r, have_r, rad1, rad2, eng);
    default: return 0.f;
    }
}

// Coulomb: returns F/r, writes r (the reference passes r on to the vdW *_r variants).
template <int CK>
__device__ __forceinline__ float az_coul_feng(const AzParams& P, float r2, float& r, float chprd, float& eng)
{
    if (CK == AZ_ELEC_DIRECT) { // direct_coul, src/cuElec.cu:27-37
        float kqq = chprd * AZ_FCOUL_SCALE;
        float ir = rsqrtf(r2);
        r = r2 * ir;
        eng += kqq * ir;
        return kqq * ir * ir * ir;
    } else if (CK == AZ_ELEC_EWALD) { // real_ewald, src/cuElec.cu:94-113
        float kqq = chprd * AZ_FCOUL_SCALE;
        float ir = rsqrtf(r2);
        r = r2 * ir;
        float ar = P.alpha * r;
        float erfcar = erfcf(ar);
        eng += kqq * erfcar * ir;
        return kqq * ir * ir * ir * (erfcar + 2.f * ar * (1.f / AZ_SQRTPI) * az_fast_exp(-ar * ar));
    } else if (CK == AZ_ELEC_FENNEL) { // fennel, src/cuElec.cu:115-134
        float kqq = chprd * AZ_FCOUL_SCALE;
        float ir = rsqrtf(r2);
        r = r2 * ir;
        float ar = P.alpha * r;
        float erfcar = erfcf(ar);
        eng += kqq * (erfcar * ir - P.elC1 + P.elC2 * (r - P.rElec));
        return kqq * ir * ((erfcar * ir * ir + P.daipi2 * az_fast_exp(-ar * ar) * ir) - P.elC2);
    }
    return 0.f;
}

// pair_1 (src/cuPairs.cu:117-167) for one (i,j) candidate, written branch-free: every term is
// evaluated and the reference's gates (vdw != NULL, r2 <= vdw->r2cut) become selects, so a warp
// never diverges inside the pair loop.  The caller applies the r2 <= r2Max gate the same way
// (garbage from r2 == 0 or out-of-range r2 is discarded by its select, never accumulated).
// FOLDCUT: the caller already gates on min(r2Max, r2cut) (single table entry, no Coulomb), so the vdW cut
// select is redundant.
template <int VK, int CK, bool FOLDCUT = false>
__device__ __forceinline__ float az_pair(const AzParams& P, const AzVdw& v, float r2, float rad_i,
                                         float rad_j, bool same_type, float& evdw, float& ecoul)
{
    float f = 0.f, r = 0.f;
    if (CK != AZ_ELEC_NONE) f += az_coul_feng<CK>(P, r2, r, v.chprod, ecoul);
    if (VK != AZ_VDW_NONE) {
        float ra = rad_i, rb = rad_j;
        if ((VK == AZ_VDW_SURK || VK == AZ_VK_MIXED) && same_type) { // symmetric same-species surk
            float lo = fminf(ra, rb), hi = fmaxf(ra, rb); ra = lo; rb = hi;
        }
        float e = 0.f, fv;
        if (VK == AZ_VK_MIXED) fv = az_vdw_feng_any(v, r2, r, CK != AZ_ELEC_NONE, ra, rb, e);
        else fv = az_vdw_feng<VK>(v, r2, r, CK != AZ_ELEC_NONE, ra, rb, e);
        if (FOLDCUT) { f += fv; evdw += e; }
        else {
            const bool vin = (VK == AZ_VK_MIXED ? v.kind != AZ_VDW_NONE : true) && r2 <= v.r2cut;
            f += vin ? fv : 0.f;
            evdw += vin ? e : 0.f;
        }
    }
    // force_autocap, src/cuPairs.cu:94-115: clamp F/r to +-MX_FRC (two FMNMX; a NaN, which the reference would
    // pass through, becomes +-MX_FRC here -- it only arises for coincident atoms and is masked by the caller)
    return fminf(fmaxf(f, -AZ_MX_FRC), AZ_MX_FRC);
}
