// az_kernels.cuh -- bin / scan / reorder / integrate / thermostat kernels (sm_100a).
//
// Data layout in HBM (per atom, SoA streams, 16-byte aligned, LDG.128/STG.128):
//   pos[i]   = float4 (x, y, z, bitcast(type))
//   vel[i]   = float4 (vx, vy, vz, 0)
//   frc[i]   = float4 (fx, fy, fz, 0)          current (sorted) order only
//   therm[i] = float2 (U thermal energy, radius)  radiative thermostat / surk only
//   id[i]    = int     persistent id (index in atoms.xyz)
// One thread per atom, coalesced -- the reference gives each thread `atPerThread`
// consecutive AoS atoms (src/cuInit.cu:963-972), i.e. uncoalesced.
#pragma once
#include "az_device.cuh"

#define AZ_TPB 256

__device__ __forceinline__ double az_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of up to NV doubles; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void az_block_sum(double (&v)[NV], double* smem /* [NV*32] */)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = az_warp_sum(v[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) smem[k * 32 + w] = v[k];
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double t = lane < nw ? smem[k * 32 + lane] : 0.0;
            v[k] = az_warp_sum(t);
        }
    }
    __syncthreads();
}

// Device-resident step scalars (what reset_quantities / calc_quantities keep in cudaMD).
struct AzScalars {
    double posMom[3], negMom[3];              // cumulative (src/cuMDfunc.cu:72-104)
    double posMomPend[3][3], negMomPend[3][3]; // crossings of the step in flight (slot = step % 3); az_k_finalize folds them in
    int acPos[AZ_MAX_SPECIES][3], acNeg[AZ_MAX_SPECIES][3];
    double engElecField;                      // per step
    double engKin, engVdW, engCoul1, engTemp; // per step, reduced
    double engCoul2, engCoul3;                // Ewald: reciprocal-space energy (az_k_ewald_reduce, every step), constant term (set once)
    double engBond, engAngle;                 // constant bonds / valence angles (az_k_bonded_reduce, every step)
    float engKin_f;                           // after after_tscale
    float posMomBuf[AZ_N_MOM][3], negMomBuf[AZ_N_MOM][3];
    int iMom;
    float pressure, engTot, engPot;
    int step;
    int overflow, dbg_count, dbg_atom; // sticky error flags (+ worst list row count and the atom it belongs to)
    // flags:: 1 skin exceeded, 2 cell larger than the pair tile, 4 NaN energy, 8 halo slots, 16 neighbour-list rows
    float tscale_k;
    float keGlob;                    // bricks + Nose-Hoover: kinetic energy of the whole system after the last application
    float chit, consInt, nose_scale; // Nose-Hoover friction, conserved-quantity integral, velocity factor of the block in flight
    float x_first; // pos[0].x at publish time (print_stat's x1)
    unsigned long long nPairs; // unique in-range pairs of the last energy-reducing force pass
    unsigned long long nVisits; // the same, counted from both sides (a brick sees cross-brick pairs from one side only)
    unsigned vmax2_bits[3]; // max |v|^2 of a stale-cell step (float bits), slot = step % 3: the fused kick+drift kernel writes
                            // the next step's slot before this step's az_k_finalize has consumed its own, and az_k_finalize of
                            // step n may still be running (on its own stream) while the pass of step n + 1 writes slot n + 2
    float drift;         // upper bound of any atom's displacement since the last rebuild
};
#define AZ_FLAG_SKIN 1
#define AZ_FLAG_TILE 2
#define AZ_FLAG_NAN 4
#define AZ_FLAG_HALO 8
#define AZ_FLAG_NBR 16
#define AZ_FLAG_FOOT 32 // tile lists: a footprint larger than the shared-memory allotment (or too many runs)

// ---------------------------------------------------------------------------------------
// K1: verlet_1stage (src/cuMDfunc.cu:333-518) + put_periodic (:35-106) + count_cell
// (src/cuSort.cu:114-128).  MODE 0: stale cells between rebuilds (no wrap, no bin); 1: wrap + bin
// (one GPU); 2: wrap only (domain decomposition: migration comes between wrap and bin, az_k_bin bins).
#define AZ_IA_STALE 0
#define AZ_IA_BIN 1
#define AZ_IA_WRAP 2
// m |v|^2 of verlet_2stage's kinetic-energy sum (src/cuMDfunc.cu:521-599) with the contraction written out, so that
// every kernel that hosts the second half kick rounds it the same way
__device__ __forceinline__ float az_ke_term(const float4 v, float m) { return fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x)) * m; }

// per-atom body: returns |v|^2 after the kick; eF receives q * (x . E)
template <int MODE>
__device__ __forceinline__ float az_kick_drift(const AzParams& P, int i, int slot, float4& p, float4& v, const float4 f, int* __restrict__ cellOf,
                                               int* __restrict__ slotOf, int* __restrict__ cellCount, AzScalars* __restrict__ sc, double& eF)
{
    int t = __float_as_int(p.w);
    float h = P.rMasshdT[t];
    // explicit FMAs: what nvcc contracts the reference's `v += rMasshdT * f; x += v * tSt` into
    v.x = fmaf(h, f.x, v.x); v.y = fmaf(h, f.y, v.y); v.z = fmaf(h, f.z, v.z);
    if (!P.frozen[t]) { p.x = fmaf(v.x, P.tSt, p.x); p.y = fmaf(v.y, P.tSt, p.y); p.z = fmaf(v.z, P.tSt, p.z); }
    const float v2 = v.x * v.x + v.y * v.y + v.z * v.z;
    if (MODE != AZ_IA_STALE) {
        float m = P.mass[t];
        float* pc[3] = { &p.x, &p.y, &p.z };
        float vc[3] = { v.x, v.y, v.z };
        int ci[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            int nx = (int)floor((double)*pc[a] * (double)P.revL[a]); // the reference's double floor
            float x = *pc[a] - (float)nx * P.L[a];
            if (x >= P.L[a]) x = 0.f;
            *pc[a] = x;
            if (nx > 0) { atomicAdd(&sc->acPos[t][a], 1); atomicAdd(&sc->posMomPend[slot][a], (double)(m * vc[a])); }
            else if (nx < 0) { atomicAdd(&sc->acNeg[t][a], 1); atomicAdd(&sc->negMomPend[slot][a], (double)(m * (-vc[a]))); }
            int c = (int)floor((double)x * (double)P.g.crev[a]);
            c = max(0, min(c, P.g.n[a] - 1));
            ci[a] = c;
        }
        if (MODE == AZ_IA_BIN) { // storage cell id (== global id on one GPU: x*cnYZ + y*cNumber.z + z)
            int lc = ((ci[0] - P.g.goff[0]) * P.g.nl[1] + (ci[1] - P.g.goff[1])) * P.g.nl[2] + (ci[2] - P.g.goff[2]);
            cellOf[i] = lc;
            slotOf[i] = atomicAdd(&cellCount[lc], 1);
        }
    }
    if (P.has_efield) {
        float q = P.charge[t];
        eF = (double)(q * (p.x * P.efield[0] + p.y * P.efield[1] + p.z * P.efield[2]));
    }
    return v2;
}

// block-level tails: the fastest atom bounds the drift since the last rebuild (stale cells); field energy
template <int MODE>
__device__ __forceinline__ void az_kick_drift_tail(const AzParams& P, float v2, double eF, int slot, AzScalars* __restrict__ sc)
{
    if (MODE == AZ_IA_STALE) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v2 = fmaxf(v2, __shfl_down_sync(0xffffffffu, v2, o));
        // one atomic per warp only while the running maximum is still smaller (a handful per launch)
        if ((threadIdx.x & 31) == 0 && __float_as_uint(v2) > *(volatile unsigned*)&sc->vmax2_bits[slot]) atomicMax(&sc->vmax2_bits[slot], __float_as_uint(v2));
    }
    if (P.has_efield) {
        __shared__ double sm[32];
        double v1[1] = { eF };
        az_block_sum<1>(v1, sm);
        if (threadIdx.x == 0) atomicAdd(&sc->engElecField, v1[0]);
    }
}

template <int MODE>
__global__ void __launch_bounds__(AZ_TPB)
az_k_integrate_a(const __grid_constant__ AzParams P, int slot, float4* __restrict__ pos, float4* __restrict__ vel,
                 const float4* __restrict__ frc, int* __restrict__ cellOf, int* __restrict__ slotOf,
                 int* __restrict__ cellCount, AzScalars* __restrict__ sc)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double eF = 0.0;
    float v2 = 0.f;
    if (i < az_n_owned(P)) {
        float4 p = pos[i], v = vel[i];
        v2 = az_kick_drift<MODE>(P, i, slot, p, v, frc[i], cellOf, slotOf, cellCount, sc, eF);
        pos[i] = p; vel[i] = v;
    }
    az_kick_drift_tail<MODE>(P, v2, eF, slot, sc);
}

// ---------------------------------------------------------------------------------------
// K2: exclusive scan of nAtInCell -> firstAtomInCell (calc_firstAtomInCell, src/cuSort.cu:130-143,
// a single-thread loop in the reference).  Two launches: tile sums, then each tile adds the
// sums of the tiles before it and scans its own items.  Up to AZ_SCAN_ONE_LAUNCH tiles (the 40 000- and 4 000-atom
// decks: 2 197 and 1 728 cells) the second kernel alone sums the counts before its tile itself: one launch less
// on a latency-bound step.
#define AZ_SCAN_ONE_LAUNCH 8
#define AZ_SCAN_TPB 256
#define AZ_SCAN_IPT 8
#define AZ_SCAN_TILE (AZ_SCAN_TPB * AZ_SCAN_IPT)

static __global__ void __launch_bounds__(AZ_SCAN_TPB)
az_k_scan_tiles(const int* __restrict__ count, int n, int* __restrict__ tileSum)
{
    int base = blockIdx.x * AZ_SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int k = 0; k < AZ_SCAN_IPT; k++) {
        int idx = base + k * AZ_SCAN_TPB + threadIdx.x;
        if (idx < n) s += count[idx];
    }
    __shared__ int sm[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = threadIdx.x < (AZ_SCAN_TPB / 32) ? sm[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) tileSum[blockIdx.x] = t;
    }
}

static __global__ void __launch_bounds__(AZ_SCAN_TPB)
az_k_scan_apply(const int* __restrict__ count, int n, const int* __restrict__ tileSum,
                int* __restrict__ first /* [n+1] */)
{
    __shared__ int sm[32];
    __shared__ int s_off;
    az_pdl_trigger(); az_pdl_wait(); // (programmatic dependent launch: no-ops on a plain launch)
    // offset of this tile = sum of the tile sums before it
    int off = 0;
    if (tileSum) { for (int k = threadIdx.x; k < (int)blockIdx.x; k += AZ_SCAN_TPB) off += tileSum[k]; }
    else { for (int k = threadIdx.x; k < (int)blockIdx.x * AZ_SCAN_TILE; k += AZ_SCAN_TPB) off += count[k]; } // few tiles: one launch, no tile sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) off += __shfl_down_sync(0xffffffffu, off, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = off;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < AZ_SCAN_TPB / 32; k++) t += sm[k]; s_off = t; }
    __syncthreads();
    int base = blockIdx.x * AZ_SCAN_TILE + threadIdx.x * AZ_SCAN_IPT; // thread owns IPT consecutive cells
    int v[AZ_SCAN_IPT], tsum = 0;
#pragma unroll
    for (int k = 0; k < AZ_SCAN_IPT; k++) { v[k] = (base + k < n) ? count[base + k] : 0; tsum += v[k]; }
    // block exclusive scan of tsum
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; k++) woff += sm[k];
    int run = s_off + woff + inc - tsum;
#pragma unroll
    for (int k = 0; k < AZ_SCAN_IPT; k++) {
        if (base + k < n) first[base + k] = run;
        run += v[k];
    }
    if (base <= n - 1 && n - 1 < base + AZ_SCAN_IPT) first[n] = run; // total, after the last cell
}

// ---------------------------------------------------------------------------------------
// K3: sort_atoms (src/cuSort.cu:145-197) made deterministic.  Phase a drops each atom's
// persistent id into its cell segment at the atomicAdd slot (arbitrary order); phase b ranks the
// atom inside its cell by id (cells hold ~10-20 atoms) and scatters the payload streams, so the
// within-cell order is canonical (ascending id) whatever the atomics did.
static __global__ void __launch_bounds__(AZ_TPB)
az_k_sort_keys(int n, const int* __restrict__ n_dyn, const int* __restrict__ cellOf, const int* __restrict__ slotOf,
               const int* __restrict__ first, const int* __restrict__ id, int* __restrict__ keyTmp)
{
    az_pdl_trigger(); az_pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dyn) n = *n_dyn;
#ifdef AZ_PROBE_MORTON // probe build only (profiles/r2): the cell's members are ranked by a sub-cell Morton code, ties by id
    if (i < n && cellOf[i] >= 0) keyTmp[first[cellOf[i]] + slotOf[i]] = i;
#else
    if (i < n && cellOf[i] >= 0) keyTmp[first[cellOf[i]] + slotOf[i]] = id[i]; // cell -1: the atom left this brick
#endif
}

template <bool THERM>
__global__ void __launch_bounds__(AZ_TPB)
az_k_sort_scatter(int n, const int* __restrict__ n_dyn, const int* __restrict__ cellOf, const int* __restrict__ first,
                  const int* __restrict__ count, const int* __restrict__ keyTmp,
                  const float4* __restrict__ pos, const float4* __restrict__ vel,
                  const int* __restrict__ id, const float2* __restrict__ therm,
                  float4* __restrict__ pos_o, float4* __restrict__ vel_o, int* __restrict__ id_o,
                  float2* __restrict__ therm_o, int* __restrict__ cell_o, int* __restrict__ newIndex, int* __restrict__ whereId
#ifdef AZ_PROBE_MORTON
                  , float subx, float suby, float subz
#endif
                  )
{
    az_pdl_trigger(); az_pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dyn) n = *n_dyn;
    if (i >= n) return;
    int c = cellOf[i];
    if (c < 0) return;
    int b = first[c], m = count[c];
    int my = id[i];
    int rank = 0;
#ifdef AZ_PROBE_MORTON
    auto code = [&](const float4 p) {
        const int a = (int)(p.x * subx) & 3, bb = (int)(p.y * suby) & 3, cc = (int)(p.z * subz) & 3;
        return ((cc >> 1) << 5) | ((bb >> 1) << 4) | ((a >> 1) << 3) | ((cc & 1) << 2) | ((bb & 1) << 1) | (a & 1);
    };
    const int mc = code(pos[i]);
    for (int k = 0; k < m; k++) {
        const int o = keyTmp[b + k];
        const int oc = code(pos[o]), oid = id[o];
        rank += (oc < mc) | ((oc == mc) & (oid < my));
    }
#else
    for (int k = 0; k < m; k++) rank += (keyTmp[b + k] < my);
#endif
    int d = b + rank;
    pos_o[d] = pos[i];
    vel_o[d] = vel[i];
    id_o[d] = my;
    cell_o[d] = c;
    if (THERM) therm_o[d] = therm[i];
    if (newIndex) newIndex[i] = d; // domain decomposition: send lists and ghost slots follow the atoms
    if (whereId) whereId[my] = d;  // bonds / angles name persistent ids (sort_bonds / sort_angles, src/cuSort.cu:199-235)
}

// ---------------------------------------------------------------------------------------
// Radiative thermostat for one atom: tstat_radi9 body (src/cuTemp.cu:709-763) with
// adsorb_rand_photon (:484-507), radiate_photon3 (:631-685) and get_angled_vector (:395-453).
__device__ __forceinline__ float az_f3sqr(float x, float y, float z) { return x * x + y * y + z * z; }

__device__ __forceinline__ void az_angled_vector(float ix, float iy, float iz, float cos_phi, float theta,
                                                 float& ox, float& oy, float& oz)
{
    float l1 = sqrtf(az_f3sqr(ix, iy, iz));
    float v1x = ix / l1, v1y = iy / l1, v1z = iz / l1;
    float v2x, v2y, v2z;
    if (v1x != 0.f) { v2y = 1.f; v2z = 1.f; v2x = -(v1y * v2y + v1z * v2z) / v1x; }
    else if (v1y != 0.f) { v2x = 1.f; v2z = 1.f; v2y = -(v1z * v2z) / v1y; }
    else { v2x = 1.f; v2y = 0.f; v2z = 0.f; }
    float v3x = v1y * v2z - v1z * v2y;
    float v3y = -v1x * v2z + v1z * v2x;
    float v3z = v1x * v2y - v1y * v2x;
    float l2 = sqrtf(az_f3sqr(v2x, v2y, v2z)), l3 = sqrtf(az_f3sqr(v3x, v3y, v3z));
    v2x /= l2; v2y /= l2; v2z /= l2;
    v3x /= l3; v3y /= l3; v3z /= l3;
    float sinPhi = sqrtf(1 - cos_phi * cos_phi);
    float sinTh, cosTh;
    sincosf(theta, &sinTh, &cosTh);
    ox = v1x * cos_phi + sinPhi * (cosTh * v2x + sinTh * v3x);
    oy = v1y * cos_phi + sinPhi * (cosTh * v2y + sinTh * v3y);
    oz = v1z * cos_phi + sinPhi * (cosTh * v2z + sinTh * v3z);
}

// Not inlined on purpose: az_k_integrate_b, az_k_integrate_ba and az_k_tscale then run the SAME machine code, so a
// step gives bitwise the same velocities / radii whichever kernel hosts it (inlined copies get different FMA
// contractions from their surroundings).  Once per atom and step: the call costs nothing measurable.
#ifdef AZ_EMU
__device__ __forceinline__
#else
static __device__ __noinline__
#endif
void az_tstat_radi_atom(const AzParams& P, int step, int pid, int tp, float pe,
                                                   const float* __restrict__ uv, float& vx, float& vy,
                                                   float& vz, float& U, float& rad)
{
    float mass = P.mass[tp];
    uint32_t rn[4];
    az_philox((uint32_t)pid, (uint32_t)step, 0x52414449u, 0u, P.seed_lo, P.seed_hi, rn);
    const float* u = uv + 3 * (rn[1] % AZ_N_UVECT);
    float ux = u[0], uy = u[1], uz = u[2];
    { // adsorb_rand_photon
        float v02 = az_f3sqr(vx, vy, vz);
        float ermc = pe * AZ_REV_LIGHT / mass;
        vx += ermc * ux; vy += ermc * uy; vz += ermc * uz;
        float v12 = az_f3sqr(vx, vy, vz);
        U += pe + 0.5f * mass * (v02 - v12);
    }
    if ((double)U > 1e-4) { // radiate_photon3
        float u0 = U;
        float v02 = az_f3sqr(vx, vy, vz);
        float v0 = sqrtf(v02);
        float ph_eng = 0.9f * u0;
        float ermc = ph_eng * AZ_REV_LIGHT / mass;
        float ermcv0 = ermc / v0;
        float rx, ry, rz;
        if (!(ermcv0 < 1.f)) {
            if (v0 > 0.f) { rx = -vx / v0; ry = -vy / v0; rz = -vz / v0; }
            else { rx = ux; ry = uy; rz = uz; }
        } else {
            float cos_phi = (float)(rn[2] % 2048u) / 1024.f * (1.f - ermcv0);
            cos_phi -= 1.f;
            float theta = (float)(rn[3] % 2048u) / 1024.f * AZ_NUM_PI;
            az_angled_vector(vx, vy, vz, cos_phi, theta, rx, ry, rz);
        }
        vx += ermc * rx; vy += ermc * ry; vz += ermc * rz;
        float v12 = az_f3sqr(vx, vy, vz);
        U -= (ph_eng + 0.5f * mass * (v12 - v02));
    }
    float restrE = fminf(U, P.mxEng[tp]);
    rad = P.radA[tp] / (P.radB[tp] - restrE);
}

// ---------------------------------------------------------------------------------------
// K5: verlet_2stage (src/cuMDfunc.cu:521-599) with the kinetic-energy reduction, and -- when
// TSTAT -- tstat_radi9 fused behind it (the reference's KE is taken before the thermostat acts).
// Partial sums go to per-block slots; az_k_finalize reduces them in a fixed order.
template <bool TSTAT>
__global__ void __launch_bounds__(AZ_TPB)
az_k_integrate_b(const __grid_constant__ AzParams P, int step_k, const float4* __restrict__ pos,
                 float4* __restrict__ vel, const float4* __restrict__ frc, const int* __restrict__ id,
                 float2* __restrict__ therm, const float* __restrict__ photons,
                 const float* __restrict__ uv, double* __restrict__ partKE, double* __restrict__ partU)
{
    const int step = AZ_STEP(P, step_k);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double ke = 0.0, ut = 0.0;
    if (i < az_n_owned(P)) {
        float4 v = vel[i], f = frc[i];
        int t = __float_as_int(pos[i].w);
        float h = P.rMasshdT[t];
        v.x = fmaf(h, f.x, v.x); v.y = fmaf(h, f.y, v.y); v.z = fmaf(h, f.z, v.z);
        ke = (double)az_ke_term(v, P.mass[t]);
        if (TSTAT) {
            float2 th = therm[i];
            az_tstat_radi_atom(P, step, id[i], t, photons[i], uv, v.x, v.y, v.z, th.x, th.y);
            therm[i] = th;
            ut = (double)th.x;
        }
        vel[i] = v;
    }
    __shared__ double sm[64];
    double vv[2] = { ke, ut };
    az_block_sum<2>(vv, sm);
    if (threadIdx.x == 0) { partKE[blockIdx.x] = 0.5 * vv[0]; if (TSTAT) partU[blockIdx.x] = vv[1]; }
}

// K5 of step n fused with K1 of step n+1 (same arithmetic in the same order, so the trajectory is bitwise
// the one of the separate kernels): one pass reads pos, vel, frc (48 B) and writes pos, vel (32 B) instead
// of 64 B + 80 B for the two passes.  Used inside az_engine_step for consecutive steps of one call.
template <bool TSTAT, int MODE>
__global__ void __launch_bounds__(AZ_TPB)
az_k_integrate_ba(const __grid_constant__ AzParams P, int step_k, int slot_next, float4* __restrict__ pos, float4* __restrict__ vel,
                  const float4* __restrict__ frc, const int* __restrict__ id, float2* __restrict__ therm,
                  const float* __restrict__ photons, const float* __restrict__ uv, double* __restrict__ partKE,
                  double* __restrict__ partU, int* __restrict__ cellOf, int* __restrict__ slotOf, int* __restrict__ cellCount,
                  AzScalars* __restrict__ sc)
{
    az_pdl_trigger(); // (the scan behind this kernel may be brought in early; it parks until the binning is complete)
    const int step = AZ_STEP(P, step_k);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double ke = 0.0, ut = 0.0, eF = 0.0;
    float v2 = 0.f;
    if (i < az_n_owned(P)) {
        float4 p = pos[i], v = vel[i];
        const float4 f = frc[i];
        int t = __float_as_int(p.w);
        float h = P.rMasshdT[t];
        v.x = fmaf(h, f.x, v.x); v.y = fmaf(h, f.y, v.y); v.z = fmaf(h, f.z, v.z); // verlet_2stage of step n
        ke = (double)az_ke_term(v, P.mass[t]);
        if (TSTAT) {
            float2 th = therm[i];
            az_tstat_radi_atom(P, step, id[i], t, photons[i], uv, v.x, v.y, v.z, th.x, th.y);
            therm[i] = th;
            ut = (double)th.x;
        }
        v2 = az_kick_drift<MODE>(P, i, slot_next, p, v, f, cellOf, slotOf, cellCount, sc, eF); // verlet_1stage of step n + 1
        pos[i] = p; vel[i] = v;
    }
    __shared__ double sm[64];
    double vv[2] = { ke, ut };
    az_block_sum<2>(vv, sm);
    if (threadIdx.x == 0) { partKE[blockIdx.x] = 0.5 * vv[0]; if (TSTAT) partU[blockIdx.x] = vv[1]; }
    az_kick_drift_tail<MODE>(P, v2, eF, slot_next, sc);
}

// temp_scale (src/cuTemp.cu:77-107) [+ tstat_radi9] on an equilibration step.
template <bool TSTAT>
__global__ void __launch_bounds__(AZ_TPB)
az_k_tscale(const __grid_constant__ AzParams P, int step_k, const float4* __restrict__ pos,
            float4* __restrict__ vel, const int* __restrict__ id, float2* __restrict__ therm,
            const float* __restrict__ photons, const float* __restrict__ uv,
            const AzScalars* __restrict__ sc, double* __restrict__ partU, int use_k)
{
    const int step = AZ_STEP(P, step_k);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double ut = 0.0;
    float k = use_k ? sc->tscale_k : 1.f; // use_k == 0: thermostat only (a 'reset_vels' step: zero_vel, then apply_tstat)
    if (i < az_n_owned(P)) {
        float4 v = vel[i];
        v.x *= k; v.y *= k; v.z *= k;
        if (TSTAT) {
            int t = __float_as_int(pos[i].w);
            float2 th = therm[i];
            az_tstat_radi_atom(P, step, id[i], t, photons[i], uv, v.x, v.y, v.z, th.x, th.y);
            therm[i] = th;
            ut = (double)th.x;
        }
        vel[i] = v;
    }
    if (TSTAT) {
        __shared__ double sm[32];
        double vv[1] = { ut };
        az_block_sum<1>(vv, sm);
        if (threadIdx.x == 0) partU[blockIdx.x] = vv[0];
    }
}

// Fixed-order reduction of the per-block partials.  mode bits: 1 = KE, 2 = pair energies,
// 4 = thermal energy, 8 = this is an equilibration T-scale step (computes tscale_k, after_tscale),
// 16 = calc_quantities (src/main.cu:121-162) and publish, 32 = rebuild step, 64 = stale-cell step.
static __global__ void __launch_bounds__(1024)
az_k_finalize(const __grid_constant__ AzParams P, int step_k, int slot, int mode, const double* __restrict__ partKE, int nKE,
              const double* __restrict__ partV, const double* __restrict__ partC, int nPair,
              const double* __restrict__ partU, int nU, const unsigned long long* __restrict__ partN, const float4* __restrict__ pos, AzScalars* __restrict__ gsc,
              float revAx, float revAy, float revAz)
{
    const int step = AZ_STEP(P, step_k);
    __shared__ double sm[4 * 32];
    // The scalar block is staged in shared memory: the bookkeeping below is ~60 dependent read-modify-writes by one
    // thread, which cost a global-memory round trip each when done in place (16 us per step) and nothing here.
    __shared__ AzScalars ssc;
    AzScalars* sc = &ssc;
    {
        const int nw = (int)(sizeof(AzScalars) / sizeof(int));
        const int* src = reinterpret_cast<const int*>(gsc);
        int* dst = reinterpret_cast<int*>(&ssc);
        for (int k = threadIdx.x; k < nw; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    double v[4] = { 0, 0, 0, 0 };
    // eight independent loads in flight per thread: a plain `v += part[k]` loop is one DRAM round trip per element
    // (30 elements per thread at 4 M atoms = 20 us for this one-block kernel)
    auto strided_sum = [&](const double* __restrict__ part, int n) {
        double acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int k = threadIdx.x; k < n; k += 8 * blockDim.x) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const int idx = k + u * (int)blockDim.x; x[u] = idx < n ? part[idx] : 0.0; }
#pragma unroll
            for (int u = 0; u < 8; u++) acc[u] += x[u];
        }
        return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    };
    if (mode & 1) v[0] = strided_sum(partKE, nKE);
    if (mode & 2) { v[1] = strided_sum(partV, nPair); v[2] = strided_sum(partC, nPair); }
    if (mode & 4) v[3] = strided_sum(partU, nU);
    az_block_sum<4>(v, sm);
    if (mode & 2) { // pair visits (both sides) -> unique pairs; exact in double up to 2^53
        double c[1] = { 0.0 };
        for (int k = threadIdx.x; k < nPair; k += 4 * blockDim.x) {
            unsigned long long x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const int idx = k + u * (int)blockDim.x; x[u] = idx < nPair ? partN[idx] : 0ull; }
            c[0] += ((double)x[0] + (double)x[1]) + ((double)x[2] + (double)x[3]);
        }
        az_block_sum<1>(c, sm);
        if (threadIdx.x == 0) { sc->nPairs = (unsigned long long)(0.5 * c[0] + 0.25); sc->nVisits = (unsigned long long)(c[0] + 0.5); }
    }
    int new_flags = 0, ring_idx = -1;
    if (threadIdx.x == 0) {
    if (mode & (32 | 64)) { // this step's face crossings join the cumulative momenta (put_periodic, src/cuMDfunc.cu:72-104)
        for (int a = 0; a < 3; a++) {
            sc->posMom[a] += sc->posMomPend[slot][a]; sc->posMomPend[slot][a] = 0.0;
            sc->negMom[a] += sc->negMomPend[slot][a]; sc->negMomPend[slot][a] = 0.0;
        }
    }
    if (mode & 32) { sc->drift = 0.f; sc->vmax2_bits[slot] = 0u; } // rebuild step: cells are exact again
    if (mode & 64) { // stale-cell step: every atom moved by at most vmax*dt
        sc->drift += sqrtf(__uint_as_float(sc->vmax2_bits[slot])) * P.tSt;
        sc->vmax2_bits[slot] = 0u;
        if (2.f * sc->drift > P.skin) new_flags |= AZ_FLAG_SKIN;
    }
    if ((mode & 1) && !(v[0] == v[0])) new_flags |= AZ_FLAG_NAN;
    if (mode & 1) { sc->engKin = v[0]; sc->engKin_f = (float)v[0]; }
    if (mode & 2) { sc->engVdW = v[1]; sc->engCoul1 = v[2]; }
    if (mode & 4) sc->engTemp = v[3];
    if (mode & 8) {
        float k = 1.f;
        if (sc->engKin_f != 0.f) {
            float c = (P.tstat == AZ_TSTAT_RADI) ? 0.25f : 1.f;
            k = sqrtf(c * P.teKin / sc->engKin_f);
        }
        sc->tscale_k = k;
        sc->engKin_f = P.teKin; sc->engKin = (double)P.teKin; // after_tscale, src/cuTemp.cu:109-113
    }
    if (mode & 16) {
        float engCoulTot = (float)sc->engCoul1 + (float)sc->engCoul2 + (float)sc->engCoul3; // src/main.cu:127
        sc->engPot = engCoulTot + (float)sc->engVdW + (float)sc->engBond + (float)sc->engAngle; // src/main.cu:128
        sc->engTot = sc->engPot + sc->engKin_f;
        float pm[3] = { (float)sc->posMom[0], (float)sc->posMom[1], (float)sc->posMom[2] };
        float nm[3] = { (float)sc->negMom[0], (float)sc->negMom[1], (float)sc->negMom[2] };
        float ra[3] = { revAx, revAy, revAz };
        if (step >= AZ_N_MOM - 1) {
            int i = sc->iMom;
            float k = (float)(2.f * 1.58e6 / (double)(P.tSt * (float)(AZ_N_MOM - 1)));
            float tot = 0.f, pp[3], np[3];
            for (int a = 0; a < 3; a++) {
                pp[a] = k * (pm[a] - sc->posMomBuf[i][a]) * ra[a];
                np[a] = k * (nm[a] - sc->negMomBuf[i][a]) * ra[a];
            }
            i--; if (i < 0) i = AZ_N_MOM - 1;
            ring_idx = i;
            for (int a = 0; a < 3; a++) { sc->posMomBuf[i][a] = pm[a]; sc->negMomBuf[i][a] = nm[a]; }
            sc->iMom++; if (sc->iMom >= AZ_N_MOM) sc->iMom = 0;
            tot = (pp[0] + pp[1] + pp[2] + np[0] + np[1] + np[2]) / 6.f;
            sc->pressure = tot;
        } else if (step >= 0) {
            ring_idx = step;
            for (int a = 0; a < 3; a++) { sc->posMomBuf[step][a] = pm[a]; sc->negMomBuf[step][a] = nm[a]; }
        }
        sc->step = step;
        sc->x_first = pos[0].x;
    }
    // Write back ONLY what this kernel owns.  Counters and flags that other kernels update with atomics (face-crossing
    // counters, the pending momenta and the speed bound of OTHER slots, the sticky flags, the list diagnostics) are left
    // alone: the pass of the next step may be running beside this kernel (az_engine.cu: finalize on its own stream).
    if (mode & (32 | 64)) {
        for (int a = 0; a < 3; a++) {
            gsc->posMom[a] = sc->posMom[a]; gsc->negMom[a] = sc->negMom[a];
            gsc->posMomPend[slot][a] = 0.0; gsc->negMomPend[slot][a] = 0.0;
        }
        gsc->drift = sc->drift; gsc->vmax2_bits[slot] = 0u;
    }
    if (new_flags) atomicOr(&gsc->overflow, new_flags);
    if (mode & (1 | 8)) { gsc->engKin = sc->engKin; gsc->engKin_f = sc->engKin_f; }
    if (mode & 2) { gsc->engVdW = sc->engVdW; gsc->engCoul1 = sc->engCoul1; gsc->nPairs = sc->nPairs; gsc->nVisits = sc->nVisits; }
    if (mode & 4) gsc->engTemp = sc->engTemp;
    if (mode & 8) gsc->tscale_k = sc->tscale_k;
    if (mode & 16) {
        gsc->engPot = sc->engPot; gsc->engTot = sc->engTot; gsc->pressure = sc->pressure; gsc->iMom = sc->iMom;
        gsc->step = sc->step; gsc->x_first = sc->x_first;
        if (ring_idx >= 0)
            for (int a = 0; a < 3; a++) { gsc->posMomBuf[ring_idx][a] = sc->posMomBuf[ring_idx][a]; gsc->negMomBuf[ring_idx][a] = sc->negMomBuf[ring_idx][a]; }
    }
    }
}

// ---------------------------------------------------------------------------------------
// constant bonds (apply_const_bonds, src/cuBonds.cu:709-797; potentials :1123-1224) and valence angles
// (apply_angles + angle_hcos, src/cuAngles.cu:169-283).  The reference walks the bond list and scatters both forces of a
// bond with float atomics; here one thread owns one atom and walks the bonds / angles that touch it (a CSR by persistent
// id built at create time), so every force component is summed in a fixed order by the thread that owns it: no atomics,
// bitwise reproducible.  A bond is evaluated by both of its atoms from the same delta = x(at1) - x(at2), an angle by all
// three of its atoms; the energy is counted by the first atom / the centre.  `whereId` takes a persistent id to the
// atom's current place in the sorted streams (the reference rewrites the lists after every sort instead).
struct AzBondType { float p0, p1, p2, p3, p4; int kind; };
struct AzAngleType { float k, cos0; };
struct AzBondedArgs {
    const float4* pos; const int* id; const int* whereId; float4* frc;
    const int* bStart; const int2* bEnt;    // per persistent id: [bStart[id], bStart[id+1]) entries (partner id, type | first << 16)
    const AzBondType* bTypes;
    const int* aStart; const int4* aEnt;    // entries (id A, id B, type | role << 16, 0): role 0 = I am the centre (A, B = ligands 1, 2),
    const AzAngleType* aTypes;              // 1 = I am ligand 1 (A = centre, B = ligand 2), 2 = I am ligand 2 (A = centre, B = ligand 1)
    double *partB, *partA;
};

__device__ __forceinline__ float az_delta_periodic(float d, float L) // delta_periodic, src/cuMDfunc.cu:178-203
{
    const float h = 0.5f * L;
    if (d > h) d -= L; else if (d < -h) d += L;
    return d;
}

// force_eng of the five bond kinds: returns -dU/dr / r, adds U
__device__ __forceinline__ float az_bond_feng(const AzBondType& b, float r2, float r, float& eng)
{
    switch (b.kind) {
    case AZ_BOND_HARM: { const float x = r - b.p1; eng += 0.5f * b.p0 * x * x; return -b.p0 / r * x; }
    case AZ_BOND_MORSE: {
        float x = r - b.p2; x = expf(-b.p1 * x); const float y = 1.f - x;
        eng += b.p0 * y * y - b.p3;
        return (float)(-2.0 * b.p0 * b.p1 * x * y / r);
    }
    case AZ_BOND_PEDONE: {
        float x = r - b.p2; x = expf(-b.p1 * x); const float y = 1.f - x;
        const float ir2 = 1.f / r2; float irn = ir2 * ir2; irn = irn * irn * irn;
        eng += b.p0 * y * y - b.p3 - b.p4 * irn;
        return (float)(-2.0 * b.p0 * b.p1 * x * y / r - 12.0 * b.p4 * irn * ir2);
    }
    case AZ_BOND_BUCK: {
        const float ir2 = (float)(1.0 / r2), irn = ir2 * ir2, ex = expf(-r / b.p1);
        eng += b.p0 * ex - b.p2 * irn * ir2;
        return b.p0 * ex / r / b.p1 - 6.f * b.p2 * irn * irn;
    }
    case AZ_BOND_E612: {
        const float ir2 = (float)(1.0 / r2), irn = ir2 * ir2, ex = expf(-r / b.p1);
        eng += b.p0 * ex - b.p2 * irn * ir2 - b.p3 * irn * irn - b.p4 * irn * irn * irn;
        return (float)(b.p0 * ex / r / b.p1 - 6.f * b.p2 * irn * irn - 8.f * b.p3 * irn * irn * ir2 - 12.0 * b.p4 * irn * irn * irn * ir2);
    }
    }
    return 0.f;
}

static __global__ void __launch_bounds__(AZ_TPB)
az_k_bonded(const __grid_constant__ AzParams P, const __grid_constant__ AzBondedArgs A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double eb = 0.0, ea = 0.0;
    if (i < az_n_owned(P)) {
        const float4 p = A.pos[i];
        const int me = A.id[i];
        float fx = 0.f, fy = 0.f, fz = 0.f;
        if (A.bStart)
            for (int k = A.bStart[me]; k < A.bStart[me + 1]; k++) {
                const int2 en = A.bEnt[k];
                const float4 q = A.pos[A.whereId[en.x]];
                const bool first = (en.y >> 16) != 0;
                // delta = x(at1) - x(at2), whichever of the two I am: both ends then see the same F/r
                const float dx = az_delta_periodic(first ? p.x - q.x : q.x - p.x, P.L[0]);
                const float dy = az_delta_periodic(first ? p.y - q.y : q.y - p.y, P.L[1]);
                const float dz = az_delta_periodic(first ? p.z - q.z : q.z - p.z, P.L[2]);
                const float r2 = dx * dx + dy * dy + dz * dz, r = sqrtf(r2);
                float e = 0.f;
                const float f = az_bond_feng(A.bTypes[en.y & 0xffff], r2, r, e);
                if (first) { fx += f * dx; fy += f * dy; fz += f * dz; eb += (double)e; }
                else { fx -= f * dx; fy -= f * dy; fz -= f * dz; }
            }
        if (A.aStart)
            for (int k = A.aStart[me]; k < A.aStart[me + 1]; k++) {
                const int4 en = A.aEnt[k];
                const int role = en.z >> 16;
                const float4 qa = A.pos[A.whereId[en.x]], qb = A.pos[A.whereId[en.y]];
                // (centre, ligand 1, ligand 2)
                const float4 c = role == 0 ? p : qa, l1 = role == 0 ? qa : (role == 1 ? p : qb), l2 = role == 0 ? qb : (role == 1 ? qb : p);
                const float ijx = az_delta_periodic(l1.x - c.x, P.L[0]), ijy = az_delta_periodic(l1.y - c.y, P.L[1]), ijz = az_delta_periodic(l1.z - c.z, P.L[2]);
                const float ikx = az_delta_periodic(l2.x - c.x, P.L[0]), iky = az_delta_periodic(l2.y - c.y, P.L[1]), ikz = az_delta_periodic(l2.z - c.z, P.L[2]);
                const float r2ij = ijx * ijx + ijy * ijy + ijz * ijz, rij = sqrtf(r2ij);
                const float r2ik = ikx * ikx + iky * iky + ikz * ikz, rik = sqrtf(r2ik);
                const float cos_th = (ijx * ikx + ijy * iky + ijz * ikz) / rij / rik;
                const AzAngleType t = A.aTypes[en.z & 0xffff];
                const float dCos = cos_th - t.cos0;
                const float c1 = -t.k * dCos;
                const float c2 = (float)(1.0 / rij / rik);
                if (role == 0) {
                    fx += -c1 * (ikx * c2 + ijx * c2 - cos_th * (ijx / r2ij + ikx / r2ik));
                    fy += -c1 * (iky * c2 + ijy * c2 - cos_th * (ijy / r2ij + iky / r2ik));
                    fz += -c1 * (ikz * c2 + ijz * c2 - cos_th * (ijz / r2ij + ikz / r2ik));
                    ea += (double)(0.5f * t.k * dCos * dCos);
                } else if (role == 1) {
                    fx += c1 * (ikx * c2 - cos_th * ijx / r2ij); fy += c1 * (iky * c2 - cos_th * ijy / r2ij); fz += c1 * (ikz * c2 - cos_th * ijz / r2ij);
                } else {
                    fx += c1 * (ijx * c2 - cos_th * ikx / r2ik); fy += c1 * (ijy * c2 - cos_th * iky / r2ik); fz += c1 * (ijz * c2 - cos_th * ikz / r2ik);
                }
            }
        float4 f = A.frc[i];
        f.x += fx; f.y += fy; f.z += fz;
        A.frc[i] = f;
    }
    __shared__ double sm[64];
    double vv[2] = { eb, ea };
    az_block_sum<2>(vv, sm);
    if (threadIdx.x == 0) { A.partB[blockIdx.x] = vv[0]; A.partA[blockIdx.x] = vv[1]; }
}

// fixed-order sum of the per-block bonded energies (one block)
static __global__ void __launch_bounds__(256)
az_k_bonded_reduce(const double* __restrict__ partB, const double* __restrict__ partA, int n, AzScalars* __restrict__ sc)
{
    __shared__ double sm[64];
    double v[2] = { 0.0, 0.0 };
    for (int k = threadIdx.x; k < n; k += blockDim.x) { v[0] += partB[k]; v[1] += partA[k]; }
    az_block_sum<2>(v, sm);
    if (threadIdx.x == 0) { sc->engBond = v[0]; sc->engAngle = v[1]; }
}

// Nose-Hoover (src/cuTemp.cu:115-163): before_nose and after_nose(1) are pure scalar recurrences, so both run in one
// thread BEFORE the velocities are touched; az_k_vscale then applies md->tscale like tstat_nose.
// Bricks: the friction follows the kinetic energy of the WHOLE system.  glob = 1: take it from sc->keGlob (the value the
// previous application left behind); glob = 2: red[0] is this step's all-reduced kinetic energy, red[1] this brick's
// share.  Every rank runs the same recurrence on the same number, so chit and the scale factor agree bit for bit; the
// brick's own engKin stays its share (the statistics sum the shares).
static __global__ void az_k_nose_scalars(const __grid_constant__ AzParams P, AzScalars* __restrict__ sc, const double* __restrict__ red, int glob)
{
    float chit = sc->chit;
    float ke = glob == 2 ? (float)red[0] : (glob == 1 ? sc->keGlob : sc->engKin_f);
    float ke_own = glob == 2 ? (float)red[1] : sc->engKin_f;
    chit += P.tSt * (ke - P.teKin) * P.rQmass;             // before_nose
    const float ts = 1.f - P.tSt * chit;
    ke = ke * ts * ts;                                     // after_nose(refresh_kin = 1)
    ke_own = glob ? ke_own * ts * ts : ke;
    sc->consInt += P.tSt * chit * P.qMassTau2;
    chit += P.tSt * (ke - P.teKin) * P.rQmass;
    sc->chit = chit; sc->nose_scale = ts; sc->keGlob = ke; sc->engKin_f = ke_own; sc->engKin = (double)ke_own;
}

static __global__ void __launch_bounds__(AZ_TPB)
az_k_vscale(const __grid_constant__ AzParams P, float4* __restrict__ vel, const AzScalars* __restrict__ sc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float k = sc->nose_scale;
    if (i < az_n_owned(P)) { float4 v = vel[i]; v.x *= k; v.y *= k; v.z *= k; vel[i] = v; }
}

// zero_vel (src/cuMDfunc.cu:602-622): 'reset_vels' steps replace the second half kick; the kinetic energy becomes 0
static __global__ void __launch_bounds__(AZ_TPB)
az_k_zero_vel(const __grid_constant__ AzParams P, float4* __restrict__ vel, double* __restrict__ partKE)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < az_n_owned(P)) vel[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) partKE[blockIdx.x] = 0.0;
}

static __global__ void az_k_reset_step(AzScalars* sc) { sc->engElecField = 0.0; } // reset_quantities

static __global__ void __launch_bounds__(AZ_TPB)
az_k_gather_by_id(int n, const int* __restrict__ id, const float4* __restrict__ src, float4* __restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[id[i]] = src[i];
}


// ---------------------------------------------------------------------------------------
// az_engine_upload on the device: host SoA doubles (struct Atoms layout) staged in HBM are
// narrowed to float exactly like src/cuInit.cu:775-777 and gathered into the slot that
// currently holds each persistent id; forces are zeroed like a fresh init_cudaMD.
template <class T>
__global__ void __launch_bounds__(AZ_TPB)
az_k_upload(int n, const int* __restrict__ id, const T* __restrict__ xs, const T* __restrict__ ys,
            const T* __restrict__ zs, const T* __restrict__ vxs, const T* __restrict__ vys,
            const T* __restrict__ vzs, float4* __restrict__ pos, float4* __restrict__ vel, float4* __restrict__ frc)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = id[i];
    float4 p = pos[i];
    p.x = (float)xs[k]; p.y = (float)ys[k]; p.z = (float)zs[k];
    pos[i] = p;
    vel[i] = make_float4(vxs ? (float)vxs[k] : 0.f, vys ? (float)vys[k] : 0.f, vzs ? (float)vzs[k] : 0.f, 0.f);
    frc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// az_engine_download_f32: positions (wrapped like put_periodic) and velocities back in atoms.xyz order, as float SoA
static __global__ void __launch_bounds__(AZ_TPB)
az_k_download_f32(const __grid_constant__ AzParams P, int n, int wrap, const int* __restrict__ id, const float4* __restrict__ pos,
                  const float4* __restrict__ vel, float* __restrict__ out /* 6 arrays of n */, int with_vel)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = id[i];
    float4 p = pos[i];
    if (wrap) {
        float* c[3] = { &p.x, &p.y, &p.z };
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int nx = (int)floor((double)*c[a] * (double)P.revL[a]);
            float x = *c[a] - (float)nx * P.L[a];
            if (x >= P.L[a]) x = 0.f;
            *c[a] = x;
        }
    }
    out[k] = p.x; out[(size_t)n + k] = p.y; out[2 * (size_t)n + k] = p.z;
    if (with_vel) { const float4 v = vel[i]; out[3 * (size_t)n + k] = v.x; out[4 * (size_t)n + k] = v.y; out[5 * (size_t)n + k] = v.z; }
}

// FP32 FMA throughput probe: 8 independent FMA chains per thread, no memory traffic.
static __global__ void __launch_bounds__(256)
az_k_ffma_probe(float* out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int k = 0; k < iters; k++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
