// az_pair.cuh -- short-range pair-force kernel over the cell grid (sm_100a).
//
// Replaces cell_list5a (intra-cell, src/cuPairs.cu:2266-2394) and cell_list4b_noshared
// (inter-cell, src/cuPairs.cu:1474-1624) with one launch.  A CTA owns a run of `zc`
// consecutive z-cells of one (x,y) column; its atoms are contiguous in the sorted streams
// (cell id = x*cnYZ + y*cNumber.z + z, z fastest) and each thread keeps one i-atom in
// registers.  For every (x,y) neighbour column of the stencil the z-window of cells
// [z0-Sz, z0+zc-1+Sz] is staged once into shared memory as float4 (x,y,z,type) and swept cell by
// cell: all lanes of a warp read the same j-atom (shared-memory broadcast), the lane's own
// cell-pair existence (cell_dist_int / pair_exists_shift, src/cuCellList.cu:98-132,213-240) is a
// predicate, so the inner loop is warp-uniform and branch-free.  Full neighbour list, no
// Newton-3 atomics: each unordered pair is evaluated from both sides with a bitwise-negated
// delta, so per-atom sums equal the reference's f_A += f*delta / atomicAdd(f_B, -f*delta) up to
// summation order, with no atomics and a deterministic result.  The periodic shift is applied
// per cell pair exactly like the reference's secShifts (src/cuCellList.cu:209,238):
// delta = (x_i - x_j) - shift.
#pragma once
#include "az_kernels.cuh"

#define AZ_PAIR_TPB 128
#define AZ_PAIR_MAXW 64 // window entries (zc + 2*Sz) a CTA can index

struct AzPairArgs {
    const float4* pos;       // sorted positions (+type) of owned and ghost atoms
    const float2* therm;     // (U, radius) or NULL
    const int* cellOfSorted; // local cell id of sorted atom (owned atoms)
    const int* first;        // [ncell_local+1]
    const int* count;        // [ncell_local]
    float4* frc;             // out
    double* partV;           // per-CTA energy partials
    double* partC;
    unsigned long long* partN; // per-CTA count of in-range pair visits (ENERGY passes only)
    int* flags;              // sticky error flags (AzScalars::overflow)
    int* nbr;                // LIST build: neighbour codes, row k of atom i at nbr[k * nbr_stride + i]
    int* nbrCount;           // LIST build: neighbours of atom i within the list radius
    int nbr_stride, nbr_cap;
    int zc;                  // z-run length per CTA
    int nruns;               // ceil(owned nz / zc)
    int tile_cap;            // atoms the dynamic shared tile holds
    AzVdw single;            // the only table entry when nspec == 1
    const AzVdw* table;      // nspec*nspec entries (device memory) otherwise
};

// Unwrapped neighbour coordinate -> storage coordinate + image sign; reference image rule
// (cell_dist_int: wrap iff delt > mx/2 with integer division; each cell pair exists once).
__device__ __forceinline__ bool az_neighbor(const AzGrid& g, int a, int cglob, int o, int& store, int& sh)
{
    int n = g.n[a];
    int ao = o < 0 ? -o : o;
    sh = 0; store = 0;
    if (ao > n / 2) return false;
    int u = cglob + o;
    sh = (u < 0) ? -1 : ((u >= n) ? 1 : 0);
    if ((n & 1) == 0 && ao == n / 2 && sh != 0) return false;
    int w = u - sh * n; // wrapped global coordinate
    if (ao != 0 && w == cglob) return false;
    store = g.dec[a] ? (u - g.goff[a]) : w;
    return true;
}

// Neighbour code of the Verlet-list path: sorted index of j in the low 26 bits, image signs
// (0 none, 1 = -L, 2 = +L per axis) above them.
#define AZ_NBR_BITS 26
#define AZ_NBR_MASK ((1 << AZ_NBR_BITS) - 1)
__device__ __forceinline__ int az_shift_code(int shx, int shy, int shz)
{
    return ((shx < 0 ? 1 : (shx > 0 ? 2 : 0)) | ((shy < 0 ? 1 : (shy > 0 ? 2 : 0)) << 2) | ((shz < 0 ? 1 : (shz > 0 ? 2 : 0)) << 4));
}

// One window entry (cell) swept against the lane's i-atom: branch-free, warp-uniform trip count.
// SHIFT=false is the interior case (all three image shifts are zero: (xi - xj) - 0 == xi - xj exactly).
// `lim` is the lane's gate: r2Max (folded with r2cut when there is a single table entry and no
// Coulomb) if the lane's cell pairs with this window entry, -1 otherwise.
// LIST=true only records the neighbours inside the list radius (no pair arithmetic): the rebuild
// step of the Verlet-list path; the forces of that step come from az_k_pair_list like all others.
template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY, bool LIST, bool SHIFT>
__device__ __forceinline__ void az_sweep_cell(const AzParams& P, const AzPairArgs& A, const AzVdw* __restrict__ s_tab,
                                              const float4* __restrict__ s_tile, const float* __restrict__ s_tile_rad,
                                              int jb, int je, int self, float lim, const float4 pi, int ti, float radi,
                                              float shx, float shy, float shz, int i, int gbase, int code,
                                              float& fx, float& fy, float& fz, float& ev, float& ec, int& hits, int& nnb)
{
    constexpr bool FOLD = SINGLE && CK == AZ_ELEC_NONE;
#pragma unroll 4
    for (int j = jb; j < je; j++) {
        const float4 pj = s_tile[j];
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        if (SHIFT) { dx -= shx; dy -= shy; dz -= shz; }
        const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy)); // nvcc's contraction of dx*dx+dy*dy+dz*dz
        const bool in = r2 <= lim && j != self;
        if (LIST) {
            if (in) {
                if (nnb < A.nbr_cap) A.nbr[(size_t)nnb * A.nbr_stride + i] = (gbase + (j - jb)) | (code << AZ_NBR_BITS);
                nnb++;
            }
        } else {
            const int tj = __float_as_int(pj.w);
            float radj = 0.f;
            if (RADII) radj = s_tile_rad[j];
            float e1 = 0.f, e2 = 0.f, f;
            if (SINGLE) f = az_pair<VK, CK, FOLD>(P, A.single, r2, radi, radj, true, e1, e2);
            else f = az_pair<VK, CK, false>(P, s_tab[ti * P.nspec + tj], r2, radi, radj, ti == tj, e1, e2);
            f = in ? f : 0.f;
            fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
            if (ENERGY) { ev += in ? e1 : 0.f; ec += in ? e2 : 0.f; hits += in ? 1 : 0; }
        }
    }
}

template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY, bool LIST>
__global__ void __launch_bounds__(AZ_PAIR_TPB)
az_k_pair(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    AZ_DYN_SMEM(float4, s_tile);
    __shared__ int s_cs[AZ_PAIR_MAXW + 1]; // prefix of atom counts over window entries
    __shared__ int s_g[AZ_PAIR_MAXW];      // global index of the first atom of each window entry
    __shared__ float s_shz[AZ_PAIR_MAXW];  // z image shift of each window entry
    __shared__ AzVdw s_tab[SINGLE ? 1 : AZ_MAX_SPECIES * AZ_MAX_SPECIES];
    __shared__ double s_red[96];

    const AzGrid& g = P.g;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // CTA -> (owned column, z-run)
    const int ony = g.ohi[1] - g.olo[1];
    const int run = blockIdx.x % A.nruns;
    const int col = blockIdx.x / A.nruns;
    const int lcx = g.olo[0] + col / ony, lcy = g.olo[1] + col % ony; // local column coords
    const int lz0 = g.olo[2] + run * A.zc;
    const int zcnt = min(A.zc, g.ohi[2] - lz0);
    const int gcx = lcx + g.goff[0], gcy = lcy + g.goff[1], gz0 = lz0 + g.goff[2];
    const int Sz = g.sten[2];
    const int W = zcnt + 2 * Sz;

    if (!SINGLE) {
        for (int k = tid; k < P.nspec * P.nspec; k += AZ_PAIR_TPB) s_tab[k] = A.table[k];
    }

    const int cell0 = (lcx * g.nl[1] + lcy) * g.nl[2] + lz0;
    const int iBeg = A.first[cell0];
    const int iEnd = A.first[cell0 + zcnt - 1] + A.count[cell0 + zcnt - 1];

    float* s_tile_rad = (float*)(s_tile + A.tile_cap); // radii of the staged atoms (RADII only)

    double eV = 0.0, eC = 0.0;
    unsigned long long nHit = 0;
    // the pair gate: list radius when recording neighbours, else r2Max (folded with the only r2cut when possible)
    const float gate = LIST ? P.r2list : ((SINGLE && CK == AZ_ELEC_NONE) ? fminf(P.r2max, A.single.r2cut) : P.r2max);
    for (int ib = iBeg; ib < iEnd; ib += AZ_PAIR_TPB) {
        const int i = ib + tid;
        const bool act = i < iEnd;
        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
        int lzi = lz0, ti = 0, rank_i = 0;
        float radi = 0.f;
        if (act) {
            pi = A.pos[i];
            ti = __float_as_int(pi.w);
            int ci = A.cellOfSorted[i];
            lzi = ci % g.nl[2];
            rank_i = i - A.first[ci];
            if (RADII) radi = A.therm[i].y;
        }
        const int gzi = lzi + g.goff[2];
        float fx = 0.f, fy = 0.f, fz = 0.f, ev = 0.f, ec = 0.f;
        int hits = 0, nnb = 0;

        for (int ox = -g.sten[0]; ox <= g.sten[0]; ox++) {
            int sx, shxi;
            if (!az_neighbor(g, 0, gcx, ox, sx, shxi)) continue;
            const int aox = ox < 0 ? -ox : ox;
            const float mx = aox ? (float)(aox - 1) * g.csize[0] : 0.f;
            for (int oy = -g.sten[1]; oy <= g.sten[1]; oy++) {
                int sy, shyi;
                if (!az_neighbor(g, 1, gcy, oy, sy, shyi)) continue;
                const int aoy = oy < 0 ? -oy : oy;
                const float my = aoy ? (float)(aoy - 1) * g.csize[1] : 0.f;
                const float mxy2 = mx * mx + my * my;
                if (mxy2 > P.r2list) continue;
                const float shx = (float)shxi * P.L[0], shy = (float)shyi * P.L[1];
                const bool centre = (ox == 0 && oy == 0);
                const int colbase = (sx * g.nl[1] + sy) * g.nl[2];

                __syncthreads(); // previous tile fully consumed; window tables free
                // ---- window table: entry w <-> unwrapped global z  u = gz0 - Sz + w
                if (tid < W) {
                    const int u = gz0 - Sz + tid;
                    const int n = g.n[2];
                    int sh = 0, cnt = 0, gi = 0;
                    if (u >= -n && u < 2 * n) { // farther images are never valid neighbours
                        sh = (u < 0) ? -1 : ((u >= n) ? 1 : 0);
                        const int st = g.dec[2] ? (u - g.goff[2]) : (u - sh * n);
                        if (st >= 0 && st < g.nl[2]) { cnt = A.count[colbase + st]; gi = A.first[colbase + st]; }
                    }
                    s_cs[tid + 1] = cnt;
                    s_g[tid] = gi;
                    s_shz[tid] = (float)sh * P.L[2];
                }
                __syncthreads();
                if (warp == 0) { // inclusive scan of the counts (W <= 64: two items per lane)
                    int a0 = (lane < W) ? s_cs[lane + 1] : 0;
                    int a1 = (lane + 32 < W) ? s_cs[lane + 33] : 0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        int t0 = __shfl_up_sync(0xffffffffu, a0, o);
                        int t1 = __shfl_up_sync(0xffffffffu, a1, o);
                        if (lane >= o) { a0 += t0; a1 += t1; }
                    }
                    int tot0 = __shfl_sync(0xffffffffu, a0, 31);
                    if (lane < W) s_cs[lane + 1] = a0;
                    if (lane + 32 < W) s_cs[lane + 33] = a1 + tot0;
                    if (lane == 0) s_cs[0] = 0;
                }
                __syncthreads();

                // ---- stage the window (in chunks of whole cells if it exceeds the tile) and sweep
                int w0 = 0;
                while (w0 < W) {
                    const int base = s_cs[w0];
                    int w1 = w0;
                    while (w1 < W && s_cs[w1 + 1] - base <= A.tile_cap) w1++;
                    if (w1 == w0) { // one cell larger than the tile: report, skip it
                        if (tid == 0) atomicOr(A.flags, AZ_FLAG_TILE);
                        w0++;
                        continue;
                    }
                    for (int w = w0 + warp; w < w1; w += AZ_PAIR_TPB / 32) {
                        const int b = s_cs[w] - base, cnt = s_cs[w + 1] - s_cs[w], gi = s_g[w];
                        for (int k = lane; k < cnt; k += 32) {
                            s_tile[b + k] = A.pos[gi + k];
                            if (RADII) s_tile_rad[b + k] = A.therm[gi + k].y;
                        }
                    }
                    __syncthreads();
                    for (int w = w0; w < w1; w++) {
                        const int jb = s_cs[w] - base, je = s_cs[w + 1] - base;
                        if (jb == je) continue;
                        // does the cell pair (lane's cell, window entry w) exist?
                        const int oz = (gz0 - Sz + w) - gzi;
                        bool valid = false;
                        if (act) {
                            int st, sh;
                            if (az_neighbor(g, 2, gzi, oz, st, sh)) {
                                const int aoz = oz < 0 ? -oz : oz;
                                const float mz = aoz ? (float)(aoz - 1) * g.csize[2] : 0.f;
                                valid = (mxy2 + mz * mz <= P.r2list);
                            }
                        }
                        if (!__any_sync(0xffffffffu, valid)) continue;
                        const float shz = s_shz[w];
                        const int self = (centre && oz == 0 && valid) ? jb + rank_i : -1;
                        const float lim = valid ? gate : -1.f;
                        const int shzi = (shz < 0.f) ? -1 : ((shz > 0.f) ? 1 : 0);
                        const int code = LIST ? az_shift_code(shxi, shyi, shzi) : 0;
                        if (shxi == 0 && shyi == 0 && shzi == 0)
                            az_sweep_cell<VK, CK, SINGLE, RADII, ENERGY, LIST, false>(P, A, s_tab, s_tile, s_tile_rad, jb, je, self, lim, pi, ti, radi,
                                                                                      0.f, 0.f, 0.f, i, s_g[w], code, fx, fy, fz, ev, ec, hits, nnb);
                        else
                            az_sweep_cell<VK, CK, SINGLE, RADII, ENERGY, LIST, true>(P, A, s_tab, s_tile, s_tile_rad, jb, je, self, lim, pi, ti, radi,
                                                                                     shx, shy, shz, i, s_g[w], code, fx, fy, fz, ev, ec, hits, nnb);
                    }
                    w0 = w1;
                    if (w0 < W) __syncthreads(); // tile is about to be overwritten
                }
            }
        }
        if (act) {
            if (LIST) {
                A.nbrCount[i] = min(nnb, A.nbr_cap);
                if (nnb > A.nbr_cap) atomicOr(A.flags, AZ_FLAG_NBR);
            } else {
                const float q = P.charge[ti];
                // verlet_1stage leaves f = -q*E before the pair kernels add to it (src/cuMDfunc.cu:474)
                A.frc[i] = make_float4(-q * P.efield[0] + fx, -q * P.efield[1] + fy, -q * P.efield[2] + fz, 0.f);
                eV += (double)ev; eC += (double)ec; nHit += (unsigned long long)hits;
            }
        }
    }
    if (ENERGY && !LIST) {
        __syncthreads();
        double vv[3] = { 0.5 * eV, 0.5 * eC, (double)nHit }; // every pair was visited from both sides
        az_block_sum<3>(vv, s_red);
        if (tid == 0) { A.partV[blockIdx.x] = vv[0]; A.partC[blockIdx.x] = vv[1]; A.partN[blockIdx.x] = (unsigned long long)(vv[2] + 0.5); }
    }
}


// ---------------------------------------------------------------------------------------
// Verlet-list force pass, used between cell-list rebuilds (opts.rebuild_every > 1, BASELINE config 3):
// the rebuild step's tile pass (LIST=true above) stored, per atom, every j within rc + skin at
// nbr[k * stride + i] (coalesced over i); the stale steps gather those j directly instead of sweeping
// 27 cells.  One thread per atom, no atomics, same pair arithmetic and the same r2 <= r2Max gate on
// current positions, so the pair set per step is identical to the every-step-rebuild schedule as
// long as no atom moved more than skin/2 (tracked by az_k_integrate_a<false> / az_k_finalize).
#define AZ_LIST_TPB 128
template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
__global__ void __launch_bounds__(AZ_LIST_TPB)
az_k_pair_list(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    __shared__ AzVdw s_tab[SINGLE ? 1 : AZ_MAX_SPECIES * AZ_MAX_SPECIES];
    __shared__ double s_red[96];
    const int tid = threadIdx.x;
    if (!SINGLE) {
        for (int k = tid; k < P.nspec * P.nspec; k += AZ_LIST_TPB) s_tab[k] = A.table[k];
        __syncthreads();
    }
    const int i = blockIdx.x * AZ_LIST_TPB + tid;
    const bool act = i < P.n_local;
    float fx = 0.f, fy = 0.f, fz = 0.f, ev = 0.f, ec = 0.f;
    int hits = 0, ti = 0;
    if (act) {
        const float4 pi = A.pos[i];
        ti = __float_as_int(pi.w);
        float radi = 0.f;
        if (RADII) radi = A.therm[i].y;
        const int cnt = A.nbrCount[i];
        const int* __restrict__ row = A.nbr + i;
        const size_t stride = (size_t)A.nbr_stride;
        constexpr bool FOLD = SINGLE && CK == AZ_ELEC_NONE;
        const float gate = FOLD ? fminf(P.r2max, A.single.r2cut) : P.r2max;
        // one neighbour: decode, (rare) image shift, pair term, masked accumulate
        auto one = [&](const int c, const float4 pj, const float radj) {
            float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            if ((unsigned)c > (unsigned)AZ_NBR_MASK) { // image signs above the index bits; zero for every interior pair
                const int sc = (int)((unsigned)c >> AZ_NBR_BITS);
                dx -= ((float)((sc >> 1) & 1) - (float)(sc & 1)) * P.L[0];
                dy -= ((float)((sc >> 3) & 1) - (float)((sc >> 2) & 1)) * P.L[1];
                dz -= ((float)((sc >> 5) & 1) - (float)((sc >> 4) & 1)) * P.L[2];
            }
            const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            const bool in = r2 <= gate;
            const int tj = __float_as_int(pj.w);
            float e1 = 0.f, e2 = 0.f, f;
            if (SINGLE) f = az_pair<VK, CK, FOLD>(P, A.single, r2, radi, radj, true, e1, e2);
            else f = az_pair<VK, CK, false>(P, s_tab[ti * P.nspec + tj], r2, radi, radj, ti == tj, e1, e2);
            f = in ? f : 0.f;
            fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
            if (ENERGY) { ev += in ? e1 : 0.f; ec += in ? e2 : 0.f; hits += in ? 1 : 0; }
        };
        // software pipeline by hand: the index loads of chunk k+1 (HBM stream) are in flight while the
        // gathers of chunk k (L1/L2) are consumed
        int k = 0;
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (cnt >= 4) { c0 = row[0]; c1 = row[stride]; c2 = row[2 * stride]; c3 = row[3 * stride]; }
        for (; k + 4 <= cnt; k += 4) {
            const float4 p0 = A.pos[c0 & AZ_NBR_MASK], p1 = A.pos[c1 & AZ_NBR_MASK];
            const float4 p2 = A.pos[c2 & AZ_NBR_MASK], p3 = A.pos[c3 & AZ_NBR_MASK];
            float r0 = 0.f, r1 = 0.f, r2_ = 0.f, r3 = 0.f;
            if (RADII) { r0 = A.therm[c0 & AZ_NBR_MASK].y; r1 = A.therm[c1 & AZ_NBR_MASK].y; r2_ = A.therm[c2 & AZ_NBR_MASK].y; r3 = A.therm[c3 & AZ_NBR_MASK].y; }
            int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
            if (k + 8 <= cnt) {
                n0 = row[(size_t)(k + 4) * stride]; n1 = row[(size_t)(k + 5) * stride];
                n2 = row[(size_t)(k + 6) * stride]; n3 = row[(size_t)(k + 7) * stride];
            }
            one(c0, p0, r0); one(c1, p1, r1); one(c2, p2, r2_); one(c3, p3, r3);
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        }
        for (; k < cnt; k++) {
            const int c = row[(size_t)k * stride];
            const float4 pj = A.pos[c & AZ_NBR_MASK];
            float radj = 0.f;
            if (RADII) radj = A.therm[c & AZ_NBR_MASK].y;
            one(c, pj, radj);
        }
        const float q = P.charge[ti];
        A.frc[i] = make_float4(-q * P.efield[0] + fx, -q * P.efield[1] + fy, -q * P.efield[2] + fz, 0.f);
    }
    if (ENERGY) {
        double vv[3] = { 0.5 * (double)ev, 0.5 * (double)ec, (double)hits };
        az_block_sum<3>(vv, s_red);
        if (tid == 0) { A.partV[blockIdx.x] = vv[0]; A.partC[blockIdx.x] = vv[1]; A.partN[blockIdx.x] = (unsigned long long)(vv[2] + 0.5); }
    }
}
