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

template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
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
        int hits = 0;

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
                        if (tid == 0) atomicOr(A.flags, 2);
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
                        for (int j = jb; j < je; j++) {
                            const float4 pj = s_tile[j];
                            const float dx = pi.x - pj.x - shx;
                            const float dy = pi.y - pj.y - shy;
                            const float dz = pi.z - pj.z - shz;
                            const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy)); // nvcc's contraction of dx*dx+dy*dy+dz*dz
                            if (valid && r2 <= P.r2max && j != self) {
                                const int tj = __float_as_int(pj.w);
                                float radj = 0.f;
                                if (RADII) radj = s_tile_rad[j];
                                float e1 = 0.f, e2 = 0.f;
                                float f;
                                if (SINGLE) f = az_pair<VK, CK>(P, A.single, r2, radi, radj, true, e1, e2);
                                else f = az_pair<VK, CK>(P, s_tab[ti * P.nspec + tj], r2, radi, radj, ti == tj, e1, e2);
                                fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
                                if (ENERGY) { ev += e1; ec += e2; hits++; }
                            }
                        }
                    }
                    w0 = w1;
                    if (w0 < W) __syncthreads(); // tile is about to be overwritten
                }
            }
        }
        if (act) {
            const float q = P.charge[ti];
            // verlet_1stage leaves f = -q*E before the pair kernels add to it (src/cuMDfunc.cu:474)
            A.frc[i] = make_float4(-q * P.efield[0] + fx, -q * P.efield[1] + fy, -q * P.efield[2] + fz, 0.f);
            eV += (double)ev; eC += (double)ec; nHit += (unsigned long long)hits;
        }
    }
    if (ENERGY) {
        __syncthreads();
        double vv[3] = { 0.5 * eV, 0.5 * eC, (double)nHit }; // every pair was visited from both sides
        az_block_sum<3>(vv, s_red);
        if (tid == 0) { A.partV[blockIdx.x] = vv[0]; A.partC[blockIdx.x] = vv[1]; A.partN[blockIdx.x] = (unsigned long long)(vv[2] + 0.5); }
    }
}
