// az_pair.cuh -- short-range pair-force kernel over the cell grid (sm_100a).
//
// Replaces cell_list5a (intra-cell, src/cuPairs.cu:2266-2394) and cell_list4b_noshared
// (inter-cell, src/cuPairs.cu:1474-1624) with one launch.  A CTA owns a run of `zc`
// consecutive z-cells of one (x,y) column; its atoms are contiguous in the sorted streams
// (cell id = x*cnYZ + y*cNumber.z + z, z fastest) and each thread keeps one i-atom in
// registers.  For every (x,y) neighbour column of the stencil the z-window of cells
// [z0-S, z0+zc-1+S] is staged once into shared memory as float4 (x,y,z,type) -- because z is
// the fastest cell index a window is at most a few contiguous pieces of the sorted stream --
// and every thread sweeps the sub-range of the window that belongs to its own cell's
// neighbours.  Full neighbour list, no Newton-3 atomics: each unordered pair is evaluated
// from both sides with a bitwise-negated delta, so per-atom sums equal the reference's
// f_A += f*delta / atomicAdd(f_B, -f*delta) up to summation order, with no atomics and a
// deterministic result.  The periodic shift is applied per cell pair exactly like the
// reference's secShifts (src/cuCellList.cu:209,238): delta = (x_i - x_j) - shift.
#pragma once
#include "az_kernels.cuh"

#define AZ_PAIR_TPB 128
#define AZ_PAIR_MAXW 96      // window entries (zc + 2*Sz) a CTA can index
#define AZ_PAIR_MAXPIECE AZ_PAIR_MAXW // contiguous pieces of one window (wraps / ghost layers)

struct AzPairArgs {
    const float4* pos;      // sorted positions (+type) of owned and ghost atoms
    const float2* therm;    // (U, radius) or NULL
    const int* cellOfSorted;// cell id of sorted atom (owned atoms)
    const int* first;       // [ncell_local+1]
    const int* count;       // [ncell_local]
    float4* frc;            // out
    double* partV;          // per-CTA energy partials
    double* partC;
    int zc;                 // z-run length per CTA
    int nruns;              // ceil(owned nz / zc)
    int tile_cap;           // float4 slots in the dynamic shared tile
    AzVdw single;           // the only table entry when nspec == 1
    const AzVdw* table;     // nspec*nspec entries (device memory) otherwise
};

// Unwrapped neighbour coordinate -> storage coordinate + image sign; reference image rule
// (cell_dist_int: wrap iff delt > mx/2 with integer division; each cell pair exists once).
__device__ __forceinline__ bool az_neighbor(const AzGrid& g, int a, int cglob, int o, int& store, int& sh)
{
    int n = g.n[a];
    int ao = o < 0 ? -o : o;
    if (ao > n / 2) return false;
    int u = cglob + o;
    sh = (u < 0) ? -1 : ((u >= n) ? 1 : 0);
    if ((n & 1) == 0 && ao == n / 2 && sh != 0) return false;
    int w = u - sh * n; // wrapped global coordinate
    if (ao != 0 && w == cglob) return false;
    store = g.dec[a] ? (u - g.goff[a]) : w;
    return true;
}

template <int VK, int CK, bool SINGLE, bool RADII>
__global__ void __launch_bounds__(AZ_PAIR_TPB)
az_k_pair(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    extern __shared__ float4 s_tile[];
    __shared__ float s_rad[1]; // placeholder to keep layout simple when !RADII
    __shared__ int s_cs[AZ_PAIR_MAXW + 1];     // prefix of atom counts over window entries
    __shared__ float s_shz[AZ_PAIR_MAXW];      // z image shift of each window entry
    __shared__ int s_pc_g[AZ_PAIR_MAXPIECE], s_pc_c[AZ_PAIR_MAXPIECE], s_pc_n[AZ_PAIR_MAXPIECE];
    __shared__ int s_np;
    __shared__ AzVdw s_tab[SINGLE ? 1 : AZ_MAX_SPECIES * AZ_MAX_SPECIES];
    __shared__ double s_red[64];
    (void)s_rad;

    const AzGrid& g = P.g;
    const int tid = threadIdx.x;
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
    for (int ib = iBeg; ib < iEnd; ib += AZ_PAIR_TPB) {
        const int i = ib + tid;
        const bool act = i < iEnd;
        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
        int lzi = lz0, ti = 0;
        float radi = 0.f;
        if (act) {
            pi = A.pos[i];
            ti = __float_as_int(pi.w);
            lzi = A.cellOfSorted[i] % g.nl[2];
            if (RADII) radi = A.therm[i].y;
        }
        const int gzi = lzi + g.goff[2];
        float fx = 0.f, fy = 0.f, fz = 0.f, ev = 0.f, ec = 0.f;

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
                if (mxy2 > P.r2max) continue;
                const float shx = (float)shxi * P.L[0], shy = (float)shyi * P.L[1];
                const bool centre = (ox == 0 && oy == 0);
                const int colbase = (sx * g.nl[1] + sy) * g.nl[2];

                __syncthreads(); // previous tile fully consumed; window tables free
                // ---- window table: entry w <-> unwrapped z = gz0 - Sz + w
                if (tid < W) {
                    int u = gz0 - Sz + tid;
                    int n = g.n[2];
                    int sh = 0, cnt = 0;
                    // images beyond one box length cannot be valid neighbours (|o| <= n/2)
                    if (u < -n || u >= 2 * n) { sh = 0; cnt = 0; }
                    else {
                        sh = (u < 0) ? -1 : ((u >= n) ? 1 : 0);
                        int st = g.dec[2] ? (u - g.goff[2]) : (u - sh * n);
                        if (st >= 0 && st < g.nl[2]) cnt = A.count[colbase + st];
                    }
                    s_cs[tid + 1] = cnt;
                    s_shz[tid] = (float)sh * P.L[2];
                }
                if (tid == 0) s_cs[0] = 0;
                __syncthreads();
                if (tid == 0) {
                    // prefix + contiguous pieces (consecutive storage z => contiguous in memory)
                    int acc = 0, np = 0;
                    int prev_st = -1000000;
                    for (int w = 0; w < W; w++) {
                        int cnt = s_cs[w + 1];
                        int u = gz0 - Sz + w, n = g.n[2];
                        int sh = (u < 0) ? -1 : ((u >= n) ? 1 : 0);
                        int st = g.dec[2] ? (u - g.goff[2]) : (u - sh * n);
                        bool valid = !(u < -n || u >= 2 * n) && st >= 0 && st < g.nl[2];
                        if (valid && cnt > 0) {
                            if (np > 0 && st == prev_st + 1 && s_pc_g[np - 1] + s_pc_n[np - 1] == A.first[colbase + st]) {
                                s_pc_n[np - 1] += cnt;
                            } else {
                                s_pc_g[np] = A.first[colbase + st]; s_pc_c[np] = acc; s_pc_n[np] = cnt; np++;
                            }
                            prev_st = st;
                        }
                        s_cs[w] = acc;
                        acc += cnt;
                    }
                    s_cs[W] = acc;
                    s_np = np;
                }
                __syncthreads();
                const int tot = s_cs[W];
                // this lane's sub-range of the window: its cell's z neighbours
                int lo = 0, hi = 0, selfj = -1;
                if (act) {
                    int wl = -1, wh = -1;
                    for (int oz = -Sz; oz <= Sz; oz++) {
                        int st, sh;
                        if (!az_neighbor(g, 2, gzi, oz, st, sh)) continue;
                        // cell-pair existence: sum of per-axis rmin^2 <= rmax^2 (pair_exists_shift)
                        int aoz = oz < 0 ? -oz : oz;
                        float mz = aoz ? (float)(aoz - 1) * g.csize[2] : 0.f;
                        if (mxy2 + mz * mz > P.r2max) continue;
                        int w = (gzi + oz) - (gz0 - Sz);
                        if (wl < 0) wl = w;
                        wh = w;
                    }
                    if (wl >= 0) { lo = s_cs[wl]; hi = s_cs[wh + 1]; }
                    if (centre) selfj = s_cs[gzi - (gz0 - Sz)] + (i - A.first[A.cellOfSorted[i]]);
                }
                // ---- stage the window in chunks of tile_cap atoms and sweep
                for (int c0 = 0; c0 < tot; c0 += A.tile_cap) {
                    if (c0 > 0) __syncthreads();
                    const int c1 = min(tot, c0 + A.tile_cap);
                    for (int p = 0; p < s_np; p++) {
                        int pc = s_pc_c[p], pn = s_pc_n[p], pg = s_pc_g[p];
                        int a = max(pc, c0), b = min(pc + pn, c1);
                        for (int k = a + tid; k < b; k += AZ_PAIR_TPB) {
                            s_tile[k - c0] = A.pos[pg + (k - pc)];
                            if (RADII) s_tile_rad[k - c0] = A.therm[pg + (k - pc)].y;
                        }
                    }
                    __syncthreads();
                    const int jl = max(lo, c0), jh = min(hi, c1);
                    // window entry of jl, to track the z shift as j advances
                    int w = 0;
                    if (jl < jh) { while (s_cs[w + 1] <= jl) w++; }
                    int wend = (jl < jh) ? s_cs[w + 1] : 0;
                    float shz = (jl < jh) ? s_shz[w] : 0.f;
                    for (int j = jl; j < jh; j++) {
                        while (j >= wend) { w++; wend = s_cs[w + 1]; shz = s_shz[w]; }
                        const float4 pj = s_tile[j - c0];
                        const float dx = pi.x - pj.x - shx;
                        const float dy = pi.y - pj.y - shy;
                        const float dz = pi.z - pj.z - shz;
                        const float r2 = dx * dx + dy * dy + dz * dz;
                        if (r2 <= P.r2max && j != selfj) {
                            const int tj = __float_as_int(pj.w);
                            float radj = 0.f;
                            if (RADII) radj = s_tile_rad[j - c0];
                            float f;
                            if (SINGLE) f = az_pair<VK, CK>(P, A.single, r2, radi, radj, true, ev, ec);
                            else f = az_pair<VK, CK>(P, s_tab[ti * P.nspec + tj], r2, radi, radj, ti == tj, ev, ec);
                            fx += f * dx; fy += f * dy; fz += f * dz;
                        }
                    }
                }
            }
        }
        if (act) {
            float q = P.charge[ti];
            // verlet_1stage leaves f = -q*E before the pair kernels add to it (src/cuMDfunc.cu:474)
            A.frc[i] = make_float4(-q * P.efield[0] + fx, -q * P.efield[1] + fy, -q * P.efield[2] + fz, 0.f);
            eV += (double)ev; eC += (double)ec;
        }
    }
    __syncthreads();
    double vv[2] = { 0.5 * eV, 0.5 * eC }; // every pair was visited from both sides
    az_block_sum<2>(vv, s_red);
    if (tid == 0) { A.partV[blockIdx.x] = vv[0]; A.partC[blockIdx.x] = vv[1]; }
}
