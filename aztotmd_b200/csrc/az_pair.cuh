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
    int* nbr;                // Verlet list, blocked by warp: row k of atom i at nbr[((i >> 5) * nbr_cap + k) * 32 + (i & 31)]
    int* nbrCount;           // LIST build: neighbours of atom i within the list radius
    int nbr_stride, nbr_cap;
    // tile lists (az_tlist.cuh): 16-bit rows that index the tile's footprint, two per word; per tile the number of runs and
    // the runs (first atom in the sorted stream, offset in the footprint), closed by the entry (0, footprint size)
    unsigned* nbr16;
    int* tileHdr;
    int4* tileRuns;          // [tile][AZ_TL_RUNS]: (first atom in the sorted stream, place in the footprint, length, 0)
    int* tlStat;             // [0] largest footprint, [1] most runs, [2] most candidate ranges, [3] longest row column seen by a rebuild
    int tl_rows_smem;        // az_k_build_tlist: rows per thread held in shared memory (even)
    int tl_foot_cap;         // az_k_pair_tlist: atoms the shared-memory footprint holds
    int tl_stride;           // words per atom of that footprint (3: x y z; 5: + type, radius); a row holds stride * place
    int n_total;             // atoms in the sorted stream (owned + ghosts: capacity on bricks)
    // az_k_pair_list with the integrator riding on it (fuse != 0): second half kick of this step, kinetic energy, and the
    // next step's kick + drift (+ wrap, + bin) applied to atom i right after its force is known; the new position and
    // velocity go to the alternate buffers (neighbours still read the current ones), forces are never written
    int fuse;                // 0 = off, else 1 + AZ_IA_* mode of the next step's drift
    int slot_next;
    const float4* vel; float4* pos_next; float4* vel_next;
    int *cellOf, *slotOf, *cellCount;
    AzScalars* sc;
    double* partKE;
    int* blockFlag;          // bricks: 1 if any atom of this 128-atom block lists a ghost (written by az_k_build_list)
    const int* blockOrder;   // direct halo (phase 4): block of the launch -> block of atoms (ghost-free blocks first), or NULL
    int phase;               // list pass on bricks: 4 = direct halo: all blocks in one launch, the ones that list ghosts wait for the
                             // arrival flags first (the neighbours store into this rank's position stream, az_halo.cuh); 0 = all blocks, 1 = blocks without ghosts (runs while the halo is
                             // in flight), 2 = blocks with ghosts (after the halo), 3 = blocks with ghosts in the FUSED halo: the
                             // GHOSTS instantiation, on the halo stream beside phase 1; it waits for the arrival flags of the
                             // peer-memory halo itself and reads ghost positions straight from the receive staging
    const float4* ghostStage;// phase 3: this step's half of the receive staging (az_halo.cuh)
    const int* ghostEntry;   // phase 3: sorted ghost index - n_owned -> entry of the staging
    const int* haloFlags;    // phase 3: arrival flags [27]
    int haloSeq, haloDirs;   // phase 3: the value the flags must reach; bit d set = direction d is in use
    int dbg;                 // diagnostics of the list pass: 1 = synthesise positions (no gathers), 2 = synthesise indices (no list stream)
    const int2* stenTab;     // warp sweep: the stencil of every owned cell (az_k_stencil_table), stenS entries per cell; NULL = work it out per atom
    int stenS;
    int prune;               // warp sweep: drop stencil cells the atom itself cannot reach (cells binned this step)
    float r2_safe;           // packed LJ list pass: |F/r| < MX_FRC for every r2 in [r2_safe, gate] (+inf: always take the clamped form)
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
// (0 none, 1 = -L, 2 = +L per axis) above them (imaged neighbours only, see az_k_build_list).
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
template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY, bool SHIFT>
__device__ __forceinline__ void az_sweep_cell(const AzParams& P, const AzPairArgs& A, const AzVdw* __restrict__ s_tab,
                                              const float4* __restrict__ s_tile, const float* __restrict__ s_tile_rad,
                                              int jb, int je, int self, float lim, const float4 pi, int ti, float radi,
                                              float shx, float shy, float shz,
                                              float& fx, float& fy, float& fz, float& ev, float& ec, int& hits)
{
    constexpr bool FOLD = SINGLE && CK == AZ_ELEC_NONE;
#pragma unroll 4
    for (int j = jb; j < je; j++) {
        const float4 pj = s_tile[j];
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        if (SHIFT) { dx -= shx; dy -= shy; dz -= shz; }
        const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy)); // nvcc's contraction of dx*dx+dy*dy+dz*dz
        const bool in = r2 <= lim && j != self;
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

    // owned cells are numbered first, z fastest inside the owned block: the run's cells and atoms are contiguous
    const int onz = g.ohi[2] - g.olo[2];
    const int cell0 = ((lcx - g.olo[0]) * ony + (lcy - g.olo[1])) * onz + (lz0 - g.olo[2]);
    const int iBeg = A.first[cell0];
    const int iEnd = A.first[cell0 + zcnt - 1] + A.count[cell0 + zcnt - 1];

    float* s_tile_rad = (float*)(s_tile + A.tile_cap); // radii of the staged atoms (RADII only)

    double eV = 0.0, eC = 0.0;
    unsigned long long nHit = 0;
    // the pair gate: r2Max (folded with the only r2cut when possible)
    const float gate = (SINGLE && CK == AZ_ELEC_NONE) ? fminf(P.r2max, A.single.r2cut) : P.r2max;
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
            lzi = g.olo[2] + ci % onz;
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
                        if (st >= 0 && st < g.nl[2]) { const int cid = az_cell_id(g, colbase + st); cnt = A.count[cid]; gi = A.first[cid]; }
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
                        if (shxi == 0 && shyi == 0 && shz == 0.f)
                            az_sweep_cell<VK, CK, SINGLE, RADII, ENERGY, false>(P, A, s_tab, s_tile, s_tile_rad, jb, je, self, lim, pi, ti, radi,
                                                                                0.f, 0.f, 0.f, fx, fy, fz, ev, ec, hits);
                        else
                            az_sweep_cell<VK, CK, SINGLE, RADII, ENERGY, true>(P, A, s_tab, s_tile, s_tile_rad, jb, je, self, lim, pi, ti, radi,
                                                                               shx, shy, shz, fx, fy, fz, ev, ec, hits);
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


// ---------------------------------------------------------------------------------------
// Verlet-list path (opts.rebuild_every > 1, BASELINE config 3).  Layout: atoms are taken in blocks of
// 32 (one warp); row k of atom i lives at nbr[((i >> 5) * cap + k) * 32 + (i & 31)], so a warp reads
// one 128-byte line per row and streams through ONE contiguous region of cap * 128 bytes (a layout
// with row stride n_atoms touches a different DRAM page for every row).  Neighbours that need no
// periodic image fill the rows from 0 upwards and hold the bare sorted index of j; the few that do
// (atoms next to a box face) fill the rows from cap-1 downwards and carry the image signs above the
// index bits, so the hot loop of the force pass never looks at a shift.
// nbrCount[i] = plain | imaged << 16.
//
// az_k_build_list: the rebuild step.  One thread per atom walks the cells of its stencil straight
// from the sorted stream (neighbouring lanes sit in the same or the next cell, so the float4 loads
// are L1 broadcasts); runs of cells that are adjacent in memory (the z neighbours of one column) are
// merged into one j-range.  delta and r2 are formed exactly like the force pass forms them, so with
// skin == 0 the recorded set is the in-range set.
#ifndef AZ_LIST_TPB
#define AZ_LIST_TPB 128
#endif
// az_k_pair_list asks for 16 resident blocks (32 registers, 64 warps per SM) in its plain Lennard-Jones form: the kernel
// is bound by gather latency and more warps hide it (ncu: 0.357 ms at 32 registers, 0.366 ms at 48); the forms with
// exp/erfc, a species table or per-atom radii would spill at 32 registers and ask for 10 blocks (48 registers)
#define AZ_NBR_CNT_BITS 16

__device__ __forceinline__ size_t az_nbr_base(int i, int cap) { return ((size_t)(i >> 5) * (size_t)cap) * 32 + (size_t)(i & 31); }

// FAST (every axis has more than 2*sten cells, so every stencil offset is a distinct valid cell): the z-cells of
// one column that can hold neighbours form an interval whose atoms are one contiguous range
// [first[lo], first[hi + 1]) -- two loads per column instead of two per cell.  Cells are kept or dropped by the
// cell-to-cell distance of the reference (cell_dist_int), NOT by the atom's own position: measured on B200, pruning
// per atom makes the lanes of a warp (atoms of one cell) walk different ranges, which costs more in divergence
// and lost L1 broadcasts (22 active lanes per instruction instead of 30) than the 30 % fewer candidates save.
#ifndef AZ_BUILD_UNROLL
#define AZ_BUILD_UNROLL 2 // candidate pairs per trip of the list build's inner loop (A/B knob)
#endif
// A/B knob: resident blocks asked of ptxas for the list build.  Left undefined by default -- naming even 1 block makes ptxas
// spend 66 registers instead of 56 on this kernel.
#ifdef AZ_BUILD_MINB
#define AZ_BUILD_BOUNDS __launch_bounds__(AZ_LIST_TPB, AZ_BUILD_MINB)
#else
#define AZ_BUILD_BOUNDS __launch_bounds__(AZ_LIST_TPB)
#endif
template <bool FAST>
__global__ void AZ_BUILD_BOUNDS
az_k_build_list(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    const AzGrid& g = P.g;
    const int i = blockIdx.x * AZ_LIST_TPB + threadIdx.x;
    const int n_owned = az_n_owned(P);
    __shared__ int s_ghost; // bricks: does any atom of this block see a ghost?
    if (A.blockFlag) { if (threadIdx.x == 0) s_ghost = 0; __syncthreads(); }
    int jmax = -1; // end of the farthest range swept by this thread (ghosts sit behind the owned atoms)
    if (i < n_owned) {
    const float4 pi = A.pos[i];
    // owned cells are numbered first, z fastest inside the owned block
    const int ony = g.ohi[1] - g.olo[1], onz = g.ohi[2] - g.olo[2];
    const int ci = A.cellOfSorted[i];
    const int cxy = ci / onz;
    const int gcz = g.olo[2] + (ci - cxy * onz) + g.goff[2];
    const int gcy = g.olo[1] + cxy % ony + g.goff[1];
    const int gcx = g.olo[0] + cxy / ony + g.goff[0];
    const int cap = A.nbr_cap;
    int* __restrict__ rows = A.nbr + az_nbr_base(i, cap);
#ifndef AZ_EMU
    asm volatile("" : "+l"(rows)); // keep the row base in a register pair: one IMAD.WIDE per store instead of re-deriving it
#endif
    int nA = 0, nB = 0;
    const float lim = P.r2list;

    // all j of one contiguous range against atom i
    auto sweep = [&](const int rb, const int re, const int code) {
        jmax = max(jmax, re - 1);
        if (code == 0) {
            const int room = cap - nB;
            if ((i < rb || i >= re) && nA + (re - rb) <= room) {
                // the common case (8 of the 9 columns of an atom): atom i is not in the range and the rows cannot run out, so
                // the loop body is the distance test and a predicated store + pointer bump.  Two candidates per FADD2 / FFMA2;
                // each half is rounded like the scalar fmaf chain of the force pass, so the recorded set is the same.
                int j = rb, o = nA * 32; // o: offset of the next free row, in ints (one IMAD.WIDE + ST + IADD per hit)
#if AZ_BUILD_UNROLL == 1
#pragma unroll 1
#elif AZ_BUILD_UNROLL == 3
#pragma unroll 3
#elif AZ_BUILD_UNROLL == 4
#pragma unroll 4
#else
#pragma unroll 2
#endif
                for (; j + 2 <= re; j += 2) {
                    const float4 pa = __ldg(A.pos + j), pb = __ldg(A.pos + j + 1);
                    const az_f2 dx = az_pk(pi.x - pa.x, pi.x - pb.x), dy = az_pk(pi.y - pa.y, pi.y - pb.y), dz = az_pk(pi.z - pa.z, pi.z - pb.z);
                    const az_f2 r2 = az_fma2(dz, dz, az_fma2(dx, dx, az_mul2(dy, dy)));
                    if (az_lo(r2) <= lim) { rows[o] = j; o += 32; }
                    if (az_hi(r2) <= lim) { rows[o] = j + 1; o += 32; }
                }
                if (j < re) {
                    const float4 pj = __ldg(A.pos + j);
                    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                    if (fmaf(dz, dz, fmaf(dx, dx, dy * dy)) <= lim) { rows[o] = j; o += 32; }
                }
                nA = o >> 5;
                return;
            }
#pragma unroll 4
            for (int j = rb; j < re; j++) {
                const float4 pj = __ldg(A.pos + j);
                const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                const bool hit = (r2 <= lim) & (j != i);
                if (hit & (nA < room)) rows[nA * 32] = j;
                nA += hit ? 1 : 0;
            }
        } else {
            const float shx = ((float)((code >> 1) & 1) - (float)(code & 1)) * P.L[0];
            const float shy = ((float)((code >> 3) & 1) - (float)((code >> 2) & 1)) * P.L[1];
            const float shz = ((float)((code >> 5) & 1) - (float)((code >> 4) & 1)) * P.L[2];
            const int room = cap - nA;
            for (int j = rb; j < re; j++) {
                const float4 pj = __ldg(A.pos + j);
                const float dx = (pi.x - pj.x) - shx, dy = (pi.y - pj.y) - shy, dz = (pi.z - pj.z) - shz;
                const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                const bool hit = (r2 <= lim) & (j != i);
                if (hit & (nB < room)) rows[(cap - 1 - nB) * 32] = j | (code << AZ_NBR_BITS);
                nB += hit ? 1 : 0;
            }
        }
    };

    int rb = 0, re = 0, rcode = 0; // pending range
    auto add_range = [&](const int jb, const int je, const int code) {
        if (jb == re && code == rcode) re = je;
        else {
            if (re > rb) sweep(rb, re, rcode);
            rb = jb; re = je; rcode = code;
        }
    };
    if (FAST) {
        const int nz = g.n[2];
        for (int ox = -g.sten[0]; ox <= g.sten[0]; ox++) {
            const int ux = gcx + ox;
            const int shxi = (ux < 0) ? -1 : ((ux >= g.n[0]) ? 1 : 0);
            const int sx = g.dec[0] ? (ux - g.goff[0]) : (ux - shxi * g.n[0]);
            const int aox = ox < 0 ? -ox : ox;
            const float dx = aox ? (float)(aox - 1) * g.csize[0] : 0.f;
            for (int oy = -g.sten[1]; oy <= g.sten[1]; oy++) {
                const int uy = gcy + oy;
                const int shyi = (uy < 0) ? -1 : ((uy >= g.n[1]) ? 1 : 0);
                const int sy = g.dec[1] ? (uy - g.goff[1]) : (uy - shyi * g.n[1]);
                const int aoy = oy < 0 ? -oy : oy;
                const float dy = aoy ? (float)(aoy - 1) * g.csize[1] : 0.f;
                const float dxy2 = dx * dx + dy * dy;
                if (dxy2 > lim) continue;
                // |oz| = k is kept iff dxy2 + ((k-1)*csize)^2 <= lim  (checked exactly below: sqrt and floor round)
                int kz = min(g.sten[2], 1 + (int)floorf(sqrtf(lim - dxy2) * g.crev[2]));
                while (kz < g.sten[2] && dxy2 + ((float)kz * g.csize[2]) * ((float)kz * g.csize[2]) <= lim) kz++;
                while (kz > 0 && dxy2 + ((float)(kz - 1) * g.csize[2]) * ((float)(kz - 1) * g.csize[2]) > lim) kz--;
                const int khi = kz, klo = kz;
                const int ua = gcz - klo, ub = gcz + khi;
                const int colbase = (sx * g.nl[1] + sy) * g.nl[2];
#pragma unroll 1
                for (int shzi = -1; shzi <= 1; shzi++) { // the interval, split where it crosses a periodic face
                    const int a = max(ua, shzi * nz), b = min(ub, shzi * nz + nz - 1);
                    if (a > b) continue;
                    const int off = g.dec[2] ? g.goff[2] : shzi * nz;
                    const int cid_a = az_cell_id(g, colbase + a - off), cid_b = az_cell_id(g, colbase + b - off);
                    const int code = az_shift_code(shxi, shyi, shzi);
                    if (cid_b - cid_a == b - a) add_range(A.first[cid_a], A.first[cid_b + 1], code);
                    else
                        for (int u = a; u <= b; u++) { // ghost cells of a brick are not numbered along z
                            const int cid = az_cell_id(g, colbase + u - off);
                            const int jb = A.first[cid];
                            add_range(jb, jb + A.count[cid], code);
                        }
                }
            }
        }
    } else {
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
            if (mxy2 > lim) continue;
            const int colbase = (sx * g.nl[1] + sy) * g.nl[2];
            for (int oz = -g.sten[2]; oz <= g.sten[2]; oz++) {
                int sz, shzi;
                if (!az_neighbor(g, 2, gcz, oz, sz, shzi)) continue;
                const int aoz = oz < 0 ? -oz : oz;
                const float mz = aoz ? (float)(aoz - 1) * g.csize[2] : 0.f;
                if (mxy2 + mz * mz > lim) continue;
                const int cid = az_cell_id(g, colbase + sz);
                const int jb = A.first[cid], je = jb + A.count[cid];
                add_range(jb, je, az_shift_code(shxi, shyi, shzi));
            }
        }
    }
    }
    if (re > rb) sweep(rb, re, rcode);
    if (nA + nB > cap) { atomicOr(A.flags, AZ_FLAG_NBR); atomicMax(A.flags + 1, nA + nB); nB = min(nB, cap); nA = min(nA, cap - nB); }
    A.nbrCount[i] = nA | (nB << AZ_NBR_CNT_BITS);
    }
    if (A.blockFlag) { // ghosts sit behind the owned atoms in the sorted streams; the 128-byte line they share with the last
        // owned atoms counts as ghost territory (direct halo: no block that does not wait may pull it into L1)
        const int gl = n_owned & ~7;
        if (jmax >= gl || (i >= gl && i < n_owned)) atomicOr(&s_ghost, 1);
        __syncthreads();
        if (threadIdx.x == 0) A.blockFlag[blockIdx.x] = s_ghost;
    }
}

// warp-wide integer max / min, the same value in every lane (REDUX on the device)
__device__ __forceinline__ int az_warp_max_int(int v)
{
#ifdef AZ_EMU
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
    return __shfl_sync(0xffffffffu, v, 0);
#else
    return __reduce_max_sync(0xffffffffu, v);
#endif
}
__device__ __forceinline__ int az_warp_min_int(int v) { return -az_warp_max_int(-v); }

// az_k_build_list_staged: the rebuild step on grids where every stencil offset is a distinct cell (the FAST case of
// az_k_build_list; large systems).  Same thread-per-atom walk, same candidate ranges, same distance test -- what changes is
// where the hits go.  ncu on B200 (profiles/r2/): az_k_build_list is bound by the L1 data pipe (l1tex__data_pipe_lsu_wavefronts
// 73 % of peak), and three quarters of those wavefronts are its predicated row stores: the lanes of a warp sit at different
// row counts, so one `rows[nA * 32] = j` touches ~7 different 128-byte lines.  Here a hit goes to a shared-memory ring
// (s_rows[row & 63][thread]: conflict-free, one wavefront) and the warp flushes the ring in lockstep over the ROW index, so
// every global store instruction writes one line.  For that the lanes of a warp must stay together: the walk over
// columns / z pieces has the same trip counts in every lane, every range is swept in chunks of AZ_BL_CHUNK candidates up
// to the longest range of the warp (shorter ones idle), and the flush decision is a warp vote.  Neighbours through a
// periodic face (rare) and ranges that are not contiguous in memory (ghost cells of a brick) keep the direct stores.
// MEASURED AND REJECTED (B200, 4 M atoms, profiles/r2/ncu_full_lj4m_staged_build_summary.txt): the L1 data pipe falls from 73 % to
// 43 % and the store sectors from 203 M to 48 M as intended, but the kernel takes 2.60 ms against 1.49 ms: the lockstep
// bookkeeping (ring index, self test on every candidate, chunk votes, idle lanes of shorter ranges) raises the instruction
// count from 1.01e9 to 1.70e9 warp instructions and the 32 KB ring cuts the resident warps from 31 to 21 per SM (L1 hit
// rate 87 % -> 56 %).  Off by default (az_engine_set_option "list_staged" 1 switches it on for the A/B); results are identical.
#define AZ_BL_RING 64
#define AZ_BL_CHUNK 16
static __global__ void __launch_bounds__(AZ_LIST_TPB)
az_k_build_list_staged(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    __shared__ int s_rows[AZ_BL_RING * AZ_LIST_TPB];
    __shared__ int s_ghost;
    const AzGrid& g = P.g;
    const int tid = threadIdx.x;
    const int i = blockIdx.x * AZ_LIST_TPB + tid;
    const int n_owned = az_n_owned(P);
    if (A.blockFlag) { if (tid == 0) s_ghost = 0; __syncthreads(); }
    const bool act = i < n_owned;
    const int ia = act ? i : max(0, n_owned - 1); // idle lanes walk along with the last atom's cell and record nothing
    const float4 pi = A.pos[ia];
    const int ony = g.ohi[1] - g.olo[1], onz = g.ohi[2] - g.olo[2];
    const int ci = A.cellOfSorted[ia];
    const int cxy = ci / onz;
    const int gcz = g.olo[2] + (ci - cxy * onz) + g.goff[2];
    const int gcy = g.olo[1] + cxy % ony + g.goff[1];
    const int gcx = g.olo[0] + cxy / ony + g.goff[0];
    const int cap = A.nbr_cap;
    int* __restrict__ rows = A.nbr + az_nbr_base(ia, cap);
    int nA = 0, fA = 0, nB = 0, jmax = -1; // rows recorded / already flushed (plain), imaged rows, farthest j swept
    const float lim = P.r2list;
    int* __restrict__ ring = s_rows + tid;

    auto flush = [&]() { // lockstep over the row index: one 128-byte line per store instruction
        const int kmin = az_warp_min_int(fA), kmax = az_warp_max_int(nA);
        for (int k = kmin; k < kmax; k++)
            if (act && k >= fA && k < nA && k < cap - nB) rows[k * 32] = ring[(k & (AZ_BL_RING - 1)) * AZ_LIST_TPB];
        fA = nA;
    };
    // one contiguous range per lane, swept by the whole warp in lockstep (lanes with a shorter or empty range idle)
    auto sweep_lock = [&](const int rb, const int re) {
        const int len = act ? max(0, re - rb) : 0;
        if (len > 0) jmax = max(jmax, re - 1);
        const int maxlen = az_warp_max_int(len);
        const bool self = act && i >= rb && i < re; // the own column holds atom i itself
        for (int p0 = 0; p0 < maxlen; p0 += AZ_BL_CHUNK) {
            const int a = rb + p0, b = min(rb + len, a + AZ_BL_CHUNK);
            int j = a;
            for (; j + 2 <= b; j += 2) {
                const float4 pa = __ldg(A.pos + j), pb = __ldg(A.pos + j + 1);
                const az_f2 dx = az_pk(pi.x - pa.x, pi.x - pb.x), dy = az_pk(pi.y - pa.y, pi.y - pb.y), dz = az_pk(pi.z - pa.z, pi.z - pb.z);
                const az_f2 r2 = az_fma2(dz, dz, az_fma2(dx, dx, az_mul2(dy, dy))); // each half rounds like the force pass's fmaf chain
                if ((az_lo(r2) <= lim) & !(self & (j == i))) { ring[(nA & (AZ_BL_RING - 1)) * AZ_LIST_TPB] = j; nA++; }
                if ((az_hi(r2) <= lim) & !(self & (j + 1 == i))) { ring[(nA & (AZ_BL_RING - 1)) * AZ_LIST_TPB] = j + 1; nA++; }
            }
            if (j < b) {
                const float4 pj = __ldg(A.pos + j);
                const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                if ((fmaf(dz, dz, fmaf(dx, dx, dy * dy)) <= lim) & !(self & (j == i))) { ring[(nA & (AZ_BL_RING - 1)) * AZ_LIST_TPB] = j; nA++; }
            }
            if (__any_sync(0xffffffffu, nA - fA > AZ_BL_RING - AZ_BL_CHUNK)) flush();
        }
    };
    // direct stores (no staging): imaged neighbours and the few non-contiguous ranges; the caller has flushed the ring
    auto sweep_direct = [&](const int rb, const int re, const int code) {
        if (!act || re <= rb) return;
        jmax = max(jmax, re - 1);
        if (code == 0) {
            const int room = cap - nB;
            for (int j = rb; j < re; j++) {
                const float4 pj = __ldg(A.pos + j);
                const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                const bool hit = (fmaf(dz, dz, fmaf(dx, dx, dy * dy)) <= lim) & (j != i);
                if (hit & (nA < room)) rows[nA * 32] = j;
                nA += hit ? 1 : 0;
            }
            fA = nA;
        } else {
            const float shx = ((float)((code >> 1) & 1) - (float)(code & 1)) * P.L[0];
            const float shy = ((float)((code >> 3) & 1) - (float)((code >> 2) & 1)) * P.L[1];
            const float shz = ((float)((code >> 5) & 1) - (float)((code >> 4) & 1)) * P.L[2];
            for (int j = rb; j < re; j++) {
                const float4 pj = __ldg(A.pos + j);
                const float dx = (pi.x - pj.x) - shx, dy = (pi.y - pj.y) - shy, dz = (pi.z - pj.z) - shz;
                const bool hit = (fmaf(dz, dz, fmaf(dx, dx, dy * dy)) <= lim) & (j != i);
                if (hit & (nB < cap - nA)) rows[(cap - 1 - nB) * 32] = j | (code << AZ_NBR_BITS);
                nB += hit ? 1 : 0;
            }
        }
    };

    const int nz = g.n[2];
    for (int ox = -g.sten[0]; ox <= g.sten[0]; ox++) {
        const int ux = gcx + ox;
        const int shxi = (ux < 0) ? -1 : ((ux >= g.n[0]) ? 1 : 0);
        const int sx = g.dec[0] ? (ux - g.goff[0]) : (ux - shxi * g.n[0]);
        const int aox = ox < 0 ? -ox : ox;
        const float ddx = aox ? (float)(aox - 1) * g.csize[0] : 0.f;
        for (int oy = -g.sten[1]; oy <= g.sten[1]; oy++) {
            const int uy = gcy + oy;
            const int shyi = (uy < 0) ? -1 : ((uy >= g.n[1]) ? 1 : 0);
            const int sy = g.dec[1] ? (uy - g.goff[1]) : (uy - shyi * g.n[1]);
            const int aoy = oy < 0 ? -oy : oy;
            const float ddy = aoy ? (float)(aoy - 1) * g.csize[1] : 0.f;
            const float dxy2 = ddx * ddx + ddy * ddy;
            if (dxy2 > lim) continue; // the same in every lane: depends on the offsets only
            // |oz| = k is kept iff dxy2 + ((k-1)*csize)^2 <= lim  (cell_dist_int; checked exactly: sqrt and floor round)
            int kz = min(g.sten[2], 1 + (int)floorf(sqrtf(lim - dxy2) * g.crev[2]));
            while (kz < g.sten[2] && dxy2 + ((float)kz * g.csize[2]) * ((float)kz * g.csize[2]) <= lim) kz++;
            while (kz > 0 && dxy2 + ((float)(kz - 1) * g.csize[2]) * ((float)(kz - 1) * g.csize[2]) > lim) kz--;
            const int ua = gcz - kz, ub = gcz + kz;
            const int colbase = (sx * g.nl[1] + sy) * g.nl[2];
#pragma unroll 1
            for (int shzi = -1; shzi <= 1; shzi++) { // the interval, split where it crosses a periodic face
                const int a = max(ua, shzi * nz), b = min(ub, shzi * nz + nz - 1);
                const bool some = a <= b;
                const int off = g.dec[2] ? g.goff[2] : shzi * nz;
                int rb = 0, re = 0;
                bool contig = true;
                if (some) {
                    const int cid_a = az_cell_id(g, colbase + a - off), cid_b = az_cell_id(g, colbase + b - off);
                    contig = (cid_b - cid_a == b - a);
                    if (contig) { rb = A.first[cid_a]; re = A.first[cid_b + 1]; }
                }
                const int code = az_shift_code(shxi, shyi, shzi);
                const bool lock = contig && code == 0;
                if (__all_sync(0xffffffffu, lock || !some || !act)) sweep_lock(lock && some ? rb : 0, lock && some ? re : 0);
                else { // some lane of the warp needs the direct stores for this piece: everybody takes them
                    flush();
                    if (some) {
                        if (contig) sweep_direct(rb, re, code);
                        else
                            for (int u = a; u <= b; u++) { // ghost cells of a brick are not numbered along z
                                const int cid = az_cell_id(g, colbase + u - off);
                                const int jb = A.first[cid];
                                sweep_direct(jb, jb + A.count[cid], code);
                            }
                    }
                }
            }
        }
    }
    flush();
    if (act) {
        if (nA + nB > cap) { atomicOr(A.flags, AZ_FLAG_NBR); atomicMax(A.flags + 1, nA + nB); nB = min(nB, cap); nA = min(nA, cap - nB); }
        A.nbrCount[i] = nA | (nB << AZ_NBR_CNT_BITS);
    }
    if (A.blockFlag) { // ghosts (and the line they share with the last owned atoms) sit behind the owned atoms in the sorted streams
        const int gl = n_owned & ~7;
        if (act && (jmax >= gl || i >= gl)) atomicOr(&s_ghost, 1);
        __syncthreads();
        if (tid == 0) A.blockFlag[blockIdx.x] = s_ghost;
    }
}

// ghost positions written by a neighbour GPU: read them from L2 (a line of the same staging half may sit stale in L1)
__device__ __forceinline__ float4 az_ld_cg4g(const float4* p)
{
#ifdef AZ_EMU
    __sync_synchronize();
    return *p;
#else
    return __ldcg(p);
#endif
}

// read-once global load: no L1 allocation
__device__ __forceinline__ int az_ld_stream(const int* p)
{
#ifdef AZ_EMU
    return *p;
#else
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
}

// az_k_pair_list: the force pass of every step of the list path (the rebuild step's forces come from
// it too).  One thread per atom, no atomics, same pair arithmetic and the same r2 <= r2Max gate on
// current positions as the cell sweep, so the pair set per step is identical to the every-step-rebuild
// schedule as long as no atom moved more than skin/2 (tracked by az_k_integrate_a / az_k_finalize).
// GHOSTS = true is the instantiation for the fused halo of a brick (A.phase == 3): it works on the blocks that list ghosts only,
// waits for the arrival flags of the peer-memory halo and takes ghost positions from the receive staging.  It runs on the halo
// stream beside the GHOSTS = false launch (phase 1: the ghost-free blocks) of the same step.  Two instantiations rather than one
// kernel with a block-level branch: the extra live values of the ghost path spilled 108 bytes into the 32-register row loop of
// EVERY block and made the whole pass 1.8x slower (B200, profiles/r2/bench_n2_fused_spill_v1.json).
template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY, bool GHOSTS = false>
// AZ_LIST_PACKED = 1 selects the f32x2 form of the Lennard-Jones row loop (FADD2 / FMUL2 / FFMA2: 17 issue slots per
// row instead of 26).  Measured on B200 (profiles/r2/): it is SLOWER -- 0.427 ms per pass at 40 registers / 48 warps per SM,
// 0.438 ms at 48 / 40, 0.73 ms at 32 / 64 with spills, against 0.385 ms for the scalar loop at 32 registers / 64 warps --
// because the pass is bound by the L1 data pipe, not by issue slots: every gather LDG.128 touches ~6 distinct 128-byte
// lines (ncu: l1tex__data_pipe_lsu_wavefronts 80 % of peak, issue 42 %), and with fewer resident warps less of that
// latency is hidden.  The scalar loop stays the default; the packed loop is kept for the measurement.
#ifndef AZ_LIST_PACKED
#define AZ_LIST_PACKED 0
#endif
#ifndef AZ_LIST_CHUNK
#define AZ_LIST_CHUNK 4
#endif
#ifndef AZ_LIST_LJ_WARPS
#define AZ_LIST_LJ_WARPS (AZ_LIST_PACKED ? 48 : 64) // resident warps per SM the plain Lennard-Jones form asks for (40 / 32 registers)
#endif
__global__ void __launch_bounds__(AZ_LIST_TPB, (VK == AZ_VDW_LJ && CK == AZ_ELEC_NONE && SINGLE && !RADII && !ENERGY && !GHOSTS) ? AZ_LIST_LJ_WARPS * 32 / AZ_LIST_TPB : 1280 / AZ_LIST_TPB)
az_k_pair_list(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    __shared__ AzVdw s_tab[SINGLE ? 1 : AZ_MAX_SPECIES * AZ_MAX_SPECIES];
    __shared__ double s_red[96];
    __shared__ double s_kw[AZ_LIST_TPB / 32];
    __shared__ int s_done;
    const int tid = threadIdx.x;
    if (!SINGLE || A.fuse) { // all warps are here at the same time: this barrier costs nothing
        for (int k = tid; k < P.nspec * P.nspec && !SINGLE; k += AZ_LIST_TPB) s_tab[k] = A.table[k];
        if (tid == 0) s_done = 0;
        __syncthreads();
    }
    const int ob = (A.phase == 4) ? A.blockOrder[blockIdx.x] : (int)blockIdx.x; // direct halo: launch order, bit 30 = lists ghosts
    const int blk = ob & 0x3fffffff;
    if (!GHOSTS && A.phase == 4) { // direct halo: the neighbours' stores into my ghost slots for this step must have landed
        if (ob >> 30) {
            if (tid < 27 && ((A.haloDirs >> tid) & 1)) az_wait_flag(A.haloFlags + tid, A.haloSeq, A.flags);
            __syncthreads();
        }
    } else
    if (GHOSTS) {
        if (A.blockFlag[blk] == 0) return; // phase 1 (the other launch) owns the blocks that list no ghost
        // the neighbours' stores for this step must have landed (az_k_ghost_put raises the flags over NVLink)
        if (tid < 27 && ((A.haloDirs >> tid) & 1)) az_wait_flag(A.haloFlags + tid, A.haloSeq, A.flags);
        __syncthreads();
    } else if (A.phase && ((A.phase == 1) == (A.blockFlag[blk] != 0))) return; // the other phase owns this block
    const int n_own = az_n_owned(P);
    const int i = blk * AZ_LIST_TPB + tid;
    const bool act = i < n_own;
    float fx = 0.f, fy = 0.f, fz = 0.f, ev = 0.f, ec = 0.f, v2 = 0.f;
    double ke = 0.0;
    int hits = 0, ti = 0;
    if (act) {
        const float4 pi = A.pos[i];
        ti = __float_as_int(pi.w);
        float radi = 0.f;
        if (RADII) radi = A.therm[i].y;
        const int cnt2 = A.nbrCount[i];
        const int cnt = cnt2 & ((1 << AZ_NBR_CNT_BITS) - 1), cntB = cnt2 >> AZ_NBR_CNT_BITS;
        constexpr int stride = 32; // ints between consecutive rows of one atom (one 128-byte line per warp and row)
        constexpr bool FOLD = SINGLE && CK == AZ_ELEC_NONE;
#if defined(AZ_LIST_DBG) || (!AZ_LIST_PACKED && !defined(AZ_EMU))
        constexpr bool PACKED = false;
#else
        constexpr bool PACKED = FOLD && VK == AZ_VDW_LJ && !RADII; // the emulator build always exercises the packed form's logic
#endif
        const float gate = FOLD ? fminf(P.r2max, A.single.r2cut) : P.r2max;
        const int* __restrict__ const rows = A.nbr + az_nbr_base(i, A.nbr_cap);
        // ---- scalar form: every (vdW, Coulomb) combination; the reference's force_autocap clamp on every pair
        auto scalar_rows = [&](auto ldpos) { // ldpos(c): position of neighbour c (plain gather, or ghost-aware in a fused-halo block)
        // one neighbour: pair term from the finished delta, masked accumulate
        auto one = [&](const float dx, const float dy, const float dz, const float4 pj, const float radj) {
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
        auto plain = [&](const float4 pj, const float radj) { one(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z, pj, radj); };
        const int* __restrict__ row = rows;
        int k = 0;
        // software pipeline by hand: the index loads of chunk k+1 (HBM stream) are in flight while the
        // gathers of chunk k (L1/L2) are consumed
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#ifdef AZ_LIST_DBG // diagnostic build only (bench.py --list-dbg): 1 = synthesise positions (no gathers), 2 = synthesise indices (no list stream)
        const int dbg = A.dbg;
        auto ldidx = [&](const int* p, int kk) { return dbg == 2 ? max(0, min(az_n_owned(P) - 1, i + kk - (cnt >> 1))) : *p; };
        auto ldpos_dbg = ldpos;
        auto ldpos2 = [&](int c) { return dbg == 1 ? make_float4(pi.x + 3.f + (float)c * 1e-12f, pi.y + 2.f, pi.z + 1.f, pi.w) : ldpos_dbg(c); };
#define ldpos ldpos2
#else
        // the list is a pure stream (every row is read once per step): keep it out of L1 so that the gathered
        // positions stay there (B200: 0.3624 -> 0.3575 ms per pass; __ldcs alone 0.3616)
        auto ldidx = [&](const int* p, int) { return az_ld_stream(p); };
#endif
        if (cnt >= 4) { c0 = ldidx(row, 0); c1 = ldidx(row + stride, 1); c2 = ldidx(row + 2 * stride, 2); c3 = ldidx(row + 3 * stride, 3); }
        for (; k + 4 <= cnt; k += 4) {
            row += 4 * stride;
            const float4 p0 = ldpos(c0), p1 = ldpos(c1), p2 = ldpos(c2), p3 = ldpos(c3);
            float r0 = 0.f, r1 = 0.f, r2_ = 0.f, r3 = 0.f;
            if (RADII) { r0 = A.therm[c0].y; r1 = A.therm[c1].y; r2_ = A.therm[c2].y; r3 = A.therm[c3].y; }
            int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
            if (k + 8 <= cnt) { n0 = ldidx(row, k + 4); n1 = ldidx(row + stride, k + 5); n2 = ldidx(row + 2 * stride, k + 6); n3 = ldidx(row + 3 * stride, k + 7); }
            plain(p0, r0); plain(p1, r1); plain(p2, r2_); plain(p3, r3);
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        }
        for (; k < cnt; k++) {
            const int c = ldidx(row, k); row += stride;
            const float4 pj = ldpos(c);
            float radj = 0.f;
            if (RADII) radj = A.therm[c].y;
            plain(pj, radj);
        }
        // neighbours through a periodic face: rows cap-1 downwards, image signs above the index bits
        const int* __restrict__ rowB = rows + (A.nbr_cap - 1) * stride;
        for (int kb = 0; kb < cntB; kb++) {
            const int c = *rowB; rowB -= stride;
            const int j = c & AZ_NBR_MASK, sc = (int)((unsigned)c >> AZ_NBR_BITS);
            const float4 pj = ldpos(j);
            float radj = 0.f;
            if (RADII) radj = A.therm[j].y;
            const float sx = ((float)((sc >> 1) & 1) - (float)(sc & 1)) * P.L[0];
            const float sy = ((float)((sc >> 3) & 1) - (float)((sc >> 2) & 1)) * P.L[1];
            const float sz = ((float)((sc >> 5) & 1) - (float)((sc >> 4) & 1)) * P.L[2];
            one((pi.x - pj.x) - sx, (pi.y - pj.y) - sy, (pi.z - pj.z) - sz, pj, radj);
        }
#ifdef AZ_LIST_DBG
#undef ldpos
#endif
        };
        // the plain gather, or -- GHOSTS -- ghost positions from the receive staging (L2 loads: a neighbour GPU wrote them over NVLink)
        auto ld_plain = [&](int c) { return __ldg(A.pos + c); };
        auto ld_ghost = [&](int c) { return c >= n_own ? az_ld_cg4g(A.ghostStage + A.ghostEntry[c - n_own]) : __ldg(A.pos + c); };
        if (GHOSTS) scalar_rows(ld_ghost);
        else if (PACKED) { // (a block that reads ghosts from the halo staging takes the scalar form)
            // ---- packed form (one Lennard-Jones entry, no Coulomb): two list rows per FADD2 / FMUL2 / FFMA2.  Each half is
            // the scalar arithmetic of az_vdw_feng<LJ> rounded the same way; the clamp of force_autocap is left out of the
            // loop and the smallest r2 is tracked instead: F/r of this entry stays inside +-MX_FRC for every r2 >= r2_safe
            // (worked out by the host from the table entry), so the clamp is the identity unless min r2 < r2_safe -- a hot
            // collision -- and then the atom is redone in the scalar, clamped form below.
            const az_f2 kp1 = az_pk1(A.single.p1), kp2 = az_pk1(A.single.p2), kp0 = az_pk1(A.single.p0);
            const az_f2 ktwo = az_pk1(2.f), kmone = az_pk1(-1.f);
            az_f2 fx2 = az_pk1(0.f), fy2 = fx2, fz2 = fx2, ev2 = fx2;
            bool hot = false; // some r2 below r2_safe (kept in a predicate register)
            const float safe = A.r2_safe;
            auto two = [&](const az_f2 dx, const az_f2 dy, const az_f2 dz) {
                const az_f2 r2 = az_fma2(dz, dz, az_fma2(dx, dx, az_mul2(dy, dy)));
                const float ra = az_lo(r2), rb = az_hi(r2);
                hot |= (ra < safe) | (rb < safe);
                // the gate acts on 1/r2: a row outside it carries sr6 = 0, hence F/r = -0 and U = -0, which leave the sums as they are
                const bool ina = ra <= gate, inb = rb <= gate;
                const az_f2 r2i = az_pk(ina ? az_fast_rcp(ra) : 0.f, inb ? az_fast_rcp(rb) : 0.f);
                const az_f2 sr2 = az_mul2(kp1, r2i);
                const az_f2 sr6 = az_mul2(az_mul2(sr2, sr2), sr2);
                const az_f2 f = az_mul2(az_mul2(az_mul2(kp2, r2i), sr6), az_fma2(ktwo, sr6, kmone));
                fx2 = az_fma2(f, dx, fx2); fy2 = az_fma2(f, dy, fy2); fz2 = az_fma2(f, dz, fz2);
                if (ENERGY) {
                    ev2 = az_fma2(az_mul2(kp0, sr6), az_add2(sr6, kmone), ev2);
                    hits += (ina ? 1 : 0) + (inb ? 1 : 0);
                }
            };
            auto pair2 = [&](const float4 pa, const float4 pb) { two(az_pk(pi.x - pa.x, pi.x - pb.x), az_pk(pi.y - pa.y, pi.y - pb.y), az_pk(pi.z - pa.z, pi.z - pb.z)); };
            constexpr float FAR = 1.0e15f; // second half of an odd row: r2 = 3e30, outside every gate
            const int* __restrict__ row = rows;
            int k = 0;
            int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#if AZ_LIST_CHUNK == 2
            if (cnt >= 2) { c0 = az_ld_stream(row); c1 = az_ld_stream(row + stride); }
#pragma unroll 1
            for (; k + 2 <= cnt; k += 2) {
                row += 2 * stride;
                const float4 p0 = __ldg(A.pos + c0), p1 = __ldg(A.pos + c1);
                if (k + 4 <= cnt) { c0 = az_ld_stream(row); c1 = az_ld_stream(row + stride); }
                pair2(p0, p1);
            }
            (void)c2; (void)c3;
#else
            if (cnt >= 4) { c0 = az_ld_stream(row); c1 = az_ld_stream(row + stride); c2 = az_ld_stream(row + 2 * stride); c3 = az_ld_stream(row + 3 * stride); }
#pragma unroll 1
            for (; k + 4 <= cnt; k += 4) {
                row += 4 * stride;
                const float4 p0 = __ldg(A.pos + c0), p1 = __ldg(A.pos + c1), p2 = __ldg(A.pos + c2), p3 = __ldg(A.pos + c3);
                int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
                if (k + 8 <= cnt) { n0 = az_ld_stream(row); n1 = az_ld_stream(row + stride); n2 = az_ld_stream(row + 2 * stride); n3 = az_ld_stream(row + 3 * stride); }
                pair2(p0, p1); pair2(p2, p3);
                c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            }
#endif
            for (; k < cnt; k++) {
                const int c = az_ld_stream(row); row += stride;
                const float4 pj = __ldg(A.pos + c);
                two(az_pk(pi.x - pj.x, FAR), az_pk(pi.y - pj.y, FAR), az_pk(pi.z - pj.z, FAR));
            }
            const int* __restrict__ rowB = rows + (A.nbr_cap - 1) * stride;
            for (int kb = 0; kb < cntB; kb++) {
                const int c = *rowB; rowB -= stride;
                const int j = c & AZ_NBR_MASK, sc = (int)((unsigned)c >> AZ_NBR_BITS);
                const float4 pj = __ldg(A.pos + j);
                const float sx = ((float)((sc >> 1) & 1) - (float)(sc & 1)) * P.L[0];
                const float sy = ((float)((sc >> 3) & 1) - (float)((sc >> 2) & 1)) * P.L[1];
                const float sz = ((float)((sc >> 5) & 1) - (float)((sc >> 4) & 1)) * P.L[2];
                two(az_pk((pi.x - pj.x) - sx, FAR), az_pk((pi.y - pj.y) - sy, FAR), az_pk((pi.z - pj.z) - sz, FAR));
            }
            if (!hot) {
                fx = az_lo(fx2) + az_hi(fx2); fy = az_lo(fy2) + az_hi(fy2); fz = az_lo(fz2) + az_hi(fz2);
                if (ENERGY) ev = az_lo(ev2) + az_hi(ev2);
            } else { hits = 0; scalar_rows(ld_plain); } // a pair close enough for the clamp (or coincident atoms): the clamped form decides
        } else scalar_rows(ld_plain);
        const float q = P.charge[ti];
        const float4 f = make_float4(-q * P.efield[0] + fx, -q * P.efield[1] + fy, -q * P.efield[2] + fz, 0.f);
        if (!A.fuse) A.frc[i] = f;
        else { // verlet_2stage of this step and verlet_1stage of the next one (az_k_integrate_ba's body)
            float4 p = pi, v = A.vel[i];
            const float h = P.rMasshdT[ti];
            v.x = fmaf(h, f.x, v.x); v.y = fmaf(h, f.y, v.y); v.z = fmaf(h, f.z, v.z);
            ke = (double)az_ke_term(v, P.mass[ti]);
            double eF = 0.0;
            if (A.fuse == 1 + AZ_IA_STALE) v2 = az_kick_drift<AZ_IA_STALE>(P, i, A.slot_next, p, v, f, A.cellOf, A.slotOf, A.cellCount, A.sc, eF);
            else if (A.fuse == 1 + AZ_IA_BIN) v2 = az_kick_drift<AZ_IA_BIN>(P, i, A.slot_next, p, v, f, A.cellOf, A.slotOf, A.cellCount, A.sc, eF);
            else v2 = az_kick_drift<AZ_IA_WRAP>(P, i, A.slot_next, p, v, f, A.cellOf, A.slotOf, A.cellCount, A.sc, eF);
            A.pos_next[i] = p; A.vel_next[i] = v;
        }
    }
    if (ENERGY) {
        double vv[3] = { 0.5 * (double)ev, 0.5 * (double)ec, (double)hits };
        az_block_sum<3>(vv, s_red);
        if (tid == 0) { A.partV[blk] = vv[0]; A.partC[blk] = vv[1]; A.partN[blk] = (unsigned long long)(vv[2] + 0.5); }
    }
    if (A.fuse) {
        // no block barrier (ncu: 10 % of the stall samples sat on it): every warp leaves its sum in shared memory and
        // retires; the LAST warp to arrive adds the four in a fixed order, so the block's partial is reproducible
        const double kw = az_warp_sum(ke);
        if ((tid & 31) == 0) {
            s_kw[tid >> 5] = kw;
            __threadfence_block();
            if (atomicAdd(&s_done, 1) == AZ_LIST_TPB / 32 - 1) {
                __threadfence_block();
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < AZ_LIST_TPB / 32; w++) t += ((volatile double*)s_kw)[w];
                A.partKE[blk] = 0.5 * t;
            }
        }
        if (A.fuse == 1 + AZ_IA_STALE) { // the fastest atom bounds the drift since the last rebuild
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v2 = fmaxf(v2, __shfl_down_sync(0xffffffffu, v2, o));
            if ((tid & 31) == 0 && __float_as_uint(v2) > *(volatile unsigned*)&A.sc->vmax2_bits[A.slot_next]) atomicMax(&A.sc->vmax2_bits[A.slot_next], __float_as_uint(v2));
        }
    }
}


// The stencil cells of one atom, enumerated by a whole warp 32 at a time, one cell per lane (z offset fastest): validity,
// image and the two loads (first, count) of all cells of a chunk are in flight together instead of one after the
// other.  Runs of lanes whose ranges are adjacent in memory with the same image (the z neighbours of one column) are
// merged, then `sweep(first_atom, end_atom, image_code)` is called by all lanes for every merged range.
// `prune` (cells binned this step only): a stencil cell is also dropped when the ATOM itself is farther than the cut-off from
// that cell's box -- exact (every atom of the cell is at least that far away; 1e-4 of a cell edge of margin for the rounding
// of the bin formula), free of divergence here because all lanes of the warp serve the same atom.  The reference's case
// study 1 bins a 4 A cut-off into 87.8 A cells: 26 of the 27 cells of an atom drop out unless it sits within 4 A of a face.
template <class F>
__device__ __forceinline__ void az_warp_stencil_ranges(const AzGrid& g, const int* sten, float r2lim, int gcx, int gcy, int gcz,
                                                       const int* __restrict__ first, const int* __restrict__ count, int lane, F&& sweep,
                                                       bool prune = false, float px = 0.f, float py = 0.f, float pz = 0.f)
{
    const int wy = 2 * sten[1] + 1, wz = 2 * sten[2] + 1;
    const int ncells = (2 * sten[0] + 1) * wy * wz;
    for (int c0 = 0; c0 < ncells; c0 += 32) {
        const int c = c0 + lane;
        int jb = -1, je = -1, code = 0;
        if (c < ncells) {
            const int oz = c % wz - sten[2], t = c / wz;
            const int oy = t % wy - sten[1], ox = t / wy - sten[0];
            int sx, sy, sz, shxi, shyi, shzi;
            if (az_neighbor(g, 0, gcx, ox, sx, shxi) && az_neighbor(g, 1, gcy, oy, sy, shyi) && az_neighbor(g, 2, gcz, oz, sz, shzi)) {
                const int aox = ox < 0 ? -ox : ox, aoy = oy < 0 ? -oy : oy, aoz = oz < 0 ? -oz : oz;
                const float mx = aox ? (float)(aox - 1) * g.csize[0] : 0.f;
                const float my = aoy ? (float)(aoy - 1) * g.csize[1] : 0.f;
                const float mz = aoz ? (float)(aoz - 1) * g.csize[2] : 0.f;
                const float mxy2 = mx * mx + my * my;
                bool keep = mxy2 <= r2lim && mxy2 + mz * mz <= r2lim;
                if (keep && prune) { // distance from the atom to the box of the (unwrapped) neighbour cell, per axis
                    auto gap = [&](int o, int gc, float x, float cs) {
                        const float d = o > 0 ? (float)(gc + o) * cs - x : (o < 0 ? x - (float)(gc + o + 1) * cs : 0.f);
                        return fmaxf(0.f, d - 1e-4f * cs);
                    };
                    const float gx = gap(ox, gcx, px, g.csize[0]), gy = gap(oy, gcy, py, g.csize[1]), gz = gap(oz, gcz, pz, g.csize[2]);
                    keep = gx * gx + gy * gy + gz * gz <= r2lim;
                }
                if (keep) {
                    const int cid = az_cell_id(g, (sx * g.nl[1] + sy) * g.nl[2] + sz);
                    jb = first[cid];
                    je = jb + count[cid];
                    code = az_shift_code(shxi, shyi, shzi);
                }
            }
        }
        const bool valid = je > jb;
        // continuation: this lane's range starts where the previous lane's ends, same image
        const int pje = __shfl_up_sync(0xffffffffu, je, 1), pcode = __shfl_up_sync(0xffffffffu, code, 1);
        const int pvalid = __shfl_up_sync(0xffffffffu, (int)valid, 1);
        const bool cont = valid && lane > 0 && pvalid && pje == jb && pcode == code;
        const unsigned contmask = __ballot_sync(0xffffffffu, cont);
        unsigned heads = __ballot_sync(0xffffffffu, valid && !cont);
        // a head's run covers the consecutive continuation lanes behind it
        const unsigned behind = lane < 31 ? (contmask >> (lane + 1)) : 0u;
        const int runlen = __ffs((int)~behind) - 1; // consecutive ones from bit 0
        const int je_run = __shfl_sync(0xffffffffu, je, min(31, lane + runlen));
        while (heads) {
            const int h = __ffs((int)heads) - 1;
            heads &= heads - 1;
            const int rb = __shfl_sync(0xffffffffu, jb, h), re = __shfl_sync(0xffffffffu, je_run, h), rcode = __shfl_sync(0xffffffffu, code, h);
            sweep(rb, re, rcode);
        }
    }
}


// The stencil of every owned cell, tabulated once at create time (the grid never changes): entry c of cell ci = what lane c of
// az_warp_stencil_ranges works out for that cell -- storage id of the neighbour cell (-1: no such cell pair), image code, and the
// three offsets (for the per-atom pruning).  The reference keeps the same kind of table (init_bypass6, src/cuCellList.cu:506-555);
// here it serves the warp-per-atom sweep only, where the per-atom bookkeeping was more than half of the instructions
// (ncu on B200, case study 1: 536 warp instructions per atom, 68 % issue utilisation).
#define AZ_STEN_OFF 8 // offsets are stored + 8 in four bits each
static __global__ void __launch_bounds__(256)
az_k_stencil_table(const __grid_constant__ AzParams P, int n_owned_cells, int S, int2* __restrict__ tab)
{
    const AzGrid& g = P.g;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_owned_cells * S) return;
    const int ci = (int)(t / S), c = (int)(t - (long long)ci * S);
    const int ony = g.ohi[1] - g.olo[1], onz = g.ohi[2] - g.olo[2];
    const int cxy = ci / onz;
    const int gcz = g.olo[2] + (ci - cxy * onz) + g.goff[2];
    const int gcy = g.olo[1] + cxy % ony + g.goff[1];
    const int gcx = g.olo[0] + cxy / ony + g.goff[0];
    const int wy = 2 * g.sten[1] + 1, wz = 2 * g.sten[2] + 1;
    const int oz = c % wz - g.sten[2], tt = c / wz;
    const int oy = tt % wy - g.sten[1], ox = tt / wy - g.sten[0];
    int2 ent = make_int2(-1, 0);
    int sx, sy, sz, shxi, shyi, shzi;
    if (az_neighbor(g, 0, gcx, ox, sx, shxi) && az_neighbor(g, 1, gcy, oy, sy, shyi) && az_neighbor(g, 2, gcz, oz, sz, shzi)) {
        const int aox = ox < 0 ? -ox : ox, aoy = oy < 0 ? -oy : oy, aoz = oz < 0 ? -oz : oz;
        const float mx = aox ? (float)(aox - 1) * g.csize[0] : 0.f;
        const float my = aoy ? (float)(aoy - 1) * g.csize[1] : 0.f;
        const float mz = aoz ? (float)(aoz - 1) * g.csize[2] : 0.f;
        const float mxy2 = mx * mx + my * my;
        if (mxy2 <= P.r2list && mxy2 + mz * mz <= P.r2list) {
            ent.x = az_cell_id(g, (sx * g.nl[1] + sy) * g.nl[2] + sz);
            ent.y = az_shift_code(shxi, shyi, shzi) | ((ox + AZ_STEN_OFF) << 6) | ((oy + AZ_STEN_OFF) << 10) | ((oz + AZ_STEN_OFF) << 14);
        }
    }
    tab[t] = ent;
}

// az_warp_stencil_ranges with the cell-level work read from the table: same cells, same order, same merged ranges.
// (A FLAT hand-out of the candidates -- warp prefix sum over the 32 cell populations of a chunk, two candidates per lane
// whatever cell they come from -- was measured on B200 and dropped: case study 2, whose 2.7 A cells hold 2 atoms, went
// from 61.8 to 65.8 us per step, the 40 000-atom gas and a 32 000-atom liquid gained 2-4 %; the pair function, executed by
// the whole warp whenever one lane has a partner in range, dominates either way.)
template <class F>
__device__ __forceinline__ void az_warp_table_ranges(const AzGrid& g, const int2* __restrict__ tab, int S, float r2lim, int ci, int gcx, int gcy, int gcz,
                                                     const int* __restrict__ first, const int* __restrict__ count, int lane, F&& sweep,
                                                     bool prune, float px, float py, float pz)
{
    const int2* __restrict__ row = tab + (size_t)ci * S;
    for (int c0 = 0; c0 < S; c0 += 32) {
        const int c = c0 + lane;
        int jb = -1, je = -1, code = 0;
        if (c < S) {
            const int2 ent = __ldg(row + c);
            bool keep = ent.x >= 0;
            if (keep && prune) { // distance from the atom to the box of the (unwrapped) neighbour cell, per axis
                const int ox = ((ent.y >> 6) & 15) - AZ_STEN_OFF, oy = ((ent.y >> 10) & 15) - AZ_STEN_OFF, oz = ((ent.y >> 14) & 15) - AZ_STEN_OFF;
                auto gap = [&](int o, int gc, float x, float cs) {
                    const float d = o > 0 ? (float)(gc + o) * cs - x : (o < 0 ? x - (float)(gc + o + 1) * cs : 0.f);
                    return fmaxf(0.f, d - 1e-4f * cs);
                };
                const float gx = gap(ox, gcx, px, g.csize[0]), gy = gap(oy, gcy, py, g.csize[1]), gz = gap(oz, gcz, pz, g.csize[2]);
                keep = gx * gx + gy * gy + gz * gz <= r2lim;
            }
            if (keep) {
                jb = first[ent.x];
                je = jb + count[ent.x];
                code = ent.y & 63;
            }
        }
        const bool valid = je > jb;
        // continuation: this lane's range starts where the previous lane's ends, same image
        const int pje = __shfl_up_sync(0xffffffffu, je, 1), pcode = __shfl_up_sync(0xffffffffu, code, 1);
        const int pvalid = __shfl_up_sync(0xffffffffu, (int)valid, 1);
        const bool cont = valid && lane > 0 && pvalid && pje == jb && pcode == code;
        const unsigned contmask = __ballot_sync(0xffffffffu, cont);
        unsigned heads = __ballot_sync(0xffffffffu, valid && !cont);
        const unsigned behind = lane < 31 ? (contmask >> (lane + 1)) : 0u;
        const int runlen = __ffs((int)~behind) - 1; // consecutive ones from bit 0
        const int je_run = __shfl_sync(0xffffffffu, je, min(31, lane + runlen));
        while (heads) {
            const int h = __ffs((int)heads) - 1;
            heads &= heads - 1;
            const int rb = __shfl_sync(0xffffffffu, jb, h), re = __shfl_sync(0xffffffffu, je_run, h), rcode = __shfl_sync(0xffffffffu, code, h);
            sweep(rb, re, rcode);
        }
    }
}

// ---------------------------------------------------------------------------------------
// az_k_pair_warp: the reference schedule (bin + sort every step, no list) with one WARP per atom.  The lanes
// stride over the atoms of the stencil cells straight from the sorted stream (coalesced float4 loads; z-adjacent
// cells are merged into one range), every lane accumulates its share of the force on atom i and a shuffle
// tree adds the 32 shares.  Same cell-pair rules (az_neighbor), same pair arithmetic and per-cell-pair image
// shift as az_k_pair.  Why it exists: the case-study decks hold 4 000 / 40 000 atoms -- one thread per atom is
// 125 / 1 250 warps on 148 SMs and each thread walks 500-800 candidates serially (measured: 790 us per step at
// 4 000 atoms); one warp per atom is 32x the parallelism and 1/32 of the serial chain.
#define AZ_WARP_TPB 256
#ifndef AZ_WARP_MINB
#define AZ_WARP_MINB 1
#endif
template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
__global__ void __launch_bounds__(AZ_WARP_TPB, AZ_WARP_MINB)
az_k_pair_warp(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    __shared__ AzVdw s_tab[SINGLE ? 1 : AZ_MAX_SPECIES * AZ_MAX_SPECIES];
    __shared__ double s_red[96];
    az_pdl_wait(); // (programmatic dependent launch behind the sort: a no-op on a plain launch)
    const AzGrid& g = P.g;
    const int tid = threadIdx.x, lane = tid & 31;
    if (!SINGLE) {
        for (int k = tid; k < P.nspec * P.nspec; k += AZ_WARP_TPB) s_tab[k] = A.table[k];
        __syncthreads();
    }
    const int i = blockIdx.x * (AZ_WARP_TPB / 32) + (tid >> 5);
    const bool act = i < az_n_owned(P);
    float fx = 0.f, fy = 0.f, fz = 0.f, ev = 0.f, ec = 0.f;
    int hits = 0, ti = 0;
    if (act) {
        const float4 pi = A.pos[i];
        ti = __float_as_int(pi.w);
        float radi = 0.f;
        if (RADII) radi = A.therm[i].y;
        const int ony = g.ohi[1] - g.olo[1], onz = g.ohi[2] - g.olo[2];
        const int ci = A.cellOfSorted[i];
        const int cxy = ci / onz;
        const int gcz = g.olo[2] + (ci - cxy * onz) + g.goff[2];
        const int gcy = g.olo[1] + cxy % ony + g.goff[1];
        const int gcx = g.olo[0] + cxy / ony + g.goff[0];
        constexpr bool FOLD = SINGLE && CK == AZ_ELEC_NONE;
        const float gate = FOLD ? fminf(P.r2max, A.single.r2cut) : P.r2max;

        // one candidate per lane: delta, gate, and -- only if some lane has a partner in range -- the pair term
        auto cand = [&](const float4 pj, const int j, const bool valid, const int code) {
            float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            if (code) { // (x_i - x_j) - shift, exactly like the reference's secShifts
                dx -= ((float)((code >> 1) & 1) - (float)(code & 1)) * P.L[0];
                dy -= ((float)((code >> 3) & 1) - (float)((code >> 2) & 1)) * P.L[1];
                dz -= ((float)((code >> 5) & 1) - (float)((code >> 4) & 1)) * P.L[2];
            }
            const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            const bool in = valid & (r2 <= gate) & (j != i);
            // dilute systems (case study 1: 490 candidates per atom, almost no pair inside 4 A): skip the pair term
            // when no lane has a partner in range
            if (!__any_sync(0xffffffffu, in)) return;
            const int tj = __float_as_int(pj.w);
            float radj = 0.f;
            if (RADII) radj = A.therm[valid ? j : i].y;
            float e1 = 0.f, e2 = 0.f, f;
            if (SINGLE) f = az_pair<VK, CK, FOLD>(P, A.single, r2, radi, radj, true, e1, e2);
            else f = az_pair<VK, CK, false>(P, s_tab[ti * P.nspec + tj], r2, radi, radj, ti == tj, e1, e2);
            f = in ? f : 0.f;
            fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
            if (ENERGY) { ev += in ? e1 : 0.f; ec += in ? e2 : 0.f; hits += in ? 1 : 0; }
        };
        auto sweep = [&](const int rb, const int re, const int code) {
            // uniform trip count (the vote needs the whole warp), two loads in flight per lane: the step of the small
            // decks is bound by the latency of these loads, not by arithmetic
            for (int j0 = rb; j0 < re; j0 += 64) {
                const int ja = j0 + lane, jb = ja + 32;
                const bool va = ja < re, vb = jb < re;
                const float4 pa = __ldg(A.pos + (va ? ja : rb)), pb = __ldg(A.pos + (vb ? jb : rb));
                cand(pa, ja, va, code);
                if (j0 + 32 < re) cand(pb, jb, vb, code);
            }
        };
        if (A.stenTab) az_warp_table_ranges(g, A.stenTab, A.stenS, P.r2list, ci, gcx, gcy, gcz, A.first, A.count, lane, sweep, A.prune != 0, pi.x, pi.y, pi.z);
        else az_warp_stencil_ranges(g, g.sten, P.r2list, gcx, gcy, gcz, A.first, A.count, lane, sweep, A.prune != 0, pi.x, pi.y, pi.z);
    }
    // the 32 shares of atom i (all lanes of a warp belong to the same atom, so the whole warp is in or out)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fx += __shfl_down_sync(0xffffffffu, fx, o);
        fy += __shfl_down_sync(0xffffffffu, fy, o);
        fz += __shfl_down_sync(0xffffffffu, fz, o);
    }
    if (act && lane == 0) {
        const float q = P.charge[ti];
        A.frc[i] = make_float4(-q * P.efield[0] + fx, -q * P.efield[1] + fy, -q * P.efield[2] + fz, 0.f);
    }
    if (ENERGY) {
        double vv[3] = { 0.5 * (double)ev, 0.5 * (double)ec, (double)hits };
        az_block_sum<3>(vv, s_red);
        if (tid == 0) { A.partV[blockIdx.x] = vv[0]; A.partC[blockIdx.x] = vv[1]; A.partN[blockIdx.x] = (unsigned long long)(vv[2] + 0.5); }
    }
}


// ---------------------------------------------------------------------------------------
// az_k_pair_thread: the reference schedule with one THREAD per atom, for systems whose cells are much wider than the cut-off
// (case study 1 of the reference: a dilute gas, 4 A cut-off binned into 87.8 A cells, 18 atoms per cell).  There an atom
// reaches 1-2 of its 27 stencil cells (the same atom-to-cell distance test as az_k_pair_warp's `prune`), i.e. ~25
// candidates: a warp per atom spends its time on the stencil bookkeeping and the shuffle tree (28 us per step at 40 000
// atoms), a thread per atom walks the few cells itself.  Same cell-pair rules (az_neighbor, cell_dist_int), same pair
// arithmetic and per-cell-pair image shift as the other sweeps.  MEASURED AND REJECTED as a default (B200, 40 000 atoms, case study
// 1 geometry, profiles/r2/bench_cs1_thread_sweep.json): 39.6 us per launch against 28 us for the warp sweep -- the 32 atoms of a
// warp keep DIFFERENT stencil cells (each face neighbour is kept by 4.5 % of the atoms, i.e. by some lane of 77 % of the
// warps), so a warp walks the union, ~6 cells instead of 1.3, with one or two lanes active.  pair_variant 4 selects it.
#define AZ_THREAD_TPB 128
template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
__global__ void __launch_bounds__(AZ_THREAD_TPB)
az_k_pair_thread(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    __shared__ AzVdw s_tab[SINGLE ? 1 : AZ_MAX_SPECIES * AZ_MAX_SPECIES];
    __shared__ double s_red[96];
    const AzGrid& g = P.g;
    const int tid = threadIdx.x;
    if (!SINGLE) {
        for (int k = tid; k < P.nspec * P.nspec; k += AZ_THREAD_TPB) s_tab[k] = A.table[k];
        __syncthreads();
    }
    const int i = blockIdx.x * AZ_THREAD_TPB + tid;
    const bool act = i < az_n_owned(P);
    float ev = 0.f, ec = 0.f;
    int hits = 0;
    if (act) {
        const float4 pi = A.pos[i];
        const int ti = __float_as_int(pi.w);
        float radi = 0.f;
        if (RADII) radi = A.therm[i].y;
        const int ony = g.ohi[1] - g.olo[1], onz = g.ohi[2] - g.olo[2];
        const int ci = A.cellOfSorted[i];
        const int cxy = ci / onz;
        const int gcz = g.olo[2] + (ci - cxy * onz) + g.goff[2];
        const int gcy = g.olo[1] + cxy % ony + g.goff[1];
        const int gcx = g.olo[0] + cxy / ony + g.goff[0];
        constexpr bool FOLD = SINGLE && CK == AZ_ELEC_NONE;
        const float gate = FOLD ? fminf(P.r2max, A.single.r2cut) : P.r2max;
        const bool prune = A.prune != 0;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        // distance from the atom to the box of the (unwrapped) neighbour cell along one axis, 1e-4 of a cell edge of margin
        auto gap = [&](const int o, const int gc, const float x, const float cs) {
            const float d = o > 0 ? (float)(gc + o) * cs - x : (o < 0 ? x - (float)(gc + o + 1) * cs : 0.f);
            return fmaxf(0.f, d - 1e-4f * cs);
        };
        for (int ox = -g.sten[0]; ox <= g.sten[0]; ox++) {
            int sx, shxi;
            if (!az_neighbor(g, 0, gcx, ox, sx, shxi)) continue;
            const int aox = ox < 0 ? -ox : ox;
            const float mx = aox ? (float)(aox - 1) * g.csize[0] : 0.f;
            const float gx = prune ? gap(ox, gcx, pi.x, g.csize[0]) : 0.f;
            if (gx * gx > P.r2list) continue;
            for (int oy = -g.sten[1]; oy <= g.sten[1]; oy++) {
                int sy, shyi;
                if (!az_neighbor(g, 1, gcy, oy, sy, shyi)) continue;
                const int aoy = oy < 0 ? -oy : oy;
                const float my = aoy ? (float)(aoy - 1) * g.csize[1] : 0.f;
                const float mxy2 = mx * mx + my * my;
                if (mxy2 > P.r2list) continue;
                const float gy = prune ? gap(oy, gcy, pi.y, g.csize[1]) : 0.f;
                const float gxy2 = gx * gx + gy * gy;
                if (gxy2 > P.r2list) continue;
                const int colbase = (sx * g.nl[1] + sy) * g.nl[2];
                for (int oz = -g.sten[2]; oz <= g.sten[2]; oz++) {
                    int sz, shzi;
                    if (!az_neighbor(g, 2, gcz, oz, sz, shzi)) continue;
                    const int aoz = oz < 0 ? -oz : oz;
                    const float mz = aoz ? (float)(aoz - 1) * g.csize[2] : 0.f;
                    if (mxy2 + mz * mz > P.r2list) continue;
                    const float gz = prune ? gap(oz, gcz, pi.z, g.csize[2]) : 0.f;
                    if (gxy2 + gz * gz > P.r2list) continue;
                    const int cid = az_cell_id(g, colbase + sz);
                    const int jb = A.first[cid], je = jb + A.count[cid];
                    const bool img = (shxi | shyi | shzi) != 0;
                    const float shx = (float)shxi * P.L[0], shy = (float)shyi * P.L[1], shz = (float)shzi * P.L[2];
#pragma unroll 2
                    for (int j = jb; j < je; j++) {
                        const float4 pj = __ldg(A.pos + j);
                        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                        if (img) { dx -= shx; dy -= shy; dz -= shz; } // (x_i - x_j) - shift, exactly like the reference's secShifts
                        const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                        const bool in = (r2 <= gate) & (j != i);
                        if (!in) continue; // dilute: most candidates are out of range, and there is no warp-wide work to keep uniform
                        const int tj = __float_as_int(pj.w);
                        float radj = 0.f;
                        if (RADII) radj = A.therm[j].y;
                        float e1 = 0.f, e2 = 0.f, f;
                        if (SINGLE) f = az_pair<VK, CK, FOLD>(P, A.single, r2, radi, radj, true, e1, e2);
                        else f = az_pair<VK, CK, false>(P, s_tab[ti * P.nspec + tj], r2, radi, radj, ti == tj, e1, e2);
                        fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
                        if (ENERGY) { ev += e1; ec += e2; hits++; }
                    }
                }
            }
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


// ---------------------------------------------------------------------------------------
// az_k_rdf: radial distribution histogram from the cell grid (replaces the O(N^2) brute_rdf,
// src/cuStat.cu:436-512).  One warp per atom, lanes stride over the atoms of the stencil cells of the RDF radius; every
// unordered pair inside the radius is counted once (j > i) in a per-CTA shared histogram of integers, flushed with
// integer atomics: the result is exact and independent of the execution order (the reference accumulates
// volume / (N_a N_b) per pair with float atomics).  Bin and species-pair index are the reference's:
// iR = (int)(sqrt((double)r2) * idRDF), iPair = mn * (nSpec - 1) + mn * (1 - mn) / 2 + mx.
struct AzRdfArgs {
    const float4* pos; const int* cellOfSorted; const int* first; const int* count;
    const int* id;            // bricks: persistent ids of the sorted atoms, ghosts included (NULL on one GPU)
    unsigned long long* hist; // [nbins * npair], bin-major like md->rdf
    int sten[3];
    int nbins, npair;
    float r2max, idrdf, r2stencil;
};

#define AZ_RDF_TPB 256
static __global__ void __launch_bounds__(AZ_RDF_TPB)
az_k_rdf(const __grid_constant__ AzParams P, const __grid_constant__ AzRdfArgs A)
{
    AZ_DYN_SMEM(unsigned, s_hist); // nbins * npair counters
    const AzGrid& g = P.g;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nh = A.nbins * A.npair;
    for (int k = tid; k < nh; k += AZ_RDF_TPB) s_hist[k] = 0u;
    __syncthreads();
    const int i = blockIdx.x * (AZ_RDF_TPB / 32) + (tid >> 5);
    if (i < az_n_owned(P)) {
        const float4 pi = A.pos[i];
        const int ti = __float_as_int(pi.w);
        const int ony = g.ohi[1] - g.olo[1], onz = g.ohi[2] - g.olo[2];
        const int ci = A.cellOfSorted[i];
        const int cxy = ci / onz;
        const int gcz = g.olo[2] + (ci - cxy * onz) + g.goff[2];
        const int gcy = g.olo[1] + cxy % ony + g.goff[1];
        const int gcx = g.olo[0] + cxy / ony + g.goff[0];
        const int m = P.nspec - 1;
        const int idi = A.id ? A.id[i] : 0;
        auto sweep = [&](const int rb, const int re, const int code) {
            const float shx = ((float)((code >> 1) & 1) - (float)(code & 1)) * P.L[0];
            const float shy = ((float)((code >> 3) & 1) - (float)((code >> 2) & 1)) * P.L[1];
            const float shz = ((float)((code >> 5) & 1) - (float)((code >> 4) & 1)) * P.L[2];
            // every unordered pair once: j > i on one GPU; on bricks the pair belongs to the brick that owns the atom with
            // the lower persistent id (the other brick sees the same pair with the roles of owner and ghost exchanged)
            for (int j = (A.id ? rb : max(rb, i + 1)) + lane; j < re; j += 32) {
                if (A.id && A.id[j] <= idi) continue;
                const float4 pj = __ldg(A.pos + j);
                float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                if (code) { dx -= shx; dy -= shy; dz -= shz; }
                const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy)); // nvcc's contraction of dx*dx + dy*dy + dz*dz, as in the force kernels
                if (r2 < A.r2max) {
                    const int iR = (int)(sqrt((double)r2) * (double)A.idrdf);
                    const int tj = __float_as_int(pj.w);
                    const int mn = min(ti, tj), mx = max(ti, tj);
                    const int iPair = mn * m + mn * (1 - mn) / 2 + mx;
                    if (iR < A.nbins) atomicAdd(&s_hist[iR * A.npair + iPair], 1u);
                }
            }
        };
        az_warp_stencil_ranges(g, A.sten, A.r2stencil, gcx, gcy, gcz, A.first, A.count, lane, sweep);
    }
    __syncthreads();
    for (int k = tid; k < nh; k += AZ_RDF_TPB) {
        const unsigned v = s_hist[k];
        if (v) atomicAdd(&A.hist[k], (unsigned long long)v);
    }
}

#include "az_tlist.cuh"
