// az_tlist.cuh -- tile lists: the Verlet-list path with the neighbourhood of a CTA staged in shared memory (sm_100a).
//
// Why.  ncu on B200 (profiles/r2/): az_k_pair_list is bound by the L1 data pipe -- every list row is one
// LDG.128 gather with 32 unrelated addresses, ~6 distinct 128-byte lines per request, ~13 cycles of the pipe per row -- and
// az_k_build_list by the same pipe, three quarters of it the predicated row stores (the lanes of a warp sit at different row
// counts, so one store instruction touches ~7 lines).  The microbenchmarks profiles/r2/micro/tile_gather_probe{,2}.cu measured
// what a row costs when the j positions come from a shared-memory copy instead: one LDS.128 per row is no better (8 bank
// quads), three LDS.32 from a structure-of-arrays copy are (32 banks; the k-th neighbours of a warp fall into a window of
// ~46 consecutive atoms, i.e. at most two words per bank): 0.35 -> 0.27 ms for the rows of BASELINE config 3.
//
// What.  A tile = the AZ_LIST_TPB consecutive sorted atoms of one CTA (the same blocks as az_k_pair_list, so block flags,
// partial sums and the riding integrator are shared).  At a rebuild, az_k_build_tlist
//   1. collects the candidate ranges of the tile's cells (the cell-pair rule of the reference, cell_dist_int, unchanged:
//      az_list_ranges enumerates exactly what az_k_build_list sweeps), merges them into the tile's FOOTPRINT: a few runs
//      of the sorted stream (9 columns x one z-interval for a tile inside one column), written to tileRuns;
//   2. sweeps the candidates per atom like az_k_build_list (same distance test, same order) and records a hit as its
//      16-bit position INSIDE the footprint.  Rows are appended to a shared-memory column per thread (conflict-free, one
//      wavefront per store) and flushed once, in lockstep over the row index, two rows per 32-bit word: every global
//      store instruction writes one full 128-byte line, and the list is half the bytes of the 32-bit format.
// Every step, az_k_pair_tlist copies the footprint (x, y, z[, type][, radius] as separate arrays) into shared memory --
// coalesced float4 loads, ~11x the tile's own atoms, served by L2 -- and runs the row loop on it: one coalesced 32-bit
// index load per two rows, three LDS.32 per row, the pair arithmetic of az_k_pair_list instruction for instruction, so
// the forces are BITWISE those of the global-row format (tests: tile_list_bitwise).  Neighbours through a periodic face
// keep the 32-bit rows of az_k_pair_list (global index + image code, rows cap-1 downwards in A.nbr): they are few.
// On bricks the footprint simply contains ghost atoms; in the fused-halo launch (phase 3) the staging loop takes them
// from the peer-memory receive buffer, so the hot loop is the same code for every block.
#pragma once

#define AZ_TL_RUNS 64  // runs of the sorted stream a footprint may consist of
#ifndef AZ_TL_MINB
#define AZ_TL_MINB 10  // resident CTAs per SM the row loop is compiled for (48 registers); shared memory usually allows 9-11
#endif

// The stencil of storage cell `ci` as ranges [rb, re) of the sorted stream with their image code, adjacent ranges merged:
// exactly the ranges az_k_build_list sweeps, in the same order (FAST: z-intervals per column; else cell by cell).
// FAST only: zspan > 0 asks for the union over the cells ci .. ci + zspan of the same column (z-interval stretched upwards).
template <bool FAST, class F>
__device__ __forceinline__ void az_list_ranges(const AzParams& P, const int* __restrict__ first, const int* __restrict__ count, const int ci, const int zspan, F&& emit)
{
    const AzGrid& g = P.g;
    const float lim = P.r2list;
    const int ony = g.ohi[1] - g.olo[1], onz = g.ohi[2] - g.olo[2];
    const int cxy = ci / onz;
    const int gcz = g.olo[2] + (ci - cxy * onz) + g.goff[2];
    const int gcy = g.olo[1] + cxy % ony + g.goff[1];
    const int gcx = g.olo[0] + cxy / ony + g.goff[0];
    int rb = 0, re = 0, rcode = 0; // pending range
    auto add_range = [&](const int jb, const int je, const int code) {
        if (jb == re && code == rcode) re = je;
        else {
            if (re > rb) emit(rb, re, rcode);
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
                const int ua = gcz - kz, ub = gcz + zspan + kz;
                const int colbase = (sx * g.nl[1] + sy) * g.nl[2];
#pragma unroll 1
                for (int shzi = -1; shzi <= 1; shzi++) { // the interval, split where it crosses a periodic face
                    const int a = max(ua, shzi * nz), b = min(ub, shzi * nz + nz - 1);
                    if (a > b) continue;
                    const int off = g.dec[2] ? g.goff[2] : shzi * nz;
                    const int cid_a = az_cell_id(g, colbase + a - off), cid_b = az_cell_id(g, colbase + b - off);
                    const int code = az_shift_code(shxi, shyi, shzi);
                    if (cid_b - cid_a == b - a) add_range(first[cid_a], first[cid_b + 1], code);
                    else
                        for (int u = a; u <= b; u++) { // ghost cells of a brick are not numbered along z
                            const int cid = az_cell_id(g, colbase + u - off);
                            const int jb = first[cid];
                            add_range(jb, jb + count[cid], code);
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
                    const int jb = first[cid], je = jb + count[cid];
                    add_range(jb, je, az_shift_code(shxi, shyi, shzi));
                }
            }
        }
    }
    if (re > rb) emit(rb, re, rcode);
}

// 16-bit rows, two per word: row k of atom i is half (k & 1) of word nbr16[((i >> 5) * cap2 + (k >> 1)) * 32 + (i & 31)]
__device__ __forceinline__ size_t az_tl_word_base(int i, int cap2) { return ((size_t)(i >> 5) * (size_t)cap2) * 32 + (size_t)(i & 31); }

// az_k_build_tlist: the rebuild step of the tile-list path (see the file header).  Dynamic shared memory: the row columns,
// A.tl_rows_smem (even) rows x AZ_LIST_TPB threads x 2 bytes; a thread with more rows than that stores the rest straight
// to global memory (correct, slow, rare: the host sizes the columns from the density).
template <bool FAST>
__global__ void __launch_bounds__(AZ_LIST_TPB)
az_k_build_tlist(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    AZ_DYN_SMEM(unsigned short, s_rows);
    __shared__ int s_runS[AZ_TL_RUNS + 1], s_runE[AZ_TL_RUNS + 1], s_runO[AZ_TL_RUNS + 1];
    __shared__ int s_nraw, s_nrun, s_ghost, s_bad;
    const int tid = threadIdx.x, tile = (int)blockIdx.x;
    const int i = tile * AZ_LIST_TPB + tid;
    const int n_owned = az_n_owned(P);
    const bool act = i < n_owned;
    // the candidate ranges are collected in the (not yet used) row columns: S rows x 256 bytes = S * 32 entries, a power of two of them
    int2* __restrict__ const s_raw = (int2*)s_rows;
    const int rawcap = 1 << (31 - __clz(A.tl_rows_smem * (AZ_LIST_TPB / 4)));
    if (tid == 0) { s_nraw = 0; s_nrun = 0; s_ghost = 0; s_bad = 0; }
    __syncthreads();

    // ---- 1. candidate ranges of the tile's cells.  All atoms of a cell see the same ranges, and on a FAST grid the cells of one
    // column see z-intervals that overlap: one thread per distinct column of the tile (else: per distinct cell) speaks
    int ci = 0;
    if (act) {
        ci = A.cellOfSorted[i];
        bool leader;
        int zspan = 0;
        if (FAST) {
            const int onz = P.g.ohi[2] - P.g.olo[2];
            const int col = ci / onz;
            leader = tid == 0 || A.cellOfSorted[i - 1] / onz != col;
            if (leader) { // the tile's cells of this column: up to the tile's last atom, or to the end of the column
                const int cl = A.cellOfSorted[min(n_owned, (tile + 1) * AZ_LIST_TPB) - 1];
                zspan = (cl / onz == col) ? cl - ci : onz - 1 - (ci - col * onz);
            }
        } else leader = tid == 0 || A.cellOfSorted[i - 1] != ci;
        if (leader)
            az_list_ranges<FAST>(P, A.first, A.count, ci, zspan, [&](const int rb, const int re, const int code) {
                if (code != 0) return; // neighbours through a periodic face keep the 32-bit rows
                const int k = atomicAdd(&s_nraw, 1);
                if (k < rawcap) s_raw[k] = make_int2(rb, re);
            });
    }
    __syncthreads();
    // ---- 2. their union as sorted runs: bitonic sort by first atom, then one pass that opens a new run wherever a range starts
    // behind everything seen so far (ranges that overlap or touch merge)
    int nraw = s_nraw;
    if (nraw > rawcap && A.n_total <= A.tl_foot_cap) { // too many ranges to merge, but the whole (small) system fits: hold it whole
        __syncthreads();
        if (tid == 0) s_raw[0] = make_int2(0, A.n_total);
        nraw = 1;
        __syncthreads();
    }
    const int R = min(nraw, rawcap);
    if (R <= 32) {
        // the usual case (9 ranges per column of the tile): one warp sorts them in registers and finds the runs with shuffles
        if (tid < 32) {
            int2 v = tid < R ? s_raw[tid] : make_int2(0x7fffffff, 0x7fffffff);
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const int ox = __shfl_xor_sync(0xffffffffu, v.x, j), oy = __shfl_xor_sync(0xffffffffu, v.y, j);
                    const bool up = (tid & k) == 0, low = (tid & j) == 0;
                    // the lower lane of a pair keeps the smaller key when sorting upwards
                    const bool take = (low == up) ? (ox < v.x) : (ox > v.x);
                    if (take) { v.x = ox; v.y = oy; }
                }
            // exclusive running maximum of the ends: a range opens a run iff it starts behind everything before it
            int pm = v.y;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pm, o); if (tid >= o) pm = max(pm, t); }
            const int incl = pm; // inclusive maximum
            int excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (tid == 0) excl = -1;
            const bool valid = tid < R;
            const bool isS = valid && (tid == 0 || v.x > excl);
            const unsigned ms = __ballot_sync(0xffffffffu, isS);
            const int nrun = __popc(ms);
            const int rid = __popc(ms & ((2u << tid) - 1u)) - 1; // run of this lane's range
            if (isS && rid < AZ_TL_RUNS) { s_runS[rid] = v.x; if (rid > 0) s_runE[rid - 1] = excl; }
            if (valid && tid == R - 1 && nrun >= 1 && nrun <= AZ_TL_RUNS) s_runE[nrun - 1] = incl;
            if (tid == 0) s_nrun = nrun;
        }
        __syncthreads();
    } else {
    int N = 2;
    while (N < R) N <<= 1;
    for (int k = R + tid; k < N; k += AZ_LIST_TPB) s_raw[k] = make_int2(0x7fffffff, 0x7fffffff);
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (N >> 1); t += AZ_LIST_TPB) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                const int2 a = s_raw[lo], b = s_raw[hi];
                if ((a.x > b.x) == ((lo & k) == 0)) { s_raw[lo] = b; s_raw[hi] = a; }
            }
            __syncthreads();
        }
    if (tid == 0) {
        int nrun = 0, curE = -1;
        for (int r = 0; r < R; r++) {
            const int2 v = s_raw[r];
            if (nrun == 0 || v.x > curE) {
                if (nrun > 0 && nrun <= AZ_TL_RUNS) s_runE[nrun - 1] = curE;
                if (nrun < AZ_TL_RUNS) s_runS[nrun] = v.x;
                nrun++;
                curE = v.y;
            } else curE = max(curE, v.y);
        }
        if (nrun > 0 && nrun <= AZ_TL_RUNS) s_runE[nrun - 1] = curE;
        s_nrun = nrun;
    }
    __syncthreads();
    }
    if (tid == 0) {
        const int nrun = s_nrun;
        bool bad = nraw > rawcap || nrun > AZ_TL_RUNS;
        int off = 0;
        if (!bad) {
            for (int r = 0; r < nrun; r++) { s_runO[r] = off; off += s_runE[r] - s_runS[r]; }
            s_runO[nrun] = off;
            if (off > A.tl_foot_cap || (off - 1) * A.tl_stride > 65535) bad = true;
        }
        if (A.tlStat) { // largest footprint / run count seen (for the error message and for sizing)
            if (off > *(volatile int*)&A.tlStat[0]) atomicMax(&A.tlStat[0], off);
            if (nrun > *(volatile int*)&A.tlStat[1]) atomicMax(&A.tlStat[1], nrun);
            if (nraw > *(volatile int*)&A.tlStat[2]) atomicMax(&A.tlStat[2], nraw);
        }
        if (bad) atomicOr(A.flags, AZ_FLAG_FOOT);
        s_bad = bad ? 1 : 0;
        s_nrun = nrun;
        A.tileHdr[tile] = bad ? 0 : nrun;
    }
    __syncthreads();
    const bool bad = s_bad != 0; // the tile is left without neighbours; the sticky flag turns into AZ_ERR_OVERFLOW at the next az_engine_stats
    const int nrun = bad ? 0 : s_nrun;
    if (!bad) { // (first atom in the sorted stream, place in the footprint, length)
        int4* __restrict__ runs = A.tileRuns + (size_t)tile * AZ_TL_RUNS;
        for (int r = tid; r < nrun; r += AZ_LIST_TPB) runs[r] = make_int4(s_runS[r], s_runO[r], s_runO[r + 1] - s_runO[r], 0);
    }

    // ---- 3. the sweep of az_k_build_list; a hit is recorded as its place in the footprint
    const int cap = A.nbr_cap, cap2 = (cap + 1) >> 1;
    const int S = A.tl_rows_smem;
    unsigned short* __restrict__ const mine = s_rows + tid;
    int nA = 0, nB = 0, jmax = -1;
    if (act && !bad) {
        const float4 pi = A.pos[i];
        int* __restrict__ rowsB = A.nbr + az_nbr_base(i, cap);
        unsigned short* __restrict__ const far16 = (unsigned short*)(A.nbr16 + az_tl_word_base(i, cap2)); // row k: far16[(k >> 1) * 64 + (k & 1)]
        const float lim = P.r2list;
        auto put = [&](const int k, const int v) {
            if (k < S) mine[k * AZ_LIST_TPB] = (unsigned short)v;
            else far16[(k >> 1) * 64 + (k & 1)] = (unsigned short)v;
        };
        az_list_ranges<FAST>(P, A.first, A.count, ci, 0, [&](const int rb, const int re, const int code) {
            jmax = max(jmax, re - 1);
            if (code == 0) {
                int r = 0;
                while (r + 1 < nrun && s_runS[r + 1] <= rb) r++; // the run that holds this range
                const int delta = s_runO[r] - s_runS[r];
                const int str = A.tl_stride; // a row holds the WORD offset of the neighbour in the pass kernel's shared copy
                const int room = cap - nB;
                if ((i < rb || i >= re) && nA + (re - rb) <= min(room, S)) {
                    // the common case: atom i is not in the range and neither the rows nor the shared column can run out.
                    // Two candidates per FADD2 / FFMA2; each half is rounded like the scalar fmaf chain of the force pass.
                    int j = rb, o = nA * AZ_LIST_TPB;
                    const int jd = delta;
#pragma unroll 2
                    for (; j + 2 <= re; j += 2) {
                        const float4 pa = __ldg(A.pos + j), pb = __ldg(A.pos + j + 1);
                        const az_f2 dx = az_pk(pi.x - pa.x, pi.x - pb.x), dy = az_pk(pi.y - pa.y, pi.y - pb.y), dz = az_pk(pi.z - pa.z, pi.z - pb.z);
                        const az_f2 r2 = az_fma2(dz, dz, az_fma2(dx, dx, az_mul2(dy, dy)));
                        if (az_lo(r2) <= lim) { mine[o] = (unsigned short)((j + jd) * str); o += AZ_LIST_TPB; }
                        if (az_hi(r2) <= lim) { mine[o] = (unsigned short)((j + 1 + jd) * str); o += AZ_LIST_TPB; }
                    }
                    if (j < re) {
                        const float4 pj = __ldg(A.pos + j);
                        const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                        if (fmaf(dz, dz, fmaf(dx, dx, dy * dy)) <= lim) { mine[o] = (unsigned short)((j + jd) * str); o += AZ_LIST_TPB; }
                    }
                    nA = o / AZ_LIST_TPB;
                    return;
                }
#pragma unroll 4
                for (int j = rb; j < re; j++) {
                    const float4 pj = __ldg(A.pos + j);
                    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                    const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                    const bool hit = (r2 <= lim) & (j != i);
                    if (hit & (nA < room)) put(nA, (j + delta) * str);
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
                    if (hit & (nB < room)) rowsB[(cap - 1 - nB) * 32] = j | (code << AZ_NBR_BITS);
                    nB += hit ? 1 : 0;
                }
            }
        });
        if (nA + nB > cap) { atomicOr(A.flags, AZ_FLAG_NBR); atomicMax(A.flags + 1, nA + nB); nB = min(nB, cap); nA = min(nA, cap - nB); }
        A.nbrCount[i] = nA | (nB << AZ_NBR_CNT_BITS);
    } else if (act) A.nbrCount[i] = 0;
    __syncthreads();

    // ---- 4. flush the shared columns: the warp walks the row index in lockstep, one 128-byte line per store instruction
    {
        const int nS = min(nA, S);
        const int kmax = az_warp_max_int(nS);
        if (A.tlStat) { // longest row column of this rebuild (sizes the shared columns)
            const int rmax = az_warp_max_int(nA + nB);
            if ((tid & 31) == 0 && rmax > *(volatile int*)&A.tlStat[3]) atomicMax(&A.tlStat[3], rmax);
        }
        unsigned* __restrict__ w = A.nbr16 + az_tl_word_base(i, cap2);
        for (int k = 0; k < kmax; k += 2)
            if (k < nS) w[(k >> 1) * 32] = (unsigned)mine[k * AZ_LIST_TPB] | ((unsigned)mine[(k + 1) * AZ_LIST_TPB] << 16);
    }
    if (A.blockFlag) { // ghosts (and the line they share with the last owned atoms) sit behind the owned atoms in the sorted streams
        const int gl = n_owned & ~7;
        if (jmax >= gl || (i >= gl && i < n_owned)) atomicOr(&s_ghost, 1);
        __syncthreads();
        if (tid == 0) A.blockFlag[tile] = s_ghost;
    }
}

__device__ __forceinline__ unsigned az_ld_stream_u32(const unsigned* p)
{
#ifdef AZ_EMU
    return *p;
#else
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
}

// az_k_pair_tlist: the force pass of every step of the tile-list path.  Same blocks, same arguments, same pair arithmetic in
// the same order as az_k_pair_list (bitwise the same forces); the j positions of the plain rows come from the shared-memory
// copy of the tile's footprint.  Dynamic shared memory: tl_foot_cap x (3 or 5) words.
// A.phase as in az_k_pair_list; phase 3 (fused halo) differs from the others only in where the staging loop finds a ghost.
template <int VK, int CK, bool SINGLE, bool RADII, bool ENERGY>
__global__ void __launch_bounds__(AZ_LIST_TPB, AZ_TL_MINB)
az_k_pair_tlist(const __grid_constant__ AzParams P, const __grid_constant__ AzPairArgs A)
{
    // the footprint in shared memory: STR words per atom -- x, y, z [, type, radius].  An odd stride keeps the lanes of a gather
    // in distinct banks exactly when their atoms differ mod 32 (like separate arrays would), and the three loads of a row
    // share one register: a list row holds STR * (place in the footprint), so its loads are LDS [R.X4], [R.X4+4], [R.X4+8]
    constexpr int STR = (SINGLE && !RADII) ? 3 : 5;
    AZ_DYN_SMEM(float, s_f);
    __shared__ AzVdw s_tab[SINGLE ? 1 : AZ_MAX_SPECIES * AZ_MAX_SPECIES];
    __shared__ double s_red[96];
    __shared__ double s_kw[AZ_LIST_TPB / 32];
    __shared__ int s_done;
    const int tid = threadIdx.x;
    const int ob = (A.phase == 4) ? A.blockOrder[blockIdx.x] : (int)blockIdx.x; // direct halo: launch order, bit 30 = lists ghosts
    const int blk = ob & 0x3fffffff;
    const bool ghosts = A.phase == 3;
    if (A.phase == 4) { // direct halo: the neighbours store this step's ghosts straight into my position stream
        if ((ob >> 30) && tid < 27 && ((A.haloDirs >> tid) & 1)) az_wait_flag(A.haloFlags + tid, A.haloSeq, A.flags);
    } else if (ghosts) {
        if (A.blockFlag[blk] == 0) return; // phase 1 (the other launch) owns the blocks that list no ghost
        // the neighbours' stores for this step must have landed (az_k_ghost_put raises the flags over NVLink)
        if (tid < 27 && ((A.haloDirs >> tid) & 1)) az_wait_flag(A.haloFlags + tid, A.haloSeq, A.flags);
    } else if (A.phase && ((A.phase == 1) == (A.blockFlag[blk] != 0))) return; // the other phase owns this block
    for (int k = tid; k < P.nspec * P.nspec && !SINGLE; k += AZ_LIST_TPB) s_tab[k] = A.table[k];
    if (tid == 0) s_done = 0;
    __syncthreads(); // (the arrival flags have been seen)
    const int n_own = az_n_owned(P);
    const int i = blk * AZ_LIST_TPB + tid;
    const bool act = i < n_own;
    // everything the row loop starts with is requested BEFORE the staging, so that these loads (HBM: the list is a pure
    // stream) are in flight while the footprint is copied: own position, row counts, the first eight rows
    const int cap2 = (A.nbr_cap + 1) >> 1;
    const unsigned* __restrict__ row = A.nbr16 + az_tl_word_base(act ? i : 0, cap2);
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt2 = 0;
    unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if (act) {
        pi = A.pos[i];
        cnt2 = A.nbrCount[i];
        w0 = az_ld_stream_u32(row); w1 = az_ld_stream_u32(row + 32); w2 = az_ld_stream_u32(row + 64); w3 = az_ld_stream_u32(row + 96); // (cap2 >= 4 rows of slack exist)
    }
    { // ---- stage the footprint: one warp per run, coalesced float4 loads, four in flight per lane
        const int nrun = A.tileHdr[blk];
        const int4* __restrict__ runs = A.tileRuns + (size_t)blk * AZ_TL_RUNS;
        const int lane = tid & 31, warp = tid >> 5;
        // the first three runs of every warp are fetched without waiting for nrun (12 runs cover the usual 9-10)
        int4 q[3];
#pragma unroll
        for (int u = 0; u < 3; u++) q[u] = __ldg(runs + warp + u * (AZ_LIST_TPB / 32));
        auto stage = [&](const int4 a) {
            for (int k0 = 0; k0 < a.z; k0 += 128) {
                float4 p[4];
                float rd[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = k0 + u * 32 + lane;
                    const int c = a.x + (k < a.z ? k : 0);
                    // fused halo: a neighbour GPU wrote this step's ghost positions into the receive staging (L2 loads)
                    p[u] = (ghosts && c >= n_own) ? az_ld_cg4g(A.ghostStage + A.ghostEntry[c - n_own]) : __ldg(A.pos + c);
                    if (RADII) rd[u] = A.therm[c].y;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = k0 + u * 32 + lane;
                    if (k < a.z) {
                        float* __restrict__ d = s_f + (a.y + k) * STR;
                        d[0] = p[u].x; d[1] = p[u].y; d[2] = p[u].z;
                        if (STR == 5) { d[3] = p[u].w; d[4] = RADII ? rd[u] : 0.f; }
                    }
                }
            }
        };
#pragma unroll
        for (int u = 0; u < 3; u++) if (warp + u * (AZ_LIST_TPB / 32) < nrun) stage(q[u]);
        for (int r = warp + 3 * (AZ_LIST_TPB / 32); r < nrun; r += AZ_LIST_TPB / 32) stage(__ldg(runs + r));
    }
    __syncthreads();
    float fx = 0.f, fy = 0.f, fz = 0.f, ev = 0.f, ec = 0.f, v2 = 0.f;
    double ke = 0.0;
    int hits = 0, ti = 0;
    if (act) {
        ti = __float_as_int(pi.w);
        float radi = 0.f;
        if (RADII) radi = A.therm[i].y;
        const int cnt = cnt2 & ((1 << AZ_NBR_CNT_BITS) - 1), cntB = cnt2 >> AZ_NBR_CNT_BITS;
        constexpr bool FOLD = SINGLE && CK == AZ_ELEC_NONE;
        const float gate = FOLD ? fminf(P.r2max, A.single.r2cut) : P.r2max;
        // one neighbour: pair term from the finished delta, masked accumulate (az_k_pair_list's `one`, instruction for instruction)
        auto one = [&](const float dx, const float dy, const float dz, const int tj, const float radj) {
            const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
            const bool in = r2 <= gate;
            float e1 = 0.f, e2 = 0.f, f;
            if (SINGLE) f = az_pair<VK, CK, FOLD>(P, A.single, r2, radi, radj, true, e1, e2);
            else f = az_pair<VK, CK, false>(P, s_tab[ti * P.nspec + tj], r2, radi, radj, ti == tj, e1, e2);
            f = in ? f : 0.f;
            fx = fmaf(f, dx, fx); fy = fmaf(f, dy, fy); fz = fmaf(f, dz, fz);
            if (ENERGY) { ev += in ? e1 : 0.f; ec += in ? e2 : 0.f; hits += in ? 1 : 0; }
        };
        auto local = [&](const unsigned w) { // a plain row: word offset of the neighbour in the footprint
            const float* __restrict__ q = s_f + w;
            one(pi.x - q[0], pi.y - q[1], pi.z - q[2], STR == 5 ? __float_as_int(q[3]) : 0, RADII ? q[4] : 0.f);
        };
        const int npair = cnt >> 1;
        int kp = 0;
        // software pipeline by hand: the index words of the NEXT eight rows are requested before the current eight are
        // evaluated (unconditionally: words behind the last row exist -- the allocation carries slack -- and are never used)
        for (; kp + 4 <= npair; kp += 4) {
            row += 128;
            const unsigned c0 = w0, c1 = w1, c2 = w2, c3 = w3;
            w0 = az_ld_stream_u32(row); w1 = az_ld_stream_u32(row + 32); w2 = az_ld_stream_u32(row + 64); w3 = az_ld_stream_u32(row + 96);
            local(c0 & 0xffffu); local(c0 >> 16); local(c1 & 0xffffu); local(c1 >> 16);
            local(c2 & 0xffffu); local(c2 >> 16); local(c3 & 0xffffu); local(c3 >> 16);
        }
        { // the last 0..7 rows are in w0..w3
            const int left = cnt - 2 * kp;
            if (left > 0) local(w0 & 0xffffu);
            if (left > 1) local(w0 >> 16);
            if (left > 2) local(w1 & 0xffffu);
            if (left > 3) local(w1 >> 16);
            if (left > 4) local(w2 & 0xffffu);
            if (left > 5) local(w2 >> 16);
            if (left > 6) local(w3 & 0xffffu);
        }
        // neighbours through a periodic face: 32-bit rows cap-1 downwards, image signs above the index bits
        const int* __restrict__ rowB = A.nbr + az_nbr_base(i, A.nbr_cap) + (A.nbr_cap - 1) * 32;
        for (int kb = 0; kb < cntB; kb++) {
            const int c = *rowB; rowB -= 32;
            const int j = c & AZ_NBR_MASK, sc = (int)((unsigned)c >> AZ_NBR_BITS);
            const float4 pj = (ghosts && j >= n_own) ? az_ld_cg4g(A.ghostStage + A.ghostEntry[j - n_own]) : __ldg(A.pos + j);
            float radj = 0.f;
            if (RADII) radj = A.therm[j].y;
            const float sx = ((float)((sc >> 1) & 1) - (float)(sc & 1)) * P.L[0];
            const float sy = ((float)((sc >> 3) & 1) - (float)((sc >> 2) & 1)) * P.L[1];
            const float sz = ((float)((sc >> 5) & 1) - (float)((sc >> 4) & 1)) * P.L[2];
            one((pi.x - pj.x) - sx, (pi.y - pj.y) - sy, (pi.z - pj.z) - sz, __float_as_int(pj.w), radj);
        }
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
        // every warp leaves its sum in shared memory and retires; the LAST warp to arrive adds the four in a fixed order
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
