// az_halo.cuh -- device side of the spatial domain decomposition (new functionality, SURVEY.md 8e;
// the reference is single-GPU).  The periodic box is cut into px*py*pz bricks aligned to the cell
// grid of split_cells; a rank owns the atoms whose cell (count_cell's formula, bit-exact) lies in its
// brick and keeps read-only ghost copies of the atoms in the sten[a] cell layers around it.  Every
// message goes straight to one of the <= 26 neighbouring bricks (no forwarding), so a step needs one
// pack kernel, one grouped exchange and one unpack kernel.
//
//   rebuild step : wrap -> classify/migrate leavers -> select + exchange ghosts -> bin all -> sort
//                  (owned cells are numbered first, so owned atoms stay a contiguous prefix) ->
//                  remap the send lists / ghost slots through the sort permutation
//   other steps  : pack current positions of the listed atoms -> exchange -> overwrite ghost slots
#pragma once
#include "az_kernels.cuh"

#define AZ_NDIR 27
#define AZ_DIR_SELF 13

struct AzDirTable { int off[AZ_NDIR + 1]; }; // prefix offsets of the per-direction segments

struct AzMigRec { // one migrating atom
    float4 pos;   // w = type bits
    float4 vel;   // w = persistent id bits
    float2 therm; // (U, radius)
    float2 pad;
};

__device__ __forceinline__ void az_global_cell(const AzParams& P, const float4 p, int ci[3])
{
    const float x[3] = { p.x, p.y, p.z };
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int c = (int)floor((double)x[a] * (double)P.g.crev[a]); // count_cell, src/cuSort.cu:119
        ci[a] = max(0, min(c, P.g.n[a] - 1));
    }
}

// which neighbouring brick (if any) owns the atom after this step's move
static __global__ void __launch_bounds__(AZ_TPB)
az_k_mig_classify(const __grid_constant__ AzParams P, int n_owned, const float4* __restrict__ pos,
                  int* __restrict__ migCount, int* __restrict__ migList, int migCap, int* __restrict__ stay, int* __restrict__ flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned) return;
    int ci[3];
    az_global_cell(P, pos[i], ci);
    int d[3] = { 0, 0, 0 };
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (!P.g.dec[a]) continue;
        const int n = P.g.n[a], clo = P.g.goff[a] + P.g.olo[a], owned = P.g.ohi[a] - P.g.olo[a];
        int delta = ci[a] - clo;
        if (delta < 0) delta += n; // 0 .. n-1 relative to the brick's first cell
        if (delta >= owned) d[a] = (delta - owned < n - delta) ? 1 : -1; // nearer face decides
    }
    const int dir = (d[0] + 1) * 9 + (d[1] + 1) * 3 + (d[2] + 1);
    if (dir == AZ_DIR_SELF) { stay[i] = 1; return; }
    stay[i] = 0;
    const int slot = atomicAdd(&migCount[dir], 1);
    if (slot < migCap) migList[dir * migCap + slot] = i; else atomicOr(flags, AZ_FLAG_HALO);
}

__device__ __forceinline__ int az_dir_of(const AzDirTable& T, int t)
{
    int d = 0;
#pragma unroll 1
    while (d < AZ_NDIR - 1 && t >= T.off[d + 1]) d++;
    return d;
}

static __global__ void __launch_bounds__(AZ_TPB)
az_k_mig_pack(const __grid_constant__ AzDirTable T, const int* __restrict__ migList, int migCap, const float4* __restrict__ pos,
              const float4* __restrict__ vel, const int* __restrict__ id, const float2* __restrict__ therm, AzMigRec* __restrict__ out)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T.off[AZ_NDIR]) return;
    const int d = az_dir_of(T, t);
    const int i = migList[d * migCap + (t - T.off[d])];
    AzMigRec r;
    r.pos = pos[i];
    r.vel = vel[i]; r.vel.w = __int_as_float(id[i]);
    r.therm = therm ? therm[i] : make_float2(0.f, 0.f);
    r.pad = make_float2(0.f, 0.f);
    out[t] = r;
}

// stayers and arrivals are appended to the other ping-pong buffers in arbitrary order (the sort that
// follows orders by cell and persistent id)
static __global__ void __launch_bounds__(AZ_TPB)
az_k_mig_compact(int n_owned, const int* __restrict__ stay, const float4* __restrict__ pos, const float4* __restrict__ vel,
                 const int* __restrict__ id, const float2* __restrict__ therm, int* __restrict__ counter,
                 float4* __restrict__ pos_o, float4* __restrict__ vel_o, int* __restrict__ id_o, float2* __restrict__ therm_o)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned || !stay[i]) return;
    const int d = atomicAdd(counter, 1);
    pos_o[d] = pos[i]; vel_o[d] = vel[i]; id_o[d] = id[i];
    if (therm) therm_o[d] = therm[i];
}

static __global__ void __launch_bounds__(AZ_TPB)
az_k_mig_unpack(int n_in, const AzMigRec* __restrict__ in, int* __restrict__ counter, int cap, float4* __restrict__ pos_o,
                float4* __restrict__ vel_o, int* __restrict__ id_o, float2* __restrict__ therm_o, int* __restrict__ flags)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_in) return;
    const AzMigRec r = in[t];
    const int d = atomicAdd(counter, 1);
    if (d >= cap) { atomicOr(flags, AZ_FLAG_HALO); return; }
    pos_o[d] = r.pos;
    id_o[d] = __float_as_int(r.vel.w);
    vel_o[d] = make_float4(r.vel.x, r.vel.y, r.vel.z, 0.f);
    if (therm_o) therm_o[d] = r.therm;
}

// owned atoms inside the sten[a] outermost cell layers go to the bricks that need them as ghosts
static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_select(const __grid_constant__ AzParams P, int n_owned, const float4* __restrict__ pos, int* __restrict__ sendCount,
                  int* __restrict__ sendList, int cap, int* __restrict__ flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned) return;
    int ci[3];
    az_global_cell(P, pos[i], ci);
    bool lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int clo = P.g.goff[a] + P.g.olo[a], owned = P.g.ohi[a] - P.g.olo[a];
        const int lc = ci[a] - clo;
        lo[a] = P.g.dec[a] && lc < P.g.sten[a];
        hi[a] = P.g.dec[a] && lc >= owned - P.g.sten[a];
    }
    if (!(lo[0] | hi[0] | lo[1] | hi[1] | lo[2] | hi[2])) return;
    for (int dir = 0; dir < AZ_NDIR; dir++) {
        if (dir == AZ_DIR_SELF) continue;
        const int d[3] = { dir / 9 - 1, (dir / 3) % 3 - 1, dir % 3 - 1 };
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; a++) ok = ok && (d[a] == 0 || (d[a] < 0 ? lo[a] : hi[a]));
        if (!ok) continue;
        const int slot = atomicAdd(&sendCount[dir], 1);
        if (slot < cap) sendList[dir * cap + slot] = i; else atomicOr(flags, AZ_FLAG_HALO);
    }
}

// compact the per-direction lists [dir*cap + k] into one list with the prefix offsets of T
static __global__ void __launch_bounds__(AZ_TPB)
az_k_list_compact(const __grid_constant__ AzDirTable T, const int* __restrict__ lists, int cap, int* __restrict__ out)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T.off[AZ_NDIR]) return;
    const int d = az_dir_of(T, t);
    out[t] = lists[d * cap + (t - T.off[d])];
}

// message = float4 position (w = type) per listed atom, plus (id, radius) when FULL
template <bool FULL>
__global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_pack(int total, const int* __restrict__ sendIdx, const float4* __restrict__ pos, const int* __restrict__ id,
                const float2* __restrict__ therm, float4* __restrict__ out_pos, float2* __restrict__ out_aux)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int i = sendIdx[t];
    out_pos[t] = pos[i];
    if (FULL) out_aux[t] = make_float2(__int_as_float(id[i]), therm ? therm[i].y : 0.f);
    else if (therm) out_aux[t] = make_float2(0.f, therm[i].y);
}

// rebuild: ghosts are appended behind the owned atoms in message order
static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_append(int total, int base, const float4* __restrict__ in_pos, const float2* __restrict__ in_aux, float4* __restrict__ pos,
                  float4* __restrict__ vel, int* __restrict__ id, float2* __restrict__ therm)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    pos[base + t] = in_pos[t];
    vel[base + t] = make_float4(0.f, 0.f, 0.f, 0.f);
    id[base + t] = __float_as_int(in_aux[t].x);
    if (therm) therm[base + t] = make_float2(0.f, in_aux[t].y);
}

// other steps: overwrite the slots the ghosts were sorted into
static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_update(int total, const int* __restrict__ slot, const float4* __restrict__ in_pos, const float2* __restrict__ in_aux,
                  float4* __restrict__ pos, float2* __restrict__ therm)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int s = slot[t];
    pos[s] = in_pos[t];
    if (therm) therm[s].y = in_aux[t].y;
}

// count_cell for owned and ghost atoms of a brick: global cell from the (wrapped) position, then the
// periodic image that falls inside the local grid
static __global__ void __launch_bounds__(AZ_TPB)
az_k_bin(const __grid_constant__ AzParams P, int n_tot, const float4* __restrict__ pos, int* __restrict__ cellOf,
         int* __restrict__ slotOf, int* __restrict__ cellCount, int* __restrict__ flags)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tot) return;
    int ci[3], l[3];
    az_global_cell(P, pos[i], ci);
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int v = ci[a] - P.g.goff[a];
        if (P.g.dec[a]) { if (v < 0) v += P.g.n[a]; else if (v >= P.g.nl[a]) v -= P.g.n[a]; }
        ok = ok && v >= 0 && v < P.g.nl[a];
        l[a] = v;
    }
    if (!ok) { atomicOr(flags, AZ_FLAG_HALO); cellOf[i] = 0; slotOf[i] = atomicAdd(&cellCount[0], 1); return; }
    const int cid = az_cell_id(P.g, (l[0] * P.g.nl[1] + l[1]) * P.g.nl[2] + l[2]);
    cellOf[i] = cid;
    slotOf[i] = atomicAdd(&cellCount[cid], 1);
}

// send lists and ghost slots follow the atoms through the sort
static __global__ void __launch_bounds__(AZ_TPB)
az_k_remap(int n, int* __restrict__ list, const int* __restrict__ newIndex)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) list[t] = newIndex[list[t]];
}
static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_slots(int n_ghost, int base, const int* __restrict__ newIndex, int* __restrict__ slot)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_ghost) slot[t] = newIndex[base + t];
}
