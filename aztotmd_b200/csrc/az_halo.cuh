// az_halo.cuh -- device side of the spatial domain decomposition (new functionality, SURVEY.md 8e;
// the reference is single-GPU).  The periodic box is cut into px*py*pz bricks aligned to the cell
// grid of split_cells; a rank owns the atoms whose cell (count_cell's formula, bit-exact) lies in its
// brick and keeps read-only ghost copies of the atoms in the sten[a] cell layers around it.  Every
// message goes straight to one of the <= 26 neighbouring bricks (no forwarding).
//
// Nothing here needs the host: every count lives in device memory (AzDDCounts) and travels in the
// header entry of its message, the messages have a fixed capacity per direction (agreed between the
// ranks at create time), and kernels are launched over capacities.  A step is therefore a fixed
// sequence of launches and NCCL calls on one stream -- no host synchronisation, CUDA-graph friendly.
//
//   rebuild step : wrap -> classify leavers, pack, exchange, append arrivals -> select ghosts of the
//                  new owned set, pack, exchange, append ghosts -> bin all live atoms -> scan ->
//                  sort (owned cells are numbered first, so owned atoms are a contiguous prefix) ->
//                  remap the send lists and the ghost slots through the sort permutation
//   other steps  : pack current positions of the listed atoms -> exchange -> overwrite ghost slots
//
// Message layout for direction d: entries [off[d], off[d+1]), entry off[d] is the header (count in
// the bits of .x), entries off[d]+1+k carry atom k.  Lists (send indices, ghost slots) use the same
// offsets so that message entry t and list entry t describe the same atom.
#pragma once
#include "az_kernels.cuh"

#define AZ_NDIR 27
#define AZ_DIR_SELF 13

struct AzDDCounts { // device resident
    int n_owned;      // owned atoms: contiguous prefix of the sorted streams (must stay the first member)
    int n_tot;        // owned + ghost atoms
    int n_arrived;    // migrants received in the current rebuild
    int ghost_base;   // index of the first ghost before the sort (= old owned + arrived)
    int migCount[AZ_NDIR];
    int sendCount[AZ_NDIR]; // ghosts sent per direction, fixed until the next rebuild
    int recvCount[AZ_NDIR]; // ghosts received per direction
};

struct AzDD { // by-value kernel parameter, fixed at create time
    int gs[AZ_NDIR + 1]; // ghost messages I send: entry offsets (capacity[d] + 1 entries each)
    int gr[AZ_NDIR + 1]; // ghost messages I receive from direction d (= the peer's send layout for 26-d)
    int ms[AZ_NDIR + 1]; // migration messages, send side
    int mr[AZ_NDIR + 1]; // migration messages, receive side
    int cap_owned, cap_total;
    AzDDCounts* cnt;
};

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

// entry t of a message layout -> direction (k = t - off[d]; k == 0 is the header)
__device__ __forceinline__ int az_dir_of(const int* off, int t)
{
    int d = 0;
#pragma unroll 1
    while (d < AZ_NDIR - 1 && t >= off[d + 1]) d++;
    return d;
}
__device__ __forceinline__ int az_hdr_count(float x, int capacity) { return max(0, min(__float_as_int(x), capacity)); }

// cell coordinate relative to the first owned cell of the brick, folded into (-n/2, n/2]-ish around it
__device__ __forceinline__ int az_rel_cell(const AzGrid& g, int a, int c)
{
    const int n = g.n[a], clo = g.goff[a] + g.olo[a], owned = g.ohi[a] - g.olo[a];
    int rel = c - clo;
    if (rel < 0) rel += n; // 0 .. n-1 from the brick's first cell
    if (rel >= owned && n - rel <= rel - owned + 1) rel -= n; // nearer to the low face: negative
    return rel;
}

// Which neighbouring brick (if any) owns the atom after this step's move.
static __global__ void __launch_bounds__(AZ_TPB)
az_k_mig_classify(const __grid_constant__ AzParams P, const __grid_constant__ AzDD D, const float4* __restrict__ pos,
                  int* __restrict__ migList, int* __restrict__ stay, int* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.cnt->n_owned) return;
    int ci[3];
    az_global_cell(P, pos[i], ci);
    int d[3] = { 0, 0, 0 };
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (!P.g.dec[a]) continue;
        const int owned = P.g.ohi[a] - P.g.olo[a];
        const int rel = az_rel_cell(P.g, a, ci[a]);
        d[a] = rel < 0 ? -1 : (rel >= owned ? 1 : 0);
    }
    const int dir = (d[0] + 1) * 9 + (d[1] + 1) * 3 + (d[2] + 1);
    if (dir == AZ_DIR_SELF) { stay[i] = 1; return; }
    const int slot = atomicAdd(&D.cnt->migCount[dir], 1);
    const int capacity = D.ms[dir + 1] - D.ms[dir] - 1;
    if (slot < capacity) { stay[i] = 0; migList[D.ms[dir] + 1 + slot] = i; }
    else { stay[i] = 0; atomicOr(flags, AZ_FLAG_HALO); } // the atom is lost: reported through the sticky flag
}

static __global__ void __launch_bounds__(AZ_TPB)
az_k_mig_pack(const __grid_constant__ AzDD D, const int* __restrict__ migList, const float4* __restrict__ pos,
              const float4* __restrict__ vel, const int* __restrict__ id, const float2* __restrict__ therm, AzMigRec* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= D.ms[AZ_NDIR]) return;
    const int d = az_dir_of(D.ms, t), k = t - D.ms[d];
    const int capacity = D.ms[d + 1] - D.ms[d] - 1;
    const int cnt = min(D.cnt->migCount[d], capacity);
    if (k == 0) { AzMigRec h; memset(&h, 0, sizeof(h)); h.pos.x = __int_as_float(cnt); out[t] = h; return; }
    if (k - 1 >= cnt) return;
    const int i = migList[t];
    AzMigRec r;
    r.pos = pos[i];
    r.vel = vel[i]; r.vel.w = __int_as_float(id[i]);
    r.therm = therm ? therm[i] : make_float2(0.f, 0.f);
    r.pad = make_float2(0.f, 0.f);
    out[t] = r;
}

// arrivals are appended behind the owned atoms in message order
// Messages of a rebuild that arrive over peer memory (az_k_peer_send): the consumer waits for the senders' flags first and reads
// the buffers from L2 (a neighbour GPU wrote them; lines of the previous rebuild may sit in L1)
__device__ __forceinline__ void az_wait_messages(const int* waitFlags, int epoch, int dirs, int* err_flags)
{
    if (waitFlags) {
        if (threadIdx.x < AZ_NDIR && ((dirs >> threadIdx.x) & 1)) az_wait_flag(waitFlags + threadIdx.x, epoch, err_flags);
        __syncthreads();
    }
}
#ifdef AZ_EMU
__device__ __forceinline__ float4 az_ld_l2(const float4* p) { __sync_synchronize(); return *p; }
__device__ __forceinline__ float2 az_ld_l2(const float2* p) { __sync_synchronize(); return *p; }
#else
__device__ __forceinline__ float4 az_ld_l2(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ float2 az_ld_l2(const float2* p) { return __ldcg(p); }
#endif

static __global__ void __launch_bounds__(AZ_TPB)
az_k_mig_unpack(const __grid_constant__ AzDD D, const AzMigRec* __restrict__ in, float4* __restrict__ pos, float4* __restrict__ vel,
                int* __restrict__ id, float2* __restrict__ therm, int* __restrict__ stay, int* __restrict__ flags,
                const int* __restrict__ waitFlags, int epoch, int dirs)
{
    az_wait_messages(waitFlags, epoch, dirs, flags);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= D.mr[AZ_NDIR]) return;
    const int d = az_dir_of(D.mr, t), k = t - D.mr[d];
    int before = 0, total = 0;
    for (int q = 0; q < AZ_NDIR; q++) {
        const int capq = D.mr[q + 1] - D.mr[q] - 1;
        if (capq <= 0) continue;
        const int c = az_hdr_count(az_ld_l2(&in[D.mr[q]].pos).x, capq);
        if (q < d) before += c;
        total += c;
    }
    if (t == 0) D.cnt->n_arrived = total;
    if (k == 0) return;
    const int capacity = D.mr[d + 1] - D.mr[d] - 1;
    if (k - 1 >= az_hdr_count(az_ld_l2(&in[D.mr[d]].pos).x, capacity)) return;
    const int dst = D.cnt->n_owned + before + (k - 1);
    if (dst >= D.cap_owned) { atomicOr(flags, AZ_FLAG_HALO); return; }
    AzMigRec r;
    r.pos = az_ld_l2(&in[t].pos); r.vel = az_ld_l2(&in[t].vel); r.therm = az_ld_l2(&in[t].therm);
    pos[dst] = r.pos;
    id[dst] = __float_as_int(r.vel.w);
    vel[dst] = make_float4(r.vel.x, r.vel.y, r.vel.z, 0.f);
    if (therm) therm[dst] = r.therm;
    stay[dst] = 1;
}

// Owned atoms inside the sten[a] outermost cell layers go to the bricks that need them as ghosts.
static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_select(const __grid_constant__ AzParams P, const __grid_constant__ AzDD D, const float4* __restrict__ pos,
                  const int* __restrict__ stay, int* __restrict__ sendList, int* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < min(D.cnt->n_owned + D.cnt->n_arrived, D.cap_owned) && stay[i];
    bool lo[3] = { false, false, false }, hi[3] = { false, false, false };
    if (live) {
        int ci[3];
        az_global_cell(P, pos[i], ci);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int owned = P.g.ohi[a] - P.g.olo[a];
            const int rel = az_rel_cell(P.g, a, ci[a]);
            lo[a] = P.g.dec[a] && rel < P.g.sten[a];
            hi[a] = P.g.dec[a] && rel >= owned - P.g.sten[a];
        }
    }
    const bool any = live && (lo[0] | hi[0] | lo[1] | hi[1] | lo[2] | hi[2]);
    if (!__any_sync(0xffffffffu, any)) return; // (whole warps of interior atoms leave here)
    const int lane = threadIdx.x & 31;
    for (int dir = 0; dir < AZ_NDIR; dir++) {
        const int capacity = D.gs[dir + 1] - D.gs[dir] - 1;
        if (dir == AZ_DIR_SELF || capacity <= 0) continue;
        const int d[3] = { dir / 9 - 1, (dir / 3) % 3 - 1, dir % 3 - 1 };
        bool ok = any;
#pragma unroll
        for (int a = 0; a < 3; a++) ok = ok && (d[a] == 0 || (d[a] < 0 ? lo[a] : hi[a]));
        // one atomic per warp and direction: the atoms of a warp (neighbours in the sorted stream) get CONSECUTIVE entries, so
        // the puts of the following steps store runs of consecutive ghost slots on the neighbour (its sort keeps the order of
        // mine) -- whole 128-byte lines over NVLink instead of one 16-byte packet per atom
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (!m) continue;
        int base = 0;
        if (lane == __ffs((int)m) - 1) base = atomicAdd(&D.cnt->sendCount[dir], __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
        if (ok) {
            const int slot = base + __popc(m & ((1u << lane) - 1u));
            if (slot < capacity) sendList[D.gs[dir] + 1 + slot] = i; else atomicOr(flags, AZ_FLAG_HALO);
        }
    }
}

// message = float4 position (w = type) per listed atom, plus (id, radius) in the aux stream.
// FULL (rebuild) also writes the headers.
template <bool FULL>
__global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_pack(const __grid_constant__ AzDD D, const int* __restrict__ sendList, const float4* __restrict__ pos, const int* __restrict__ id,
                const float2* __restrict__ therm, float4* __restrict__ out_pos, float2* __restrict__ out_aux)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= D.gs[AZ_NDIR]) return;
    const int d = az_dir_of(D.gs, t), k = t - D.gs[d];
    const int capacity = D.gs[d + 1] - D.gs[d] - 1;
    const int cnt = min(D.cnt->sendCount[d], capacity);
    if (k == 0) { if (FULL) out_pos[t] = make_float4(__int_as_float(cnt), 0.f, 0.f, 0.f); return; }
    if (k - 1 >= cnt) return;
    const int i = sendList[t];
    out_pos[t] = pos[i];
    if (FULL) out_aux[t] = make_float2(__int_as_float(id[i]), therm ? therm[i].y : 0.f);
    else if (therm) out_aux[t] = make_float2(0.f, therm[i].y);
}

// rebuild: ghosts are appended behind owned + arrived atoms in message order
static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_append(const __grid_constant__ AzDD D, const float4* __restrict__ in_pos, const float2* __restrict__ in_aux, float4* __restrict__ pos,
                  float4* __restrict__ vel, int* __restrict__ id, float2* __restrict__ therm, int* __restrict__ stay, int* __restrict__ flags,
                  const int* __restrict__ waitFlags, int epoch, int dirs)
{
    az_wait_messages(waitFlags, epoch, dirs, flags);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= D.gr[AZ_NDIR]) return;
    const int d = az_dir_of(D.gr, t), k = t - D.gr[d];
    int before = 0, total = 0;
    for (int q = 0; q < AZ_NDIR; q++) {
        const int capq = D.gr[q + 1] - D.gr[q] - 1;
        if (capq <= 0) continue;
        const int c = az_hdr_count(az_ld_l2(in_pos + D.gr[q]).x, capq);
        if (q < d) before += c;
        total += c;
    }
    const int base = min(D.cnt->n_owned + D.cnt->n_arrived, D.cap_owned);
    const int capacity = D.gr[d + 1] - D.gr[d] - 1;
    const int cnt = capacity > 0 ? az_hdr_count(az_ld_l2(in_pos + D.gr[d]).x, capacity) : 0;
    if (t == 0) { D.cnt->ghost_base = base; D.cnt->n_tot = min(base + total, D.cap_total); }
    if (k == 0) { D.cnt->recvCount[d] = cnt; return; }
    if (k - 1 >= cnt) return;
    const int dst = base + before + (k - 1);
    if (dst >= D.cap_total) { atomicOr(flags, AZ_FLAG_HALO); return; }
    pos[dst] = az_ld_l2(in_pos + t);
    vel[dst] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float2 aux = az_ld_l2(in_aux + t);
    id[dst] = __float_as_int(aux.x);
    if (therm) therm[dst] = make_float2(0.f, aux.y);
    stay[dst] = 1;
}

// other steps: overwrite the slots the ghosts were sorted into
static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_update(const __grid_constant__ AzDD D, const int* __restrict__ ghostSlot, const float4* __restrict__ in_pos,
                  const float2* __restrict__ in_aux, float4* __restrict__ pos, float2* __restrict__ therm)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= D.gr[AZ_NDIR]) return;
    const int d = az_dir_of(D.gr, t), k = t - D.gr[d];
    if (k == 0 || k - 1 >= D.cnt->recvCount[d]) return;
    const int s = ghostSlot[t];
    pos[s] = in_pos[t];
    if (therm) therm[s].y = in_aux[t].y;
}

// ---------------------------------------------------------------------------------------
// Ghost refresh over PEER MEMORY (NVLink 5 / NVSwitch): the steps between rebuilds without a NCCL call.  At create time every
// rank maps the receive staging and the arrival flags of its neighbours (AzComm::share -> CUDA IPC).  az_k_ghost_put is
// the pack kernel with the stores pointed at the neighbour: entry k of my message for direction d lands in entry
// gr[26-d] + k of the neighbour's staging buffer, exactly where ncclRecv would have put it; when the last block of the
// grid has finished, it raises the neighbour's arrival flag for that direction to `seq` (a counter all ranks advance together).
// az_k_ghost_update_wait is az_k_ghost_update behind a wait for the flags of the directions in use.  Two staging
// buffers alternate (seq & 1): a neighbour that is one step ahead writes the other one.  [Why that is enough: a rank posts
// put(n) only after its force pass n-1, which consumed the neighbour's put(n-1), which the neighbour posted after ITS update(n-2)
// -- so nobody still reads buffer (n & 1) when put(n) arrives.]
struct AzPeer {
    float4* pos[AZ_NDIR];  // neighbour's staging (both halves, 2 * gr[AZ_NDIR] entries), NULL = direction unused
    float2* aux[AZ_NDIR];
    int* flags[AZ_NDIR];   // neighbour's arrival flags [2 * AZ_NDIR]: ghost positions of a step, then slot tables of a rebuild
    // direct halo: the neighbour's three position streams (they rotate like mine: the same buffer number is "current" on
    // every rank at every step) and its table of where my atoms sit in them
    float4* pbuf[AZ_NDIR][3];
    int* slot[AZ_NDIR];
};

__device__ __forceinline__ int az_ld_volatile(const int* p) { return *(const volatile int*)p; }

template <bool AUX>
__global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_put(const __grid_constant__ AzDD D, const __grid_constant__ AzPeer R, const int* __restrict__ sendList, const float4* __restrict__ pos,
               const float2* __restrict__ therm, int seq, int* __restrict__ done, int dbg = 0)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < D.gs[AZ_NDIR]) {
        const int d = az_dir_of(D.gs, t), k = t - D.gs[d];
        const int capacity = D.gs[d + 1] - D.gs[d] - 1;
        if (k >= 1 && k - 1 < min(D.cnt->sendCount[d], capacity) && R.pos[d]) {
            const int i = sendList[t];
            const size_t dst = (size_t)(seq & 1) * D.gr[AZ_NDIR] + D.gr[AZ_NDIR - 1 - d] + k; // the neighbour's entry for "from direction 26-d"
            R.pos[d][dst] = pos[i];
            if (AUX) R.aux[d][dst] = make_float2(0.f, therm[i].y);
        }
    }
    // completion: the last block to finish raises the flags.  One system-scope fence per BLOCK, by thread 0 behind the block
    // barrier (fences are cumulative: the stores of the other threads, ordered before the barrier, are covered -- the pattern
    // of a cooperative grid barrier); a fence in every thread made this kernel take 37 us for 2 MB of stores.
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!(dbg & 1)) __threadfence_system(); else __threadfence(); // dbg (diagnostics only): device-scope fence per block
        const int ticket = atomicAdd(done, 1);
        if (ticket == (int)gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            for (int d = 0; d < AZ_NDIR; d++)
                if (R.flags[d]) *(volatile int*)(R.flags[d] + (AZ_NDIR - 1 - d)) = seq;
            __threadfence_system();
        }
    }
}

// ---------------------------------------------------------------------------------------
// DIRECT halo (peer memory + Verlet list, positions only): my boundary atoms are stored straight into the ghost slots of the
// neighbour's sorted position stream -- no staging, no scatter kernel, no ghost-aware loads in the force pass.  The list pass
// is the plain kernel in ONE launch; the blocks that list ghosts (blockFlag, ordered last by az_k_block_order) wait in
// their prologue for the arrival flags of this step.
// * Where: at a rebuild every receiver knows the slot each received atom was sorted into (ghostSlot); az_k_dd_remap stores
//   that number into the SENDER's remoteSlot table over NVLink and raises the sender's slot flag of that direction to the
//   rebuild epoch.  The first put after a rebuild waits for those flags.
// * Which buffer: the position streams rotate (sort: pos[c] -> pos[1-c]; the riding integrator: pos[c] <-> pos_alt) in the
//   same way on every rank, so "the buffer that is current at step n" has the same number everywhere.  put(n) reads my
//   current buffer and writes the neighbour's buffer of the same number.  A neighbour that is still at step n-1 has that
//   buffer as the TARGET of its riding integrator, which writes owned slots only; it cannot be at n-2 or n+1 (it needs my
//   put(n-1) to finish n-1, and I need its put(n) before I finish n).
// * L1: peer stores land in L2.  A block reads ghost slots only behind its wait, but a 128-byte line that holds the last
//   owned atoms AND the first ghosts could have been pulled into L1 by a block that does not wait; the list build therefore
//   flags every block that touches that line (az_k_build_list: index >= n_owned & ~7).
template <int DUMMY>
__global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_put_direct(const __grid_constant__ AzDD D, const __grid_constant__ AzPeer R, const int* __restrict__ sendList,
                      const int* __restrict__ remoteSlot, const float4* __restrict__ pos, int buf, int seq, int epoch,
                      const int* __restrict__ myFlags, int* __restrict__ done, int* __restrict__ err_flags)
{
    if (threadIdx.x < AZ_NDIR && R.slot[threadIdx.x]) az_wait_flag(myFlags + AZ_NDIR + threadIdx.x, epoch, err_flags);
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < D.gs[AZ_NDIR]) {
        const int d = az_dir_of(D.gs, t), k = t - D.gs[d];
        const int capacity = D.gs[d + 1] - D.gs[d] - 1;
        if (k >= 1 && k - 1 < min(D.cnt->sendCount[d], capacity) && R.pbuf[d][buf]) {
            const int s = ((const volatile int*)remoteSlot)[t]; // written by the neighbour (L2)
            if (s >= 0 && s < D.cap_total) R.pbuf[d][buf][s] = pos[sendList[t]];
            else atomicOr(err_flags, AZ_FLAG_HALO);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const int ticket = atomicAdd(done, 1);
        if (ticket == (int)gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            for (int d = 0; d < AZ_NDIR; d++)
                if (R.flags[d]) *(volatile int*)(R.flags[d] + (AZ_NDIR - 1 - d)) = seq;
            __threadfence_system();
        }
    }
}

// A rebuild's messages over peer memory instead of ncclSend / ncclRecv: entry k of my section for direction d goes to entry k of
// the neighbour's section 26-d (header + the live entries only: the counts live on the device), for up to two parallel streams
// (positions + aux); the last block raises the neighbour's flags [flagBase + 26-d] to the rebuild epoch.  At 8 GPUs the two
// exchanges of a rebuild were 52 + 104 NCCL point-to-point operations, ~330 us per rebuild.
template <class T> struct AzPeerArr { T* p[AZ_NDIR]; };
struct AzOffsets { int o[AZ_NDIR + 1]; };
template <class T, class U>
__global__ void __launch_bounds__(AZ_TPB)
az_k_peer_send(const __grid_constant__ AzOffsets off, const __grid_constant__ AzOffsets offR /* receive layout: the same table on every rank */, const int* __restrict__ counts,
               const T* __restrict__ sendA, const __grid_constant__ AzPeerArr<T> remoteA, const U* __restrict__ sendB, const __grid_constant__ AzPeerArr<U> remoteB,
               const __grid_constant__ AzPeerArr<int> remoteFlags, int flagBase, int epoch, int* __restrict__ done)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < off.o[AZ_NDIR]) {
        const int d = az_dir_of(off.o, t), k = t - off.o[d];
        const int capacity = off.o[d + 1] - off.o[d] - 1;
        if (capacity > 0 && remoteA.p[d] && k <= min(counts[d], capacity)) {
            const int dst = offR.o[AZ_NDIR - 1 - d] + k;
            remoteA.p[d][dst] = sendA[t];
            if (sendB && k >= 1) remoteB.p[d][dst] = sendB[t];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const int ticket = atomicAdd(done, 1);
        if (ticket == (int)gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            for (int d = 0; d < AZ_NDIR; d++)
                if (remoteFlags.p[d]) *(volatile int*)(remoteFlags.p[d] + flagBase + (AZ_NDIR - 1 - d)) = epoch;
            __threadfence_system();
        }
    }
}

#define AZ_ORDER_GHOSTS 0x40000000
// blocks of the list pass in launch order: the ones that list no ghost first (they can start at once), the others behind them
static __global__ void __launch_bounds__(1024)
az_k_block_order(const int* __restrict__ flag, int n, int* __restrict__ order)
{
    __shared__ int s_w[32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int b0 = min(n, tid * per), b1 = min(n, b0 + per);
    int nu = 0;
    for (int b = b0; b < b1; b++) nu += flag[b] ? 0 : 1;
    int incl = nu;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    if (w == 0) {
        int v = s_w[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
        s_w[lane] = iv - v;
        if (lane == 31) s_total = iv;
    }
    __syncthreads();
    int ku = s_w[w] + incl - nu;          // unflagged blocks before mine
    int kf = s_total + (b0 - ku);         // flagged blocks before mine, behind all unflagged ones
    for (int b = b0; b < b1; b++) {
        if (flag[b]) order[kf++] = b | AZ_ORDER_GHOSTS; else order[ku++] = b; // (the flag travels with the entry: one load per block)
    }
}

#ifdef AZ_EMU
__device__ __forceinline__ float4 az_ld_cg4(const float4* p) { __sync_synchronize(); return *p; }
__device__ __forceinline__ float2 az_ld_cg2(const float2* p) { __sync_synchronize(); return *p; }
#else
__device__ __forceinline__ float4 az_ld_cg4(const float4* p) { return __ldcg(p); } // L2 only: the line may sit stale in L1 from two steps ago
__device__ __forceinline__ float2 az_ld_cg2(const float2* p) { return __ldcg(p); }
#endif

static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_update_wait(const __grid_constant__ AzDD D, const int* __restrict__ ghostSlot, const float4* __restrict__ stage_pos,
                       const float2* __restrict__ stage_aux, float4* __restrict__ pos, float2* __restrict__ therm, const int* __restrict__ flags,
                       int seq, int* __restrict__ err_flags)
{
    // every block waits for the messages of all directions in use (a direction is in use iff its capacity is not empty)
    if (threadIdx.x < AZ_NDIR) {
        const int d = threadIdx.x;
        if (D.gr[d + 1] - D.gr[d] - 1 > 0) az_wait_flag(flags + d, seq, err_flags);
    }
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= D.gr[AZ_NDIR]) return;
    const int d = az_dir_of(D.gr, t), k = t - D.gr[d];
    if (k == 0 || k - 1 >= D.cnt->recvCount[d]) return;
    const int s = ghostSlot[t];
    const size_t src = (size_t)(seq & 1) * D.gr[AZ_NDIR] + t;
    pos[s] = az_ld_cg4(stage_pos + src);
    if (therm) therm[s].y = az_ld_cg2(stage_aux + src).y;
}

// fused halo: sorted ghost index - n_owned -> entry of the receive staging (the inverse of ghostSlot), rebuilt with the list
static __global__ void __launch_bounds__(AZ_TPB)
az_k_ghost_entry(const __grid_constant__ AzDD D, const int* __restrict__ ghostSlot, int* __restrict__ ghostEntry)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= D.gr[AZ_NDIR]) return;
    const int d = az_dir_of(D.gr, t), k = t - D.gr[d];
    if (k == 0 || k - 1 >= D.cnt->recvCount[d]) return;
    const int s = ghostSlot[t] - D.cnt->n_owned;
    if (s >= 0 && s < D.gr[AZ_NDIR]) ghostEntry[s] = t;
}

// count_cell for the owned and ghost atoms of a brick: global cell from the (wrapped) position, then the
// periodic image that falls inside the local grid.  Leavers (stay == 0) are dead: cell -1, skipped by the sort.
static __global__ void __launch_bounds__(AZ_TPB)
az_k_bin(const __grid_constant__ AzParams P, const __grid_constant__ AzDD D, const float4* __restrict__ pos, const int* __restrict__ stay,
         int* __restrict__ cellOf, int* __restrict__ slotOf, int* __restrict__ cellCount, int* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.cnt->n_tot) return;
    if (!stay[i]) { cellOf[i] = -1; return; }
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
    if (!ok) { atomicOr(flags, AZ_FLAG_HALO); cellOf[i] = -1; return; } // moved farther than the ghost shell
    const int cid = az_cell_id(P.g, (l[0] * P.g.nl[1] + l[1]) * P.g.nl[2] + l[2]);
    cellOf[i] = cid;
    slotOf[i] = atomicAdd(&cellCount[cid], 1);
}


// send lists and ghost slots follow the atoms through the sort; then publish the new counts
// Direct halo (R != NULL): the slot of every ghost also goes into the remoteSlot table of the rank that sent it (entry k of my
// direction d is entry k of its direction 26-d), and the last block raises that rank's slot flags to `epoch`.
static __global__ void __launch_bounds__(AZ_TPB)
az_k_dd_remap(const __grid_constant__ AzDD D, const int* __restrict__ newIndex, int* __restrict__ sendList, int* __restrict__ ghostSlot,
              const __grid_constant__ AzPeer R, int direct, int epoch, int* __restrict__ done)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < D.gs[AZ_NDIR]) {
        const int d = az_dir_of(D.gs, t), k = t - D.gs[d];
        const int capacity = D.gs[d + 1] - D.gs[d] - 1;
        if (k >= 1 && k - 1 < min(D.cnt->sendCount[d], capacity)) sendList[t] = newIndex[sendList[t]];
    }
    if (t < D.gr[AZ_NDIR]) {
        const int d = az_dir_of(D.gr, t), k = t - D.gr[d];
        if (k >= 1 && k - 1 < D.cnt->recvCount[d]) {
            int before = 0;
            for (int q = 0; q < d; q++) before += D.cnt->recvCount[q];
            const int slot = newIndex[D.cnt->ghost_base + before + (k - 1)];
            ghostSlot[t] = slot;
            if (direct && R.slot[d]) R.slot[d][D.gs[AZ_NDIR - 1 - d] + k] = slot;
        }
    }
    if (direct) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const int ticket = atomicAdd(done, 1);
            if (ticket == (int)gridDim.x - 1) {
                *done = 0;
                __threadfence_system();
                for (int d = 0; d < AZ_NDIR; d++)
                    if (R.flags[d]) *(volatile int*)(R.flags[d] + AZ_NDIR + (AZ_NDIR - 1 - d)) = epoch;
                __threadfence_system();
            }
        }
    }
}

static __global__ void az_k_dd_publish(const __grid_constant__ AzDD D, const int* __restrict__ first, int n_owned_cells, int n_cells)
{
    D.cnt->n_owned = first[n_owned_cells];
    D.cnt->n_tot = first[n_cells];
    D.cnt->n_arrived = 0;
}

static __global__ void az_k_dd_begin_rebuild(const __grid_constant__ AzDD D)
{
    const int t = threadIdx.x;
    if (t < AZ_NDIR) { D.cnt->migCount[t] = 0; D.cnt->sendCount[t] = 0; D.cnt->recvCount[t] = 0; }
    if (t == 0) D.cnt->n_arrived = 0;
}
