// az_engine_dd.cuh -- host side of the spatial domain decomposition (included by az_engine.cu).
// New functionality: the reference runs on device 0 only (src/cuInit.cu:693-753).
//
// Bricks: the cell grid of split_cells is cut into pg[0] x pg[1] x pg[2] bricks; rank r sits at
// (r / (pg1*pg2), (r / pg2) % pg1, r % pg2) and owns the cells [n*p/P, n*(p+1)/P) per axis.  An axis
// with one brick is not decomposed: the pair kernels wrap it locally as on one GPU.  The local grid
// of a rank is its owned block plus sten[a] ghost layers on each side of a decomposed axis.
#pragma once
#include "az_halo.cuh"
#include "az_comm.h"

struct AzDDHost {
    AzComm* comm = nullptr;
    int pg[3] = { 1, 1, 1 }, pc[3] = { 0, 0, 0 };
    int peer[AZ_NDIR];
    AzDD D;
    int n_owned_cells = 0;
    int* cellmap = nullptr;
    int *stay = nullptr, *newIndex = nullptr, *sendList = nullptr, *ghostSlot = nullptr, *migList = nullptr;
    float4 *gSendPos = nullptr, *gRecvPos = nullptr;
    float2 *gSendAux = nullptr, *gRecvAux = nullptr;
    AzMigRec *mSend = nullptr, *mRecv = nullptr;
    AzDDCounts* cnt = nullptr;
    double* red = nullptr; // 128 doubles of all-reduce scratch
    bool replicated = true;
    std::vector<int> keep; // indices into the create-time arrays of the atoms this rank took
    // halo / compute overlap on the steps between rebuilds: the ghost refresh runs on its own stream while the list
    // blocks that see no ghost are already computing
    cudaStream_t halo_stream = nullptr;
    cudaEvent_t ev_moved = nullptr, ev_ghosts = nullptr;
    int* blockFlag = nullptr;
    bool overlap = true;
    // ghost refresh over peer memory (AzComm::share): no NCCL call on the steps between rebuilds
    bool p2p = false;
    AzPeer peerPtr;                       // the neighbours' staging buffers and arrival flags, per direction
    float4* stagePos = nullptr;           // my receive staging: two halves of gr[AZ_NDIR] entries (alternating with halo_seq & 1)
    float2* stageAux = nullptr;
    int *haloFlags = nullptr, *putDone = nullptr;
    int halo_seq = 0;                     // advanced by every refresh; the same on all ranks
    // fused halo (peer memory + Verlet list): ONE list-pass launch per step; ghost-touching blocks come last in blockOrder, wait
    // for the arrival flags and read the ghosts from the staging; nothing scatters them into pos[] unless somebody asks
    bool fused = false;
    int* ghostEntry = nullptr;
    cudaEvent_t ev_put = nullptr;
    // direct halo (az_halo.cuh): boundary atoms are stored straight into the neighbours' sorted position streams
    bool direct = false;
    int* remoteSlot = nullptr;            // [gs total] where my k-th atom for direction d sits in that neighbour's streams
    int* blockOrder = nullptr;            // launch order of the list blocks: ghost-free ones first
    int slot_epoch = 0;                   // rebuilds so far; the same on all ranks
    int prev_step_buf = -1;
    // a rebuild's messages (migration, ghost selection) over peer memory too: no NCCL call on the whole stepping path
    bool peer_rebuild = false;
    AzPeerArr<AzMigRec> peerMig;
    AzPeerArr<float4> peerGPos;
    AzPeerArr<float2> peerGAux;
    AzPeerArr<int> peerFlags;               // number of the position stream the previous step's force pass read
    float4* pos_bufs[3] = { nullptr, nullptr, nullptr }; // my three position streams by number
    bool pos_ghosts_stale = false;        // pos[] still holds the ghost positions of the last scatter (rebuild or dd_sync_ghosts)
    int halo_dirs = 0;
    int n_create = 0;      // length of the create-time arrays (az_engine_upload takes arrays of the same shape)
    int *keepSrc = nullptr, *keepGid = nullptr, *keepType = nullptr; // device copies for az_engine_upload
};

static int dd_read_counts(az_engine* e, AzDDCounts* out);

extern "C" int az_comm_unique_id(void* id128)
{
    if (!id128) return AZ_ERR_INVALID;
    std::string err;
    int rc = az_comm_make_id(id128, &err);
    if (rc) g_create_error = err;
    return rc;
}

static void dd_free(az_engine* e)
{
    AzDDHost* h = e->dd;
    if (!h) return;
    cudaFree(h->cellmap); cudaFree(h->stay); cudaFree(h->newIndex); cudaFree(h->sendList); cudaFree(h->ghostSlot); cudaFree(h->migList);
    cudaFree(h->gSendPos); cudaFree(h->gRecvPos); cudaFree(h->gSendAux); cudaFree(h->gRecvAux); cudaFree(h->mSend); cudaFree(h->mRecv);
    cudaFree(h->blockFlag); cudaFree(h->remoteSlot); cudaFree(h->blockOrder);
    cudaFree(h->stagePos); cudaFree(h->stageAux); cudaFree(h->haloFlags); cudaFree(h->putDone);
    cudaFree(h->ghostEntry);
    if (h->ev_put) cudaEventDestroy(h->ev_put);
    if (h->ev_moved) cudaEventDestroy(h->ev_moved);
    if (h->ev_ghosts) cudaEventDestroy(h->ev_ghosts);
    if (h->halo_stream) cudaStreamDestroy(h->halo_stream);
    cudaFree(h->cnt); cudaFree(h->red); cudaFree(h->keepSrc); cudaFree(h->keepGid); cudaFree(h->keepType);
    delete h->comm;
    delete h;
    e->dd = nullptr;
}

static int dd_allreduce_host(az_engine* e, double* v, int n, bool take_max)
{
    AzDDHost* h = e->dd;
    if (n > 128) return AZ_ERR_INVALID;
    AZ_CUDA(e, cudaMemcpyAsync(h->red, v, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    int rc = h->comm->allreduce(h->red, n, take_max, e->stream);
    if (rc) { e->err = h->comm->err; return rc; }
    AZ_CUDA(e, cudaMemcpyAsync(v, h->red, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    return AZ_OK;
}

// in place, device memory, asynchronous on `st`: sum over the ranks
static int dd_allreduce_dev(az_engine* e, double* dev, int n, cudaStream_t st)
{
    int rc = e->dd->comm->allreduce(dev, n, false, st);
    if (rc) e->err = e->dd->comm->err;
    return rc;
}

// the brick grid with the smallest halo surface among the factorisations of n_ranks that fit
static bool dd_pick_grid(const AzGrid& g, int nr, int pg[3])
{
    double best = -1.0;
    for (int a = 1; a <= nr; a++) {
        if (nr % a) continue;
        for (int b = 1; b <= nr / a; b++) {
            if ((nr / a) % b) continue;
            const int c = nr / a / b, p[3] = { a, b, c };
            bool ok = true;
            double w[3];
            for (int k = 0; k < 3; k++) {
                w[k] = (double)g.n[k] / p[k];
                if (p[k] > 1 && (g.n[k] / p[k] < 2 * g.sten[k] || g.n[k] / p[k] < 1)) ok = false;
            }
            if (!ok) continue;
            double surf = 0.0;
            if (a > 1) surf += w[1] * w[2];
            if (b > 1) surf += w[0] * w[2];
            if (c > 1) surf += w[0] * w[1];
            if (best < 0.0 || surf < best) { best = surf; pg[0] = a; pg[1] = b; pg[2] = c; }
        }
    }
    return best >= 0.0;
}

// put_periodic's wrap followed by count_cell (src/cuMDfunc.cu:35-106, src/cuSort.cu:114-128) on the host
static int dd_host_cell(const AzParams& P, int a, double x_in)
{
    float x = (float)x_in;
    int nx = (int)floor((double)x * (double)P.revL[a]);
    x = x - (float)nx * P.L[a];
    if (x >= P.L[a]) x = 0.f;
    int c = (int)floor((double)x * (double)P.g.crev[a]);
    if (c < 0) c = 0;
    if (c > P.g.n[a] - 1) c = P.g.n[a] - 1;
    return c;
}

static int dd_rel(const AzGrid& g, int a, int c) // host twin of az_rel_cell
{
    const int n = g.n[a], clo = g.goff[a] + g.olo[a], owned = g.ohi[a] - g.olo[a];
    int rel = c - clo;
    if (rel < 0) rel += n;
    if (rel >= owned && n - rel <= rel - owned + 1) rel -= n;
    return rel;
}

// Decide the brick layout, turn P.g into the local grid, pick this rank's atoms and agree on the
// message capacities with the other ranks.
static int dd_setup(az_engine* e, const az_system_desc* d, const az_cuda_opts* o)
{
    AzDDHost* h = e->dd = new AzDDHost();
    AzParams& P = e->P;
    AzGrid& g = P.g;
    int rc = AZ_OK;
    // Every check that can fail comes BEFORE the communicator is created: the checks depend only on
    // arguments that are the same on all ranks, so either every rank fails here or none does (a rank
    // that entered ncclCommInitRank alone would wait for the others forever).
    if (e->rank < 0 || e->rank >= e->n_ranks) { AZ_SET_ERR(e, "rank %d outside 0..%d", e->rank, e->n_ranks - 1); return AZ_ERR_INVALID; }
    if (!o->nccl_id) { AZ_SET_ERR(e, "multi-GPU create needs opts.nccl_id (128 bytes from az_comm_unique_id, the same on all ranks)"); return AZ_ERR_INVALID; }

    if (o->pgrid[0] > 0 && o->pgrid[0] * o->pgrid[1] * o->pgrid[2] == e->n_ranks) { for (int a = 0; a < 3; a++) h->pg[a] = o->pgrid[a]; }
    else if (o->pgrid[0] > 0) { AZ_SET_ERR(e, "pgrid %dx%dx%d does not match %d ranks", o->pgrid[0], o->pgrid[1], o->pgrid[2], e->n_ranks); return AZ_ERR_INVALID; }
    else if (!dd_pick_grid(g, e->n_ranks, h->pg)) { AZ_SET_ERR(e, "no brick layout of %d ranks fits %dx%dx%d cells with stencil %d,%d,%d", e->n_ranks, g.n[0], g.n[1], g.n[2], g.sten[0], g.sten[1], g.sten[2]); return AZ_ERR_UNSUPPORTED; }
    h->pc[0] = e->rank / (h->pg[1] * h->pg[2]); h->pc[1] = (e->rank / h->pg[2]) % h->pg[1]; h->pc[2] = e->rank % h->pg[2];

    for (int a = 0; a < 3; a++) {
        const int n = g.n[a], Pn = h->pg[a], p = h->pc[a];
        const int lo = (int)((long long)n * p / Pn), hi = (int)((long long)n * (p + 1) / Pn);
        g.dec[a] = Pn > 1;
        if (g.dec[a]) {
            // every brick of the axis must hold the ghost shell of its neighbours without aliasing
            for (int q = 0; q < Pn; q++) {
                const int w = (int)((long long)n * (q + 1) / Pn) - (int)((long long)n * q / Pn);
                if (w < g.sten[a] || w + 2 * g.sten[a] > n) {
                    AZ_SET_ERR(e, "axis %d: %d cells over %d bricks is too fine for a stencil of %d cells", a, n, Pn, g.sten[a]);
                    return AZ_ERR_UNSUPPORTED;
                }
            }
            g.olo[a] = g.sten[a]; g.ohi[a] = g.sten[a] + (hi - lo); g.nl[a] = (hi - lo) + 2 * g.sten[a]; g.goff[a] = lo - g.sten[a];
        } else {
            g.olo[a] = 0; g.ohi[a] = n; g.nl[a] = n; g.goff[a] = 0;
        }
    }
    long long ncl = (long long)g.nl[0] * g.nl[1] * g.nl[2];
    if (ncl > 0x3fffffff) { AZ_SET_ERR(e, "too many local cells (%lld)", ncl); return AZ_ERR_INVALID; }
    {
        std::string cerr;
        rc = az_comm_create(o->nccl_id, e->rank, e->n_ranks, e->device, &h->comm, &cerr);
        if (rc) { e->err = cerr; return rc; }
    }
    g.ncell_local = (int)ncl;
    const int on[3] = { g.ohi[0] - g.olo[0], g.ohi[1] - g.olo[1], g.ohi[2] - g.olo[2] };
    h->n_owned_cells = on[0] * on[1] * on[2];

    // storage numbering: owned cells first (z fastest inside the owned block), ghost cells behind them
    {
        std::vector<int> map((size_t)ncl);
        int next = h->n_owned_cells;
        for (int x = 0; x < g.nl[0]; x++)
            for (int y = 0; y < g.nl[1]; y++)
                for (int z = 0; z < g.nl[2]; z++) {
                    const bool own = x >= g.olo[0] && x < g.ohi[0] && y >= g.olo[1] && y < g.ohi[1] && z >= g.olo[2] && z < g.ohi[2];
                    const size_t lin = ((size_t)x * g.nl[1] + y) * g.nl[2] + z;
                    map[lin] = own ? ((x - g.olo[0]) * on[1] + (y - g.olo[1])) * on[2] + (z - g.olo[2]) : next++;
                }
        AZ_ALLOC(e, h->cellmap, map.size());
        AZ_CUDA(e, cudaMemcpy(h->cellmap, map.data(), sizeof(int) * map.size(), cudaMemcpyHostToDevice));
        g.cellmap = h->cellmap;
    }

    // neighbours: only along decomposed axes
    for (int dir = 0; dir < AZ_NDIR; dir++) {
        const int dv[3] = { dir / 9 - 1, (dir / 3) % 3 - 1, dir % 3 - 1 };
        bool ok = dir != AZ_DIR_SELF;
        int q[3];
        for (int a = 0; a < 3; a++) {
            if (dv[a] != 0 && !g.dec[a]) ok = false;
            q[a] = (h->pc[a] + dv[a] + h->pg[a]) % h->pg[a];
        }
        h->peer[dir] = ok ? (q[0] * h->pg[1] + q[1]) * h->pg[2] + q[2] : -1;
    }

    // this rank's atoms
    h->n_create = d->n_atoms;
    h->replicated = (d->global_ids == nullptr);
    const double* xyz[3] = { d->xs, d->ys, d->zs };
    std::vector<int> rel((size_t)3 * d->n_atoms);
    h->keep.clear();
    for (int i = 0; i < d->n_atoms; i++) {
        bool mine = true;
        for (int a = 0; a < 3; a++) {
            const int r = dd_rel(g, a, dd_host_cell(P, a, xyz[a][i]));
            rel[(size_t)3 * i + a] = r;
            if (g.dec[a] && (r < 0 || r >= on[a])) mine = false;
        }
        if (mine || !h->replicated) h->keep.push_back(i);
    }

    // message capacities: 1.5 x what the initial configuration needs (+ slack), the maximum over all ranks
    double need[2 * AZ_NDIR];
    for (int k = 0; k < 2 * AZ_NDIR; k++) need[k] = 0.0;
    for (int i : h->keep) {
        bool lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            const int r = rel[(size_t)3 * i + a];
            lo[a] = g.dec[a] && r < g.sten[a];
            hi[a] = g.dec[a] && r >= on[a] - g.sten[a];
        }
        if (!(lo[0] | hi[0] | lo[1] | hi[1] | lo[2] | hi[2])) continue;
        for (int dir = 0; dir < AZ_NDIR; dir++) {
            if (h->peer[dir] < 0) continue;
            const int dv[3] = { dir / 9 - 1, (dir / 3) % 3 - 1, dir % 3 - 1 };
            bool ok = true;
            for (int a = 0; a < 3; a++) ok = ok && (dv[a] == 0 || (dv[a] < 0 ? lo[a] : hi[a]));
            if (ok) need[dir] += 1.0;
        }
    }
    double glob[2 * AZ_NDIR + 2];
    for (int dir = 0; dir < AZ_NDIR; dir++) glob[dir] = need[dir];
    glob[AZ_NDIR] = (double)h->keep.size();
    AZ_ALLOC(e, h->red, 128);
    rc = dd_allreduce_host(e, glob, AZ_NDIR + 1, true);
    if (rc) return rc;
    int gcap[AZ_NDIR], mcap[AZ_NDIR];
    for (int dir = 0; dir < AZ_NDIR; dir++) {
        // a direction is in use on every rank or on none (the brick grid is the same everywhere)
        gcap[dir] = h->peer[dir] >= 0 ? (int)(1.5 * glob[dir]) + 256 : -1;
        mcap[dir] = h->peer[dir] >= 0 ? (int)(0.25 * glob[dir]) + 256 : -1;
    }
    AzDD& D = h->D;
    D.gs[0] = D.gr[0] = D.ms[0] = D.mr[0] = 0;
    for (int dir = 0; dir < AZ_NDIR; dir++) {
        const int opp = AZ_NDIR - 1 - dir;
        D.gs[dir + 1] = D.gs[dir] + (gcap[dir] + 1);
        D.gr[dir + 1] = D.gr[dir] + (gcap[opp] + 1); // what the peer in direction dir sends towards me
        D.ms[dir + 1] = D.ms[dir] + (mcap[dir] + 1);
        D.mr[dir + 1] = D.mr[dir] + (mcap[opp] + 1);
    }
    const int n_max = (int)glob[AZ_NDIR];
    D.cap_owned = (int)(1.15 * n_max) + 1024 + D.mr[AZ_NDIR];
    D.cap_total = D.cap_owned + D.gr[AZ_NDIR];
    if (!h->replicated && (int)h->keep.size() > D.cap_owned) D.cap_owned = (int)h->keep.size(), D.cap_total = D.cap_owned + D.gr[AZ_NDIR];
    AZ_ALLOC(e, h->cnt, 1);
    D.cnt = h->cnt;
    P.n_dyn = &h->cnt->n_owned;
    P.n_local = D.cap_owned;
    e->cap_total_dd = D.cap_total;
    if (!h->replicated && d->n_atoms_global <= 0) { // the global atom count is the sum of the shares
        double tot[1] = { (double)h->keep.size() };
        rc = dd_allreduce_host(e, tot, 1, false);
        if (rc) return rc;
        e->n_global = (int)(tot[0] + 0.5);
    }
    return AZ_OK;
}

static const std::vector<int>* dd_keep(az_engine* e) { return e->dd ? &e->dd->keep : nullptr; }

// az_engine_upload on a brick: the arrays have the shape of the create-time arrays of THIS rank; the
// rank takes the same atoms it took at create time (leavers migrate in the next step's rebuild).
static __global__ void __launch_bounds__(AZ_TPB)
az_k_upload_dd(int n, const int* __restrict__ src, const int* __restrict__ gid, const int* __restrict__ type, const double* __restrict__ xs,
               const double* __restrict__ ys, const double* __restrict__ zs, const double* __restrict__ vxs, const double* __restrict__ vys,
               const double* __restrict__ vzs, float4* __restrict__ pos, float4* __restrict__ vel, int* __restrict__ id, float4* __restrict__ frc)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = src[k];
    pos[k] = make_float4((float)xs[i], (float)ys[i], (float)zs[i], __int_as_float(type[k]));
    vel[k] = make_float4(vxs ? (float)vxs[i] : 0.f, vys ? (float)vys[i] : 0.f, vzs ? (float)vzs[i] : 0.f, 0.f);
    id[k] = gid[k];
    frc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
}

static __global__ void az_k_dd_set_counts(const __grid_constant__ AzDD D, int n)
{
    D.cnt->n_owned = n; D.cnt->n_tot = n; D.cnt->n_arrived = 0;
}

static int dd_store_keep(az_engine* e, const az_system_desc* d)
{
    AzDDHost* h = e->dd;
    const int n = (int)h->keep.size();
    std::vector<int> gid(n), type(n);
    for (int k = 0; k < n; k++) { const int i = h->keep[k]; gid[k] = d->global_ids ? d->global_ids[i] : i; type[k] = d->types[i]; }
    AZ_ALLOC(e, h->keepSrc, n); AZ_ALLOC(e, h->keepGid, n); AZ_ALLOC(e, h->keepType, n);
    if (n > 0) {
        AZ_CUDA(e, cudaMemcpy(h->keepSrc, h->keep.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(h->keepGid, gid.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(h->keepType, type.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    }
    return AZ_OK;
}

static int dd_upload(az_engine* e, const double* const src[6], double* dst[6], int* slot)
{
    AzDDHost* h = e->dd;
    if (e->use_therm) { AZ_SET_ERR(e, "az_engine_upload on bricks does not carry the thermostat state"); return AZ_ERR_UNSUPPORTED; }
    const int nc = h->n_create, n = (int)h->keep.size(), c = e->cur;
    int rc = stage_upload(e, (const void* const*)src, nc, sizeof(double), (void**)dst, slot);
    if (rc) return rc;
    AZ_LAUNCH(az_k_upload_dd, (n + AZ_TPB - 1) / AZ_TPB, AZ_TPB, 0, e->stream, n, h->keepSrc, h->keepGid, h->keepType, dst[0], dst[1], dst[2], dst[3], dst[4], dst[5],
              e->pos[c], e->vel[c], e->id[c], e->frc);
    AZ_LAUNCH(az_k_dd_set_counts, 1, 1, 0, e->stream, h->D, n); // stream ordered: the previous step may still be running
    e->launches += 2;
    return AZ_OK;
}

static int dd_init_counts(az_engine* e, int n_keep)
{
    AzDDCounts c;
    memset(&c, 0, sizeof(c));
    c.n_owned = n_keep; c.n_tot = n_keep;
    AZ_CUDA(e, cudaMemcpy(e->dd->cnt, &c, sizeof(c), cudaMemcpyHostToDevice));
    return AZ_OK;
}

static int dd_owned_count(az_engine* e, int* n)
{
    AzDDCounts c;
    int rc = dd_read_counts(e, &c);
    if (rc) return rc;
    *n = c.n_owned;
    return AZ_OK;
}

// az_engine_stats on a brick: the scalars of all ranks are summed (energies, momenta, pressure -- every
// one of them is linear in per-atom or per-pair terms), flags are OR-ed.  Collective.
static int dd_reduce_stats(az_engine* e, az_stats* o, int* flags)
{
    const int ns = e->P.nspec;
    double v[128];
    int k = 0;
    v[k++] = o->engKin_d; v[k++] = o->engVdW_d; v[k++] = o->engCoul1_d; v[k++] = o->engTemp_d;
    v[k++] = (double)o->engElecField; v[k++] = (double)o->pressure; v[k++] = (double)o->n_pairs;
    for (int a = 0; a < 3; a++) { v[k++] = (double)o->posMom[a]; v[k++] = (double)o->negMom[a]; }
    for (int t = 0; t < ns; t++) for (int a = 0; a < 3; a++) { v[k++] = (double)o->specAcBoxPos[t][a]; v[k++] = (double)o->specAcBoxNeg[t][a]; }
    for (int b = 0; b < 5; b++) v[k++] = ((*flags >> b) & 1) ? 1.0 : 0.0;
    int rc = dd_allreduce_host(e, v, k, false);
    if (rc) return rc;
    k = 0;
    o->engKin_d = v[k++]; o->engVdW_d = v[k++]; o->engCoul1_d = v[k++]; o->engTemp_d = v[k++];
    o->engElecField = (float)v[k++]; o->pressure = (float)v[k++]; o->n_pairs = (long long)(0.5 * v[k++] + 0.25); // visits from both sides, summed over the bricks
    for (int a = 0; a < 3; a++) { o->posMom[a] = (float)v[k++]; o->negMom[a] = (float)v[k++]; }
    for (int t = 0; t < ns; t++) for (int a = 0; a < 3; a++) { o->specAcBoxPos[t][a] = (int)(v[k++] + 0.5); o->specAcBoxNeg[t][a] = (int)(v[k++] + 0.5); }
    int f = 0;
    for (int b = 0; b < 5; b++) if (v[k++] > 0.5) f |= 1 << b;
    *flags = f;
    o->engKin = (float)o->engKin_d; o->engVdW = (float)o->engVdW_d; o->engCoul1 = (float)o->engCoul1_d; o->engTemp = (float)o->engTemp_d;
    o->engPot = o->engCoul1 + o->engCoul2 + o->engCoul3 + o->engVdW; // calc_quantities, src/main.cu:121-162 (the k-space terms are global on every rank)
    o->engTot = o->engPot + o->engKin;
    return AZ_OK;
}

static int dd_alloc(az_engine* e)
{
    AzDDHost* h = e->dd;
    const AzDD& D = h->D;
    AZ_ALLOC(e, h->stay, D.cap_total); AZ_ALLOC(e, h->newIndex, D.cap_total);
    AZ_ALLOC(e, h->sendList, D.gs[AZ_NDIR]); AZ_ALLOC(e, h->ghostSlot, D.gr[AZ_NDIR]); AZ_ALLOC(e, h->migList, D.ms[AZ_NDIR]);
    AZ_ALLOC(e, h->gSendPos, D.gs[AZ_NDIR]); AZ_ALLOC(e, h->gRecvPos, D.gr[AZ_NDIR]);
    AZ_ALLOC(e, h->gSendAux, D.gs[AZ_NDIR]); AZ_ALLOC(e, h->gRecvAux, D.gr[AZ_NDIR]);
    AZ_ALLOC(e, h->mSend, D.ms[AZ_NDIR]); AZ_ALLOC(e, h->mRecv, D.mr[AZ_NDIR]);
    AZ_ALLOC(e, h->blockFlag, e->list_nblocks > 0 ? e->list_nblocks : 1);
    { // highest priority: the halo's small kernels and NCCL's copy kernels take SM slots as soon as blocks of the
      // (much larger) list kernel retire, instead of queueing behind all of them
        int lo = 0, hi = 0;
        AZ_CUDA(e, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        AZ_CUDA(e, cudaStreamCreateWithPriority(&h->halo_stream, cudaStreamNonBlocking, hi));
    }
    AZ_CUDA(e, cudaEventCreateWithFlags(&h->ev_moved, cudaEventDisableTiming));
    AZ_CUDA(e, cudaEventCreateWithFlags(&h->ev_ghosts, cudaEventDisableTiming));
    // Peer memory for the per-step ghost refresh.  Every rank offers its receive staging (two halves) and its arrival flags;
    // if the transport cannot map them (no NVLink / IPC, or the callback transport of the tests) the NCCL send/recv path stays.
    memset(&h->peerPtr, 0, sizeof(h->peerPtr));
    const char* no_p2p = getenv("AZTOT_NO_PEER_HALO");
    if (!(no_p2p && no_p2p[0] == '1')) {
        AZ_ALLOC(e, h->stagePos, (size_t)2 * D.gr[AZ_NDIR]); AZ_ALLOC(e, h->stageAux, (size_t)2 * D.gr[AZ_NDIR]);
        AZ_ALLOC(e, h->haloFlags, 4 * AZ_NDIR); AZ_ALLOC(e, h->putDone, 4);
        AZ_ALLOC(e, h->remoteSlot, D.gs[AZ_NDIR] > 0 ? D.gs[AZ_NDIR] : 1);
        AZ_ALLOC(e, h->blockOrder, e->list_nblocks > 0 ? e->list_nblocks : 1);
        h->pos_bufs[0] = e->pos[0]; h->pos_bufs[1] = e->pos[1]; h->pos_bufs[2] = e->pos_alt ? e->pos_alt : e->pos[0];
        const int NSH = 10;
        void* local[NSH] = { h->stagePos, h->stageAux, h->haloFlags, h->remoteSlot, h->pos_bufs[0], h->pos_bufs[1], h->pos_bufs[2], h->mRecv, h->gRecvPos, h->gRecvAux };
        std::vector<void*> remote;
        memset(&h->peerMig, 0, sizeof(h->peerMig)); memset(&h->peerGPos, 0, sizeof(h->peerGPos));
        memset(&h->peerGAux, 0, sizeof(h->peerGAux)); memset(&h->peerFlags, 0, sizeof(h->peerFlags));
        const int rc = h->comm->share(local, NSH, &remote);
        if (rc == AZ_OK) {
            for (int dir = 0; dir < AZ_NDIR; dir++) {
                const int pr = h->peer[dir];
                if (pr < 0) continue;
                h->peerPtr.pos[dir] = (float4*)remote[(size_t)pr * NSH + 0];
                h->peerPtr.aux[dir] = (float2*)remote[(size_t)pr * NSH + 1];
                h->peerPtr.flags[dir] = (int*)remote[(size_t)pr * NSH + 2];
                h->peerPtr.slot[dir] = (int*)remote[(size_t)pr * NSH + 3];
                for (int b = 0; b < 3; b++) h->peerPtr.pbuf[dir][b] = (float4*)remote[(size_t)pr * NSH + 4 + b];
                h->peerMig.p[dir] = (AzMigRec*)remote[(size_t)pr * NSH + 7];
                h->peerGPos.p[dir] = (float4*)remote[(size_t)pr * NSH + 8];
                h->peerGAux.p[dir] = (float2*)remote[(size_t)pr * NSH + 9];
                h->peerFlags.p[dir] = (int*)remote[(size_t)pr * NSH + 2];
            }
            h->p2p = true;
            { const char* no_pr = getenv("AZTOT_NO_PEER_REBUILD"); h->peer_rebuild = !(no_pr && no_pr[0] == '1'); }
            for (int dir = 0; dir < AZ_NDIR; dir++) if (h->peer[dir] >= 0) h->halo_dirs |= 1 << dir;
            // the fused single-launch list pass: needs the list (per-block ghost flags) and positions only (radii travel in the aux stream)
            const char* no_fuse = getenv("AZTOT_NO_FUSED_HALO");
            if (e->use_list && !e->use_therm && !(no_fuse && no_fuse[0] == '1')) {
                AZ_ALLOC(e, h->ghostEntry, D.gr[AZ_NDIR] > 0 ? D.gr[AZ_NDIR] : 1);
                AZ_CUDA(e, cudaEventCreateWithFlags(&h->ev_put, cudaEventDisableTiming));
                h->fused = true;
                // ... or, better, no staging at all: the direct halo (positions only, the integrator riding on the list kernel
                // or not -- the buffer rotation is the same on all ranks either way)
                const char* no_direct = getenv("AZTOT_NO_DIRECT_HALO");
                if (!(no_direct && no_direct[0] == '1')) h->direct = true;
            }
        } else if (rc != AZ_ERR_UNSUPPORTED) { e->err = h->comm->err; return rc; }
    }
    return AZ_OK;
}

// One grouped exchange: for every direction in use, `esize`-byte entries [soff[d], soff[d+1]) of `send` go
// to the peer in direction d, and the message that peer sends towards me lands in [roff[d], roff[d+1]) of
// `recv`.  Receives are posted in descending direction order: between two ranks the k-th send (ascending d
// on the sender) is the message for the receiver's direction 26-d, i.e. descending on the receiver --
// this matters when both neighbours along an axis are the same rank (two bricks on that axis).
static void dd_add_ops(const AzDDHost* h, std::vector<AzP2P>& ops, const void* send, const int* soff, void* recv, const int* roff, size_t esize)
{
    for (int dir = 0; dir < AZ_NDIR; dir++)
        if (h->peer[dir] >= 0) ops.push_back({ h->peer[dir], (const char*)send + (size_t)soff[dir] * esize, (size_t)(soff[dir + 1] - soff[dir]) * esize, nullptr, 0 });
    for (int dir = AZ_NDIR - 1; dir >= 0; dir--)
        if (h->peer[dir] >= 0) ops.push_back({ h->peer[dir], nullptr, 0, (char*)recv + (size_t)roff[dir] * esize, (size_t)(roff[dir + 1] - roff[dir]) * esize });
}

static int dd_exchange(az_engine* e, std::vector<AzP2P>& ops)
{
    int rc = e->dd->comm->exchange(ops, e->stream);
    if (rc) e->err = e->dd->comm->err;
    return rc;
}

// rebuild step, after the wrap: migration, ghost exchange, bin, scan, sort, remap.  Reads the streams of
// buffer c and leaves the sorted streams in buffer 1-c.
static int dd_rebuild(az_engine* e, int c)
{
    AzDDHost* h = e->dd;
    const AzDD& D = h->D;
    const AzParams& P = e->P;
    cudaStream_t st = e->stream;
    const int nbo = (D.cap_owned + AZ_TPB - 1) / AZ_TPB, nbt = (D.cap_total + AZ_TPB - 1) / AZ_TPB;
    auto blocks = [](int n) { return (n + AZ_TPB - 1) / AZ_TPB; };
    int* flags = &e->sc->overflow;
    float2* th = e->use_therm ? e->therm[c] : nullptr;
    {
        Timer t(e, AZ_TIMER_MIGRATE);
        AZ_LAUNCH(az_k_dd_begin_rebuild, 1, 32, 0, st, D);
        AZ_LAUNCH(az_k_mig_classify, nbo, AZ_TPB, 0, st, P, D, e->pos[c], h->migList, h->stay, flags);
        AZ_LAUNCH(az_k_mig_pack, blocks(D.ms[AZ_NDIR]), AZ_TPB, 0, st, D, h->migList, e->pos[c], e->vel[c], e->id[c], th, h->mSend);
        std::vector<AzP2P> ops;
        int rc;
        const bool pr = h->peer_rebuild;
        const int epoch = h->slot_epoch + 1; // (advanced below, with the slot tables)
        const int* wait = pr ? h->haloFlags : nullptr;
        AzOffsets om, og, omr, ogr;
        memcpy(om.o, D.ms, sizeof(om.o)); memcpy(og.o, D.gs, sizeof(og.o)); memcpy(omr.o, D.mr, sizeof(omr.o)); memcpy(ogr.o, D.gr, sizeof(ogr.o));
        if (pr) {
            AzPeerArr<float2> none; memset(&none, 0, sizeof(none));
            AZ_LAUNCH((az_k_peer_send<AzMigRec, float2>), blocks(D.ms[AZ_NDIR]), AZ_TPB, 0, st, om, omr, h->cnt->migCount, h->mSend, h->peerMig, (const float2*)nullptr, none,
                      h->peerFlags, 2 * AZ_NDIR, epoch, h->putDone + 2);
        } else {
            dd_add_ops(h, ops, h->mSend, D.ms, h->mRecv, D.mr, sizeof(AzMigRec));
            rc = dd_exchange(e, ops);
            if (rc) return rc;
        }
        AZ_LAUNCH(az_k_mig_unpack, blocks(D.mr[AZ_NDIR]), AZ_TPB, 0, st, D, h->mRecv, e->pos[c], e->vel[c], e->id[c], th, h->stay, flags,
                  wait ? wait + 2 * AZ_NDIR : nullptr, epoch, h->halo_dirs);
        AZ_LAUNCH(az_k_ghost_select, nbo, AZ_TPB, 0, st, P, D, e->pos[c], h->stay, h->sendList, flags);
        AZ_LAUNCH(az_k_ghost_pack<true>, blocks(D.gs[AZ_NDIR]), AZ_TPB, 0, st, D, h->sendList, e->pos[c], e->id[c], th, h->gSendPos, h->gSendAux);
        if (pr) {
            AZ_LAUNCH((az_k_peer_send<float4, float2>), blocks(D.gs[AZ_NDIR]), AZ_TPB, 0, st, og, ogr, h->cnt->sendCount, h->gSendPos, h->peerGPos, (const float2*)h->gSendAux, h->peerGAux,
                      h->peerFlags, 3 * AZ_NDIR, epoch, h->putDone + 3);
        } else {
            ops.clear();
            dd_add_ops(h, ops, h->gSendPos, D.gs, h->gRecvPos, D.gr, sizeof(float4));
            std::vector<AzP2P> ops2;
            dd_add_ops(h, ops2, h->gSendAux, D.gs, h->gRecvAux, D.gr, sizeof(float2));
            ops.insert(ops.end(), ops2.begin(), ops2.end());
            rc = dd_exchange(e, ops);
            if (rc) return rc;
        }
        AZ_LAUNCH(az_k_ghost_append, blocks(D.gr[AZ_NDIR]), AZ_TPB, 0, st, D, h->gRecvPos, h->gRecvAux, e->pos[c], e->vel[c], e->id[c], th, h->stay, flags,
                  wait ? wait + 3 * AZ_NDIR : nullptr, epoch, h->halo_dirs);
        AZ_COUNT(e, AZ_TIMER_MIGRATE, 7);
    }
    const int nc = P.g.ncell_local;
    {
        Timer t(e, AZ_TIMER_SCAN);
        AZ_CUDA(e, cudaMemsetAsync(e->cellCount, 0, sizeof(int) * (nc + 1), st));
        AZ_LAUNCH(az_k_bin, nbt, AZ_TPB, 0, st, P, D, e->pos[c], h->stay, e->cellOf, e->slotOf, e->cellCount, flags);
        AZ_LAUNCH(az_k_scan_tiles, e->n_scan_tiles, AZ_SCAN_TPB, 0, st, e->cellCount, nc, e->tileSum);
        AZ_LAUNCH(az_k_scan_apply, e->n_scan_tiles, AZ_SCAN_TPB, 0, st, e->cellCount, nc, e->tileSum, e->first);
        AZ_COUNT(e, AZ_TIMER_SCAN, 3);
    }
    {
        Timer t(e, AZ_TIMER_SORT);
        const int* ntot = &h->cnt->n_tot;
        AZ_LAUNCH(az_k_sort_keys, nbt, AZ_TPB, 0, st, D.cap_total, ntot, e->cellOf, e->slotOf, e->first, e->id[c], e->keyTmp);
        if (e->use_therm)
            AZ_LAUNCH(az_k_sort_scatter<true>, nbt, AZ_TPB, 0, st, D.cap_total, ntot, e->cellOf, e->first, e->cellCount, e->keyTmp, e->pos[c], e->vel[c],
                      e->id[c], e->therm[c], e->pos[1 - c], e->vel[1 - c], e->id[1 - c], e->therm[1 - c], e->cellSorted, h->newIndex, (int*)nullptr AZ_PROBE_SORT_ARGS);
        else
            AZ_LAUNCH(az_k_sort_scatter<false>, nbt, AZ_TPB, 0, st, D.cap_total, ntot, e->cellOf, e->first, e->cellCount, e->keyTmp, e->pos[c], e->vel[c],
                      e->id[c], (const float2*)nullptr, e->pos[1 - c], e->vel[1 - c], e->id[1 - c], (float2*)nullptr, e->cellSorted, h->newIndex, (int*)nullptr AZ_PROBE_SORT_ARGS);
        const int nmax = D.gs[AZ_NDIR] > D.gr[AZ_NDIR] ? D.gs[AZ_NDIR] : D.gr[AZ_NDIR];
        h->slot_epoch++;
        AZ_LAUNCH(az_k_dd_remap, blocks(nmax), AZ_TPB, 0, st, D, h->newIndex, h->sendList, h->ghostSlot, h->peerPtr, h->direct ? 1 : 0, h->slot_epoch, h->putDone + 1);
        AZ_LAUNCH(az_k_dd_publish, 1, 1, 0, st, D, e->first, h->n_owned_cells, nc);
        if (h->fused) { AZ_LAUNCH(az_k_ghost_entry, blocks(D.gr[AZ_NDIR]), AZ_TPB, 0, st, D, h->ghostSlot, h->ghostEntry); AZ_COUNT(e, AZ_TIMER_SORT, 1); }
        h->pos_ghosts_stale = false; // the ghosts were appended to pos[] just now
        AZ_COUNT(e, AZ_TIMER_SORT, 4);
    }
    return AZ_OK;
}

// steps between rebuilds: the ghosts get the current positions (and radii) of their originals
static int dd_refresh_ghosts(az_engine* e, int c, bool on_halo_stream)
{
    AzDDHost* h = e->dd;
    const AzDD& D = h->D;
    cudaStream_t st = on_halo_stream ? h->halo_stream : e->stream;
    if (on_halo_stream) AZ_CUDA(e, cudaStreamWaitEvent(h->halo_stream, h->ev_moved, 0)); // recorded by dd_mark_moved
    auto blocks = [](int n) { return (n + AZ_TPB - 1) / AZ_TPB; };
    float2* th = e->use_therm ? e->therm[c] : nullptr;
    Timer t(e, AZ_TIMER_HALO, st); // events on the stream the halo kernels run on
    if (h->p2p) { // stores straight into the neighbours' staging over NVLink + arrival flags; no NCCL call, no host-side group
        const int seq = ++h->halo_seq;
        if (th) AZ_LAUNCH(az_k_ghost_put<true>, blocks(D.gs[AZ_NDIR]), AZ_TPB, 0, st, D, h->peerPtr, h->sendList, e->pos[c], th, seq, h->putDone);
        else AZ_LAUNCH(az_k_ghost_put<false>, blocks(D.gs[AZ_NDIR]), AZ_TPB, 0, st, D, h->peerPtr, h->sendList, e->pos[c], th, seq, h->putDone);
        AZ_LAUNCH(az_k_ghost_update_wait, blocks(D.gr[AZ_NDIR]), AZ_TPB, 0, st, D, h->ghostSlot, h->stagePos, h->stageAux, e->pos[c], th, h->haloFlags, seq, &e->sc->overflow);
        if (on_halo_stream) AZ_CUDA(e, cudaEventRecord(h->ev_ghosts, h->halo_stream));
        AZ_COUNT(e, AZ_TIMER_HALO, 2);
        return AZ_OK;
    }
    AZ_LAUNCH(az_k_ghost_pack<false>, blocks(D.gs[AZ_NDIR]), AZ_TPB, 0, st, D, h->sendList, e->pos[c], e->id[c], th, h->gSendPos, h->gSendAux);
    std::vector<AzP2P> ops;
    dd_add_ops(h, ops, h->gSendPos, D.gs, h->gRecvPos, D.gr, sizeof(float4));
    if (th) {
        std::vector<AzP2P> ops2;
        dd_add_ops(h, ops2, h->gSendAux, D.gs, h->gRecvAux, D.gr, sizeof(float2));
        ops.insert(ops.end(), ops2.begin(), ops2.end());
    }
    int rc = h->comm->exchange(ops, st);
    if (rc) { e->err = h->comm->err; return rc; }
    AZ_LAUNCH(az_k_ghost_update, blocks(D.gr[AZ_NDIR]), AZ_TPB, 0, st, D, h->ghostSlot, h->gRecvPos, h->gRecvAux, e->pos[c], th);
    if (on_halo_stream) AZ_CUDA(e, cudaEventRecord(h->ev_ghosts, h->halo_stream));
    AZ_COUNT(e, AZ_TIMER_HALO, 2);
    return AZ_OK;
}

// Fused halo, steps between rebuilds: only the put, on the ENGINE's stream right behind the kick + drift -- while the GPU is
// still empty, so that the neighbours see their flags a few microseconds later.  (On the halo stream beside the running list
// pass the put took 60 us: its blocks had to wait for list blocks to retire, and every neighbour's ghost blocks waited with them.)
// The list pass that follows takes the ghosts from the staging; *seq_out / *stage_out tell it which flags value and which half.
static int dd_put_ghosts(az_engine* e, int c, int* seq_out, const float4** stage_out)
{
    AzDDHost* h = e->dd;
    const AzDD& D = h->D;
    auto blocks = [](int n) { return (n + AZ_TPB - 1) / AZ_TPB; };
    Timer t(e, AZ_TIMER_HALO);
    const int seq = ++h->halo_seq;
    static const int put_dbg = getenv("AZTOT_PUT_DBG") ? atoi(getenv("AZTOT_PUT_DBG")) : 0; // diagnostics of the put kernel's cost
    AzPeer R = h->peerPtr;
    if (put_dbg & 2) for (int d = 0; d < AZ_NDIR; d++) if (R.pos[d]) R.pos[d] = h->stagePos; // stores stay on this GPU (results are NOT valid)
    AZ_LAUNCH(az_k_ghost_put<false>, blocks(D.gs[AZ_NDIR]), AZ_TPB, 0, e->stream, D, R, h->sendList, e->pos[c], (const float2*)nullptr, seq, h->putDone, put_dbg);
    AZ_COUNT(e, AZ_TIMER_HALO, 1);
    h->pos_ghosts_stale = true;
    *seq_out = seq;
    *stage_out = h->stagePos + (size_t)(seq & 1) * D.gr[AZ_NDIR];
    return AZ_OK;
}

// Direct halo, steps between rebuilds: my boundary atoms go straight into the ghost slots of the neighbours' current position
// stream (the one with the same number as mine), then the arrival flags.  On the engine's stream behind the kick + drift.
static int dd_put_direct(az_engine* e, int c, AzPairArgs* A)
{
    AzDDHost* h = e->dd;
    const AzDD& D = h->D;
    int buf = -1;
    for (int b = 0; b < 3; b++) if (h->pos_bufs[b] == e->pos[c]) buf = b;
    if (buf < 0) { AZ_SET_ERR(e, "direct halo: unknown position stream"); return AZ_ERR_INVALID; }
    // On the halo stream, beside the list pass that starts right away on the engine's stream: the put reads this step's
    // (final) positions, which nobody writes during this step; the ghost-touching blocks of the neighbours' list passes come
    // last in their launches, so a put that has to wait a few microseconds for SM slots delays nobody.  (On the engine's
    // stream the put cost 20 us of every step: launch, NVLink round trips of the fences, flags.)
    AZ_CUDA(e, cudaEventRecord(h->ev_moved, e->stream));
    AZ_CUDA(e, cudaStreamWaitEvent(h->halo_stream, h->ev_moved, 0));
    {
        Timer t(e, AZ_TIMER_HALO, h->halo_stream);
        const int seq = ++h->halo_seq;
        AZ_LAUNCH(az_k_ghost_put_direct<0>, (D.gs[AZ_NDIR] + AZ_TPB - 1) / AZ_TPB, AZ_TPB, 0, h->halo_stream, D, h->peerPtr, h->sendList, h->remoteSlot, e->pos[c], buf, seq,
                  h->slot_epoch, h->haloFlags, h->putDone, &e->sc->overflow);
        AZ_COUNT(e, AZ_TIMER_HALO, 1);
        A->haloSeq = seq;
    }
    AZ_CUDA(e, cudaEventRecord(h->ev_put, h->halo_stream));
    A->haloFlags = h->haloFlags; A->haloDirs = h->halo_dirs; A->blockOrder = h->blockOrder;
    return AZ_OK;
}
// ... and whatever overwrites this step's positions next (the riding integrator of the NEXT step's list pass) comes behind the put
static int dd_join_put(az_engine* e) { AZ_CUDA(e, cudaStreamWaitEvent(e->stream, e->dd->ev_put, 0)); return AZ_OK; }
// the position stream this step's force pass reads, by number; direct puts are allowed when it differs from the previous step's
static int dd_buf_number(const az_engine* e, int c)
{
    for (int b = 0; b < 3; b++) if (e->dd->pos_bufs[b] == e->pos[c]) return b;
    return -1;
}
static bool dd_direct_ok(az_engine* e, int c) { const int b = dd_buf_number(e, c); return b >= 0 && b != e->dd->prev_step_buf; }
static void dd_note_step_buffer(az_engine* e, int c) { if (e->dd) e->dd->prev_step_buf = dd_buf_number(e, c); }
static int dd_order_blocks(az_engine* e) // behind every list build
{
    AzDDHost* h = e->dd;
    AZ_LAUNCH(az_k_block_order, 1, 1024, 0, e->stream, h->blockFlag, e->list_nblocks, h->blockOrder);
    e->launches++;
    return AZ_OK;
}

// Readers of ghost positions other than the list pass (RDF sampling): scatter the last received ghosts into pos[].  The flags
// of that refresh have been seen by the list pass that ran on this stream, so no wait is needed.
static int dd_sync_ghosts(az_engine* e)
{
    AzDDHost* h = e->dd;
    if (!h || !h->fused || !h->pos_ghosts_stale) return AZ_OK;
    const AzDD& D = h->D;
    AZ_LAUNCH(az_k_ghost_update_wait, (D.gr[AZ_NDIR] + AZ_TPB - 1) / AZ_TPB, AZ_TPB, 0, e->stream, D, h->ghostSlot, h->stagePos, h->stageAux, e->pos[e->cur],
              (float2*)nullptr, h->haloFlags, h->halo_seq, &e->sc->overflow);
    e->launches++;
    h->pos_ghosts_stale = false;
    return AZ_OK;
}

// positions of this step are final (the kick + drift kernel is the last thing on the engine's stream so far)
static int dd_mark_moved(az_engine* e)
{
    AZ_CUDA(e, cudaEventRecord(e->dd->ev_moved, e->stream));
    return AZ_OK;
}

static int dd_wait_ghosts(az_engine* e)
{
    AZ_CUDA(e, cudaStreamWaitEvent(e->stream, e->dd->ev_ghosts, 0));
    return AZ_OK;
}
static bool dd_overlap(const az_engine* e) { return e->dd && e->dd->overlap && e->use_list; }
static bool dd_peer_halo(const az_engine* e) { return e->dd && e->dd->p2p; }
static bool dd_direct(const az_engine* e) { return e->dd && e->dd->direct && e->dd->overlap; }
static bool dd_fused(const az_engine* e) { return e->dd && e->dd->fused && e->dd->overlap && !e->dd->direct; }
static int dd_fused_args(az_engine* e, AzPairArgs& A)
{
    AzDDHost* h = e->dd;
    A.ghostEntry = h->ghostEntry; A.haloFlags = h->haloFlags; A.haloDirs = h->halo_dirs;
    e->pair_stream = h->halo_stream;
    return AZ_OK;
}
// the halo stream starts its GHOSTS launch behind the put (and the clear_clist memset) of the engine's stream
static int dd_fork_halo_stream(az_engine* e)
{
    AzDDHost* h = e->dd;
    AZ_CUDA(e, cudaEventRecord(h->ev_moved, e->stream));
    AZ_CUDA(e, cudaStreamWaitEvent(h->halo_stream, h->ev_moved, 0));
    return AZ_OK;
}
// the GHOSTS launch is the last thing on the halo stream: the engine's stream goes on behind it
static int dd_join_halo_stream(az_engine* e)
{
    AzDDHost* h = e->dd;
    AZ_CUDA(e, cudaEventRecord(h->ev_ghosts, h->halo_stream));
    AZ_CUDA(e, cudaStreamWaitEvent(e->stream, h->ev_ghosts, 0));
    return AZ_OK;
}
static int dd_n_owned_cells(const az_engine* e) { return e->dd ? e->dd->n_owned_cells : 0; }
static int* dd_block_flags(const az_engine* e) { return e->dd ? e->dd->blockFlag : nullptr; }
static void dd_set_overlap(az_engine* e, bool on) { if (e->dd) e->dd->overlap = on; }

static int dd_read_counts(az_engine* e, AzDDCounts* out)
{
    AZ_CUDA(e, cudaMemcpyAsync(out, e->dd->cnt, sizeof(AzDDCounts), cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    return AZ_OK;
}

// equilibration step on bricks: temp_scale (src/cuTemp.cu:77-107) needs the kinetic energy of the whole system
static __global__ void az_k_dd_ke_out(const AzScalars* __restrict__ sc, double* __restrict__ red) { red[0] = sc->engKin; red[1] = sc->engKin; }
static __global__ void az_k_dd_tscale_k(const __grid_constant__ AzParams P, AzScalars* __restrict__ sc, const double* __restrict__ red)
{
    const float keg = (float)red[0];
    float k = 1.f;
    if (keg != 0.f) {
        const float c = (P.tstat == AZ_TSTAT_RADI) ? 0.25f : 1.f;
        k = sqrtf(c * P.teKin / keg);
    }
    sc->tscale_k = k;
    // after_tscale (src/cuTemp.cu:109-113) sets engKin := teKin; each brick keeps its share so that the sum is teKin
    const double share = red[0] != 0.0 ? red[1] / red[0] : 0.0;
    sc->engKin = (double)P.teKin * share; sc->engKin_f = (float)sc->engKin;
    sc->keGlob = P.teKin;
}

// Nose-Hoover on bricks, second application of the step: all-reduce this step's kinetic energy, then the scalar recurrence
static int dd_global_nose(az_engine* e)
{
    AzDDHost* h = e->dd;
    AZ_LAUNCH(az_k_dd_ke_out, 1, 1, 0, e->stream, e->sc, h->red);
    int rc = h->comm->allreduce(h->red, 1, false, e->stream); // red[1] keeps the local value
    if (rc) { e->err = h->comm->err; return rc; }
    AZ_LAUNCH(az_k_nose_scalars, 1, 1, 0, e->stream, e->P, e->sc, (const double*)h->red, 2);
    e->launches += 2;
    return AZ_OK;
}

static int dd_global_tscale(az_engine* e)
{
    AzDDHost* h = e->dd;
    AZ_LAUNCH(az_k_dd_ke_out, 1, 1, 0, e->stream, e->sc, h->red);
    int rc = h->comm->allreduce(h->red, 1, false, e->stream); // red[1] keeps the local value
    if (rc) { e->err = h->comm->err; return rc; }
    AZ_LAUNCH(az_k_dd_tscale_k, 1, 1, 0, e->stream, e->P, e->sc, h->red);
    e->launches += 2;
    return AZ_OK;
}
