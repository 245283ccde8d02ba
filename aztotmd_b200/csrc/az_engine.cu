// az_engine.cu -- the C ABI of include/aztot_b200.h: engine life cycle, the per-step launch
// sequence and the host<->device plumbing.  This file replaces the reference's host->device seam
// (init_cudaMD src/cuInit.cu:756-1199, the loop body src/main.cu:281-383, iter_fastCellList
// src/cuPairs.cu:2519-2568, apply_tstat src/cuTemp.cu:775-803, md_to_host src/cuInit.cu:1212-1261,
// free_device_md src/cuInit.cu:1263-1373).  There is no CPU fallback: without a CUDA device
// az_engine_create fails with AZ_ERR_NO_DEVICE.
#define AZ_EWALD_KERNELS
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>

#include "az_pair_launch.cuh"

// ----------------------------------------------------------------------------------------
static thread_local std::string g_create_error;

#define AZ_SET_ERR(e, ...)                                   \
    do {                                                     \
        char buf_[512];                                      \
        snprintf(buf_, sizeof(buf_), __VA_ARGS__);           \
        if (e) (e)->err = buf_; else g_create_error = buf_;  \
    } while (0)

#define AZ_CUDA(e, call)                                                                       \
    do {                                                                                       \
        cudaError_t err_ = (call);                                                             \
        if (err_ != cudaSuccess) {                                                             \
            AZ_SET_ERR(e, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err_), __FILE__, __LINE__); \
            return AZ_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

static void host_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The 3072 preset photon directions of read_tstat (src/temperature.cpp:164-223): three axis
// permutations x 32 azimuths x 16 polar angles, each followed by its opposite; doubles, exactly the host arithmetic of
// the reference (tests/test_oracle_pinning.py compares the table bit for bit with what read_tstat builds).
extern "C" int az_default_uvects(double* ux, double* uy, double* uz)
{
    if (!ux || !uy || !uz) return AZ_ERR_INVALID;
    const double pi = 3.14159265359, twopi = 2.0 * pi; // src/const.h:11-12
    int k = 0;
    for (int perm = 0; perm < 3; perm++)
        for (int i = 0; i < 32; i++) {
            double phi = (double)i / 32 * twopi;
            for (int j = 0; j < 16; j++) {
                double theta = (double)j / 16 * pi;
                double st = sin(theta), ct = cos(theta), sp = sin(phi), cp = cos(phi);
                double a, b, c;
                if (perm == 0) { c = st; b = sp * ct; a = cp * ct; }
                else if (perm == 1) { b = st; c = sp * ct; a = cp * ct; }
                else { a = st; b = sp * ct; c = cp * ct; }
                ux[k] = a; uy[k] = b; uz[k] = c;
                ux[k + 1] = -a; uy[k + 1] = -b; uz[k + 1] = -c;
                k += 2;
            }
        }
    return AZ_OK;
}

static void default_uvects(float* uv)
{
    std::vector<double> t(3 * AZ_N_UVECT);
    az_default_uvects(t.data(), t.data() + AZ_N_UVECT, t.data() + 2 * AZ_N_UVECT);
    for (int k = 0; k < AZ_N_UVECT; k++) { uv[3 * k] = (float)t[k]; uv[3 * k + 1] = (float)t[AZ_N_UVECT + k]; uv[3 * k + 2] = (float)t[2 * AZ_N_UVECT + k]; }
}

// split_cells, div_type 1 (src/cuCellList.cu:9-34 as called from src/cuInit.cu:979)
static void split_cells(const float L[3], float r, int cnum[3], float csize[3], float crev[3])
{
    for (int a = 0; a < 3; a++) {
        int n = (int)floorf(L[a] / r);
        if (n < 1) n = 1;
        cnum[a] = n;
        csize[a] = L[a] / (float)n;
        crev[a] = (float)n / L[a];
    }
}

// largest cell offset whose minimal distance (delt-1)*csize is still inside rc, capped at n/2
// (cell_dist_int, src/cuCellList.cu:98-132)
static int stencil_half(int n, float csize, float rc)
{
    int s = 1;
    while ((float)s * csize <= rc && s < n) s++;
    if (s > n / 2) s = n / 2;
    return s;
}

template <class T> static int dev_alloc(az_engine* e, T** p, size_t count)
{
    AZ_CUDA(e, cudaMalloc(p, (count ? count : 1) * sizeof(T)));
    // the engine's stream is non-blocking: a memset on the legacy stream is not ordered against it,
    // so the clear is finished before the buffer can be used anywhere
    AZ_CUDA(e, cudaMemset(*p, 0, (count ? count : 1) * sizeof(T)));
    AZ_CUDA(e, cudaStreamSynchronize(0));
    e->dev_bytes += (count ? count : 1) * sizeof(T);
    return AZ_OK;
}
#define AZ_ALLOC(e, p, n) do { int rc_ = dev_alloc(e, &(p), (size_t)(n)); if (rc_) return rc_; } while (0)

#ifdef AZ_PROBE_MORTON // probe build only: sub-cell scale of the Morton in-cell order (az_k_sort_scatter)
#define AZ_PROBE_SORT_ARGS , 4.f * e->P.g.crev[0], 4.f * e->P.g.crev[1], 4.f * e->P.g.crev[2]
#else
#define AZ_PROBE_SORT_ARGS
#endif
static int dd_setup(az_engine* e, const az_system_desc* d, const az_cuda_opts* o);
static int ewald_setup(az_engine* e, const az_system_desc* d, double qs[2]);
static int bonded_setup(az_engine* e, const az_system_desc* d);
static int dd_alloc(az_engine* e);
static int dd_store_keep(az_engine* e, const az_system_desc* d);
static void dd_free(az_engine* e);
static int dd_rebuild(az_engine* e, int c);
static int dd_refresh_ghosts(az_engine* e, int c, bool on_halo_stream);
static int dd_wait_ghosts(az_engine* e);
static int dd_mark_moved(az_engine* e);
static bool dd_overlap(const az_engine* e);
static bool dd_peer_halo(const az_engine* e);
static bool dd_fused(const az_engine* e);
static bool dd_direct(const az_engine* e);
static int dd_n_owned_cells(const az_engine* e);
static int dd_put_direct(az_engine* e, int c, AzPairArgs* A);
static int dd_order_blocks(az_engine* e);
static int dd_join_put(az_engine* e);
static bool dd_direct_ok(az_engine* e, int c);
static void dd_note_step_buffer(az_engine* e, int c);
static int dd_put_ghosts(az_engine* e, int c, int* seq_out, const float4** stage_out);
static int dd_fork_halo_stream(az_engine* e);
static int dd_sync_ghosts(az_engine* e);
static int dd_fused_args(az_engine* e, AzPairArgs& A);
static int dd_join_halo_stream(az_engine* e);
static int* dd_block_flags(const az_engine* e);
static void dd_set_overlap(az_engine* e, bool on);
static int dd_init_counts(az_engine* e, int n_keep);
static int dd_reduce_stats(az_engine* e, az_stats* o, int* flags);
static int dd_owned_count(az_engine* e, int* n);
static const std::vector<int>* dd_keep(az_engine* e);
static int dd_global_tscale(az_engine* e);
static int stage_upload(az_engine* e, const void* const src[6], int len, size_t esize, void* dst[6], int* slot_out);
static int dd_upload(az_engine* e, const double* const src[6], double* dst[6], int* slot);

// ---------------------------------------------------------------------------------------- ABI
extern "C" int az_abi_version(void) { return AZ_ABI_VERSION; }

extern "C" const char* az_last_error(const az_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" int az_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static int build_params(az_engine* e, const az_system_desc* d, const az_cuda_opts* o)
{
    AzParams& P = e->P;
    memset(&P, 0, sizeof(P));
    P.n_local = d->n_atoms;
    P.nspec = d->n_species;
    P.L[0] = (float)d->la; P.L[1] = (float)d->lb; P.L[2] = (float)d->lc; // src/cuInit.cu:936
    for (int a = 0; a < 3; a++) {
        if (!(P.L[a] > 0.f)) { AZ_SET_ERR(e, "box edge %d is not positive", a); return AZ_ERR_INVALID; }
        P.revL[a] = 1.f / P.L[a];
    }
    e->revArea[0] = 1.f / (P.L[1] * P.L[2]); // revEdgeArea, src/cuInit.cu:939-942
    e->revArea[1] = 1.f / (P.L[0] * P.L[2]);
    e->revArea[2] = 1.f / (P.L[0] * P.L[1]);
    P.tSt = (float)d->timestep;
    P.r2max = (float)d->r2_max;
    P.skin = (o && o->rebuild_every > 1) ? o->skin : 0.f;
    if (P.skin < 0.f) { AZ_SET_ERR(e, "negative skin"); return AZ_ERR_INVALID; }
    if (o && o->rebuild_every > 1 && !(o->skin > 0.f)) { // a list reused for R steps with no margin is invalid after the first move
        AZ_SET_ERR(e, "rebuild %d needs a positive 'skin' (cuda.txt): cells and neighbour lists are reused for %d steps, so every atom needs room to move skin/2",
                   o->rebuild_every, o->rebuild_every);
        return AZ_ERR_INVALID;
    }
    float rc = sqrtf(P.r2max);
    P.r2list = P.skin > 0.f ? (rc + P.skin) * (rc + P.skin) : P.r2max;
    for (int a = 0; a < 3; a++) P.efield[a] = (float)d->efield[a];
    P.has_efield = (P.efield[0] != 0.f || P.efield[1] != 0.f || P.efield[2] != 0.f);
    P.alpha = (float)d->alpha; P.daipi2 = (float)d->daipi2; P.elC1 = (float)d->el_c1; P.elC2 = (float)d->el_c2;
    P.rElec = (float)d->r_real;
    for (int i = 0; i < P.nspec; i++) {
        const az_species& s = d->species[i];
        P.mass[i] = (float)s.mass;
        if (!(P.mass[i] > 0.f)) { AZ_SET_ERR(e, "species %d has non-positive mass", i); return AZ_ERR_INVALID; }
        P.charge[i] = (float)s.charge;
        P.rMasshdT[i] = 0.5f * (float)d->timestep / P.mass[i]; // src/cuInit.cu:779
        P.radA[i] = (float)s.radA; P.radB[i] = (float)s.radB; P.mxEng[i] = (float)s.mxEng;
        P.frozen[i] = s.frozen;
    }
    P.tstat = d->tstat_kind;
    P.teKin = (float)d->t_kin;
    if (P.tstat == AZ_TSTAT_NOSE) { // src/sys_init.cpp:1107-1111, narrowed like init_cuda_tstat (src/cuTemp.cu:20-23)
        P.rQmass = (float)(0.5 / d->t_kin / d->tstat_tau / d->tstat_tau);
        P.qMassTau2 = (float)(2.0 * d->t_kin);
    }
    e->reset_vels = d->reset_vels;
    P.seed_lo = (uint32_t)d->seed; P.seed_hi = (uint32_t)(d->seed >> 32);

    // cell grid (split_cells) and stencil
    AzGrid& g = P.g;
    if (!(d->cell_size > 0.0)) { AZ_SET_ERR(e, "cell_list size must be positive (the CUDA path needs 'cell_list')"); return AZ_ERR_INVALID; }
    split_cells(P.L, (float)d->cell_size, g.n, g.csize, g.crev);
    float rlist = sqrtf(P.r2list);
    for (int a = 0; a < 3; a++) {
        g.sten[a] = stencil_half(g.n[a], g.csize[a], rlist);
        g.nl[a] = g.n[a]; g.goff[a] = 0; g.dec[a] = 0; g.olo[a] = 0; g.ohi[a] = g.n[a];
    }
    long long nc = (long long)g.n[0] * g.n[1] * g.n[2];
    if (nc > 0x7fffffff / 2) { AZ_SET_ERR(e, "too many cells (%lld)", nc); return AZ_ERR_INVALID; }
    g.ncell_local = (int)nc;
    return AZ_OK;
}

static int pick_pair_config(az_engine* e, const az_system_desc* d, const az_cuda_opts* o)
{
    const AzGrid& g = e->P.g;
    int ns = d->n_species;
    // table of prepared entries (cudaVdW + chProd, src/cuInit.cu:794-904)
    e->h_table.assign((size_t)ns * ns, AzVdw());
    bool any_radii = false, all_same = true, any = false;
    int kind0 = -1;
    for (int i = 0; i < ns; i++)
        for (int j = 0; j < ns; j++) {
            const az_vdw_entry& v = d->vdw[i * ns + j];
            AzVdw& t = e->h_table[i * ns + j];
            t.p0 = (float)v.p0; t.p1 = (float)v.p1; t.p2 = (float)v.p2; t.p3 = (float)v.p3; t.p4 = (float)v.p4;
            t.r2cut = (float)v.r2cut;
            t.kind = v.kind;
            t.chprod = (float)(d->species[i].charge * d->species[j].charge); // src/cuInit.cu:823
            if (v.kind == AZ_VDW_P746) { AZ_SET_ERR(e, "vdw kind p746 exists only in the reference's serial build"); return AZ_ERR_UNSUPPORTED; }
            if (v.kind < 0 || v.kind > AZ_VDW_SURK) { AZ_SET_ERR(e, "unknown vdw kind %d", v.kind); return AZ_ERR_INVALID; }
            if (v.kind != AZ_VDW_NONE) {
                any = true;
                if (kind0 < 0) kind0 = v.kind; else if (kind0 != v.kind) all_same = false;
                if (v.kind == AZ_VDW_SURK || v.use_radii) any_radii = true;
            } else all_same = all_same && false;
        }
    if (!any) all_same = false;
    e->pair_radii = any_radii;
    e->pair_single = (ns == 1);
    e->pair_ck = d->elec_kind;
    if (e->pair_ck < 0 || e->pair_ck > AZ_ELEC_FENNEL) { AZ_SET_ERR(e, "unknown elec kind %d", e->pair_ck); return AZ_ERR_INVALID; }
    // fast paths: one kind everywhere; everything else goes through the per-pair switch
    if (all_same && kind0 == AZ_VDW_LJ && e->pair_ck == AZ_ELEC_NONE) e->pair_vk = AZ_VDW_LJ;
    else if (all_same && kind0 == AZ_VDW_BUCK) { e->pair_vk = AZ_VDW_BUCK; e->pair_single = false; }
    else if (all_same && kind0 == AZ_VDW_SURK && e->pair_ck == AZ_ELEC_NONE) e->pair_vk = AZ_VDW_SURK;
    else { e->pair_vk = AZ_VK_MIXED; e->pair_single = false; }

    // Packed list pass (one Lennard-Jones entry, no Coulomb): the r2 above which force_autocap (src/cuPairs.cu:94-115) is
    // the identity.  F/r = p2/r2 * s (2 s - 1), s = (p1/r2)^3, falls monotonically through the repulsive range; find where
    // it reaches 0.9 MX_FRC and check the attractive side on a grid.  No such r2 (absurd parameters): the kernel always
    // takes the clamped scalar form.
    e->r2_safe = 3.4e38f;
    if (e->pair_vk == AZ_VDW_LJ && e->pair_single && e->pair_ck == AZ_ELEC_NONE && !e->pair_radii) {
        const AzVdw& t = e->h_table[0];
        auto fr = [&](double r2) { const double s = pow((double)t.p1 / r2, 3.0); return (double)t.p2 / r2 * s * (2.0 * s - 1.0); };
        const double lim = 0.9 * (double)AZ_MX_FRC, gate = std::min((double)e->P.r2max, (double)t.r2cut);
        double lo = 1e-12 * (double)t.p1, hi = (double)t.p1; // r = 1e-6 sigma ... sigma
        if (t.p1 > 0.f && t.p2 > 0.f && fr(hi) < lim && fr(lo) > lim) {
            for (int it = 0; it < 200; it++) { const double mid = sqrt(lo * hi); if (fr(mid) > lim) lo = mid; else hi = mid; }
            bool ok = true;
            for (int k = 0; k <= 4096 && ok; k++) { const double r2 = hi * pow(std::max(gate, hi) / hi, k / 4096.0); if (fabs(fr(r2)) > lim) ok = false; }
            if (ok) e->r2_safe = (float)(hi * 1.0001);
        }
    }

    // CTA shape: a z-run of cells holding about one thread block of atoms
    const double owned_cells = (double)(g.ohi[0] - g.olo[0]) * (g.ohi[1] - g.olo[1]) * (g.ohi[2] - g.olo[2]);
    double avg = (double)e->n / owned_cells;
    int zc = (o && o->cells_per_block > 0) ? o->cells_per_block : (int)floor((double)AZ_PAIR_TPB / (avg > 0.05 ? avg : 0.05));
    int zmax = AZ_PAIR_MAXW - 2 * g.sten[2];
    if (zmax < 1) { AZ_SET_ERR(e, "cell size too small for the cutoff: z stencil half width %d", g.sten[2]); return AZ_ERR_UNSUPPORTED; }
    if (zc > zmax) zc = zmax;
    int onz = g.ohi[2] - g.olo[2];
    if (zc > onz) zc = onz;
    if (zc < 1) zc = 1;
    e->pair_zc = zc;
    e->pair_nruns = (onz + zc - 1) / zc;
    e->pair_nblocks = (g.ohi[0] - g.olo[0]) * (g.ohi[1] - g.olo[1]) * e->pair_nruns;
    double want = (zc + 2 * g.sten[2]) * avg * 3.0 + 64.0;
    int cap = 512;
    while (cap < want && cap < 8192) cap *= 2;
    e->pair_tile_cap = cap;
    e->pair_smem = (size_t)cap * sizeof(float4) + (e->pair_radii ? (size_t)cap * sizeof(float) : 0);
    return AZ_OK;
}

extern "C" int az_engine_create(const az_system_desc* d, const az_cuda_opts* o, az_engine** out)
{
    g_create_error.clear();
    if (!out) { AZ_SET_ERR((az_engine*)0, "out is NULL"); return AZ_ERR_INVALID; }
    *out = nullptr;
    if (!d) { AZ_SET_ERR((az_engine*)0, "desc is NULL"); return AZ_ERR_INVALID; }
    if (d->abi_version != AZ_ABI_VERSION) { AZ_SET_ERR((az_engine*)0, "abi_version %d != %d", d->abi_version, AZ_ABI_VERSION); return AZ_ERR_INVALID; }
    if (d->n_atoms <= 0 || d->n_species <= 0 || d->n_species > AZ_MAX_SPECIES) { AZ_SET_ERR((az_engine*)0, "bad n_atoms/n_species"); return AZ_ERR_INVALID; }
    if (!d->xs || !d->ys || !d->zs || !d->types || !d->species || !d->vdw) { AZ_SET_ERR((az_engine*)0, "missing atom/species/vdw arrays"); return AZ_ERR_INVALID; }
    if (d->tstat_kind == AZ_TSTAT_NOSE && !(d->tstat_tau > 0.0 && d->t_kin > 0.0)) { AZ_SET_ERR((az_engine*)0, "Nose-Hoover thermostat needs tstat_tau > 0 and t_kin > 0"); return AZ_ERR_INVALID; }
    if (d->reset_vels < 0) { AZ_SET_ERR((az_engine*)0, "reset_vels < 0"); return AZ_ERR_INVALID; }
    if (d->tstat_kind < AZ_TSTAT_NONE || d->tstat_kind > AZ_TSTAT_RADI) { AZ_SET_ERR((az_engine*)0, "unknown thermostat kind"); return AZ_ERR_INVALID; }
    if (d->tstat_kind == AZ_TSTAT_RADI && !d->photons) { AZ_SET_ERR((az_engine*)0, "radiative thermostat needs photon energies"); return AZ_ERR_INVALID; }
    for (int i = 0; i < d->n_atoms; i++)
        if (d->types[i] < 0 || d->types[i] >= d->n_species) { AZ_SET_ERR((az_engine*)0, "atom %d has type %d", i, d->types[i]); return AZ_ERR_INVALID; }
    int ndev = az_device_count();
    if (ndev <= 0) { AZ_SET_ERR((az_engine*)0, "no CUDA device (there is no CPU fallback)"); return AZ_ERR_NO_DEVICE; }
    int dev = o ? o->device : 0;
    if (dev < 0 || dev >= ndev) { AZ_SET_ERR((az_engine*)0, "device %d out of range (%d devices)", dev, ndev); return AZ_ERR_INVALID; }

    az_engine* e = new az_engine();
    e->device = dev;
    int rc = az_engine_init(e, d, o);
    if (rc != AZ_OK) {
        g_create_error = e->err;
        az_engine_destroy(e);
        return rc;
    }
    *out = e;
    return AZ_OK;
}

int az_engine_init(az_engine* e, const az_system_desc* d, const az_cuda_opts* o)
{
    AZ_CUDA(e, cudaSetDevice(e->device));
    cudaDeviceProp prop;
    AZ_CUDA(e, cudaGetDeviceProperties(&prop, e->device));
    e->n_sm = prop.multiProcessorCount;
    if (o && o->n_ranks > 1 && d->tstat_kind == AZ_TSTAT_RADI && !d->photons) { AZ_SET_ERR(e, "radiative thermostat needs photon energies"); return AZ_ERR_INVALID; }
    e->n = d->n_atoms;
    e->n_global = d->n_atoms_global > 0 ? (int)d->n_atoms_global : d->n_atoms;
    e->rebuild_every = (o && o->rebuild_every > 1) ? o->rebuild_every : 1;
    e->use_graph = o && o->use_graph != 0 && !(o->n_ranks > 1); // bricks: NCCL calls sit between the kernels of a step
    e->n_ranks = (o && o->n_ranks > 1) ? o->n_ranks : 1;
    e->rank = o ? o->rank : 0;
    AZ_CUDA(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    AZ_CUDA(e, cudaStreamCreateWithFlags(&e->fin_stream, cudaStreamNonBlocking));
    AZ_CUDA(e, cudaEventCreateWithFlags(&e->ev_fin_fork, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) AZ_CUDA(e, cudaEventCreateWithFlags(&e->ev_fin[k], cudaEventDisableTiming));
    for (int k = 0; k < 2 * AZ_N_TIMERS; k++) AZ_CUDA(e, cudaEventCreate(&e->ev[k]));
    for (int k = 0; k < 2; k++) AZ_CUDA(e, cudaEventCreate(&e->ev_step[k]));
    int rc = build_params(e, d, o);
    if (rc) return rc;
    if (e->n_ranks > 1) { // the global grid of build_params becomes this rank's brick + ghost shell
        rc = dd_setup(e, d, o);
        if (rc) return rc;
        e->n = (int)dd_keep(e)->size();
    }
    rc = pick_pair_config(e, d, o);
    if (rc) return rc;
    e->n_eq = d->n_equil; e->eq_freq = d->eq_freq;
    e->use_therm = (d->tstat_kind == AZ_TSTAT_RADI) || e->pair_radii;

    // capacities: on one GPU exactly the atoms; on a brick the agreed owned / owned+ghost capacities
    const int n = e->n, nc = e->P.g.ncell_local;
    const int capo = e->dd ? e->P.n_local : n, capt = e->dd ? e->cap_total_dd : n;
    e->cap = capt;
    for (int k = 0; k < 2; k++) {
        AZ_ALLOC(e, e->pos[k], capt); AZ_ALLOC(e, e->vel[k], capt); AZ_ALLOC(e, e->id[k], capt);
        if (e->use_therm) AZ_ALLOC(e, e->therm[k], capt);
    }
    AZ_ALLOC(e, e->frc, capo);
    AZ_ALLOC(e, e->cellOf, capt); AZ_ALLOC(e, e->slotOf, capt); AZ_ALLOC(e, e->cellSorted, capt); AZ_ALLOC(e, e->keyTmp, capt);
    AZ_ALLOC(e, e->cellCount, nc + 1); AZ_ALLOC(e, e->first, nc + 1);
    e->n_scan_tiles = (nc + AZ_SCAN_TILE - 1) / AZ_SCAN_TILE;
    AZ_ALLOC(e, e->tileSum, e->n_scan_tiles);
    e->n_atom_blocks = (capo + AZ_TPB - 1) / AZ_TPB;
    e->n_ke_blocks = std::max(e->n_atom_blocks, (capo + AZ_LIST_TPB - 1) / AZ_LIST_TPB); // the list kernel may host the integrator
    AZ_ALLOC(e, e->partKE, (size_t)2 * e->n_ke_blocks); AZ_ALLOC(e, e->partU, e->n_ke_blocks); // (kinetic-energy partials: by step parity)
    // Verlet list between rebuilds (pair_variant: 0 auto, 1 cell tiles only, 2 list)
    // Pair path (pair_variant 0 = auto): between rebuilds always the Verlet list.  On the every-step schedule a large
    // system also goes through the list (build + force pass per step: 1.9 ms at 4 M atoms against 3.6 ms for the
    // shared-memory tile sweep), a small one through the warp-per-atom sweep (one thread per atom would leave the SMs
    // empty and walk the stencil serially).  pair_variant 1 = shared-memory cell tiles, 3 = warp-per-atom sweep.
    const int pv = o ? o->pair_variant : 0;
    e->use_list = pv == 2 || (pv == 0 && (e->rebuild_every > 1 || capo >= 131072));
    e->list_nblocks = (capo + AZ_LIST_TPB - 1) / AZ_LIST_TPB;
    e->list_fast = true;
    for (int a = 0; a < 3; a++) if (e->P.g.n[a] < 2 * e->P.g.sten[a] + 1) e->list_fast = false;
    if (e->use_list) {
        if (capt > AZ_NBR_MASK) { AZ_SET_ERR(e, "the Verlet-list path addresses at most %d atoms per GPU", AZ_NBR_MASK); return AZ_ERR_UNSUPPORTED; }
        const double rl = sqrt((double)e->P.r2list);
        const double vol = (double)e->P.L[0] * e->P.L[1] * e->P.L[2];
        int cap = (o && o->max_neigh > 0) ? o->max_neigh : 0;
        double rho = 0.0;
        const double rho_avg = (double)e->n_global / vol;
        {
            // the densest region of the initial configuration: bins of edge >= r_list
            int nb3[3]; double inv[3];
            for (int a = 0; a < 3; a++) { nb3[a] = (int)floor((double)e->P.L[a] / rl); if (nb3[a] < 1) nb3[a] = 1; if (nb3[a] > 256) nb3[a] = 256; inv[a] = nb3[a] / (double)e->P.L[a]; }
            std::vector<int> bins((size_t)nb3[0] * nb3[1] * nb3[2], 0);
            const double* xyz[3] = { d->xs, d->ys, d->zs };
            int mx = 0;
            for (int i = 0; i < d->n_atoms; i++) {
                int c[3];
                for (int a = 0; a < 3; a++) { double u = xyz[a][i] * inv[a]; u -= floor(u / nb3[a]) * nb3[a]; c[a] = (int)u; if (c[a] >= nb3[a]) c[a] = nb3[a] - 1; if (c[a] < 0) c[a] = 0; }
                int v = ++bins[((size_t)c[0] * nb3[1] + c[1]) * nb3[2] + c[2]];
                if (v > mx) mx = v;
            }
            rho = mx * inv[0] * inv[1] * inv[2];
            if (rho < rho_avg) rho = rho_avg;
        }
        if (cap <= 0) cap = (int)(1.5 * rho * 4.18879 * rl * rl * rl) + 32; // rows from that density
        cap = (cap + 1) & ~1; // tile lists pack two rows per word
        if (cap >= (1 << AZ_NBR_CNT_BITS) / 2) { AZ_SET_ERR(e, "max_neigh %d is beyond the list format (%d rows)", cap, (1 << AZ_NBR_CNT_BITS) / 2 - 1); return AZ_ERR_INVALID; }
        e->nbr_cap = cap;
        e->nbr_stride = (capo + 31) / 32 * 32;
        AZ_ALLOC(e, e->nbr, (size_t)e->nbr_cap * e->nbr_stride);
        AZ_ALLOC(e, e->nbrCount, capo);
        // Tile lists (az_tlist.cuh; list_format: 0 = auto = 1, 1 = global 32-bit rows, 2 = tile lists).  The footprint of a tile in
        // a region with `a` atoms per cell is (2sx+1)(2sy+1) columns x (128 / a + 1 + 2sz cells [+ 2sz ghost cells when z is
        // decomposed]) x a atoms; it grows with a, so the densest region sizes it.  10 % + 32 atoms of margin; a system that
        // is smaller than that is simply held whole.
        {
            const AzGrid& g = e->P.g;
            const double a_max = rho * (double)g.csize[0] * g.csize[1] * g.csize[2];
            const double zcells = 1.0 + 2.0 * g.sten[2] * (g.dec[2] ? 2.0 : 1.0);
            double foot = (2.0 * g.sten[0] + 1.0) * (2.0 * g.sten[1] + 1.0) * ((double)AZ_LIST_TPB + zcells * a_max);
            int fc = (o && o->tile_atoms > 0) ? o->tile_atoms : (int)(1.10 * foot) + 32;
            e->tl_user_cap = o && o->tile_atoms > 0;
            if (fc > capt || capt <= 4096) fc = capt;
            fc = (fc + 31) & ~31;
            e->tl_foot_alloc = (capt + 31) & ~31;
            e->tl_stride = (e->pair_single && !e->pair_radii) ? 3 : 5; // words per atom of the shared-memory footprint (az_k_pair_tlist)
            const int bytes_per_atom = 4 * e->tl_stride;
            const int lf = o ? o->list_format : 0;
            e->tl_foot_cap = fc;
            e->tl_smem_pass = (size_t)fc * bytes_per_atom;
            int rs = (int)(1.25 * rho * 4.18879 * rl * rl * rl) + 8;
            if (rs < 32) rs = 32;
            if (rs > cap) rs = cap;
            rs = (rs + 1) & ~1;
            e->tl_rows_smem = rs;
            e->tl_smem_build = (size_t)rs * AZ_LIST_TPB * sizeof(unsigned short);
            const size_t smem_max = 200 * 1024;
            const bool fits = (long long)(fc - 1) * e->tl_stride <= 65535 && e->tl_smem_pass <= smem_max && e->tl_smem_build <= smem_max;
            if (lf == 2 && !fits) { AZ_SET_ERR(e, "list_format 2 (tile lists): a footprint of %d atoms does not fit shared memory", fc); return AZ_ERR_UNSUPPORTED; }
            e->use_tlist = lf == 2 && fits; // auto = global rows: measured faster on B200 (az_tlist.cuh)
            const int cap2 = (e->nbr_cap + 1) / 2;
            if (lf == 2) { // (the A/B switch "tile_list" works on engines created with list_format 2)
            AZ_ALLOC(e, e->nbr16, (size_t)cap2 * e->nbr_stride + 512); // (+ slack: the row loop requests eight rows ahead)
            AZ_ALLOC(e, e->tileHdr, e->list_nblocks);
            AZ_ALLOC(e, e->tileRuns, (size_t)e->list_nblocks * AZ_TL_RUNS);
            AZ_ALLOC(e, e->tlStat, 4);
            AZ_CUDA(e, cudaMemset(e->tlStat, 0, 4 * sizeof(int)));
            AZ_CUDA(e, cudaMemset(e->tileHdr, 0, sizeof(int) * e->list_nblocks));
            }
        }
    }
    // pair_variant: 0 auto (list between rebuilds; warp-per-atom sweep on the every-step schedule), 1 = shared-memory
    // cell tiles only, 2 = list, 3 = warp-per-atom sweep
    e->warp_nblocks = (capo + (AZ_WARP_TPB / 32) - 1) / (AZ_WARP_TPB / 32);
    e->use_warp = !e->use_list && pv != 1;
    // ... or one thread per atom (pair_variant 4 only: measured slower than the warp sweep even where it should shine, az_pair.cuh)
    e->thread_nblocks = (capo + AZ_THREAD_TPB - 1) / AZ_THREAD_TPB;
    e->use_thread = !e->use_list && pv == 4;
    if (e->use_thread) e->use_warp = false;
    const bool ewald_rec = d->elec_kind == AZ_ELEC_EWALD && (d->ewald_k[0] > 0 || d->ewald_k[1] > 0 || d->ewald_k[2] > 0);
    e->fuse_pair = e->use_list && !e->use_therm && !e->P.has_efield && !ewald_rec; // the k-space forces come after the pair pass
    if (e->fuse_pair) { AZ_ALLOC(e, e->pos_alt, capt); AZ_ALLOC(e, e->vel_alt, capt); }
    { int np = e->pair_nblocks > e->list_nblocks ? e->pair_nblocks : e->list_nblocks; // partials serve all pair kernels
      if (e->warp_nblocks > np) np = e->warp_nblocks;
      if (e->thread_nblocks > np) np = e->thread_nblocks;
      AZ_ALLOC(e, e->partV, np); AZ_ALLOC(e, e->partC, np); AZ_ALLOC(e, e->partN, np); e->n_pair_partials = np; }
    AZ_ALLOC(e, e->sc, 1);
    { const char* no_pdl = getenv("AZTOT_NO_PDL"); e->pdl = !e->dd && capo <= 200000 && !(no_pdl && no_pdl[0] == '1'); }
    if (e->use_warp) { // the stencil of every owned cell, once (az_k_stencil_table)
        const AzGrid& g = e->P.g;
        const int S = (2 * g.sten[0] + 1) * (2 * g.sten[1] + 1) * (2 * g.sten[2] + 1);
        const int noc = e->dd ? dd_n_owned_cells(e) : g.ncell_local;
        const long long ent = (long long)noc * S;
        if (g.sten[0] < AZ_STEN_OFF && g.sten[1] < AZ_STEN_OFF && g.sten[2] < AZ_STEN_OFF && ent * (long long)sizeof(int2) <= (256ll << 20)) {
            AZ_ALLOC(e, e->stenTab, (size_t)ent);
            e->stenS = S;
            AZ_LAUNCH(az_k_stencil_table, (unsigned)((ent + 255) / 256), 256, 0, e->stream, e->P, noc, S, e->stenTab);
            AZ_CUDA(e, cudaGetLastError());
        }
    }
    if (e->use_graph) { AZ_ALLOC(e, e->step0_dev, 1); e->P.step_base = e->step0_dev; }
    AZ_ALLOC(e, e->table, e->h_table.size());
    AZ_CUDA(e, cudaMemcpy(e->table, e->h_table.data(), e->h_table.size() * sizeof(AzVdw), cudaMemcpyHostToDevice));
    if (e->dd) { rc = dd_alloc(e); if (rc) return rc; rc = dd_store_keep(e, d); if (rc) return rc; }
    const std::vector<int>* keep = dd_keep(e); // NULL on one GPU: every atom, in input order
    auto src = [&](int k) { return keep ? (*keep)[k] : k; };
    if (d->tstat_kind == AZ_TSTAT_RADI) {
        AZ_ALLOC(e, e->photons, capo); AZ_ALLOC(e, e->uv, 3 * AZ_N_UVECT);
        std::vector<float> ph(capo);
        // src/cuTemp.cu:24-33; a brick has more slots than photons: the table repeats
        for (int i = 0; i < capo; i++) ph[i] = (float)d->photons[n > 0 ? src(i % n) : 0];
        AZ_CUDA(e, cudaMemcpy(e->photons, ph.data(), sizeof(float) * capo, cudaMemcpyHostToDevice));
        std::vector<float> uv(3 * AZ_N_UVECT);
        if (d->uvect_x && d->uvect_y && d->uvect_z)
            for (int k = 0; k < AZ_N_UVECT; k++) { uv[3 * k] = (float)d->uvect_x[k]; uv[3 * k + 1] = (float)d->uvect_y[k]; uv[3 * k + 2] = (float)d->uvect_z[k]; }
        else default_uvects(uv.data());
        AZ_CUDA(e, cudaMemcpy(e->uv, uv.data(), sizeof(float) * uv.size(), cudaMemcpyHostToDevice));
    }

    // atoms: double -> float narrowing exactly like src/cuInit.cu:775-779
    std::vector<float4> hp(n), hv(n), hf(n);
    std::vector<int> hid(n);
    std::vector<float2> hth(e->use_therm ? n : 0);
    for (int k = 0; k < n; k++) {
        const int i = src(k);
        const int gid = d->global_ids ? d->global_ids[i] : i;
        hp[k] = make_float4((float)d->xs[i], (float)d->ys[i], (float)d->zs[i], __int_as_float_host(d->types[i]));
        hv[k] = make_float4(d->vxs ? (float)d->vxs[i] : 0.f, d->vys ? (float)d->vys[i] : 0.f, d->vzs ? (float)d->vzs[i] : 0.f, 0.f);
        hf[k] = make_float4(d->fxs ? (float)d->fxs[i] : 0.f, d->fys ? (float)d->fys[i] : 0.f, d->fzs ? (float)d->fzs[i] : 0.f, 0.f);
        hid[k] = gid;
        if (e->use_therm) {
            uint32_t r[4];
            host_philox((uint32_t)gid, 0xA5A5A5A5u, 0, 0, (uint32_t)d->seed, (uint32_t)(d->seed >> 32), r);
            float u = (float)(r[0] % 10000u) / 10000.0f; // rand01(), src/utils.cpp:197-201
            hth[k].x = d->init_engs ? d->init_engs[i] : 0.f;                      // engs = 0, src/cuTemp.cu:40
            hth[k].y = d->init_radii ? d->init_radii[i] : 0.577f + u * 0.0001f;   // src/cuTemp.cu:41
        }
    }
    if (n > 0) {
        AZ_CUDA(e, cudaMemcpy(e->pos[0], hp.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(e->vel[0], hv.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(e->frc, hf.data(), sizeof(float4) * n, cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(e->id[0], hid.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        if (e->use_therm) AZ_CUDA(e, cudaMemcpy(e->therm[0], hth.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    }
    if (e->dd) { rc = dd_init_counts(e, n); if (rc) return rc; }
    if (ewald_rec) {
        double qs[2] = { 0.0, 0.0 }; // sum q, sum q^2 over the atoms this rank keeps
        for (int k = 0; k < n; k++) { const double q = d->species[d->types[src(k)]].charge; qs[0] += q; qs[1] += q * q; }
        rc = ewald_setup(e, d, qs);
        if (rc) return rc;
    }
    if (d->n_bonds > 0 || d->n_angles > 0) { rc = bonded_setup(e, d); if (rc) return rc; }
    e->cur = 0;
    e->binned = false;
    e->since_rebuild = 0;
    e->last_step = -1;
    return AZ_OK;
}

// SURVEY 8f-4.  The bond / angle lists of the description (atoms.xyz indices = persistent ids) become one CSR per kind,
// keyed by persistent id: the entries of atom i are the bonds / angles that touch it, in list order, so that the thread
// owning atom i sums its bonded forces in a fixed order (az_k_bonded).  Replaces the device copies init_cudaMD makes of
// Field::at1/at2/bTypes and centrs/lig1/lig2/angTypes and the index rewriting of sort_bonds / sort_angles.
static int bonded_setup(az_engine* e, const az_system_desc* d)
{
    const int n = e->n;
    if (e->dd) { AZ_SET_ERR(e, "bonds / angles are single-GPU only (a bond may span two bricks)"); return AZ_ERR_UNSUPPORTED; }
    auto bad_atom = [&](int a) { return a < 0 || a >= n; };
    std::vector<int> start(n + 1, 0);
    if (d->n_bonds > 0) {
        if (!d->bond_types || !d->bond_at1 || !d->bond_at2 || !d->bond_type || d->n_bond_types < 2 || d->n_bond_types > 65535) { AZ_SET_ERR(e, "bonds: missing arrays or no bond types"); return AZ_ERR_INVALID; }
        std::vector<AzBondType> bt(d->n_bond_types);
        for (int t = 0; t < d->n_bond_types; t++) {
            const az_bond_type& b = d->bond_types[t];
            if (b.kind < AZ_BOND_NONE || b.kind > AZ_BOND_E612) { AZ_SET_ERR(e, "bond type %d: unknown kind %d", t, b.kind); return AZ_ERR_INVALID; }
            bt[t] = { (float)b.p0, (float)b.p1, (float)b.p2, (float)b.p3, (float)b.p4, b.kind };
        }
        for (int b = 0; b < d->n_bonds; b++) {
            const int a1 = d->bond_at1[b], a2 = d->bond_at2[b], t = d->bond_type[b];
            if (bad_atom(a1) || bad_atom(a2) || a1 == a2 || t < 0 || t >= d->n_bond_types) { AZ_SET_ERR(e, "bond %d: atoms %d %d type %d out of range", b, a1, a2, t); return AZ_ERR_INVALID; }
            if (t == 0 || bt[t].kind == AZ_BOND_NONE) continue; // empty bond, src/cuBonds.cu:731
            start[a1 + 1]++; start[a2 + 1]++;
        }
        for (int i = 0; i < n; i++) start[i + 1] += start[i];
        std::vector<int2> ent(start[n] ? start[n] : 1);
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int b = 0; b < d->n_bonds; b++) {
            const int a1 = d->bond_at1[b], a2 = d->bond_at2[b], t = d->bond_type[b];
            if (t == 0 || bt[t].kind == AZ_BOND_NONE) continue;
            ent[fill[a1]++] = make_int2(a2, t | (1 << 16));
            ent[fill[a2]++] = make_int2(a1, t);
        }
        AZ_ALLOC(e, e->bStart, n + 1); AZ_ALLOC(e, e->bEnt, ent.size()); AZ_ALLOC(e, e->bTypes, bt.size());
        AZ_CUDA(e, cudaMemcpy(e->bStart, start.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(e->bEnt, ent.data(), sizeof(int2) * ent.size(), cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(e->bTypes, bt.data(), sizeof(AzBondType) * bt.size(), cudaMemcpyHostToDevice));
    }
    if (d->n_angles > 0) {
        if (!d->angle_types || !d->angle_centre || !d->angle_lig1 || !d->angle_lig2 || !d->angle_type || d->n_angle_types < 2 || d->n_angle_types > 65535) { AZ_SET_ERR(e, "angles: missing arrays or no angle types"); return AZ_ERR_INVALID; }
        std::vector<AzAngleType> at(d->n_angle_types);
        for (int t = 0; t < d->n_angle_types; t++) {
            if (d->angle_types[t].kind != 0 && d->angle_types[t].kind != AZ_ANGLE_HCOS) { AZ_SET_ERR(e, "angle type %d: unknown kind %d", t, d->angle_types[t].kind); return AZ_ERR_INVALID; }
            at[t] = { (float)d->angle_types[t].p0, (float)d->angle_types[t].p1 };
        }
        std::fill(start.begin(), start.end(), 0);
        auto live = [&](int a) { const int t = d->angle_type[a]; return t != 0 && d->angle_types[t].kind != 0; };
        for (int a = 0; a < d->n_angles; a++) {
            const int c = d->angle_centre[a], l1 = d->angle_lig1[a], l2 = d->angle_lig2[a], t = d->angle_type[a];
            if (bad_atom(c) || bad_atom(l1) || bad_atom(l2) || c == l1 || c == l2 || l1 == l2 || t < 0 || t >= d->n_angle_types) { AZ_SET_ERR(e, "angle %d: atoms %d %d %d type %d out of range", a, c, l1, l2, t); return AZ_ERR_INVALID; }
            if (!live(a)) continue;
            start[c + 1]++; start[l1 + 1]++; start[l2 + 1]++;
        }
        for (int i = 0; i < n; i++) start[i + 1] += start[i];
        std::vector<int4> ent(start[n] ? start[n] : 1);
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int a = 0; a < d->n_angles; a++) {
            if (!live(a)) continue;
            const int c = d->angle_centre[a], l1 = d->angle_lig1[a], l2 = d->angle_lig2[a], t = d->angle_type[a];
            ent[fill[c]++] = make_int4(l1, l2, t, 0);
            ent[fill[l1]++] = make_int4(c, l2, t | (1 << 16), 0);
            ent[fill[l2]++] = make_int4(c, l1, t | (2 << 16), 0);
        }
        AZ_ALLOC(e, e->aStart, n + 1); AZ_ALLOC(e, e->aEnt, ent.size()); AZ_ALLOC(e, e->aTypes, at.size());
        AZ_CUDA(e, cudaMemcpy(e->aStart, start.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(e->aEnt, ent.data(), sizeof(int4) * ent.size(), cudaMemcpyHostToDevice));
        AZ_CUDA(e, cudaMemcpy(e->aTypes, at.data(), sizeof(AzAngleType) * at.size(), cudaMemcpyHostToDevice));
    }
    AZ_ALLOC(e, e->whereId, n);
    { std::vector<int> ident(n); for (int i = 0; i < n; i++) ident[i] = i;
      AZ_CUDA(e, cudaMemcpy(e->whereId, ident.data(), sizeof(int) * n, cudaMemcpyHostToDevice)); }
    AZ_ALLOC(e, e->partBond, e->n_atom_blocks); AZ_ALLOC(e, e->partAngle, e->n_atom_blocks);
    e->bonded = true;
    e->fuse_pair = false; // the integrator cannot ride on the pair pass: the bonded forces follow it
    return AZ_OK;
}

extern "C" void az_engine_destroy(az_engine* e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    dd_free(e);
    for (int k = 0; k < 2; k++) { cudaFree(e->pos[k]); cudaFree(e->vel[k]); cudaFree(e->id[k]); cudaFree(e->therm[k]); }
#ifndef AZ_EMU
    for (auto& g : e->graphs) cudaGraphExecDestroy((cudaGraphExec_t)g.exec);
#endif
    e->graphs.clear();
    cudaFree(e->step0_dev);
    cudaFree(e->ew_ak); cudaFree(e->ew_pairs); cudaFree(e->ew_runs); cudaFree(e->ew_part); cudaFree(e->ew_T); cudaFree(e->ew_Sd);
    cudaFree(e->frc); cudaFree(e->pos_alt); cudaFree(e->vel_alt); cudaFree(e->cellOf); cudaFree(e->slotOf); cudaFree(e->cellSorted); cudaFree(e->keyTmp);
    cudaFree(e->cellCount); cudaFree(e->first); cudaFree(e->tileSum);
    cudaFree(e->partKE); cudaFree(e->partU); cudaFree(e->partV); cudaFree(e->partC); cudaFree(e->partN); cudaFree(e->stage2[0]); cudaFree(e->stage2[1]); cudaFree(e->nbr); cudaFree(e->nbrCount); cudaFree(e->stenTab); cudaFree(e->nbr16); cudaFree(e->tileHdr); cudaFree(e->tileRuns); cudaFree(e->tlStat);
    cudaFree(e->rdf_hist); cudaFree(e->sc); cudaFree(e->table); cudaFree(e->photons); cudaFree(e->uv);
    cudaFree(e->whereId); cudaFree(e->bStart); cudaFree(e->aStart); cudaFree(e->bEnt); cudaFree(e->aEnt); cudaFree(e->bTypes); cudaFree(e->aTypes);
    cudaFree(e->partBond); cudaFree(e->partAngle); cudaFree(e->dl_buf);
    for (int k = 0; k < 2 * AZ_N_TIMERS; k++) if (e->ev[k]) cudaEventDestroy(e->ev[k]);
    for (int k = 0; k < 2; k++) if (e->ev_step[k]) cudaEventDestroy(e->ev_step[k]);
    for (auto& p : e->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (int k = 0; k < 2; k++) { if (e->ev_copied[k]) cudaEventDestroy(e->ev_copied[k]); if (e->ev_unpacked[k]) cudaEventDestroy(e->ev_unpacked[k]); }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->fin_stream) { cudaStreamSynchronize(e->fin_stream); cudaStreamDestroy(e->fin_stream); }
    if (e->ev_fin_fork) cudaEventDestroy(e->ev_fin_fork);
    for (int k = 0; k < 2; k++) if (e->ev_fin[k]) cudaEventDestroy(e->ev_fin[k]);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

// ---------------------------------------------------------------------------------------- step
// NVTX ranges around the sections of a step (tracing).  nvtx3 is header-only and resolves the injection
// library at run time: without a profiler attached a range is a no-op.  AZ_NO_NVTX (and the emulator build) compile them out.
#if !defined(AZ_EMU) && !defined(AZ_NO_NVTX) && __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
struct az_nvtx_range {
    explicit az_nvtx_range(int k)
    {
        static const char* names[AZ_N_TIMERS] = { "az:integrate_a", "az:scan", "az:sort", "az:pair", "az:integrate_b", "az:halo", "az:pair_list", "az:ewald", "az:migrate" };
        nvtxRangePushA(k >= 0 && k < AZ_N_TIMERS ? names[k] : "az:section");
    }
    ~az_nvtx_range() { nvtxRangePop(); }
};
#else
struct az_nvtx_range { explicit az_nvtx_range(int) {} };
#endif
namespace {
// profile 1: synchronous per-section timing; profile 2: events are only recorded (no host sync
// inside the step loop) and resolved by az_engine_timers after the stream has drained.
struct Timer {
    az_engine* e; int k; cudaEvent_t a = nullptr, b = nullptr; cudaStream_t s;
    az_nvtx_range nvtx; // one NVTX range per section (nsys / ncu timelines)
    // `on`: the stream the section's kernels run on (the halo refresh of a brick runs on its own stream)
    Timer(az_engine* e_, int k_, cudaStream_t on = nullptr) : e(e_), k(k_), s(on ? on : e_->stream), nvtx(k_)
    {
        if (e->profile == 1) cudaEventRecord(e->ev[2 * k], s);
        else if (e->profile == 2) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); }
    }
    ~Timer()
    {
        if (e->profile == 1) {
            cudaEventRecord(e->ev[2 * k + 1], s);
            cudaEventSynchronize(e->ev[2 * k + 1]);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e->ev[2 * k], e->ev[2 * k + 1]);
            e->t_ms[k] += ms;
        } else if (e->profile == 2) {
            cudaEventRecord(b, s);
            e->pending.push_back({ a, b, k });
        }
    }
};
} // namespace

#define AZ_COUNT(e, k, nl) do { (e)->launches += (nl); (e)->t_launch[k] += (nl); } while (0)

#include "az_engine_dd.cuh"

// everything on the engine's stream from here on comes behind the finalize kernels that ran beside the list passes
static int fin_join(az_engine* e)
{
    if (e->fin_last >= 0 && e->fin_rec[e->fin_last]) AZ_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_fin[e->fin_last], 0));
    e->fin_rec[0] = e->fin_rec[1] = false;
    e->fin_last = -1;
    return AZ_OK;
}

static bool step_rebuilds(const az_engine* e) { return !e->binned || e->rebuild_every <= 1 || e->since_rebuild + 1 >= e->rebuild_every; }

static bool is_eq_step(const az_engine* e, int step)
{
    return e->n_eq && step <= e->n_eq && e->eq_freq > 0 && (step % e->eq_freq) == 0; // src/cuTemp.cu:777-779
}

// ---- reciprocal-space Ewald: host preparation (prepare_elec, src/elec.cpp:376-397) and the per-step launches
static int ewald_setup(az_engine* e, const az_system_desc* d, double qs[2])
{
    const double kPiD = 3.14159265359, twopi = 2.0 * kPiD; // src/const.h:11-12
    const double eps = d->eps != 0.0 ? d->eps : 1.0;
    const double L[3] = { d->la, d->lb, d->lc };
    const int kx = d->ewald_k[0], ky = d->ewald_k[1], kz = d->ewald_k[2];
    if (kx < 1 || ky < 1 || kz < 1 || !(d->alpha > 0.0)) { AZ_SET_ERR(e, "elec pme: kx, ky, kz must be >= 1 and alpha > 0"); return AZ_ERR_INVALID; }
    if (kx > 255 || ky > 255 || kz > 255) { AZ_SET_ERR(e, "elec pme: at most 255 k-vectors per axis"); return AZ_ERR_UNSUPPORTED; }
    // the cut-off in k space: 1.05 * 2 pi * min(kx / a, ky / b, kz / c) (src/elec.cpp:383-393)
    double rkcut = kx / L[0];
    if (rkcut > ky / L[1]) rkcut = ky / L[1];
    if (rkcut > kz / L[2]) rkcut = kz / L[2];
    rkcut *= twopi * 1.05;
    const double rkcut2 = rkcut * rkcut, mr4a2 = -0.25 / d->alpha / d->alpha;
    // the reference's enumeration (src/elec.cpp:234-324); for one (l, m) the kept n form one interval, contiguous in the list
    std::vector<float> ak; std::vector<int4> runs, pairs;
    int mmin = 0, nmin = 1;
    for (int l = 0; l < kx; l++) {
        const double rkx = l * twopi / L[0];
        for (int m = mmin; m < ky; m++) {
            const double rky = m * twopi / L[1];
            const int am = m < 0 ? -m : m, lmcode = l | (am << 8) | (m < 0 ? 1 << 16 : 0);
            int n_lo = 0, n_hi = 0, k_lo = -1;
            for (int n = nmin; n < kz; n++) {
                const double rkz = n * twopi / L[2];
                const double rk2 = rkx * rkx + rky * rky + rkz * rkz;
                if (rk2 < rkcut2) {
                    if (k_lo < 0) { k_lo = (int)ak.size(); n_lo = n; }
                    else if (n != n_hi + 1) { AZ_SET_ERR(e, "elec pme: k-vectors of one (l, m) are not contiguous"); return AZ_ERR_INVALID; }
                    n_hi = n;
                    ak.push_back((float)(exp(rk2 * mr4a2) / rk2));
                }
            }
            if (k_lo >= 0) {
                if (n_lo < 0 && n_lo != -n_hi) { AZ_SET_ERR(e, "elec pme: asymmetric n range"); return AZ_ERR_INVALID; }
                runs.push_back(make_int4(lmcode, n_lo, n_hi, k_lo));
                for (int n = (n_lo > 0 ? n_lo : 0); n <= n_hi; n++) // (l, m, n) and, if kept, (l, m, -n)
                    pairs.push_back(make_int4(lmcode | (n << 20), k_lo + (n - n_lo), (n > 0 && -n >= n_lo) ? k_lo + (-n - n_lo) : -1, 0));
            }
            nmin = 1 - kz;
        }
        mmin = 1 - ky;
    }
    const int nk = (int)ak.size();
    if (nk == 0) { AZ_SET_ERR(e, "elec pme: no k-vector inside the cut-off"); return AZ_ERR_INVALID; }
    if (nk > AZ_EW_MAX_K || (int)pairs.size() > AZ_EW_MAX_PAIRS) { AZ_SET_ERR(e, "elec pme: %d k-vectors, at most %d are supported", nk, AZ_EW_MAX_K); return AZ_ERR_UNSUPPORTED; }
    AzEwald& E = e->ew;
    E.nk = nk; E.kx = kx > 1 ? kx : 2; E.ky = ky > 1 ? ky : 2; E.kz = kz > 1 ? kz : 2;
    for (int a = 0; a < 3; a++) E.tpl[a] = (float)(twopi / L[a]);
    const double scale = 2.0 * twopi / (L[0] * L[1] * L[2]) * (double)AZ_FCOUL_SCALE / eps; // Elec::scale, src/elec.cpp:380
    E.escale = scale; E.fscale = (float)(2.0 * scale);                                  // Elec::scale2
    E.npairs = (int)pairs.size(); E.nruns = (int)runs.size();
    AZ_ALLOC(e, e->ew_ak, nk); AZ_ALLOC(e, e->ew_pairs, pairs.size()); AZ_ALLOC(e, e->ew_runs, runs.size()); AZ_ALLOC(e, e->ew_T, nk); AZ_ALLOC(e, e->ew_Sd, nk);
    AZ_CUDA(e, cudaMemcpy(e->ew_ak, ak.data(), sizeof(float) * nk, cudaMemcpyHostToDevice));
    AZ_CUDA(e, cudaMemcpy(e->ew_pairs, pairs.data(), sizeof(int4) * pairs.size(), cudaMemcpyHostToDevice));
    AZ_CUDA(e, cudaMemcpy(e->ew_runs, runs.data(), sizeof(int4) * runs.size(), cudaMemcpyHostToDevice));
    E.ak = e->ew_ak; E.pairs = e->ew_pairs; E.runs = e->ew_runs;
    const int ntiles = (e->cap + AZ_EW_TILE - 1) / AZ_EW_TILE;
    e->ew_blocks = std::max(1, std::min(ntiles, 2 * (e->n_sm > 0 ? e->n_sm : 148)));
    AZ_ALLOC(e, e->ew_part, (size_t)e->ew_blocks * nk);
    e->ew_smem_s = sizeof(float2) * AZ_EW_TILE * (size_t)(E.kx + E.ky + E.kz);
    e->ew_smem_f = sizeof(float2) * (AZ_EW_FTPB * (size_t)(E.kx + E.ky + E.kz) + ((nk + 1) & ~1)) + sizeof(int4) * runs.size();
    if (e->ew_smem_s > 200 * 1024 || e->ew_smem_f > 200 * 1024) { AZ_SET_ERR(e, "elec pme: kx + ky + kz = %d does not fit the shared-memory tables (at most 200)", E.kx + E.ky + E.kz); return AZ_ERR_UNSUPPORTED; }
#ifndef AZ_EMU
    if (e->ew_smem_s > 48 * 1024) AZ_CUDA(e, cudaFuncSetAttribute(az_k_ewald_sfac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ew_smem_s));
    if (e->ew_smem_f > 48 * 1024) AZ_CUDA(e, cudaFuncSetAttribute(az_k_ewald_force, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->ew_smem_f));
#endif
    // ewald_const (src/elec.cpp:144-165) from the global charge sums
    if (e->dd) { int rc = dd_allreduce_host(e, qs, 2, false); if (rc) return rc; }
    const double vol = L[0] * L[1] * L[2];
    const double c3 = (double)AZ_FCOUL_SCALE * (-d->alpha / sqrt(kPiD) * qs[1] - 0.5 * kPiD * (qs[0] * qs[0] / d->alpha / d->alpha) / vol) / eps;
    AZ_CUDA(e, cudaMemcpy(&e->sc->engCoul3, &c3, sizeof(double), cudaMemcpyHostToDevice));
    return AZ_OK;
}

// add_elec of the step (src/main.cpp:99; src/main.cu:330-336 in the CUDA twin): k-space forces onto frc, engCoul2
static int ewald_step(az_engine* e, int c)
{
    Timer t(e, AZ_TIMER_EWALD);
    cudaStream_t st = e->stream;
    const AzEwald& E = e->ew;
    AZ_LAUNCH(az_k_ewald_sfac, e->ew_blocks, AZ_EW_TPB, e->ew_smem_s, st, e->P, E, e->pos[c], e->ew_part);
    if (e->dd) { // the structure factor is a sum over all atoms of all bricks
        AZ_LAUNCH(az_k_ewald_reduce, 1, 1024, 0, st, E, e->ew_part, e->ew_blocks, e->ew_Sd, e->ew_T, &e->sc->engCoul2, 0);
        int rc = dd_allreduce_dev(e, (double*)e->ew_Sd, 2 * E.nk, st);
        if (rc) return rc;
        AZ_LAUNCH(az_k_ewald_reduce, 1, 1024, 0, st, E, e->ew_part, e->ew_blocks, e->ew_Sd, e->ew_T, &e->sc->engCoul2, 2);
        AZ_COUNT(e, AZ_TIMER_EWALD, 1);
    } else
        AZ_LAUNCH(az_k_ewald_reduce, 1, 1024, 0, st, E, e->ew_part, e->ew_blocks, e->ew_Sd, e->ew_T, &e->sc->engCoul2, 1);
    const int nb = (e->cap + AZ_EW_FTPB - 1) / AZ_EW_FTPB;
    AZ_LAUNCH(az_k_ewald_force, nb, AZ_EW_FTPB, e->ew_smem_f, st, e->P, E, e->pos[c], e->ew_T, e->frc);
    AZ_COUNT(e, AZ_TIMER_EWALD, 3);
    return AZ_OK;
}

// One MD step.  skip_a: this step's first half kick + drift was already done by the fused kernel of the
// previous step.  fuse_next: do the second half kick of this step and the first half kick + drift of the
// next one in one pass (az_k_integrate_ba).  `slot` = parity of the step count: which vmax2 slot this
// step's drift bookkeeping uses.
static int one_step(az_engine* e, int step, bool want_energy, bool skip_a, bool fuse_next)
{
    const AzParams& P = e->P;
    const int n = e->n, nc = P.g.ncell_local;
    cudaStream_t st = e->stream;
    const bool rebuild = step_rebuilds(e);
    const bool pdl = e->pdl && e->profile == 0;
    const bool radi = (P.tstat == AZ_TSTAT_RADI);
    const bool eq_step = is_eq_step(e, step);
    const int nb = e->n_atom_blocks;
    int c = e->cur;
    const int n_pp = e->use_list ? e->list_nblocks : (e->use_thread ? e->thread_nblocks : (e->use_warp ? e->warp_nblocks : e->pair_nblocks)); // live pair partials this step
    // pending momenta / speed bound of the step in flight: three slots (az_kernels.cuh), numbered by the step's place INSIDE the
    // call -- nothing is pending across calls (the last step of a call does not kick the next one), so the launch sequence of a
    // call does not depend on how many steps ran before it (one captured graph per call shape, not three)
    const int slot = e->call_rel % 3;
    // az_k_finalize of a step whose integrator rode on the list pass has nothing the next pass waits for: it runs on its own
    // stream beside the pass of step n + 1 (it was 16 us of every 570 us step, one block reducing 31 250 partials).  What the
    // two share is double- or triple-buffered: the kinetic-energy partials by step parity, the pending momenta and the speed
    // bound by step % 3; az_k_finalize writes back only its own fields.  The pass of step n + 2 waits for finalize(n).  Every
    // step that does anything else with the scalars (no rider, energies, thermostats, bonds, graph capture) first
    // joins the finalize stream, and the last step of a call never rides, so nothing is pending when a call returns.
    const int kpar = e->call_rel & 1;
    double* const pke = e->partKE + (size_t)kpar * e->n_ke_blocks;
    const bool ovl_ok = e->fin_overlap && fuse_next && !is_eq_step(e, step) && !want_energy && !e->bonded && e->profile != 1;
    const bool ovl = ovl_ok && e->fuse_pair && !e->use_graph;
    // ... and without a rider (the small decks: bin + sort + sweep every step): az_k_finalize is the last of the step's six short
    // kernels; on its own stream it runs beside scan / sort / pair sweep of the next step and is joined before that step's
    // az_k_integrate_ba (the next kernel that touches the scalars and the partials), also inside a captured graph
    const bool ovl2 = ovl_ok && !e->fuse_pair && P.tstat != AZ_TSTAT_NOSE && e->reset_vels <= 0 && !P.has_efield && !e->ew.nk && !e->dd;
    if (!ovl && !ovl2) { int rcj = fin_join(e); if (rcj) return rcj; }
    const int ks = P.step_base ? step - e->call_first : step; // what the kernels get: relative when *step_base supplies the rest

    const bool nose = (P.tstat == AZ_TSTAT_NOSE);
    if (nose) { // first application of the step, before verlet_1stage (src/main.cu:291-299)
        Timer t(e, AZ_TIMER_INTEGRATE_A);
        AZ_LAUNCH(az_k_nose_scalars, 1, 1, 0, st, P, e->sc, (const double*)nullptr, e->dd ? 1 : 0);
        AZ_LAUNCH(az_k_vscale, nb, AZ_TPB, 0, st, P, e->vel[c], e->sc);
        AZ_COUNT(e, AZ_TIMER_INTEGRATE_A, 2);
    }
    if (!skip_a) { // K1: half kick + drift (+ wrap + bin)
        Timer t(e, AZ_TIMER_INTEGRATE_A);
        if (P.has_efield) { AZ_LAUNCH(az_k_reset_step, 1, 1, 0, st, e->sc); AZ_COUNT(e, AZ_TIMER_INTEGRATE_A, 1); }
        if (rebuild && e->dd) { // bricks: wrap only; migration comes between the wrap and the binning
            AZ_LAUNCH(az_k_integrate_a<AZ_IA_WRAP>, nb, AZ_TPB, 0, st, P, slot, e->pos[c], e->vel[c], e->frc, e->cellOf, e->slotOf, e->cellCount, e->sc);
        } else if (rebuild) {
            AZ_CUDA(e, cudaMemsetAsync(e->cellCount, 0, sizeof(int) * (nc + 1), st)); // clear_clist
            AZ_LAUNCH(az_k_integrate_a<AZ_IA_BIN>, nb, AZ_TPB, 0, st, P, slot, e->pos[c], e->vel[c], e->frc, e->cellOf, e->slotOf, e->cellCount, e->sc);
        } else {
            AZ_LAUNCH(az_k_integrate_a<AZ_IA_STALE>, nb, AZ_TPB, 0, st, P, slot, e->pos[c], e->vel[c], e->frc, e->cellOf, e->slotOf, e->cellCount, e->sc);
        }
        AZ_COUNT(e, AZ_TIMER_INTEGRATE_A, 1);
    }
    if (rebuild && e->dd) {
        int rc = dd_rebuild(e, c);
        if (rc) return rc;
        c = e->cur = 1 - c;
        e->binned = true;
        e->since_rebuild = 0;
    } else if (rebuild) {
        {
            Timer t(e, AZ_TIMER_SCAN);
            if (e->n_scan_tiles <= AZ_SCAN_ONE_LAUNCH) {
                AZ_LAUNCH_PDL(pdl, az_k_scan_apply, e->n_scan_tiles, AZ_SCAN_TPB, 0, st, (const int*)e->cellCount, nc, (const int*)nullptr, e->first);
                AZ_COUNT(e, AZ_TIMER_SCAN, 1);
            } else {
                AZ_LAUNCH(az_k_scan_tiles, e->n_scan_tiles, AZ_SCAN_TPB, 0, st, e->cellCount, nc, e->tileSum);
                AZ_LAUNCH(az_k_scan_apply, e->n_scan_tiles, AZ_SCAN_TPB, 0, st, e->cellCount, nc, e->tileSum, e->first);
                AZ_COUNT(e, AZ_TIMER_SCAN, 2);
            }
        }
        {
            Timer t(e, AZ_TIMER_SORT);
            AZ_LAUNCH_PDL(pdl, az_k_sort_keys, nb, AZ_TPB, 0, st, n, (const int*)nullptr, (const int*)e->cellOf, (const int*)e->slotOf, (const int*)e->first, (const int*)e->id[c], e->keyTmp);
            if (e->use_therm)
                AZ_LAUNCH_PDL(pdl, az_k_sort_scatter<true>, nb, AZ_TPB, 0, st, n, (const int*)nullptr, (const int*)e->cellOf, (const int*)e->first, (const int*)e->cellCount, (const int*)e->keyTmp,
                              (const float4*)e->pos[c], (const float4*)e->vel[c], (const int*)e->id[c], (const float2*)e->therm[c], e->pos[1 - c], e->vel[1 - c], e->id[1 - c], e->therm[1 - c],
                              e->cellSorted, (int*)nullptr, e->whereId AZ_PROBE_SORT_ARGS);
            else
                AZ_LAUNCH_PDL(pdl, az_k_sort_scatter<false>, nb, AZ_TPB, 0, st, n, (const int*)nullptr, (const int*)e->cellOf, (const int*)e->first, (const int*)e->cellCount, (const int*)e->keyTmp,
                              (const float4*)e->pos[c], (const float4*)e->vel[c], (const int*)e->id[c], (const float2*)nullptr, e->pos[1 - c], e->vel[1 - c], e->id[1 - c], (float2*)nullptr,
                              e->cellSorted, (int*)nullptr, e->whereId AZ_PROBE_SORT_ARGS);
            AZ_COUNT(e, AZ_TIMER_SORT, 2);
        }
        c = e->cur = 1 - c; // refresh_arrays (src/cuSort.cu:74-100): host-side ping-pong
        e->binned = true;
        e->since_rebuild = 0;
    } else {
        e->since_rebuild++;
        if (e->dd && dd_direct(e)) {
            // direct halo needs a position stream that was NOT the current one of the previous step (a neighbour that is one step
            // behind still reads its ghost slots there): true whenever the integrator rode on the previous list pass; the first
            // step of a call drifts in place, and then the ghosts take the staged route (peer put + scatter) for that step
            if (!dd_direct_ok(e, c)) { int rc = dd_refresh_ghosts(e, c, false); if (rc) return rc; }
        } else if (e->dd) { int rc = (dd_overlap(e) || dd_fused(e)) ? dd_mark_moved(e) : dd_refresh_ghosts(e, c, false); if (rc) return rc; }
    }
    // the integrator rides on the list kernel when the next step's kick may follow this step's forces directly
    const bool ride = fuse_next && e->fuse_pair && !eq_step;
    const bool directh = !rebuild && dd_direct(e) && dd_direct_ok(e, c); // direct halo: boundary atoms stored into the neighbours' position streams, ONE plain list launch
    const bool fusedh = !rebuild && dd_fused(e);  // peer-memory halo: one list launch, ghost-touching blocks last, waiting on arrival flags
    const bool split = !rebuild && !fusedh && !directh && dd_overlap(e);
    dd_note_step_buffer(e, c); // the ghost refresh runs on the halo stream, beside the list blocks that see no ghost
    if (ovl && e->fin_rec[kpar]) { // this pass's rider overwrites the partials and the slot that finalize(n - 2) worked on
        AZ_CUDA(e, cudaStreamWaitEvent(st, e->ev_fin[kpar], 0));
        e->fin_rec[kpar] = false;
    }
    { // K4: pair forces
        const int pair_mode = e->use_thread ? AZ_PAIR_MODE_THREAD : e->use_warp ? AZ_PAIR_MODE_WARP : (!e->use_list ? AZ_PAIR_MODE_TILE : (rebuild ? AZ_PAIR_MODE_TILE_LIST : AZ_PAIR_MODE_LIST));
        const int pair_timer = pair_mode == AZ_PAIR_MODE_LIST ? AZ_TIMER_PAIR_LIST : AZ_TIMER_PAIR;
        Timer t(e, pair_timer);
        AzPairArgs A;
        memset(&A, 0, sizeof(A));
        A.pos = e->pos[c]; A.therm = e->therm[c]; A.cellOfSorted = e->cellSorted; A.first = e->first; A.count = e->cellCount;
        A.frc = e->frc; A.partV = e->partV; A.partC = e->partC; A.partN = e->partN; A.flags = &e->sc->overflow;
        A.zc = e->pair_zc; A.nruns = e->pair_nruns; A.tile_cap = e->pair_tile_cap;
        A.single = e->h_table[0]; A.table = e->table;
        A.nbr = e->nbr; A.nbrCount = e->nbrCount; A.nbr_stride = e->nbr_stride; A.nbr_cap = e->nbr_cap; A.dbg = e->list_dbg; A.r2_safe = e->r2_safe;
        A.nbr16 = e->nbr16; A.tileHdr = e->tileHdr; A.tileRuns = e->tileRuns; A.tlStat = e->tlStat; A.tl_rows_smem = e->tl_rows_smem; A.tl_foot_cap = e->tl_foot_cap; A.tl_stride = e->tl_stride;
        A.n_total = e->dd ? e->cap_total_dd : e->n;
        A.stenTab = e->stenTab; A.stenS = e->stenS;
        A.prune = (e->warp_prune && rebuild) ? 1 : 0; // the atom-to-cell distance test needs cells that are exact for this step's positions
        A.blockFlag = dd_block_flags(e);
        if (ride) { // second half kick of this step + kick/drift of the next one inside the list kernel
            const bool rb_next = step_rebuilds(e);
            A.fuse = 1 + (rb_next ? (e->dd ? AZ_IA_WRAP : AZ_IA_BIN) : AZ_IA_STALE);
            A.slot_next = (slot + 1) % 3;
            A.vel = e->vel[c]; A.pos_next = e->pos_alt; A.vel_next = e->vel_alt;
            A.cellOf = e->cellOf; A.slotOf = e->slotOf; A.cellCount = e->cellCount; A.sc = e->sc; A.partKE = pke;
            e->clear_counts_bytes = (rb_next && !e->dd) ? sizeof(int) * (size_t)(nc + 1) : 0; // clear_clist for the next step's binning
        }
        int rc;
        if (directh) { // (the put could not come earlier: it needs the drift of this step, which may have been the previous pass's rider)
            rc = dd_put_direct(e, c, &A);
            if (rc) return rc;
            A.phase = 4;
        }
        if (fusedh) {
            // Fused halo.  Engine stream: put my boundary atoms into the neighbours' staging (the GPU is empty: a few microseconds),
            // then the ghost-free blocks of the list pass (phase 1).  Halo stream, forked behind the put: the GHOSTS instantiation for
            // the blocks that list ghosts (it waits for the neighbours' flags itself).  The two list launches run side by side;
            // the engine's stream joins the halo stream afterwards.
            if (e->clear_counts_bytes) { // the riders of BOTH launches bin the next step: clear_clist before either starts
                AZ_CUDA(e, cudaMemsetAsync(e->cellCount, 0, e->clear_counts_bytes, st));
                e->clear_counts_bytes = 0;
            }
            rc = dd_put_ghosts(e, c, &A.haloSeq, &A.ghostStage);
            if (rc) return rc;
            rc = dd_fork_halo_stream(e);
            if (rc) return rc;
            rc = dd_fused_args(e, A);
            if (rc) return rc;
            A.phase = 3;
            rc = az_launch_pair(e, A, want_energy, pair_mode);
            if (rc) return rc;
            AZ_COUNT(e, pair_timer, 1);
            A.phase = 1;
        }
        if (split) { // blocks that list no ghost run while the halo is in flight, the others once it has landed
            A.phase = 1;
            rc = az_launch_pair(e, A, want_energy, pair_mode);
            if (rc) return rc;
            rc = dd_refresh_ghosts(e, c, true); // pack -> exchange -> overwrite the ghost slots, on the halo stream
            if (rc) return rc;
            rc = dd_wait_ghosts(e);
            if (rc) return rc;
            A.phase = 2;
            AZ_COUNT(e, pair_timer, 1);
        }
        rc = az_launch_pair(e, A, want_energy, pair_mode);
        if (rc) return rc;
        if (fusedh) { rc = dd_join_halo_stream(e); if (rc) return rc; } // finalize, and whatever moves the atoms next, come after both launches
        if (directh) { rc = dd_join_put(e); if (rc) return rc; }
        if (rebuild && e->use_list && dd_direct(e)) { rc = dd_order_blocks(e); if (rc) return rc; }
        AZ_COUNT(e, pair_timer, pair_mode == AZ_PAIR_MODE_TILE_LIST ? 2 : 1);
        if (ride) { // the new positions / velocities are in the alternate buffers: swap roles
            std::swap(e->pos[c], e->pos_alt);
            std::swap(e->vel[c], e->vel_alt);
            e->clear_counts_bytes = 0;
        }
    }
    if (e->ew.nk) { int rc = ewald_step(e, c); if (rc) return rc; }
    if (e->bonded) { // apply_const_bonds / apply_angles (src/main.cu:307-311,352-362) onto the pair forces
        Timer t(e, AZ_TIMER_PAIR);
        AzBondedArgs B;
        memset(&B, 0, sizeof(B));
        B.pos = e->pos[c]; B.id = e->id[c]; B.whereId = e->whereId; B.frc = e->frc;
        B.bStart = e->bStart; B.bEnt = e->bEnt; B.bTypes = e->bTypes; B.aStart = e->aStart; B.aEnt = e->aEnt; B.aTypes = e->aTypes;
        B.partB = e->partBond; B.partA = e->partAngle;
        AZ_LAUNCH(az_k_bonded, nb, AZ_TPB, 0, st, P, B);
        AZ_LAUNCH(az_k_bonded_reduce, 1, 256, 0, st, e->partBond, e->partAngle, nb, e->sc);
        AZ_COUNT(e, AZ_TIMER_PAIR, 2);
    }
    if (ride && ovl) { // finalize on its own stream, behind this step's pass
        AZ_CUDA(e, cudaEventRecord(e->ev_fin_fork, st));
        AZ_CUDA(e, cudaStreamWaitEvent(e->fin_stream, e->ev_fin_fork, 0));
        // (not timed as a section: events around a kernel that shares the GPU with the next pass measure the pass)
        AZ_LAUNCH(az_k_finalize, 1, 1024, 0, e->fin_stream, P, ks, slot, 1 | 16 | (rebuild ? 32 : 64), pke, e->list_nblocks, e->partV, e->partC, n_pp,
                  e->partU, 0, e->partN, e->pos[c], e->sc, e->revArea[0], e->revArea[1], e->revArea[2]);
        AZ_COUNT(e, AZ_TIMER_INTEGRATE_B, 1);
        AZ_CUDA(e, cudaEventRecord(e->ev_fin[kpar], e->fin_stream));
        e->fin_rec[kpar] = true; e->fin_last = kpar;
    } else if (ride) {
        Timer t(e, AZ_TIMER_INTEGRATE_B);
        AZ_LAUNCH(az_k_finalize, 1, 1024, 0, st, P, ks, slot, 1 | (want_energy ? 2 : 0) | 16 | (rebuild ? 32 : 64), pke, e->list_nblocks, e->partV, e->partC, n_pp,
                  e->partU, 0, e->partN, e->pos[c], e->sc, e->revArea[0], e->revArea[1], e->revArea[2]);
        AZ_COUNT(e, AZ_TIMER_INTEGRATE_B, 1);
    } else
    { // K5: half kick + KE (+ thermostat) + derived quantities
        { int rcj = fin_join(e); if (rcj) return rcj; } // (the previous step's finalize, if it ran beside this step's sweep)
        Timer t(e, AZ_TIMER_INTEGRATE_B);
        const int drift_mode = rebuild ? 32 : 64;
        const int emode = want_energy ? 2 : 0;
        const bool reset_now = e->reset_vels > 0 && step % e->reset_vels == 0;
        float2* th = radi ? e->therm[c] : nullptr;
        const float* ph = radi ? e->photons : nullptr;
        const float* uv = radi ? e->uv : nullptr;
        if (!eq_step) {
            if (fuse_next) {
                // the next step's kick + drift rides on this pass; its mode follows the next step's schedule
                const bool rb_next = step_rebuilds(e);
                const int ns = (slot + 1) % 3;
#define AZ_BA(TS, MODE) AZ_LAUNCH((az_k_integrate_ba<TS, MODE>), nb, AZ_TPB, 0, st, P, ks, ns, e->pos[c], e->vel[c], e->frc, e->id[c], th, ph, uv, \
                                  pke, e->partU, e->cellOf, e->slotOf, e->cellCount, e->sc)
                if (rb_next && e->dd) { if (radi) AZ_BA(true, AZ_IA_WRAP); else AZ_BA(false, AZ_IA_WRAP); }
                else if (rb_next) {
                    AZ_CUDA(e, cudaMemsetAsync(e->cellCount, 0, sizeof(int) * (nc + 1), st)); // clear_clist (the pair pass has read the counts)
                    if (radi) AZ_BA(true, AZ_IA_BIN); else AZ_BA(false, AZ_IA_BIN);
                } else { if (radi) AZ_BA(true, AZ_IA_STALE); else AZ_BA(false, AZ_IA_STALE); }
#undef AZ_BA
            } else if (reset_now) { // zero_vel instead of verlet_2stage (src/main.cu:372-374); apply_tstat still runs the thermostat
                AZ_LAUNCH(az_k_zero_vel, nb, AZ_TPB, 0, st, P, e->vel[c], pke);
                if (radi) { AZ_LAUNCH(az_k_tscale<true>, nb, AZ_TPB, 0, st, P, ks, e->pos[c], e->vel[c], e->id[c], e->therm[c], e->photons, e->uv, e->sc, e->partU, 0); AZ_COUNT(e, AZ_TIMER_INTEGRATE_B, 1); }
            } else if (radi) {
                AZ_LAUNCH(az_k_integrate_b<true>, nb, AZ_TPB, 0, st, P, ks, e->pos[c], e->vel[c], e->frc, e->id[c], e->therm[c], e->photons, e->uv, pke, e->partU);
            } else {
                AZ_LAUNCH(az_k_integrate_b<false>, nb, AZ_TPB, 0, st, P, ks, e->pos[c], e->vel[c], e->frc, e->id[c], (float2*)nullptr, (const float*)nullptr, (const float*)nullptr, pke, e->partU);
            }
            const int radi_mode = radi ? 4 : 0;
            if (nose) { // second application of the step (apply_tstat): needs this step's kinetic energy, then the quantities
                AZ_LAUNCH(az_k_finalize, 1, 1024, 0, st, P, ks, slot, 1 | emode | drift_mode, pke, nb, e->partV, e->partC, n_pp, e->partU, nb, e->partN, e->pos[c], e->sc,
                          e->revArea[0], e->revArea[1], e->revArea[2]);
                if (e->dd) { int rc = dd_global_nose(e); if (rc) return rc; }
                else AZ_LAUNCH(az_k_nose_scalars, 1, 1, 0, st, P, e->sc, (const double*)nullptr, 0);
                AZ_LAUNCH(az_k_vscale, nb, AZ_TPB, 0, st, P, e->vel[c], e->sc);
                AZ_LAUNCH(az_k_finalize, 1, 1024, 0, st, P, ks, slot, 16, pke, nb, e->partV, e->partC, n_pp, e->partU, nb, e->partN, e->pos[c], e->sc,
                          e->revArea[0], e->revArea[1], e->revArea[2]);
                AZ_COUNT(e, AZ_TIMER_INTEGRATE_B, 3);
            } else if (ovl2) { // beside the next step's scan / sort / sweep
                AZ_CUDA(e, cudaEventRecord(e->ev_fin_fork, st));
                AZ_CUDA(e, cudaStreamWaitEvent(e->fin_stream, e->ev_fin_fork, 0));
                AZ_LAUNCH(az_k_finalize, 1, 1024, 0, e->fin_stream, P, ks, slot, 1 | emode | radi_mode | 16 | drift_mode, pke, nb, e->partV, e->partC, n_pp, e->partU, nb, e->partN, e->pos[c], e->sc,
                          e->revArea[0], e->revArea[1], e->revArea[2]);
                AZ_CUDA(e, cudaEventRecord(e->ev_fin[kpar], e->fin_stream));
                e->fin_rec[kpar] = true; e->fin_last = kpar;
            } else
            AZ_LAUNCH(az_k_finalize, 1, 1024, 0, st, P, ks, slot, 1 | emode | radi_mode | 16 | drift_mode, pke, nb, e->partV, e->partC, n_pp, e->partU, nb, e->partN, e->pos[c], e->sc,
                      e->revArea[0], e->revArea[1], e->revArea[2]);
            AZ_COUNT(e, AZ_TIMER_INTEGRATE_B, 2);
        } else {
            // equilibration step: the scale factor needs this step's kinetic energy first
            if (reset_now) AZ_LAUNCH(az_k_zero_vel, nb, AZ_TPB, 0, st, P, e->vel[c], pke);
            else AZ_LAUNCH(az_k_integrate_b<false>, nb, AZ_TPB, 0, st, P, ks, e->pos[c], e->vel[c], e->frc, e->id[c], (float2*)nullptr, (const float*)nullptr, (const float*)nullptr, pke, e->partU);
            AZ_LAUNCH(az_k_finalize, 1, 1024, 0, st, P, ks, slot, 1 | (e->dd ? 0 : 8) | emode | drift_mode, pke, nb, e->partV, e->partC, n_pp, e->partU, nb, e->partN, e->pos[c], e->sc,
                      e->revArea[0], e->revArea[1], e->revArea[2]);
            if (e->dd) { int rc = dd_global_tscale(e); if (rc) return rc; }
            if (radi) AZ_LAUNCH(az_k_tscale<true>, nb, AZ_TPB, 0, st, P, ks, e->pos[c], e->vel[c], e->id[c], e->therm[c], e->photons, e->uv, e->sc, e->partU, 1);
            else AZ_LAUNCH(az_k_tscale<false>, nb, AZ_TPB, 0, st, P, ks, e->pos[c], e->vel[c], e->id[c], (float2*)nullptr, (const float*)nullptr, (const float*)nullptr, e->sc, e->partU, 1);
            if (nose) { // apply_tstat: the rescale comes first, then the thermostat (src/cuTemp.cu:777-794)
                AZ_LAUNCH(az_k_nose_scalars, 1, 1, 0, st, P, e->sc, (const double*)nullptr, e->dd ? 1 : 0); // bricks: keGlob = teKin after the rescale
                AZ_LAUNCH(az_k_vscale, nb, AZ_TPB, 0, st, P, e->vel[c], e->sc);
                AZ_COUNT(e, AZ_TIMER_INTEGRATE_B, 2);
            }
            AZ_LAUNCH(az_k_finalize, 1, 1024, 0, st, P, ks, slot, (radi ? 4 : 0) | 16, pke, nb, e->partV, e->partC, n_pp, e->partU, nb, e->partN, e->pos[c], e->sc,
                      e->revArea[0], e->revArea[1], e->revArea[2]);
            AZ_COUNT(e, AZ_TIMER_INTEGRATE_B, 4);
        }
    }
    e->steps_done++;
    e->last_step = step;
    e->energy_valid = want_energy;
    return AZ_OK;
}

static int run_steps(az_engine* e, int first_step, int nsteps)
{
    bool a_done = false; // the fused kernel of the previous step already did this step's kick + drift
    for (int k = 0; k < nsteps; k++) {
        const int step = first_step + k;
        // Consecutive steps of one call share a pass over the atoms (second half kick of step n + first half
        // kick and drift of step n+1) unless something must happen in between: an equilibration rescale, or the
        // external field's per-step energy reset.  The last step of the call ends on a plain second half kick,
        // so downloads and statistics always see the state "after step n".
        const bool fuse = e->fuse_integrate && k + 1 < nsteps && !is_eq_step(e, step) && !e->P.has_efield && e->P.tstat != AZ_TSTAT_NOSE && e->reset_vels == 0;
        // potential energies are reduced on the last step of the call (the one az_engine_stats reports)
        e->call_rel = k;
        int rc = one_step(e, step, e->energy_every_step || k == nsteps - 1, a_done, fuse);
        if (rc) return rc;
        a_done = fuse;
    }
    return AZ_OK;
}

static __global__ void az_k_set_int(int* p, int v) { *p = v; }

#ifndef AZ_EMU
// What decides the launch sequence of a call, apart from the step numbers the kernels read from the device: the number
// of steps, where the engine stands in its rebuild / buffer ping-pong cycle, and which steps rescale or zero velocities.
static unsigned long long graph_key(const az_engine* e, int first_step, int nsteps)
{
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { h ^= v; h *= 1099511628211ull; };
    mix((unsigned long long)nsteps); mix((unsigned long long)e->since_rebuild); mix(e->binned ? 1 : 0);
    mix((unsigned long long)e->cur); mix(e->energy_every_step ? 1 : 0); mix(e->fuse_integrate ? 1 : 0);
    mix((unsigned long long)(uintptr_t)e->pos[0]); mix((unsigned long long)(uintptr_t)e->pos[1]); mix((unsigned long long)(uintptr_t)e->pos_alt);
    mix((unsigned long long)(uintptr_t)e->vel[0]); mix((unsigned long long)(uintptr_t)e->vel[1]); mix((unsigned long long)(uintptr_t)e->vel_alt);
    for (int k = 0; k < nsteps; k++) {
        const int step = first_step + k;
        const int bits = (is_eq_step(e, step) ? 1 : 0) | ((e->reset_vels > 0 && step % e->reset_vels == 0) ? 2 : 0);
        if (bits) { mix((unsigned long long)k); mix((unsigned long long)bits); }
    }
    // the pressure ring switches behaviour at step AZ_N_MOM - 1 inside the kernel: no host decision depends on it
    return h;
}

static az_engine::GraphEnd graph_end_state(const az_engine* e)
{
    az_engine::GraphEnd g;
    g.cur = e->cur; g.since_rebuild = e->since_rebuild; g.binned = e->binned;
    for (int k = 0; k < 2; k++) { g.pos[k] = e->pos[k]; g.vel[k] = e->vel[k]; }
    g.pos_alt = e->pos_alt; g.vel_alt = e->vel_alt; g.clear_counts_bytes = e->clear_counts_bytes;
    g.launches = e->launches;
    for (int k = 0; k < AZ_N_TIMERS; k++) g.t_launch[k] = e->t_launch[k];
    return g;
}
#endif

// Tile lists: fit the shared-memory allotments (footprint of the force pass, row columns of the build) to what the rebuilds
// have actually seen.  They are launch parameters, not allocations: the create-time values come from a density estimate
// that has to be generous (a lattice start makes the densest bin twice the average), and shared memory per CTA is what
// bounds the resident warps of both kernels.  Called once behind the first rebuild (one stream synchronisation, never
// again on the step path) and -- grow only -- from az_engine_stats, which synchronises anyway.  12 % + 32 atoms of
// margin over the largest footprint seen; a rebuild that outgrows it raises AZ_FLAG_FOOT.
static int tl_autotune(az_engine* e, bool shrink)
{
    if (!e->use_list || !e->use_tlist || !e->tlStat || e->tl_user_cap) return AZ_OK;
    int st[4] = { 0, 0, 0, 0 };
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    AZ_CUDA(e, cudaMemcpy(st, e->tlStat, sizeof(st), cudaMemcpyDeviceToHost));
    if (st[0] <= 0) return AZ_OK;
    const int bytes_per_atom = 4 * e->tl_stride;
    int fc = ((int)(1.12 * st[0]) + 32 + 31) & ~31;
    const int fc_max = std::min(65535 / e->tl_stride + 1, (200 * 1024) / bytes_per_atom) & ~31;
    if (fc > fc_max) fc = fc_max;
    if (fc > e->tl_foot_alloc) fc = e->tl_foot_alloc; // (never more than the whole system)
    int rs = (st[3] + 8 + 1) & ~1;
    if (rs < 32) rs = 32;
    if (rs > e->nbr_cap) rs = e->nbr_cap;
    if (!shrink) { fc = std::max(fc, e->tl_foot_cap); rs = std::max(rs, e->tl_rows_smem); }
    e->tl_foot_cap = fc; e->tl_smem_pass = (size_t)fc * bytes_per_atom;
    e->tl_rows_smem = rs; e->tl_smem_build = (size_t)rs * AZ_LIST_TPB * sizeof(unsigned short);
    e->tl_tuned = true;
    return AZ_OK;
}

extern "C" int az_engine_step(az_engine* e, int first_step, int nsteps)
{
    if (!e) return AZ_ERR_INVALID;
    if (nsteps < 0) { AZ_SET_ERR(e, "nsteps < 0"); return AZ_ERR_INVALID; }
    AZ_CUDA(e, cudaSetDevice(e->device));
    if (!e->use_graph || nsteps == 0) {
        int rc = run_steps(e, first_step, nsteps);
        if (rc) return rc;
        AZ_CUDA(e, cudaGetLastError());
        if (!e->tl_tuned && e->binned && nsteps > 0) { rc = tl_autotune(e, true); if (rc) return rc; }
        return AZ_OK;
    }
    // graph mode: the kernels take their step from the device (*step0_dev + index inside the call).
    // A long call is replayed in pieces of ~48 steps (a multiple of the rebuild interval): recording and instantiating a graph
    // costs ~80 us per kernel node -- a second for a 2 000-step call of a six-kernel step, more than the steps themselves --
    // while a piece is recorded in 25 ms, once, and replayed for every later piece of every later call.  The trajectory is the
    // one of the caller stepping in such pieces itself (bitwise the same: case_fused_kick_drift..., case_graph_replay...).
    {
        const int R = e->rebuild_every > 1 ? e->rebuild_every : 1;
        const int G = R * ((48 + R - 1) / R);
        if (nsteps > G + G / 2) {
            int done = 0;
            while (nsteps - done > G + G / 2) {
                int rc = az_engine_step(e, first_step + done, G);
                if (rc) return rc;
                done += G;
            }
            return az_engine_step(e, first_step + done, nsteps - done);
        }
    }
    e->call_first = first_step;
    AZ_LAUNCH(az_k_set_int, 1, 1, 0, e->stream, e->step0_dev, first_step);
    e->launches++;
#ifdef AZ_EMU
    int rc = run_steps(e, first_step, nsteps); // no graph API in the host emulator: same kernels, same relative steps
    if (rc) return rc;
#else
    if (e->profile != 0) { // per-section timers need the plain launches
        int rc = run_steps(e, first_step, nsteps);
        if (rc) return rc;
        AZ_CUDA(e, cudaGetLastError());
        return AZ_OK;
    }
    const unsigned long long key = graph_key(e, first_step, nsteps);
    for (auto& g : e->graphs)
        if (g.key == key) { // replay: one launch for the whole call, then put the host-side cycle where the call leaves it
            AZ_CUDA(e, cudaGraphLaunch((cudaGraphExec_t)g.exec, e->stream));
            const long long l0 = g.end.launches;
            e->cur = g.end.cur; e->since_rebuild = g.end.since_rebuild; e->binned = g.end.binned;
            for (int k = 0; k < 2; k++) { e->pos[k] = g.end.pos[k]; e->vel[k] = g.end.vel[k]; }
            e->pos_alt = g.end.pos_alt; e->vel_alt = g.end.vel_alt; e->clear_counts_bytes = g.end.clear_counts_bytes;
            e->launches += l0; // the kernels the graph holds
            for (int k = 0; k < AZ_N_TIMERS; k++) e->t_launch[k] += g.end.t_launch[k];
            e->steps_done += nsteps; e->last_step = first_step + nsteps - 1; e->energy_valid = true;
            e->graph_replays++;
            return AZ_OK;
        }
    // first call of this shape: record the launches instead of running them, then run the recording
    const az_engine::GraphEnd before = graph_end_state(e);
    const long long steps_before = e->steps_done;
    const int last_before = e->last_step;
    const bool evalid_before = e->energy_valid;
    // nothing ran while the launches were being recorded: if the recording cannot be turned into a graph the host-side
    // cycle (buffer roles, rebuild phase, step counters) goes back to where it stood and the steps run as plain launches
    auto restore = [&]() {
        e->cur = before.cur; e->since_rebuild = before.since_rebuild; e->binned = before.binned;
        for (int k = 0; k < 2; k++) { e->pos[k] = before.pos[k]; e->vel[k] = before.vel[k]; }
        e->pos_alt = before.pos_alt; e->vel_alt = before.vel_alt; e->clear_counts_bytes = before.clear_counts_bytes;
        e->launches = before.launches;
        for (int k = 0; k < AZ_N_TIMERS; k++) e->t_launch[k] = before.t_launch[k];
        e->steps_done = steps_before; e->last_step = last_before; e->energy_valid = evalid_before;
        e->fin_rec[0] = e->fin_rec[1] = false; e->fin_last = -1; // (a call starts and ends with the finalize stream joined)
    };
    AZ_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeRelaxed));
    int rc = run_steps(e, first_step, nsteps);
    cudaGraph_t graph = nullptr;
    cudaError_t cerr = cudaStreamEndCapture(e->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (!rc && cerr == cudaSuccess) cerr = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc || cerr != cudaSuccess) {
        cudaGetLastError(); // clear the capture error
        restore();
        if (rc) return rc;  // a launch-sequence error would repeat
        rc = run_steps(e, first_step, nsteps);
        if (rc) return rc;
        AZ_CUDA(e, cudaGetLastError());
        return AZ_OK;
    }
    az_engine::GraphEntry ent;
    ent.key = key; ent.exec = (void*)exec; ent.end = graph_end_state(e);
    ent.end.launches -= before.launches; // deltas
    for (int k = 0; k < AZ_N_TIMERS; k++) ent.end.t_launch[k] -= before.t_launch[k];
    if (e->graphs.size() >= 16) { cudaGraphExecDestroy((cudaGraphExec_t)e->graphs.front().exec); e->graphs.erase(e->graphs.begin()); }
    e->graphs.push_back(ent);
    AZ_CUDA(e, cudaGraphLaunch(exec, e->stream));
#endif
    AZ_CUDA(e, cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_engine_sync(az_engine* e)
{
    if (!e) return AZ_ERR_INVALID;
    AZ_CUDA(e, cudaSetDevice(e->device));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    AZ_CUDA(e, cudaGetLastError());
    return AZ_OK;
}

extern "C" void* az_engine_stream(az_engine* e) { return e ? (void*)e->stream : nullptr; }

static int fetch_scalars(az_engine* e, AzScalars* h)
{
    AZ_CUDA(e, cudaSetDevice(e->device));
    AZ_CUDA(e, cudaMemcpyAsync(h, e->sc, sizeof(AzScalars), cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    AZ_CUDA(e, cudaGetLastError());
    return AZ_OK;
}

static int check_flags(az_engine* e, int flags, int worst_row)
{
    if (flags & AZ_FLAG_SKIN) { AZ_SET_ERR(e, "atoms drifted more than skin/2 between cell-list rebuilds (skin %g): raise 'skin' or lower 'rebuild'", e->P.skin); return AZ_ERR_OVERFLOW; }
    if (flags & AZ_FLAG_TILE) { AZ_SET_ERR(e, "a cell holds more atoms than the pair kernel's shared tile (%d)", e->pair_tile_cap); return AZ_ERR_OVERFLOW; }
    if (flags & AZ_FLAG_NBR) { AZ_SET_ERR(e, "an atom has more than %d neighbours inside the list radius (worst row: %d): raise 'max_neigh'", e->nbr_cap, worst_row); return AZ_ERR_OVERFLOW; }
    if (flags & AZ_FLAG_FOOT) {
        int st[4] = { 0, 0, 0, 0 };
        if (e->tlStat) cudaMemcpy(st, e->tlStat, sizeof(st), cudaMemcpyDeviceToHost);
        AZ_SET_ERR(e, "tile lists: the neighbourhood of a 128-atom tile outgrew its shared-memory allotment (%d atoms in %d runs seen, room for %d atoms / %d runs): raise 'tile_atoms' or set 'list_format 1'",
                   st[0], st[1], e->tl_foot_cap, AZ_TL_RUNS);
        return AZ_ERR_OVERFLOW;
    }
    if (flags & AZ_FLAG_HALO) { AZ_SET_ERR(e, "a halo / migration message or the brick's atom capacity overflowed, or an atom crossed more than one ghost shell in a step"); return AZ_ERR_OVERFLOW; }
    return AZ_OK;
}

extern "C" int az_engine_stats(az_engine* e, az_stats* o)
{
    if (!e || !o) return AZ_ERR_INVALID;
    AzScalars h;
    int rc = fetch_scalars(e, &h);
    if (rc) return rc;
    if (!e->dd) { rc = check_flags(e, h.overflow, h.dbg_count); if (rc) return rc; }
    if (e->tl_tuned && !e->use_graph) { rc = tl_autotune(e, false); if (rc) return rc; } // a system that densifies: grow the allotments in time
    memset(o, 0, sizeof(*o));
    o->step = e->last_step;
    o->n_atoms_total = e->n_global;
    o->engTot = h.engTot; o->engKin = h.engKin_f; o->engVdW = (float)h.engVdW; o->engCoul1 = (float)h.engCoul1;
    o->engCoul2 = (float)h.engCoul2; o->engCoul3 = (float)h.engCoul3; o->engCoul2_d = h.engCoul2; o->engTemp = (float)h.engTemp; o->engElecField = (float)h.engElecField; o->engPot = h.engPot;
    for (int a = 0; a < 3; a++) { o->posMom[a] = (float)h.posMom[a]; o->negMom[a] = (float)h.negMom[a]; }
    o->pressure = h.pressure;
    o->x_first = h.x_first;
    for (int t = 0; t < e->P.nspec; t++)
        for (int a = 0; a < 3; a++) { o->specAcBoxPos[t][a] = h.acPos[t][a]; o->specAcBoxNeg[t][a] = h.acNeg[t][a]; }
    o->engKin_d = h.engKin; o->engVdW_d = h.engVdW; o->engCoul1_d = h.engCoul1; o->engTemp_d = h.engTemp;
    o->engBond = (float)h.engBond; o->engAngle = (float)h.engAngle; o->engBond_d = h.engBond; o->engAngle_d = h.engAngle;
    o->n_pairs = e->dd ? (long long)h.nVisits : (long long)h.nPairs;
    if (e->dd) { // collective: every rank reduces, then every rank sees the OR of the flags
        int flags = h.overflow;
        rc = dd_reduce_stats(e, o, &flags);
        if (rc) return rc;
        rc = check_flags(e, flags, h.dbg_count);
        if (rc) return rc;
    }
    return AZ_OK;
}

// wrap one coordinate into [0, L) the way put_periodic does (used for downloads between rebuilds)
static float wrap_host(float x, float L, float revL)
{
    int n = (int)floor((double)x * (double)revL);
    x -= (float)n * L;
    if (x >= L) x = 0.f;
    return x;
}

extern "C" int az_engine_download(az_engine* e, az_atoms_out* o, int order)
{
    if (!e || !o) return AZ_ERR_INVALID;
    AZ_CUDA(e, cudaSetDevice(e->device));
    int n = e->n;
    if (e->dd) { int rc = dd_owned_count(e, &n); if (rc) return rc; if (order == AZ_ORDER_ORIGINAL) order = AZ_ORDER_SORTED; } // a brick hands back the atoms it owns, with their ids
    if (o->n_atoms < n) { AZ_SET_ERR(e, "output capacity %d < %d atoms", o->n_atoms, n); return AZ_ERR_INVALID; }
    const int c = e->cur;
    std::vector<float4> hp(n), hv(n);
    std::vector<int> hid(n);
    std::vector<float2> hth(e->use_therm ? n : 0);
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    AZ_CUDA(e, cudaMemcpy(hp.data(), e->pos[c], sizeof(float4) * n, cudaMemcpyDeviceToHost));
    AZ_CUDA(e, cudaMemcpy(hv.data(), e->vel[c], sizeof(float4) * n, cudaMemcpyDeviceToHost));
    AZ_CUDA(e, cudaMemcpy(hid.data(), e->id[c], sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (e->use_therm) AZ_CUDA(e, cudaMemcpy(hth.data(), e->therm[c], sizeof(float2) * n, cudaMemcpyDeviceToHost));
    const bool wrap = e->rebuild_every > 1;
    for (int i = 0; i < n; i++) {
        int k = (order == AZ_ORDER_ORIGINAL) ? hid[i] : i;
        float x = hp[i].x, y = hp[i].y, z = hp[i].z;
        if (wrap) { x = wrap_host(x, e->P.L[0], e->P.revL[0]); y = wrap_host(y, e->P.L[1], e->P.revL[1]); z = wrap_host(z, e->P.L[2], e->P.revL[2]); }
        if (o->xs) o->xs[k] = x;
        if (o->ys) o->ys[k] = y;
        if (o->zs) o->zs[k] = z;
        if (o->vxs) o->vxs[k] = hv[i].x;
        if (o->vys) o->vys[k] = hv[i].y;
        if (o->vzs) o->vzs[k] = hv[i].z;
        if (o->types) { int t; memcpy(&t, &hp[i].w, 4); o->types[k] = t; }
        if (o->ids) o->ids[k] = hid[i];
        if (o->engs) o->engs[k] = e->use_therm ? hth[i].x : 0.f;
        if (o->radii) o->radii[k] = e->use_therm ? hth[i].y : 0.f;
    }
    o->n_written = n;
    return AZ_OK;
}

extern "C" int az_engine_debug_dump(az_engine* e, az_dump* o)
{
    if (!e || !o) return AZ_ERR_INVALID;
    int n = e->n;
    const int nc = e->P.g.ncell_local, c = e->cur;
    if (e->dd) { int rc = dd_owned_count(e, &n); if (rc) return rc; }
    if (o->n_atoms < n || o->n_cells < nc) { AZ_SET_ERR(e, "dump capacity too small (%d atoms, %d cells needed)", n, nc); return AZ_ERR_INVALID; }
    AZ_CUDA(e, cudaSetDevice(e->device));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    o->n_atoms = n; o->n_cells = nc;
    for (int a = 0; a < 3; a++) { o->cnum[a] = e->P.g.n[a]; o->csize[a] = e->P.g.csize[a]; o->crev[a] = e->P.g.crev[a]; }
    if (o->cell_indexes) AZ_CUDA(e, cudaMemcpy(o->cell_indexes, e->cellSorted, sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (o->first_atom_in_cell) AZ_CUDA(e, cudaMemcpy(o->first_atom_in_cell, e->first, sizeof(int) * nc, cudaMemcpyDeviceToHost));
    if (o->n_at_in_cell) AZ_CUDA(e, cudaMemcpy(o->n_at_in_cell, e->cellCount, sizeof(int) * nc, cudaMemcpyDeviceToHost));
    if (o->ids) AZ_CUDA(e, cudaMemcpy(o->ids, e->id[c], sizeof(int) * n, cudaMemcpyDeviceToHost));
    std::vector<float4> h(n);
    struct { float4* src; float *x, *y, *z; } sets[3] = { { e->frc, o->fx, o->fy, o->fz }, { e->pos[c], o->x, o->y, o->z }, { e->vel[c], o->vx, o->vy, o->vz } };
    for (auto& s : sets) {
        if (!s.x && !s.y && !s.z) continue;
        AZ_CUDA(e, cudaMemcpy(h.data(), s.src, sizeof(float4) * n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) { if (s.x) s.x[i] = h[i].x; if (s.y) s.y[i] = h[i].y; if (s.z) s.z[i] = h[i].z; }
    }
    return AZ_OK;
}

// H2D staging shared by the single-GPU and the brick upload: the six arrays (each `len` doubles) are copied
// on a copy stream into one of two staging buffers; the engine's stream waits for the copy.  With
// async == false the call returns after the data has left the caller's arrays (borrowed for the call only).
static int stage_upload(az_engine* e, const void* const src[6], int len, size_t esize, void* dst[6], int* slot_out)
{
    if (!e->copy_stream) {
        AZ_CUDA(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            AZ_CUDA(e, cudaEventCreateWithFlags(&e->ev_copied[k], cudaEventDisableTiming));
            AZ_CUDA(e, cudaEventCreateWithFlags(&e->ev_unpacked[k], cudaEventDisableTiming));
            AZ_ALLOC(e, e->stage2[k], (size_t)6 * len);
            AZ_CUDA(e, cudaEventRecord(e->ev_unpacked[k], e->stream));
        }
    }
    const int b = e->up_slot;
    e->up_slot ^= 1;
    AZ_CUDA(e, cudaStreamWaitEvent(e->copy_stream, e->ev_unpacked[b], 0)); // the kernel that last read this buffer is done
    for (int k = 0; k < 6; k++) { // the staging buffers are sized for doubles; floats use the front of each slice
        dst[k] = src[k] ? (void*)((char*)e->stage2[b] + (size_t)k * len * sizeof(double)) : nullptr;
        if (src[k]) AZ_CUDA(e, cudaMemcpyAsync(dst[k], src[k], esize * len, cudaMemcpyHostToDevice, e->copy_stream));
    }
    AZ_CUDA(e, cudaEventRecord(e->ev_copied[b], e->copy_stream));
    AZ_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_copied[b], 0));
    *slot_out = b;
    return AZ_OK;
}

template <class T>
static int upload_impl(az_engine* e, const T* xs, const T* ys, const T* zs, const T* vxs, const T* vys, const T* vzs, bool async)
{
    if (!e || !xs || !ys || !zs) return AZ_ERR_INVALID;
    AZ_CUDA(e, cudaSetDevice(e->device));
    const void* src[6] = { xs, ys, zs, vxs, vys, vzs };
    void* dst[6];
    int b = 0;
    if (e->dd) {
        if (sizeof(T) != sizeof(double)) { AZ_SET_ERR(e, "az_engine_upload_f32_async is single-GPU only"); return AZ_ERR_UNSUPPORTED; }
        int rc = dd_upload(e, (const double* const*)src, (double**)dst, &b);
        if (rc) return rc;
    } else {
        const int n = e->n, c = e->cur;
        int rc = stage_upload(e, src, n, sizeof(T), dst, &b);
        if (rc) return rc;
        AZ_LAUNCH(az_k_upload<T>, e->n_atom_blocks, AZ_TPB, 0, e->stream, n, e->id[c], (const T*)dst[0], (const T*)dst[1], (const T*)dst[2], (const T*)dst[3],
                  (const T*)dst[4], (const T*)dst[5], e->pos[c], e->vel[c], e->frc);
        e->launches++;
    }
    AZ_CUDA(e, cudaEventRecord(e->ev_unpacked[b], e->stream));
    if (!async) { // the caller's arrays are borrowed for the call only
        AZ_CUDA(e, cudaStreamSynchronize(e->stream));
        AZ_CUDA(e, cudaGetLastError());
    }
    e->binned = false;
    return AZ_OK;
}

extern "C" int az_engine_upload(az_engine* e, const double* xs, const double* ys, const double* zs,
                                const double* vxs, const double* vys, const double* vzs)
{
    return upload_impl(e, xs, ys, zs, vxs, vys, vzs, false);
}

extern "C" int az_engine_upload_async(az_engine* e, const double* xs, const double* ys, const double* zs,
                                      const double* vxs, const double* vys, const double* vzs)
{
    return upload_impl(e, xs, ys, zs, vxs, vys, vzs, true);
}

// float inputs: half the PCIe traffic of the reference's double arrays (the engine computes in fp32 anyway: src/cuInit.cu:775-779
// narrows every coordinate on upload)
extern "C" int az_engine_upload_f32_async(az_engine* e, const float* xs, const float* ys, const float* zs,
                                          const float* vxs, const float* vys, const float* vzs)
{
    return upload_impl(e, xs, ys, zs, vxs, vys, vzs, true);
}

// positions (and velocities) of all atoms back to the host in atoms.xyz order, float, one device-side gather + one D2H copy per
// array; the host arrays should be pinned (az_host_alloc).  Synchronous at return.  Single GPU.
extern "C" int az_engine_download_f32(az_engine* e, float* xs, float* ys, float* zs, float* vxs, float* vys, float* vzs)
{
    if (!e || !xs || !ys || !zs) return AZ_ERR_INVALID;
    if (e->dd) { AZ_SET_ERR(e, "az_engine_download_f32 is single-GPU only (a brick hands its atoms back through az_engine_download)"); return AZ_ERR_UNSUPPORTED; }
    AZ_CUDA(e, cudaSetDevice(e->device));
    const int n = e->n, c = e->cur;
    const int with_vel = (vxs && vys && vzs) ? 1 : 0;
    if (!e->dl_buf) AZ_ALLOC(e, e->dl_buf, (size_t)6 * n);
    AZ_LAUNCH(az_k_download_f32, e->n_atom_blocks, AZ_TPB, 0, e->stream, e->P, n, e->rebuild_every > 1 ? 1 : 0, e->id[c], e->pos[c], e->vel[c], e->dl_buf, with_vel);
    e->launches++;
    float* host[6] = { xs, ys, zs, vxs, vys, vzs };
    for (int k = 0; k < (with_vel ? 6 : 3); k++)
        AZ_CUDA(e, cudaMemcpyAsync(host[k], e->dl_buf + (size_t)k * n, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    AZ_CUDA(e, cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_engine_step_timed(az_engine* e, int first_step, int nsteps, float* ms)
{
    if (!e || !ms) return AZ_ERR_INVALID;
    AZ_CUDA(e, cudaSetDevice(e->device));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    AZ_CUDA(e, cudaEventRecord(e->ev_step[0], e->stream));
    int rc = az_engine_step(e, first_step, nsteps);
    if (rc) return rc;
    AZ_CUDA(e, cudaEventRecord(e->ev_step[1], e->stream));
    AZ_CUDA(e, cudaEventSynchronize(e->ev_step[1]));
    AZ_CUDA(e, cudaEventElapsedTime(ms, e->ev_step[0], e->ev_step[1]));
    AZ_CUDA(e, cudaGetLastError());
    return AZ_OK;
}

extern "C" void* az_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

extern "C" void az_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int az_measure_fp32_peak(int device, double* tflops, double* sm_mhz_effective)
{
    if (!tflops) return AZ_ERR_INVALID;
    if (az_device_count() <= device || device < 0) return AZ_ERR_NO_DEVICE;
    az_engine* e = nullptr;
    AZ_CUDA(e, cudaSetDevice(device));
    cudaDeviceProp prop;
    AZ_CUDA(e, cudaGetDeviceProperties(&prop, device));
    float* out = nullptr;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    AZ_CUDA(e, cudaMalloc(&out, sizeof(float) * blocks * threads));
    cudaStream_t st;
    AZ_CUDA(e, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t a, b;
    AZ_CUDA(e, cudaEventCreate(&a)); AZ_CUDA(e, cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        AZ_CUDA(e, cudaEventRecord(a, st));
        AZ_LAUNCH(az_k_ffma_probe, blocks, threads, 0, st, out, iters, 0.999f, 0.001f);
        AZ_CUDA(e, cudaEventRecord(b, st));
        AZ_CUDA(e, cudaEventSynchronize(b));
        float ms = 0.f;
        AZ_CUDA(e, cudaEventElapsedTime(&ms, a, b));
        double fl = 2.0 * 64.0 * (double)iters * (double)blocks * threads; // 64 FMAs per iteration per thread
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    if (sm_mhz_effective) *sm_mhz_effective = best * 1e12 / (2.0 * 128.0 * prop.multiProcessorCount) / 1e6;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaStreamDestroy(st); cudaFree(out);
    return AZ_OK;
}

// ---------------------------------------------------------------------------------------- RDF
extern "C" int az_engine_rdf_begin(az_engine* e, double r_max, double dr)
{
    if (!e) return AZ_ERR_INVALID;
    if (!(r_max > 0.0) || !(dr > 0.0)) { AZ_SET_ERR(e, "rdf: r_max and dr must be positive"); return AZ_ERR_INVALID; }
    AZ_CUDA(e, cudaSetDevice(e->device));
    const AzGrid& g = e->P.g;
    const double idrdf = 1.0 / dr;                                   // Sim::idRDF, src/rdf.cpp:22
    double minR = r_max;
    if (minR > (double)e->P.L[0]) minR = (double)e->P.L[0];          // init_rdf, src/rdf.cpp:38-45
    const int nb = (int)(minR * idrdf);
    const int np = e->P.nspec * (e->P.nspec + 1) / 2;
    if (nb <= 0) { AZ_SET_ERR(e, "rdf: dr is larger than the radius"); return AZ_ERR_INVALID; }
    if ((size_t)nb * np * sizeof(unsigned) > 200 * 1024) { AZ_SET_ERR(e, "rdf: %d bins x %d pairs do not fit the shared-memory histogram", nb, np); return AZ_ERR_UNSUPPORTED; }
    const float reach = (float)r_max + e->P.skin; // cells may be stale by up to skin / 2 per atom
    for (int a = 0; a < 3; a++) {
        int s1 = 1;
        while ((float)s1 * g.csize[a] <= reach && s1 < g.n[a]) s1++;
        if (g.n[a] < 2 * s1 + 1) { AZ_SET_ERR(e, "rdf: radius %g needs %d cell layers on axis %d, the grid has %d cells (every pair must be seen through one image only)", r_max, s1, a, g.n[a]); return AZ_ERR_UNSUPPORTED; }
        if (e->dd && g.dec[a] && s1 > g.sten[a]) { AZ_SET_ERR(e, "rdf: radius %g reaches %d cell layers on axis %d, the ghost shell of a brick has %d", r_max, s1, a, g.sten[a]); return AZ_ERR_UNSUPPORTED; }
        e->rdf_sten[a] = s1;
    }
    cudaFree(e->rdf_hist); e->rdf_hist = nullptr;
    AZ_ALLOC(e, e->rdf_hist, (size_t)nb * np);
    e->rdf_nbins = nb; e->rdf_npair = np; e->rdf_samples = 0;
    e->rdf_r2max = (float)(r_max * r_max); e->rdf_idrdf = (float)idrdf; e->rdf_r2stencil = reach * reach;
#ifndef AZ_EMU
    if ((size_t)nb * np * sizeof(unsigned) > 48 * 1024)
        AZ_CUDA(e, cudaFuncSetAttribute(az_k_rdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)nb * np * sizeof(unsigned))));
#endif
    return AZ_OK;
}

extern "C" int az_engine_rdf_sample(az_engine* e)
{
    if (!e) return AZ_ERR_INVALID;
    if (!e->rdf_hist) { AZ_SET_ERR(e, "rdf: call az_engine_rdf_begin first"); return AZ_ERR_INVALID; }
    if (!e->binned) { AZ_SET_ERR(e, "rdf: the atoms are not binned yet (sample after a step)"); return AZ_ERR_INVALID; }
    AZ_CUDA(e, cudaSetDevice(e->device));
    { int rc = dd_sync_ghosts(e); if (rc) return rc; } // fused halo: the last ghosts sit in the receive staging only
    AzRdfArgs A;
    memset(&A, 0, sizeof(A));
    const int c = e->cur;
    A.pos = e->pos[c]; A.cellOfSorted = e->cellSorted; A.first = e->first; A.count = e->cellCount; A.hist = e->rdf_hist;
    for (int a = 0; a < 3; a++) A.sten[a] = e->rdf_sten[a];
    A.nbins = e->rdf_nbins; A.npair = e->rdf_npair; A.r2max = e->rdf_r2max; A.idrdf = e->rdf_idrdf; A.r2stencil = e->rdf_r2stencil;
    A.id = e->dd ? e->id[c] : nullptr;
    const int n_slots = e->dd ? e->n_atom_blocks * AZ_TPB : e->n; // bricks: the owned count lives on the device (az_n_owned)
    const int blocks = (n_slots + (AZ_RDF_TPB / 32) - 1) / (AZ_RDF_TPB / 32);
    AZ_LAUNCH(az_k_rdf, blocks, AZ_RDF_TPB, (size_t)e->rdf_nbins * e->rdf_npair * sizeof(unsigned), e->stream, e->P, A);
    e->launches++;
    e->rdf_samples++;
    AZ_CUDA(e, cudaGetLastError());
    return AZ_OK;
}

extern "C" int az_engine_rdf_read(az_engine* e, unsigned long long* counts, int capacity, int* n_bins, int* n_pairs, int* n_samples)
{
    if (!e) return AZ_ERR_INVALID;
    if (!e->rdf_hist) { AZ_SET_ERR(e, "rdf: call az_engine_rdf_begin first"); return AZ_ERR_INVALID; }
    if (n_bins) *n_bins = e->rdf_nbins;
    if (n_pairs) *n_pairs = e->rdf_npair;
    if (n_samples) *n_samples = e->rdf_samples;
    if (counts) {
        if (capacity < e->rdf_nbins * e->rdf_npair) { AZ_SET_ERR(e, "rdf: capacity %d < %d", capacity, e->rdf_nbins * e->rdf_npair); return AZ_ERR_INVALID; }
        AZ_CUDA(e, cudaSetDevice(e->device));
        AZ_CUDA(e, cudaStreamSynchronize(e->stream));
        AZ_CUDA(e, cudaMemcpyAsync(counts, e->rdf_hist, sizeof(unsigned long long) * e->rdf_nbins * e->rdf_npair, cudaMemcpyDeviceToHost, e->stream));
        AZ_CUDA(e, cudaStreamSynchronize(e->stream)); // like fetch_scalars: no legacy-stream copy while a peer may sit in a collective
        if (e->dd) { // collective: every rank gets the histogram of the whole box (integers below 2^53 are exact in the double all-reduce)
            const int nh = e->rdf_nbins * e->rdf_npair;
            std::vector<double> v(128);
            for (int k0 = 0; k0 < nh; k0 += 128) {
                const int m = std::min(128, nh - k0);
                for (int k = 0; k < m; k++) v[k] = (double)counts[k0 + k];
                int rc = dd_allreduce_host(e, v.data(), m, false);
                if (rc) return rc;
                for (int k = 0; k < m; k++) counts[k0 + k] = (unsigned long long)(v[k] + 0.5);
            }
        }
    }
    return AZ_OK;
}

extern "C" int az_engine_profile(az_engine* e, int enable)
{
    if (!e) return AZ_ERR_INVALID;
    e->profile = enable;
    for (auto& p : e->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    e->pending.clear();
    for (int k = 0; k < AZ_N_TIMERS; k++) { e->t_ms[k] = 0.0; e->t_launch[k] = 0; }
    return AZ_OK;
}

extern "C" int az_engine_timers(az_engine* e, double ms_sum[AZ_N_TIMERS], long long launches[AZ_N_TIMERS])
{
    if (!e) return AZ_ERR_INVALID;
    if (!e->pending.empty()) {
        AZ_CUDA(e, cudaStreamSynchronize(e->stream));
        for (auto& p : e->pending) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, p.a, p.b);
            e->t_ms[p.k] += ms;
            cudaEventDestroy(p.a); cudaEventDestroy(p.b);
        }
        e->pending.clear();
    }
    for (int k = 0; k < AZ_N_TIMERS; k++) { if (ms_sum) ms_sum[k] = e->t_ms[k]; if (launches) launches[k] = e->t_launch[k]; }
    return AZ_OK;
}

extern "C" long long az_engine_launch_count(az_engine* e) { return e ? e->launches : 0; }

extern "C" int az_engine_set_option(az_engine* e, const char* key, double value)
{
    if (!e || !key) return AZ_ERR_INVALID;
    if (!strcmp(key, "energy_every_step")) { e->energy_every_step = value != 0.0; return AZ_OK; }
    if (!strcmp(key, "halo_overlap")) { dd_set_overlap(e, value != 0.0); return AZ_OK; }
    if (!strcmp(key, "fuse_integrate")) { e->fuse_integrate = value != 0.0; return AZ_OK; }
    if (!strcmp(key, "list_dbg")) { e->list_dbg = (int)value; return AZ_OK; }
    if (!strcmp(key, "list_fast")) { if (value == 0.0) e->list_fast = false; return AZ_OK; } // only ever switched off (tests)
    if (!strcmp(key, "finalize_overlap")) { int rc = fin_join(e); if (rc) return rc; e->fin_overlap = value != 0.0; return AZ_OK; } // A/B
    if (!strcmp(key, "pdl")) { e->pdl = value != 0.0 && !e->dd; return AZ_OK; } // A/B of the programmatic dependent launches
    if (!strcmp(key, "stencil_table")) { if (value == 0.0) { cudaFree(e->stenTab); e->stenTab = nullptr; } return AZ_OK; } // only ever switched off (A/B, tests)
    if (!strcmp(key, "warp_prune")) { e->warp_prune = value != 0.0; return AZ_OK; }         // A/B of the per-atom cell pruning in the warp sweep
    if (!strcmp(key, "tile_list")) { // A/B of the two list formats; before the first step (the next rebuild writes the chosen format)
        if (value != 0.0 && (!e->nbr16 || (long long)(e->tl_foot_cap - 1) * e->tl_stride > 65535 || e->tl_smem_pass > 200 * 1024)) { AZ_SET_ERR(e, "tile lists are not available for this engine"); return AZ_ERR_UNSUPPORTED; }
        e->use_tlist = value != 0.0; e->binned = false; return AZ_OK;
    }
    if (!strcmp(key, "list_staged")) { e->list_staged = value != 0.0; return AZ_OK; }       // A/B of the two FAST list builds
    AZ_SET_ERR(e, "unknown option '%s'", key);
    return AZ_ERR_INVALID;
}

extern "C" int az_engine_info(az_engine* e, az_engine_info_t* o)
{
    if (!e || !o) return AZ_ERR_INVALID;
    memset(o, 0, sizeof(*o));
    const AzGrid& g = e->P.g;
    for (int a = 0; a < 3; a++) { o->cnum[a] = g.n[a]; o->stencil[a] = g.sten[a]; o->csize[a] = g.csize[a]; o->crev[a] = g.crev[a]; }
    o->n_cells = g.ncell_local; o->n_atoms = e->n;
    o->pair_blocks = e->pair_nblocks; o->pair_threads = AZ_PAIR_TPB; o->pair_smem_bytes = (int)e->pair_smem;
    o->pair_zc = e->pair_zc; o->pair_vk = e->pair_vk; o->pair_ck = e->pair_ck;
    o->n_sm = e->n_sm; o->device_bytes = (long long)e->dev_bytes;
    o->list_rows = e->nbr_cap;
    o->tile_list = (e->use_list && e->use_tlist) ? 1 : 0; o->tile_atoms = e->tl_foot_cap; o->tile_rows_smem = e->tl_rows_smem;
    if (e->tlStat) {
        int st[4] = { 0, 0, 0, 0 };
        AZ_CUDA(e, cudaSetDevice(e->device));
        AZ_CUDA(e, cudaStreamSynchronize(e->stream));
        AZ_CUDA(e, cudaMemcpy(st, e->tlStat, sizeof(st), cudaMemcpyDeviceToHost));
        o->tile_atoms_seen = st[0]; o->tile_runs_seen = st[1];
    }
    o->capacity = e->dd ? e->P.n_local : e->n; o->n_ranks = e->n_ranks;
    o->graph_replays = e->graph_replays; o->graphs_cached = (int)e->graphs.size();
    o->peer_halo = dd_peer_halo(e) ? (e->dd->direct ? 3 : (e->dd->fused ? 2 : 1)) : 0;
    if (e->sc) {
        AZ_CUDA(e, cudaSetDevice(e->device));
        AZ_CUDA(e, cudaStreamSynchronize(e->stream));
        AzScalars h;
        AZ_CUDA(e, cudaMemcpy(&h, e->sc, sizeof(h), cudaMemcpyDeviceToHost));
        o->flags = h.overflow;
        if (e->use_list && e->binned) {
            std::vector<int> cnt(e->n);
            AZ_CUDA(e, cudaMemcpy(cnt.data(), e->nbrCount, sizeof(int) * e->n, cudaMemcpyDeviceToHost));
            int mx = 0;
            for (int v : cnt) { int t = (v & ((1 << AZ_NBR_CNT_BITS) - 1)) + (v >> AZ_NBR_CNT_BITS); if (t > mx) mx = t; }
            o->max_neigh_seen = mx;
        }
    }
    return AZ_OK;
}
