// az_engine.h -- internal: engine state shared by az_engine.cu and the pair-kernel instantiation units.
#pragma once
#include <string>
#include <vector>
#include <string.h>
#include "az_pair.cuh"
#include "az_ewald.cuh"

static inline float __int_as_float_host(int i) { float f; memcpy(&f, &i, 4); return f; }

struct AzDDHost;

struct az_engine {
    std::string err;
    AzDDHost* dd = nullptr; // spatial domain decomposition (n_ranks > 1), az_engine_dd.cuh
    int device = 0, n_sm = 0;
    int n = 0, n_global = 0, cap = 0, cap_total_dd = 0;
    int rank = 0, n_ranks = 1;
    AzParams P;
    float revArea[3] = { 0, 0, 0 };
    int n_eq = 0, eq_freq = 0, reset_vels = 0;
    int rebuild_every = 1, since_rebuild = 0;
    bool binned = false, use_therm = false, energy_every_step = false, energy_valid = false;
    int profile = 0; // 0 off, 1 synchronous sections, 2 recorded events resolved later
    struct PendingEv { cudaEvent_t a, b; int k; };
    std::vector<PendingEv> pending;
    int last_step = -1;
    int call_rel = 0; // index of the step in flight inside the current az_engine_step call
    long long steps_done = 0;
    bool fuse_integrate = true; // second half kick of step n and kick + drift of step n+1 in one pass
    // az_cuda_opts.use_graph: az_engine_step(first, n) is captured once per call shape as a CUDA graph and replayed (the
    // latency-bound decks are launch-gap bound); kernels then read the step from *step0_dev + their relative index
    bool use_graph = false;
    int* step0_dev = nullptr;
    int call_first = 0;
    struct GraphEnd { int cur, since_rebuild; bool binned; float4 *pos[2], *vel[2], *pos_alt, *vel_alt; size_t clear_counts_bytes; long long launches, t_launch[AZ_N_TIMERS]; };
    struct GraphEntry { unsigned long long key; void* exec; GraphEnd end; };
    std::vector<GraphEntry> graphs;
    long long graph_replays = 0;
    // pair kernel configuration
    int pair_vk = 0, pair_ck = 0;
    bool pair_single = false, pair_radii = false;
    int pair_zc = 1, pair_nruns = 1, pair_nblocks = 1, pair_tile_cap = 512;
    size_t pair_smem = 0;
    // Verlet-list path (rebuild_every > 1)
    bool use_list = false;
    int list_nblocks = 0, nbr_cap = 0, nbr_stride = 0;
    int warp_nblocks = 0;     // az_k_pair_warp: one warp per owned atom
    bool warp_prune = true;   // az_k_pair_warp: drop the stencil cells an atom cannot reach from where it stands
    bool pdl = false;         // small systems: scan / sort / sweep are programmatic dependent launches (az_device.cuh)
    int2* stenTab = nullptr;  // az_k_pair_warp: tabulated stencil of every owned cell
    int stenS = 0;
    int thread_nblocks = 0;   // az_k_pair_thread: one thread per owned atom
    bool use_thread = false;  // ... when every cell edge is at least twice the list radius (dilute / coarse-celled systems)
    bool use_warp = false;    // reference schedule (no list): warp-per-atom sweep instead of the shared-memory tile sweep
    int list_dbg = 0;
    float r2_safe = 3.4e38f; // packed Lennard-Jones list pass: below this r2 the reference's force_autocap clamp can act
    bool list_staged = false; // az_k_build_list_staged (rows staged in shared memory, lockstep flush): measured slower, kept for the A/B (az_pair.cuh)
    bool list_fast = false; // every axis has more than 2*sten cells: az_k_build_list prunes by atom position and walks z-intervals
    int *nbr = nullptr, *nbrCount = nullptr;
    // tile lists (az_tlist.cuh): the list format of the Verlet-list path unless list_format 1 asks for the global rows
    bool use_tlist = false;
    unsigned* nbr16 = nullptr;
    int *tileHdr = nullptr, *tlStat = nullptr;
    int4* tileRuns = nullptr;
    bool tl_user_cap = false; // az_cuda_opts.tile_atoms was given: no fitting
    int tl_foot_alloc = 0;    // the whole system, rounded up: no footprint is larger
    bool tl_tuned = false; // the shared-memory allotments have been fitted to what the first rebuild saw
    int tl_foot_cap = 0, tl_rows_smem = 0, tl_stride = 3;
    size_t tl_smem_pass = 0, tl_smem_build = 0;
    std::vector<AzVdw> h_table;
    // device state
    cudaStream_t stream = nullptr;
    // az_k_finalize beside the next step's list pass (az_engine.cu one_step)
    cudaStream_t fin_stream = nullptr;
    cudaEvent_t ev_fin_fork = nullptr, ev_fin[2] = { nullptr, nullptr };
    bool fin_rec[2] = { false, false }, fin_overlap = true;
    int fin_last = -1;
    cudaStream_t pair_stream = nullptr; // fused halo: where the ghost-touching blocks of the list pass are launched (the halo stream)
    cudaEvent_t ev[2 * AZ_N_TIMERS] = {};
    cudaEvent_t ev_step[2] = {};
    int cur = 0;
    float4 *pos[2] = { nullptr, nullptr }, *vel[2] = { nullptr, nullptr }, *frc = nullptr;
    float4 *pos_alt = nullptr, *vel_alt = nullptr; // targets of the integrator that rides on the list kernel (swapped with pos/vel[cur])
    bool fuse_pair = false;     // the list kernel also integrates (no thermostat, no per-atom radii, no external field)
    int n_ke_blocks = 0;        // capacity of partKE / partU
    size_t clear_counts_bytes = 0; // != 0: the list force pass must be preceded by clear_clist (its rider bins the next step)
    int* id[2] = { nullptr, nullptr };
    float2* therm[2] = { nullptr, nullptr };
    int *cellOf = nullptr, *slotOf = nullptr, *cellSorted = nullptr, *keyTmp = nullptr;
    int *cellCount = nullptr, *first = nullptr, *tileSum = nullptr;
    int n_scan_tiles = 0, n_atom_blocks = 0;
    double *partKE = nullptr, *partU = nullptr, *partV = nullptr, *partC = nullptr;
    unsigned long long* partN = nullptr;
    int n_pair_partials = 0;
    // reciprocal-space Ewald (az_ewald.cuh); ew.nk == 0: off
    AzEwald ew = {};
    float* ew_ak = nullptr; int4 *ew_pairs = nullptr, *ew_runs = nullptr; float2 *ew_part = nullptr, *ew_T = nullptr; double2* ew_Sd = nullptr;
    int ew_blocks = 0;
    size_t ew_smem_s = 0, ew_smem_f = 0;
    // az_engine_upload staging: two buffers of 6*n doubles filled on a copy stream, so that the H2D copy of the
    // next upload overlaps the kernels of the current step (az_engine_upload_async)
    double* stage2[2] = { nullptr, nullptr };
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {}, ev_unpacked[2] = {};
    int up_slot = 0;
    float* dl_buf = nullptr; // az_engine_download_f32: 6 * n floats in atoms.xyz order
    AzScalars* sc = nullptr;
    AzVdw* table = nullptr;
    float *photons = nullptr, *uv = nullptr;
    // radial distribution function (az_engine_rdf_*)
    unsigned long long* rdf_hist = nullptr;
    int rdf_nbins = 0, rdf_npair = 0, rdf_samples = 0, rdf_sten[3] = { 0, 0, 0 };
    float rdf_r2max = 0.f, rdf_idrdf = 0.f, rdf_r2stencil = 0.f;
    // constant bonds and valence angles (az_k_bonded); bonded == false: nothing of this is allocated
    bool bonded = false;
    int* whereId = nullptr;          // persistent id -> current place in the sorted streams (written by az_k_sort_scatter)
    int *bStart = nullptr, *aStart = nullptr;
    int2* bEnt = nullptr; int4* aEnt = nullptr;
    AzBondType* bTypes = nullptr; AzAngleType* aTypes = nullptr;
    double *partBond = nullptr, *partAngle = nullptr;
    size_t dev_bytes = 0;
    // accounting
    long long launches = 0;
    double t_ms[AZ_N_TIMERS] = {};
    long long t_launch[AZ_N_TIMERS] = {};
};

int az_engine_init(az_engine* e, const az_system_desc* d, const az_cuda_opts* o);
// launches the pair kernel instantiation chosen at create time (az_pair_inst_*.cu)
int az_launch_pair(az_engine* e, const AzPairArgs& A, bool energy, int mode);
int az_launch_pair_lj(az_engine* e, const AzPairArgs& A, bool energy, int mode);
int az_launch_pair_buck(az_engine* e, const AzPairArgs& A, bool energy, int mode);
int az_launch_pair_surk(az_engine* e, const AzPairArgs& A, bool energy, int mode);
int az_launch_pair_mixed(az_engine* e, const AzPairArgs& A, bool energy, int mode);
