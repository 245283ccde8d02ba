/*
 * aztot_b200.h -- C ABI of the B200-native azTotMD hot-path engine.
 *
 * This is the drop-in boundary for the per-step MD hot path of raadyn/aztotmd
 * (bin + reorder -> short-range pair forces -> velocity Verlet + radiative
 * thermostat + kinetic-energy reduction).  The reference has no FFI; its internal
 * seam is the host->device translation layer:
 *
 *   az_engine_create    replaces  init_cudaMD          src/cuInit.cu:756-1199
 *                                 define_global_func   src/main.cu:196-226
 *                                 define_vdw_func      src/cuVdW.cu:260-302
 *   az_engine_step      replaces  the loop body        src/main.cu:281-383
 *                                 iter_fastCellList    src/cuPairs.cu:2519-2568
 *                                 apply_tstat          src/cuTemp.cu:775-803
 *   az_engine_stats     replaces  write_stat/copy_stat src/cuStat.cu:17-71
 *                                 calc_quantities      src/main.cu:121-162
 *   az_engine_download  replaces  md_to_host           src/cuInit.cu:1212-1261
 *                                 out_thermalchar (device part)  src/main.cu:51-66
 *   az_engine_debug_dump  (new; parity tests)  cellIndexes/firstAtomInCell/nAtInCell
 *                                              of src/cuStruct.h:342-349 + forces
 *   az_engine_destroy   replaces  free_device_md       src/cuInit.cu:1263-1373
 *
 * Plain C: POD structs, plain pointers and sizes.  Host arrays passed in are
 * borrowed for the duration of the call only; outputs are caller-allocated.
 * One engine = one GPU = one owning host thread.  No exceptions cross the ABI
 * and the library never calls exit(); every entry point returns an az_status
 * (0 = ok, negative = error) and az_last_error() explains the last failure.
 * There is NO CPU fallback: without a CUDA device az_engine_create fails.
 */
#ifndef AZTOT_B200_H
#define AZTOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZ_ABI_VERSION 2
#define AZ_MAX_SPECIES 15   /* MX_SPEC, src/defines.h:15 */
#define AZ_N_UVECT 3072     /* nUvect, src/cuTemp.h:4-6 = 3 * 32 * 16 * 2 */
#define AZ_N_MOM 20         /* nMom, src/cuInit.cu:921 (pressure ring length) */
#define AZ_MX_FRC 1.0e4f    /* MX_FRC, src/defines.h:10 (AUTO_CAP) */

typedef enum az_status {
    AZ_OK = 0,
    AZ_ERR_INVALID = -1,     /* bad argument / inconsistent description          */
    AZ_ERR_NO_DEVICE = -2,   /* no CUDA device (there is no CPU fallback)        */
    AZ_ERR_CUDA = -3,        /* a CUDA runtime call or kernel failed             */
    AZ_ERR_UNSUPPORTED = -4, /* feature outside the hot path (Ewald k-space ...) */
    AZ_ERR_OVERFLOW = -5,    /* a fixed-capacity device buffer overflowed        */
    AZ_ERR_COMM = -6,        /* NCCL / multi-GPU plumbing failed                 */
    AZ_ERR_IO = -7           /* deck file missing or unparsable                  */
} az_status;

/* van der Waals kinds: numeric values are the reference's type ids, src/vdw.h:15-21 */
typedef enum az_vdw_kind {
    AZ_VDW_NONE = 0,
    AZ_VDW_LJ = 1,    /* lnjs  U = 4e[(s/r)^12-(s/r)^6]       src/cuVdW.cu:25-47   */
    AZ_VDW_BUCK = 2,  /* buck  U = A exp(-r/ro) - C/r^6       src/cuVdW.cu:69-91   */
    AZ_VDW_P746 = 3,  /* p746  host-only in the reference (src/vdw.cpp); rejected   */
    AZ_VDW_BMH = 4,   /* bmhs  U = A exp[B(s-r)]-C/r^6-D/r^8  src/cuVdW.cu:115-136 */
    AZ_VDW_ELIN = 5,  /* elin  U = A exp(-r/ro) + C r         src/cuVdW.cu:162-179 */
    AZ_VDW_EINV = 6,  /* einv  U = A exp(-r/ro) - C/r         src/cuVdW.cu:200-217 */
    AZ_VDW_SURK = 7   /* surk  radius/temperature dependent   src/cuVdW.cu:236-257 */
} az_vdw_kind;

/* electrostatics kinds: values are tpElec* of src/dataStruct.h */
typedef enum az_elec_kind {
    AZ_ELEC_NONE = 0,   /* no_coul      src/cuElec.cu:20     */
    AZ_ELEC_DIRECT = 1, /* direct_coul  src/cuElec.cu:27-37  */
    AZ_ELEC_EWALD = 2,  /* real_ewald   src/cuElec.cu:94-113; + reciprocal space if az_system_desc.ewald_k is set */
    AZ_ELEC_FENNEL = 3  /* fennel (DSF) src/cuElec.cu:115-134 */
} az_elec_kind;

/* thermostat kinds: values are tpTerm* of src/dataStruct.h */
typedef enum az_tstat_kind {
    AZ_TSTAT_NONE = 0,
    AZ_TSTAT_NOSE = 1, /* before_nose / tstat_nose / after_nose  src/cuTemp.cu:115-163 (applied twice per step, src/main.cu:291-299 and apply_tstat) */
    AZ_TSTAT_RADI = 2  /* tstat_radi9  src/cuTemp.cu:689-773 */
} az_tstat_kind;

/* constant bonds (apply_const_bonds, src/cuBonds.cu:709-797) and valence angles (apply_angles,
 * src/cuAngles.cu:169-228).  Bond kinds are the reference's type ids (define_bond_potential, src/cuBonds.cu:1236-1270;
 * parameters as read_bond leaves them, src/bonds.cpp:125-250). */
typedef enum az_bond_kind {
    AZ_BOND_NONE = 0,   /* empty entry: type 0 is reserved for "no bond" (src/cuBonds.cu:731)                    */
    AZ_BOND_HARM = 1,   /* harm  U = 1/2 k (r-r0)^2                       p0 = k, p1 = r0                        */
    AZ_BOND_MORSE = 2,  /* mors  U = D[1-exp(-a(r-r0))]^2 - C             p0 = D, p1 = a, p2 = r0, p3 = C        */
    AZ_BOND_PEDONE = 3, /* pdn   U = Morse - E/r^12                       ... p4 = E                             */
    AZ_BOND_BUCK = 4,   /* buck  U = A exp(-r/ro) - C/r^6                 p0 = A, p1 = ro, p2 = C                */
    AZ_BOND_E612 = 5    /* e612  U = A exp(-r/ro) - C/r^6 - D/r^8 - F/r^12   p0..p4 = A, ro, C, D, F             */
} az_bond_kind;

typedef struct az_bond_type {
    int kind;           /* az_bond_kind */
    int spec1, spec2;   /* species of the first / second atom (read_bond; informational) */
    int reserved;
    double p0, p1, p2, p3, p4;
} az_bond_type;

#define AZ_ANGLE_HCOS 1 /* hcos  U = 1/2 k (cos th - cos th0)^2, src/cuAngles.cu:230-283: p0 = k, p1 = cos th0 */
typedef struct az_angle_type {
    int kind;           /* 0 = empty, AZ_ANGLE_HCOS */
    int central;        /* species of the central atom (informational) */
    double p0, p1;
} az_angle_type;

/* One species; mirror of the fields of Spec that init_cudaMD reads
 * (src/cuInit.cu:843-866).  Internal units (mass already * m_scale). */
typedef struct az_species {
    char name[8];
    double mass;
    double charge;
    int frozen;
    int reserved;
    double radA, radB, mxEng; /* 'radii' section, src/sys_init.cpp:468-480 */
} az_species;

/* One entry of the nSpec x nSpec pair-potential table, already prepared the
 * way read_vdw leaves it (src/vdw.cpp:261-299): LJ p0=4e p1=s^2 p2=24e,
 * r2cut = rcut^2.  kind == AZ_VDW_NONE means "no entry" (NULL in the reference). */
typedef struct az_vdw_entry {
    int kind;
    int use_radii;
    double p0, p1, p2, p3, p4;
    double r2cut;
} az_vdw_entry;

/* Everything init_cudaMD reads from Atoms/Field/Sim/TStat/Box/Elec. */
typedef struct az_system_desc {
    int abi_version; /* = AZ_ABI_VERSION */
    int n_atoms;
    int n_species;
    int reserved0;
    /* atoms (SoA doubles as in struct Atoms, src/dataStruct.h); narrowed to
     * float exactly like src/cuInit.cu:775-779.  vxs.. / fxs.. may be NULL (=0). */
    const double *xs, *ys, *zs;
    const double *vxs, *vys, *vzs;
    const double *fxs, *fys, *fzs;
    const int *types;
    /* rectangular periodic box, src/box.cpp:9-27 */
    double la, lb, lc;
    const az_species *species;   /* [n_species] */
    const az_vdw_entry *vdw;     /* [n_species*n_species], row-major [ti][tj] */
    double max_r_vdw;            /* Field::maxRvdw, src/vdw.cpp:264 */
    /* electrostatics after prepare_elec, src/elec.cpp:371-406 */
    int elec_kind;
    int reserved1;
    double r_real;   /* Elec::rReal (kept even when the kind was downgraded) */
    double alpha;
    double daipi2;   /* 2 alpha / sqrt(pi) */
    double el_c1;    /* Elec::scale  = erfc(a rc)/rc                         */
    double el_c2;    /* Elec::scale2 = erfc(a rc)/rc^2 + daipi2 exp(-a^2rc^2)/rc */
    double eps;      /* permittivity (parsed, unused by the reference kernels) */
    /* thermostat, src/temperature.cpp:91-259, src/cuTemp.cu:11-62 */
    int tstat_kind;
    int reserved2;
    double temperature;
    double t_kin;              /* TStat::tKin, src/sys_init.cpp:1099-1106 */
    const double *photons;     /* [n_atoms] photon energies or NULL if not radi */
    const double *uvect_x, *uvect_y, *uvect_z; /* [AZ_N_UVECT] or NULL -> built-in table */
    const float *init_radii;   /* [n_atoms] or NULL -> 0.577 + u*1e-4 (seeded)  */
    const float *init_engs;    /* [n_atoms] or NULL -> 0                        */
    uint64_t seed;             /* counter-based RNG key (the reference is unseeded) */
    /* simulation scalars, src/cuInit.cu:951-961 */
    double timestep;           /* Sim::tSt */
    int n_equil, eq_freq;      /* Sim::nEq, Sim::freqEq */
    double r2_max;             /* Sim::r2Max, src/sys_init.cpp:1059-1072 */
    double cell_size;          /* Sim::desired_cell_size ('cell_list') */
    double efield[3];          /* Sim::Ux,Uy,Uz */
    /* Domain decomposition only (az_cuda_opts.n_ranks > 1; new functionality, the reference is
     * single-GPU).  global_ids == NULL: every rank passes the whole system (persistent id = index)
     * and keeps the atoms of its own brick.  Otherwise the rank passes a share of the system (any
     * atoms within one brick of its own) with these persistent ids; t_kin is the global value. */
    const int *global_ids;
    long long n_atoms_global;  /* 0 = n_atoms */
    /* Nose-Hoover thermostat (src/cuTemp.cu:115-163, host prep src/sys_init.cpp:1107-1111) and 'reset_vels' */
    double tstat_tau;          /* TStat::tau ('temperature T nose tau'); rQmass = 0.5 / tKin / tau^2, qMassTau2 = 2 tKin */
    int reset_vels;            /* Sim::reset_vels: every n-th step velocities are zeroed instead of the 2nd half kick (src/main.cu:372-377) */
    int reserved3;
    /* reciprocal-space part of the Ewald sum ('elec pme rc alpha kx ky kz', src/elec.cpp:47-51).  With
     * elec_kind == AZ_ELEC_EWALD and any ewald_k > 0 the engine adds the k-space forces and energies of ewald_rec
     * (src/elec.cpp:167-336; the CUDA twin recip_ewald / ewald_force, src/cuElec.cu:151-381, is compiled out of the
     * reference by defines.h:18) and the constant term of ewald_const (src/elec.cpp:144-165).  All zero: real-space
     * part only, as the reference's CUDA build does. */
    int ewald_k[3];            /* Elec::kx, ky, kz: l in [0,kx), m in (-ky,ky), n in (-kz,kz) */
    int reserved4;
    /* constant bonds and valence angles ('bonds' / 'bond_list' + bonds.txt, 'angles' / 'angle_list' +
     * angles.txt; src/bonds.cpp:25-110, src/angles.cpp:20-130).  Atom indices are positions in the arrays above
     * (atoms.xyz order); the engine follows the atoms through every sort itself (what sort_bonds / sort_angles,
     * src/cuSort.cu:199-235, do in the reference).  Type 0 of either table is the empty entry.  Single GPU only. */
    int n_bond_types, n_bonds;
    const az_bond_type *bond_types;                 /* [n_bond_types]                         */
    const int *bond_at1, *bond_at2, *bond_type;     /* [n_bonds]                              */
    int n_angle_types, n_angles;
    const az_angle_type *angle_types;               /* [n_angle_types]                        */
    const int *angle_centre, *angle_lig1, *angle_lig2, *angle_type; /* [n_angles]             */
} az_system_desc;

/* cuda.txt-level options plus the multi-GPU layout.  Zero-initialise for defaults. */
typedef struct az_cuda_opts {
    int device;          /* CUDA device ordinal                                       */
    int rebuild_every;   /* 0/1 = reference behaviour (bin + sort every step)         */
    float skin;          /* extra cell width when rebuild_every > 1 (Angstrom)        */
    int use_graph;       /* 1: az_engine_step(first, n) is captured as a CUDA graph the first time a call shape (n, position in
                            the rebuild cycle) occurs and replayed afterwards -- for the latency-bound decks whose step is
                            6-7 short kernels; single GPU only; per-section timers (az_engine_profile) switch it off */
    int pair_variant;    /* 0 = auto (Verlet list between rebuilds and for large systems, warp-per-atom cell sweep
                            for small systems on the every-step schedule), 1 = shared-memory cell tiles,
                            2 = list (rebuilt every step if rebuild_every <= 1), 3 = warp-per-atom sweep,
                            4 = thread-per-atom sweep (measured slower, kept for the A/B) */
    int cells_per_block; /* 0 = auto; z-run length a CTA owns in the pair kernel      */
    /* spatial domain decomposition (new functionality) */
    int rank, n_ranks;
    int pgrid[3];        /* ranks along x,y,z; product = n_ranks; 0 -> {n_ranks,1,1}  */
    int ghost_slots;     /* reserved                                                   */
    const void *nccl_id; /* 128 bytes from az_comm_unique_id, the same on all ranks   */
    int max_neigh;       /* rows of the Verlet list per atom ('max_neigh' of control.txt); 0 = from density */
    int list_format;     /* Verlet-list path: 0 = auto (= 1), 1 = global 32-bit rows gathered through L1, 2 = tile lists: 16-bit
                            rows that index a shared-memory copy of the neighbourhood of each 128-atom tile (bitwise the same
                            forces, half the list bytes; measured slower on B200, kept for the A/B -- 'list_format' of cuda.txt) */
    int tile_atoms;      /* tile lists: atoms the shared-memory neighbourhood of a tile holds; 0 = from density ('tile_atoms') */
    int reserved[5];
} az_cuda_opts;

/* Scalars of stat.dat / msd.dat for the last completed step
 * (src/cuStat.cu:232-298; float like the reference, reduced in double). */
typedef struct az_stats {
    int step;            /* iStep of the sample */
    int n_atoms_total;
    float engTot, engKin, engVdW, engCoul1, engCoul2, engTemp, engElecField, engPot;
    float posMom[3], negMom[3]; /* cumulative face momenta, src/cuMDfunc.cu:72-104 */
    float pressure;
    float x_first;       /* xyz[0].x of the sorted stream: the 'x1=' of print_stat, src/cuMDfunc.cu:320 */
    int specAcBoxPos[AZ_MAX_SPECIES][3]; /* face crossing counters */
    int specAcBoxNeg[AZ_MAX_SPECIES][3];
    double engKin_d, engVdW_d, engCoul1_d, engTemp_d; /* unrounded accumulators */
    long long n_pairs;   /* unique pairs inside r2Max seen by the last energy-reducing force pass */
    float engCoul3;      /* Ewald: constant term (ewald_const, src/elec.cpp:144-165); engCoul2 = reciprocal-space energy */
    float reserved;
    double engCoul2_d;
    float engBond, engAngle;   /* md->engBond, md->engAngl (part of engPot, src/main.cu:128) */
    double engBond_d, engAngle_d;
} az_stats;

#define AZ_ORDER_SORTED 0   /* cell-sorted order, what md_to_host hands back */
#define AZ_ORDER_ORIGINAL 1 /* order of atoms.xyz (by persistent id)         */

/* Caller-allocated outputs; any pointer may be NULL to skip that array. */
typedef struct az_atoms_out {
    int n_atoms;      /* capacity of the arrays = desc.n_atoms (global count on 1 GPU) */
    int n_written;    /* out: atoms written (owned atoms of this rank on multi-GPU)    */
    double *xs, *ys, *zs;
    double *vxs, *vys, *vzs;
    int *types;
    int *ids;         /* persistent ids (index in the original atoms.xyz) */
    float *engs;      /* radiative thermostat: thermal energy U */
    float *radii;     /* radiative thermostat: radius           */
} az_atoms_out;

/* Parity-test dump, sorted order.  Any pointer may be NULL. */
typedef struct az_dump {
    int n_atoms, n_cells;      /* capacities in; actual sizes out */
    int cnum[3];               /* out: cNumber */
    float csize[3], crev[3];   /* out: cSize, cRevSize (src/cuCellList.cu:21-22) */
    int *cell_indexes;         /* [n_atoms]  cell of sorted atom i       */
    int *first_atom_in_cell;   /* [n_cells]                              */
    int *n_at_in_cell;         /* [n_cells]                              */
    int *ids;                  /* [n_atoms]  persistent id of sorted atom i */
    float *fx, *fy, *fz;       /* [n_atoms]  forces after the pair kernel */
    float *x, *y, *z;          /* [n_atoms]  float positions              */
    float *vx, *vy, *vz;       /* [n_atoms]                               */
} az_dump;

typedef struct az_engine az_engine;

/* library / environment */
int az_abi_version(void);
const char *az_last_error(const az_engine *e); /* e may be NULL: error of the last failed create */
int az_device_count(void);

/* Multi-GPU: one process (or thread) per GPU, one engine per rank.  Rank 0 calls az_comm_unique_id
 * and hands the 128 bytes to every rank (e.g. torch.distributed broadcast, MPI_Bcast, a file); all
 * ranks then call az_engine_create with opts.rank / n_ranks / pgrid / nccl_id.  az_engine_create,
 * az_engine_step, az_engine_stats and az_engine_destroy are collective: every rank calls them with the
 * same arguments.  az_engine_download returns the atoms the rank owns. */
int az_comm_unique_id(void *id128);

/* life cycle */
int az_engine_create(const az_system_desc *desc, const az_cuda_opts *opts, az_engine **out);
void az_engine_destroy(az_engine *e); /* idempotent on NULL */

/* Advance nsteps steps; first_step is the reference's iStep of the first one
 * (it gates equilibration scaling and keys the thermostat RNG).  Asynchronous on the
 * engine's stream; az_engine_sync / stats / download synchronise. */
int az_engine_step(az_engine *e, int first_step, int nsteps);
int az_engine_sync(az_engine *e);
void *az_engine_stream(az_engine *e); /* cudaStream_t of the engine (for external timing) */

int az_engine_stats(az_engine *e, az_stats *out);
int az_engine_download(az_engine *e, az_atoms_out *out, int order);
int az_engine_debug_dump(az_engine *e, az_dump *out);
/* Replace positions/velocities (original order, doubles, host memory) without
 * re-creating the engine; forces are zeroed like a fresh create. */
int az_engine_upload(az_engine *e, const double *xs, const double *ys, const double *zs,
                     const double *vxs, const double *vys, const double *vzs);
/* The same without the final synchronisation: the copies run on the engine's copy stream into one of two
 * staging buffers, so the H2D transfer of the next input overlaps the kernels of the step in flight.  The
 * arrays (pinned memory from az_host_alloc, or the copy is not asynchronous) must stay unchanged until the
 * next az_engine_sync / az_engine_stats / az_engine_download. */
int az_engine_upload_async(az_engine *e, const double *xs, const double *ys, const double *zs,
                           const double *vxs, const double *vys, const double *vzs);
/* The same from float arrays: half the host->device bytes (the engine narrows every coordinate to float on upload anyway,
 * src/cuInit.cu:775-779).  Single GPU. */
int az_engine_upload_f32_async(az_engine *e, const float *xs, const float *ys, const float *zs,
                               const float *vxs, const float *vys, const float *vzs);
/* Positions (wrapped into the box) and, if the last three pointers are set, velocities of all atoms in atoms.xyz order as
 * float: one gather kernel and one D2H copy per array; pinned host arrays (az_host_alloc) make the copies run at PCIe
 * speed.  Synchronous at return.  Single GPU. */
int az_engine_download_f32(az_engine *e, float *xs, float *ys, float *zs, float *vxs, float *vys, float *vzs);

/* Radial distribution function from the cell grid: replaces init_cuda_rdf / brute_rdf / rdf_iter / copy_rdf
 * (src/cuStat.cu:418-594; the reference's brute force is O(N^2) per sample).  az_engine_rdf_begin allocates and clears
 * the histogram: n_bins = (int)(min(r_max, la) / dr) like init_rdf (src/rdf.cpp:38-45), one column per unordered species
 * pair in the reference's order (i <= j, row-major).  az_engine_rdf_sample adds the pairs of the current configuration
 * (call it after a step, every 'rdf' frequency steps).  az_engine_rdf_read returns the raw integer pair counts, bin-major
 * (counts[bin * n_pairs + pair]), and the number of samples; the normalisation of copy_rdf (src/cuStat.cu:514-561) is
 * host arithmetic on them: g = count * V / (N_a N_b) * 2 / (4/3 pi dr^3 n_samples) / (3 i (i + 1) + 1) * (a == b ? 1 : 1/2). */
int az_engine_rdf_begin(az_engine *e, double r_max, double dr);
int az_engine_rdf_sample(az_engine *e);
int az_engine_rdf_read(az_engine *e, unsigned long long *counts, int capacity, int *n_bins, int *n_pairs, int *n_samples);

/* Per-kernel device timers (CUDA events on the engine stream).  az_engine_profile(e, 1) times each
 * section synchronously; (e, 2) only records events inside the step loop (no host sync) and
 * az_engine_timers resolves them once the stream has drained; (e, 0) switches timing off. */
#define AZ_TIMER_INTEGRATE_A 0 /* kick+drift+wrap+bin  (verlet_1stage)     */
#define AZ_TIMER_SCAN 1        /* calc_firstAtomInCell                     */
#define AZ_TIMER_SORT 2        /* sort_atoms                               */
#define AZ_TIMER_PAIR 3        /* cell_list5a + cell_list4b_noshared       */
#define AZ_TIMER_INTEGRATE_B 4 /* verlet_2stage + thermostat + quantities  */
#define AZ_TIMER_HALO 5        /* multi-GPU: per-step ghost refresh        */
#define AZ_TIMER_PAIR_LIST 6   /* Verlet-list force pass between rebuilds  */
#define AZ_TIMER_EWALD 7       /* reciprocal-space Ewald (ewald_rec)       */
#define AZ_TIMER_MIGRATE 8     /* multi-GPU: migration + ghost selection at a rebuild */
#define AZ_N_TIMERS 9
int az_engine_profile(az_engine *e, int enable);
int az_engine_timers(az_engine *e, double ms_sum[AZ_N_TIMERS], long long launches[AZ_N_TIMERS]);
long long az_engine_launch_count(az_engine *e); /* kernels launched since create */

/* az_engine_step bracketed by CUDA events on the engine stream: *ms = device time of the nsteps. */
int az_engine_step_timed(az_engine *e, int first_step, int nsteps, float *ms);

/* Pinned host memory for the arrays handed to az_engine_create / az_engine_upload (optional;
 * pageable memory works, pinned memory lets the copies run at full PCIe speed). */
void *az_host_alloc(size_t bytes);
void az_host_free(void *p);

/* Measured FP32 FMA throughput of the device in TFLOP/s (2 flops per FMA): the roofline
 * denominator of the pair kernel (MEASURED_PEAKS.json carries no FP32 figure). */
int az_measure_fp32_peak(int device, double *tflops, double *sm_mhz_effective);

/* Engine options by name ("energy_every_step": reduce potential energies on every step, not only
 * on the last step of each az_engine_step call). */
int az_engine_set_option(az_engine *e, const char *key, double value);

/* Geometry and launch plan chosen at create time (what split_cells / init_cellList decided,
 * src/cuCellList.cu:9-34,583-636). */
typedef struct az_engine_info_t {
    int n_atoms, n_cells;
    int cnum[3], stencil[3];
    float csize[3], crev[3];
    int pair_blocks, pair_threads, pair_smem_bytes, pair_zc, pair_vk, pair_ck;
    int n_sm;
    int max_neigh_seen;  /* largest Verlet-list row count of the last rebuild (0 without a list) */
    long long device_bytes;
    int list_rows, flags;  /* list capacity per atom; raw sticky device flags */
    int capacity;          /* most atoms az_engine_download / az_engine_debug_dump can hand back (a brick's owned capacity) */
    int n_ranks;
    long long graph_replays; /* az_engine_step calls served by replaying a captured CUDA graph (az_cuda_opts.use_graph) */
    int graphs_cached;
    int peer_halo;       /* bricks: 0 = the per-step ghost refresh is ncclSend/ncclRecv; 1 = peer memory (NVLink stores + arrival flags);
                            2 = peer memory fused with the list pass (ghost-touching blocks wait on the flags and read the staging);
                            3 = direct: neighbours store their boundary atoms straight into this rank's sorted position stream,
                            one plain list launch per step whose ghost-touching blocks wait on the flags */
    int tile_list;       /* 1 = the Verlet-list path runs on tile lists (az_cuda_opts.list_format) */
    int tile_atoms, tile_rows_smem;       /* shared-memory allotment: atoms of a tile's neighbourhood, rows per atom of the build */
    int tile_atoms_seen, tile_runs_seen;  /* largest neighbourhood (atoms, runs of the sorted stream) any rebuild has seen */
    int reserved;
} az_engine_info_t;
int az_engine_info(az_engine *e, az_engine_info_t *out);

/* ---- deck surface (host side, libaztot_host): the reference's input files ---- */
typedef struct az_deck az_deck;
/* Parse atoms.xyz, field.txt, control.txt, cuda.txt from `dir` with the reference's
 * token-scan semantics (src/utils.cpp:87-195, src/sys_init.cpp:174-988). */
int az_deck_load(const char *dir, uint64_t seed, az_deck **out); /* on failure *out still carries the message: free it */
const az_system_desc *az_deck_desc(const az_deck *d);
const az_cuda_opts *az_deck_opts(const az_deck *d);
int az_deck_nsteps(const az_deck *d);
int az_deck_stat_period(const az_deck *d);
const char *az_deck_error(const az_deck *d);      /* d may be NULL: error of the last failed load */
const char *az_deck_warnings(const az_deck *d);   /* the WARNING[...] lines the reference would print */
int az_deck_nstep_stat(const az_deck *d);         /* cuda.txt 'nstep stat' (flush granularity)       */
int az_deck_species_count(const az_deck *d, int species);
/* the 'rdf rmax dr freq outfreq' directive (src/rdf.cpp:14-37); returns 0 if the deck has none */
int az_deck_rdf(const az_deck *d, double *r_max, double *dr, int *freq, int *outfreq);
/* 'rdf ... nucl': the by-nuclei table of brute_nrdf / copy_nrdf (src/cuStat.cu:614-777).  Every species names its nucleus
 * (second word of a 'spec' line, read_spec src/sys_init.cpp:78-104).  az_deck_nuclei returns the number of distinct nuclei
 * and, if asked, the nucleus index of every species; az_deck_rdf_nuclei says whether the deck asked for rdf_n.dat. */
int az_deck_nuclei(const az_deck *d, int *nucleus_of_species /* [n_species] or NULL */);
const char *az_deck_nucleus_name(const az_deck *d, int nucleus);
int az_deck_nucleus_count(const az_deck *d, int nucleus);   /* atoms carrying that nucleus (Field::nnumbers) */
int az_deck_rdf_nuclei(const az_deck *d);
/* Host-side thermostat tables of read_tstat (src/temperature.cpp:91-259), exposed so that they can be pinned on the reference:
 * the 3 x 32 x 16 x 2 = AZ_N_UVECT preset unit vectors (:164-223; what the engine uses when az_system_desc.uvect_* are NULL),
 * and photon_engs' inverse CDF for one uniform draw (:28-89; returns 0 when the reference would draw again). */
int az_default_uvects(double *ux, double *uy, double *uz);
int az_photon_energy(double x, double temperature, double prev, double *e_out);
void az_deck_free(az_deck *d);

#ifdef __cplusplus
}
#endif
#endif /* AZTOT_B200_H */
