/*
 * az_oracle.c -- CPU restatement of azTotMD's per-step hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file.  The product (libaztot_b200.so) never links or calls it and has no
 * CPU fallback.
 *
 * It restates, in plain C with float32 wherever the reference's device code is float32,
 * the algorithm of the reference's CUDA build (paths into /root/reference):
 *   grid geometry        split_cells            src/cuCellList.cu:9-34
 *   cell-pair existence  cell_dist_int          src/cuCellList.cu:98-132
 *                        pair_exists_shift      src/cuCellList.cu:213-240
 *   binning              count_cell             src/cuSort.cu:114-128
 *   scan                 calc_firstAtomInCell   src/cuSort.cu:130-143
 *   reorder              sort_atoms             src/cuSort.cu:145-197
 *   stage 1              verlet_1stage          src/cuMDfunc.cu:333-518
 *   PBC + face stats     put_periodic           src/cuMDfunc.cu:35-106
 *   pair gate / cap      pair_1, force_autocap  src/cuPairs.cu:94-167
 *   vdW forms            cu_fe(r)_lj/buck/bmh/elin/einv, surk_pot   src/cuVdW.cu:25-257
 *   Coulomb forms        direct_coul, real_ewald, fennel            src/cuElec.cu:27-134
 *   stage 2              verlet_2stage          src/cuMDfunc.cu:521-599
 *   T-scale              temp_scale/after_tscale  src/cuTemp.cu:77-113, gate :777-785
 *   radiative thermostat tstat_radi9, adsorb_rand_photon, radiate_photon3,
 *                        get_angled_vector      src/cuTemp.cu:395-453,484-507,631-773
 *   derived quantities   calc_quantities        src/main.cu:121-162
 *   host narrowing       init_cudaMD            src/cuInit.cu:768-791,934-961
 *                        init_cuda_tstat        src/cuTemp.cu:11-62
 *   unit vectors         read_tstat             src/temperature.cpp:164-223
 *   bonds / angles       bond_harm..bond_e6812, angle_hcos, delta_periodic   src/cuBonds.cu:1123-1224, src/cuAngles.cu:230-283,
 *                        src/cuMDfunc.cu:178-203  (SURVEY 8f-4 GROUNDWORK: pinned arithmetic, no kernel yet)
 *   k-space Ewald        ewald_rec, ewald_const, prepare_elec   src/elec.cpp:144-165,167-336,376-397 (the serial
 *                        twin, double precision: the CUDA twin src/cuElec.cu:151-381 is compiled out, defines.h:18)
 *
 * Deliberate, documented divergences from reference behaviour:
 *   - within-cell order is canonical (ascending persistent id) instead of atomicAdd
 *     arrival order (src/cuSort.cu:126);
 *   - per-atom radii/engs follow the atom through the sort (reference bug, src/cuSort.cu:96);
 *   - the thermostat RNG is counter based: Philox4x32-10 keyed by (seed), counter
 *     (atom id, step); the reference's xorshift state is uninitialised and racy
 *     (src/cuTemp.cu:701,770, src/cuUtils.cu:89-105);
 *   - photon energy of atom i is engPhotons[sorted index i], the in-order-block case of
 *     src/cuTemp.cu:697-722;
 *   - radiate_photon3 with e/(m c v0) >= 1 uses the anti-parallel unit vector instead of an
 *     uninitialised one (src/cuTemp.cu:663-675);
 *   - count_cell indices are clamped into the grid (src/cuSort.cu:119-126 has no clamp);
 *   - posMomBuf/negMomBuf are real float3[nMom] rings, face counters start at zero
 *     (src/cuInit.cu:927-928,986-987);
 *   - same-species surk with k != l: radii are ordered (min,max) so the pair term is symmetric.
 * Energies and momenta are accumulated in double from float pair terms.
 *
 * PARITY PINNING: the reference ships no golden vectors.  This oracle is
 * pinned (tests/test_oracle_pinning.py) against (a) the analytic known answers listed in
 * SURVEY.md 8c and (b) forces/energies of the reference's own serial build compiled from
 * /root/reference into oracle/_ref (oracle/build_ref.sh), committed as tests/golden/ fixtures.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/aztot_b200.h"

#if defined(__GNUC__)
/* built with -ffp-contract=off (oracle/Makefile): products and sums round separately */
#endif

/* ------------------------------------------------------------------ RNG ---- */
/* Philox4x32-10 (Salmon et al., SC'11), the same function the CUDA path uses. */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void azo_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                uint32_t *out) { philox4x32_10(c0, c1, c2, c3, k0, k1, out); }

/* uniform in [0,1) used for default initial radii: 0.577 + u*1e-4 (src/cuTemp.cu:41) */
static float seeded_unit(uint64_t seed, uint32_t id)
{
    uint32_t o[4];
    philox4x32_10(id, 0xA5A5A5A5u, 0, 0, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return (float)(o[0] % 10000u) / 10000.0f; /* rand01(), src/utils.cpp:197-201 */
}

/* ------------------------------------------------------------- geometry ---- */
typedef struct {
    int cnum[3];
    float csize[3], crev[3];
    int cnYZ, ncell;
} azo_grid;

/* split_cells, div_type 1 (src/cuCellList.cu:9-34, called from src/cuInit.cu:979) */
void azo_split_cells(float lx, float ly, float lz, float r, int *cnum, float *csize, float *crev)
{
    float L[3] = { lx, ly, lz };
    for (int a = 0; a < 3; a++) {
        float q = L[a] / r;
        int n = (int)floorf(q);
        if (n < 1) n = 1; /* reference would divide by zero; keep one cell */
        cnum[a] = n;
        csize[a] = L[a] / (float)n;
        crev[a] = (float)n / L[a];
    }
}

/* count_cell index formula (src/cuSort.cu:119) with the clamp divergence. */
int azo_cell_index(float x, float y, float z, const float *crev, const int *cnum)
{
    int ix = (int)floor((double)x * (double)crev[0]);
    int iy = (int)floor((double)y * (double)crev[1]);
    int iz = (int)floor((double)z * (double)crev[2]);
    if (ix < 0) ix = 0; if (ix >= cnum[0]) ix = cnum[0] - 1;
    if (iy < 0) iy = 0; if (iy >= cnum[1]) iy = cnum[1] - 1;
    if (iz < 0) iz = 0; if (iz >= cnum[2]) iz = cnum[2] - 1;
    return ix * (cnum[1] * cnum[2]) + iy * cnum[2] + iz;
}

/* Stencil half width along one axis: largest delt with rmin=(delt-1)*csize <= rc
 * (cell_dist_int, src/cuCellList.cu:98-132), capped at n/2 because each cell pair
 * exists once with the image chosen by "delt > mx/2" (integer division). */
int azo_stencil_half(int n, float csize, float rc)
{
    int s = 1;
    while ((float)s * csize <= rc && s < n) s++;
    if (s > n / 2) s = n / 2;
    return s;
}

/* Is offset o from cell coordinate c a distinct neighbour (reference image rule)?
 * Returns 1 and writes the wrapped coordinate and the image shift sign (-1,0,+1). */
int azo_neighbor(int n, int c, int o, int *cn, int *sh)
{
    int t = c + o;
    int ao = o < 0 ? -o : o;
    *sh = 0;
    if (ao > n / 2) return 0;
    if (t < 0) { t += n; *sh = -1; }
    else if (t >= n) { t -= n; *sh = 1; }
    if ((n % 2) == 0 && ao == n / 2 && *sh != 0) return 0; /* even grid: unwrapped image only */
    if (ao != 0 && t == c) return 0;
    *cn = t;
    return 1;
}

/* ------------------------------------------------------- pair potentials ---- */
typedef struct { int kind, use_radii; float p0, p1, p2, p3, p4, r2cut; } azo_vdw;

#define FCOUL_SCALE 14.3996f /* d_Fcoul_scale, src/main.cu:40 */
#define SQRTPI_F 1.772453f   /* d_sqrtpi,      src/main.cu:41 */

/* returns F/r, adds the energy; `r` is the "known r" in/out of the *_r variants. */
static float vdw_feng(const azo_vdw *v, float r2, float *r, int use_coul, float rad1, float rad2,
                      float *eng)
{
    switch (v->kind) {
    case AZ_VDW_LJ: { /* src/cuVdW.cu:25-47 */
        float r2i = 1.f / r2;
        float sr2 = v->p1 * r2i;
        float sr6 = sr2 * sr2 * sr2;
        *eng += v->p0 * sr6 * (sr6 - 1.f);
        return v->p2 * r2i * sr6 * (2.f * sr6 - 1.f);
    }
    case AZ_VDW_BUCK: { /* src/cuVdW.cu:69-91 */
        float r2i = 1.f / r2;
        float r4i = r2i * r2i;
        float rr;
        if (use_coul) { if (*r == 0.f) *r = sqrtf(r2); rr = *r; } else rr = sqrtf(r2);
        *eng += v->p0 * expf(-rr / v->p1) - v->p2 * r4i * r2i;
        if (use_coul)
            return v->p0 * expf(-rr / v->p1) / rr / v->p1 - 6.f * v->p2 * r4i * r4i;
        /* cu_fe_buck has a stray double constant, src/cuVdW.cu:90 */
        return (float)((double)(v->p0 * expf(-rr / v->p1) / rr / v->p1) - 6.0 * (double)v->p2 * (double)r4i * (double)r4i);
    }
    case AZ_VDW_BMH: { /* src/cuVdW.cu:115-136 */
        float r2i = 1.f / r2;
        float r4i = r2i * r2i;
        float rr;
        if (use_coul) { if (*r == 0.f) *r = sqrtf(r2); rr = *r; } else rr = sqrtf(r2);
        *eng += v->p0 * expf(v->p1 * (v->p2 - rr)) - v->p3 * r4i * r2i - v->p4 * r4i * r4i;
        return v->p0 * v->p1 * expf(v->p1 * (v->p2 - rr)) / rr - 6.f * v->p3 * r4i * r4i - 8.f * v->p4 * r4i * r4i * r2i;
    }
    case AZ_VDW_ELIN: { /* src/cuVdW.cu:162-179 */
        float rr;
        if (use_coul) { if (*r == 0.f) *r = sqrtf(r2); rr = *r; } else rr = sqrtf(r2);
        *eng += v->p0 * expf(-rr / v->p1) + v->p2 * rr;
        return v->p0 * expf(-rr / v->p1) / rr / v->p1 - v->p2 / rr;
    }
    case AZ_VDW_EINV: { /* src/cuVdW.cu:200-217 */
        float rr;
        if (use_coul) { if (*r == 0.f) *r = sqrtf(r2); rr = *r; } else rr = sqrtf(r2);
        *eng += v->p0 * expf(-rr / v->p1) - v->p2 / rr;
        return v->p0 * expf(-rr / v->p1) / rr / v->p1 - v->p2 / rr / r2;
    }
    case AZ_VDW_SURK: { /* surk_pot, src/cuVdW.cu:236-257 */
        float c2ir = v->p1 / (v->p2 * rad1 + v->p3 * rad2);
        float r_prod = rad1 * rad2;
        float c1ab2 = r_prod * r_prod * v->p0;
        float r6 = r2 * r2 * r2;
        float rr = sqrtf(r2);
        float ir6 = 1.f / r6;
        float ir = 1.f / rr;
        float val = r_prod * ir6 * (c1ab2 * ir - c2ir);
        *eng += val;
        return r_prod * ir6 / r2 * (7.f * c1ab2 * ir - 6.f * c2ir);
    }
    default:
        return 0.f;
    }
}

typedef struct { int kind; float alpha, daipi2, elC1, elC2, rElec; } azo_elec;

static float coul_feng(const azo_elec *e, float r2, float *r, float chprd, float *eng)
{
    switch (e->kind) {
    case AZ_ELEC_DIRECT: { /* src/cuElec.cu:27-37 */
        float kqq = chprd * FCOUL_SCALE;
        *r = sqrtf(r2);
        *eng += kqq / *r;
        return kqq / *r / r2;
    }
    case AZ_ELEC_EWALD: { /* src/cuElec.cu:94-113 */
        float kqq = chprd * FCOUL_SCALE;
        *r = sqrtf(r2);
        float ar = e->alpha * *r;
        float erfcar = erfcf(ar);
        *eng += kqq * erfcar / *r;
        return kqq / *r / r2 * (erfcar + 2 * ar / SQRTPI_F * expf(-ar * ar));
    }
    case AZ_ELEC_FENNEL: { /* src/cuElec.cu:115-134, with its double intermediates */
        float kqq = chprd * FCOUL_SCALE;
        *r = sqrtf(r2);
        double ir = 1.f / *r;
        float ar = e->alpha * *r;
        float erfcar = erfcf(ar);
        *eng += (float)(kqq * (erfcar * ir - e->elC1 + e->elC2 * (*r - e->rElec)));
        return (float)(kqq * ir * ((erfcar / r2 + e->daipi2 * expf(-ar * ar) * ir) - e->elC2));
    }
    default:
        return 0.f;
    }
}

/* Exposed single-pair evaluation for known-answer tests (pair_1 without the geometry). */
float azo_pair_eval(int vdw_kind, const float *p /*p0..p4*/, float r2cut, int elec_kind,
                    float alpha, float daipi2, float elC1, float elC2, float rElec, float chprd,
                    float r2, float r2max, float rad1, float rad2, float *e_vdw, float *e_coul)
{
    azo_vdw v = { vdw_kind, vdw_kind == AZ_VDW_SURK, p[0], p[1], p[2], p[3], p[4], r2cut };
    azo_elec el = { elec_kind, alpha, daipi2, elC1, elC2, rElec };
    float f = 0.f, r = 0.f;
    *e_vdw = 0.f; *e_coul = 0.f;
    if (!(r2 <= r2max)) return 0.f;
    if (elec_kind) f += coul_feng(&el, r2, &r, chprd, e_coul);
    if (vdw_kind != AZ_VDW_NONE && r2 <= r2cut) f += vdw_feng(&v, r2, &r, elec_kind != 0, rad1, rad2, e_vdw);
    if (f > AZ_MX_FRC) f = AZ_MX_FRC; else if (f < -AZ_MX_FRC) f = -AZ_MX_FRC;
    return f;
}

/* ---------------------------------------------------------------- state ---- */
typedef struct azo_state {
    int n, nspec;
    float L[3], revL[3], revArea[3];
    float tSt;
    float efield[3];
    float r2max;
    int nEq, freqEq;
    int tstat; float teKin;
    float rQmass, qMassTau2, chit, consInt; /* Nose-Hoover (init_cuda_tstat, src/cuTemp.cu:17-24) */
    int reset_vels;
    uint64_t seed;
    /* species */
    float mass[AZ_MAX_SPECIES], charge[AZ_MAX_SPECIES], rMasshdT[AZ_MAX_SPECIES];
    int frozen[AZ_MAX_SPECIES];
    float radA[AZ_MAX_SPECIES], radB[AZ_MAX_SPECIES], mxEng[AZ_MAX_SPECIES];
    azo_vdw vdw[AZ_MAX_SPECIES][AZ_MAX_SPECIES];
    float chprod[AZ_MAX_SPECIES][AZ_MAX_SPECIES];
    azo_elec elec;
    azo_grid g;
    int sten[3];
    /* atoms, current (sorted) order */
    float *x, *y, *z, *vx, *vy, *vz, *fx, *fy, *fz, *eng, *rad, *fscale;
    int *type, *id, *cell;
    float *photons, *uv; /* uv[3*AZ_N_UVECT] */
    int *first, *count;
    /* scratch for the sort */
    float *t_f[11]; int *t_i[2]; int *order;
    /* step scalars */
    double engKin, engVdW, engCoul1, engTemp, engElecField;
    float engKin_f; /* value after after_tscale */
    double posMom[3], negMom[3];
    int acPos[AZ_MAX_SPECIES][3], acNeg[AZ_MAX_SPECIES][3];
    float posMomBuf[AZ_N_MOM][3], negMomBuf[AZ_N_MOM][3];
    int iMom;
    float pressure, engTot, engPot;
    int last_step;
    long long n_pairs_in_range; /* unique in-cutoff pairs of the last force evaluation */
    /* reciprocal-space Ewald: Elec::kx,ky,kz, alpha, eps and the double box, as the serial twin keeps them */
    int ew_k[3];
    double ew_alpha, ew_eps, boxd[3];
    double engCoul2, engCoul3;
    float *frx, *fry, *frz; /* the k-space force of the last evaluation alone (sorted order), for the parity tests */
    /* constant bonds and valence angles, by persistent id (the reference remaps the indices through every
     * sort instead: sort_bonds / sort_angles, src/cuSort.cu:199-235) */
    int n_bonds, n_btypes, *b_at1, *b_at2, *b_type, *bt_kind; float *bt_par;
    int n_angles, n_atypes, *a_c, *a_l1, *a_l2, *a_type; float *at_par;
    int *where; /* current (sorted) index of persistent id */
    double engBond, engAngle;
} azo_state;

static void *xmalloc(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) { fprintf(stderr, "azo: oom\n"); abort(); } return p; }

void azo_default_uvects(double *ux, double *uy, double *uz);

double azo_ewald_const(int n, const int *types, const az_species *sp, double alpha, double eps, double vol);

azo_state *azo_create(const az_system_desc *d)
{
    if (!d || d->n_atoms <= 0 || d->n_species <= 0 || d->n_species > AZ_MAX_SPECIES) return NULL;
    azo_state *s = (azo_state *)xmalloc(sizeof(azo_state));
    int n = d->n_atoms;
    s->n = n; s->nspec = d->n_species;
    s->L[0] = (float)d->la; s->L[1] = (float)d->lb; s->L[2] = (float)d->lc;
    for (int a = 0; a < 3; a++) s->revL[a] = 1.f / s->L[a];
    /* edgeArea / revEdgeArea, src/cuInit.cu:943-946 */
    { float ax = s->L[1] * s->L[2], ay = s->L[0] * s->L[2], az = s->L[0] * s->L[1];
      s->revArea[0] = 1.f / ax; s->revArea[1] = 1.f / ay; s->revArea[2] = 1.f / az; }
    s->tSt = (float)d->timestep;
    for (int a = 0; a < 3; a++) s->efield[a] = (float)d->efield[a];
    s->r2max = (float)d->r2_max;
    s->nEq = d->n_equil; s->freqEq = d->eq_freq;
    s->tstat = d->tstat_kind; s->teKin = (float)d->t_kin;
    if (s->tstat == AZ_TSTAT_NOSE) { /* src/sys_init.cpp:1107-1111, narrowed like src/cuTemp.cu:20-23 */
        s->rQmass = (float)(0.5 / d->t_kin / d->tstat_tau / d->tstat_tau);
        s->qMassTau2 = (float)(2.0 * d->t_kin);
        s->chit = 0.f; s->consInt = 0.f;
    }
    s->reset_vels = d->reset_vels;
    s->seed = d->seed;
    for (int i = 0; i < s->nspec; i++) {
        s->mass[i] = (float)d->species[i].mass;
        s->charge[i] = (float)d->species[i].charge;
        s->frozen[i] = d->species[i].frozen;
        s->rMasshdT[i] = 0.5f * (float)d->timestep / s->mass[i]; /* src/cuInit.cu:779 */
        s->radA[i] = (float)d->species[i].radA; s->radB[i] = (float)d->species[i].radB;
        s->mxEng[i] = (float)d->species[i].mxEng;
        for (int j = 0; j < s->nspec; j++) {
            const az_vdw_entry *v = &d->vdw[i * s->nspec + j];
            azo_vdw *o = &s->vdw[i][j];
            o->kind = v->kind; o->use_radii = v->use_radii;
            o->p0 = (float)v->p0; o->p1 = (float)v->p1; o->p2 = (float)v->p2;
            o->p3 = (float)v->p3; o->p4 = (float)v->p4; o->r2cut = (float)v->r2cut;
            s->chprod[i][j] = (float)(d->species[i].charge * d->species[j].charge); /* src/cuInit.cu:823 */
        }
    }
    s->elec.kind = d->elec_kind; s->elec.alpha = (float)d->alpha; s->elec.daipi2 = (float)d->daipi2;
    s->elec.elC1 = (float)d->el_c1; s->elec.elC2 = (float)d->el_c2; s->elec.rElec = (float)d->r_real;
    azo_split_cells(s->L[0], s->L[1], s->L[2], (float)d->cell_size, s->g.cnum, s->g.csize, s->g.crev);
    s->g.cnYZ = s->g.cnum[1] * s->g.cnum[2];
    s->g.ncell = s->g.cnum[0] * s->g.cnYZ;
    { float rc = sqrtf(s->r2max);
      for (int a = 0; a < 3; a++) s->sten[a] = azo_stencil_half(s->g.cnum[a], s->g.csize[a], rc); }

    float **fa[] = { &s->x, &s->y, &s->z, &s->vx, &s->vy, &s->vz, &s->fx, &s->fy, &s->fz, &s->eng, &s->rad, &s->fscale };
    for (unsigned k = 0; k < sizeof(fa) / sizeof(fa[0]); k++) *fa[k] = (float *)xmalloc(sizeof(float) * n);
    for (int k = 0; k < 11; k++) s->t_f[k] = (float *)xmalloc(sizeof(float) * n);
    s->type = (int *)xmalloc(sizeof(int) * n); s->id = (int *)xmalloc(sizeof(int) * n);
    s->cell = (int *)xmalloc(sizeof(int) * n); s->order = (int *)xmalloc(sizeof(int) * n);
    s->t_i[0] = (int *)xmalloc(sizeof(int) * n); s->t_i[1] = (int *)xmalloc(sizeof(int) * n);
    s->first = (int *)xmalloc(sizeof(int) * (s->g.ncell + 1)); s->count = (int *)xmalloc(sizeof(int) * (s->g.ncell + 1));
    for (int i = 0; i < n; i++) {
        s->x[i] = (float)d->xs[i]; s->y[i] = (float)d->ys[i]; s->z[i] = (float)d->zs[i];
        s->vx[i] = d->vxs ? (float)d->vxs[i] : 0.f; s->vy[i] = d->vys ? (float)d->vys[i] : 0.f; s->vz[i] = d->vzs ? (float)d->vzs[i] : 0.f;
        s->fx[i] = d->fxs ? (float)d->fxs[i] : 0.f; s->fy[i] = d->fys ? (float)d->fys[i] : 0.f; s->fz[i] = d->fzs ? (float)d->fzs[i] : 0.f;
        s->type[i] = d->types[i]; s->id[i] = i;
        s->eng[i] = d->init_engs ? d->init_engs[i] : 0.f;
        s->rad[i] = d->init_radii ? d->init_radii[i] : 0.577f + seeded_unit(d->seed, (uint32_t)i) * 0.0001f;
    }
    if (s->tstat == AZ_TSTAT_RADI) {
        s->photons = (float *)xmalloc(sizeof(float) * n);
        for (int i = 0; i < n; i++) s->photons[i] = d->photons ? (float)d->photons[i] : 0.f;
        s->uv = (float *)xmalloc(sizeof(float) * 3 * AZ_N_UVECT);
        double *ux = (double *)xmalloc(sizeof(double) * 3 * AZ_N_UVECT), *uy = ux + AZ_N_UVECT, *uz = uy + AZ_N_UVECT;
        if (d->uvect_x && d->uvect_y && d->uvect_z) {
            memcpy(ux, d->uvect_x, sizeof(double) * AZ_N_UVECT); memcpy(uy, d->uvect_y, sizeof(double) * AZ_N_UVECT);
            memcpy(uz, d->uvect_z, sizeof(double) * AZ_N_UVECT);
        } else azo_default_uvects(ux, uy, uz);
        for (int k = 0; k < AZ_N_UVECT; k++) { s->uv[3 * k] = (float)ux[k]; s->uv[3 * k + 1] = (float)uy[k]; s->uv[3 * k + 2] = (float)uz[k]; }
        free(ux);
    }
    for (int a = 0; a < 3; a++) s->ew_k[a] = (d->elec_kind == AZ_ELEC_EWALD) ? d->ewald_k[a] : 0;
    s->ew_alpha = d->alpha; s->ew_eps = d->eps != 0.0 ? d->eps : 1.0;
    s->boxd[0] = d->la; s->boxd[1] = d->lb; s->boxd[2] = d->lc;
    if (s->ew_k[0] > 0 || s->ew_k[1] > 0 || s->ew_k[2] > 0) {
        s->frx = (float *)xmalloc(sizeof(float) * n); s->fry = (float *)xmalloc(sizeof(float) * n); s->frz = (float *)xmalloc(sizeof(float) * n);
        s->engCoul3 = azo_ewald_const(n, d->types, d->species, d->alpha, s->ew_eps, d->la * d->lb * d->lc);
    }
    if (d->n_bonds > 0 || d->n_angles > 0) s->where = (int *)xmalloc(sizeof(int) * n);
    if (d->n_bonds > 0) {
        s->n_bonds = d->n_bonds; s->n_btypes = d->n_bond_types;
        s->b_at1 = (int *)xmalloc(sizeof(int) * d->n_bonds); s->b_at2 = (int *)xmalloc(sizeof(int) * d->n_bonds); s->b_type = (int *)xmalloc(sizeof(int) * d->n_bonds);
        memcpy(s->b_at1, d->bond_at1, sizeof(int) * d->n_bonds); memcpy(s->b_at2, d->bond_at2, sizeof(int) * d->n_bonds); memcpy(s->b_type, d->bond_type, sizeof(int) * d->n_bonds);
        s->bt_kind = (int *)xmalloc(sizeof(int) * d->n_bond_types); s->bt_par = (float *)xmalloc(sizeof(float) * 5 * d->n_bond_types);
        for (int t = 0; t < d->n_bond_types; t++) { /* narrowed like init_cudaMD narrows cudaBond (src/cuInit.cu) */
            const az_bond_type *b = &d->bond_types[t];
            s->bt_kind[t] = b->kind;
            s->bt_par[5 * t] = (float)b->p0; s->bt_par[5 * t + 1] = (float)b->p1; s->bt_par[5 * t + 2] = (float)b->p2; s->bt_par[5 * t + 3] = (float)b->p3; s->bt_par[5 * t + 4] = (float)b->p4;
        }
    }
    if (d->n_angles > 0) {
        s->n_angles = d->n_angles; s->n_atypes = d->n_angle_types;
        s->a_c = (int *)xmalloc(sizeof(int) * d->n_angles); s->a_l1 = (int *)xmalloc(sizeof(int) * d->n_angles); s->a_l2 = (int *)xmalloc(sizeof(int) * d->n_angles); s->a_type = (int *)xmalloc(sizeof(int) * d->n_angles);
        memcpy(s->a_c, d->angle_centre, sizeof(int) * d->n_angles); memcpy(s->a_l1, d->angle_lig1, sizeof(int) * d->n_angles);
        memcpy(s->a_l2, d->angle_lig2, sizeof(int) * d->n_angles); memcpy(s->a_type, d->angle_type, sizeof(int) * d->n_angles);
        s->at_par = (float *)xmalloc(sizeof(float) * 2 * d->n_angle_types);
        for (int t = 0; t < d->n_angle_types; t++) { s->at_par[2 * t] = (float)d->angle_types[t].p0; s->at_par[2 * t + 1] = (float)d->angle_types[t].p1; }
    }
    s->last_step = -1;
    return s;
}

void azo_destroy(azo_state *s)
{
    if (!s) return;
    float *fa[] = { s->x, s->y, s->z, s->vx, s->vy, s->vz, s->fx, s->fy, s->fz, s->eng, s->rad, s->fscale, s->photons, s->uv };
    for (unsigned k = 0; k < sizeof(fa) / sizeof(fa[0]); k++) free(fa[k]);
    for (int k = 0; k < 11; k++) free(s->t_f[k]);
    free(s->type); free(s->id); free(s->cell); free(s->order); free(s->t_i[0]); free(s->t_i[1]);
    free(s->first); free(s->count);
    free(s->frx); free(s->fry); free(s->frz);
    free(s->b_at1); free(s->b_at2); free(s->b_type); free(s->bt_kind); free(s->bt_par);
    free(s->a_c); free(s->a_l1); free(s->a_l2); free(s->a_type); free(s->at_par); free(s->where);
    free(s);
}

/* The 3072 preset unit vectors of read_tstat (src/temperature.cpp:164-223). */
void azo_default_uvects(double *ux, double *uy, double *uz)
{
    const double pi = 3.14159265359, twopi = 2.0 * pi; /* src/const.h:11-12 */
    const int nTh = 16, nPhi = 32;
    int k = 0;
    for (int perm = 0; perm < 3; perm++)
        for (int i = 0; i < nPhi; i++) {
            double phi = (double)i / nPhi * twopi;
            for (int j = 0; j < nTh; j++) {
                double theta = (double)j / nTh * pi;
                double st = sin(theta), ct = cos(theta), sp = sin(phi), cp = cos(phi);
                double a, b, c; /* (x,y,z) of this variant */
                if (perm == 0) { c = st; b = sp * ct; a = cp * ct; }        /* z=sin th, y=sin ph cos th, x=cos ph cos th */
                else if (perm == 1) { b = st; c = sp * ct; a = cp * ct; }   /* y=sin th, z=sin ph cos th, x=cos ph cos th */
                else { a = st; b = sp * ct; c = cp * ct; }                  /* x=sin th, y=sin ph cos th, z=cos ph cos th */
                ux[k] = a; uy[k] = b; uz[k] = c;
                ux[k + 1] = -a; uy[k + 1] = -b; uz[k + 1] = -c;
                k += 2;
            }
        }
}

/* ------------------------------------------------------------- stage 1 ---- */
/* one axis of put_periodic (src/cuMDfunc.cu:35-60): returns the wrapped coordinate, *n = number of box lengths removed */
float azo_wrap_coord(float x, float L, float revL, int *n)
{
    int nx = (int)floor((double)x * (double)revL);
    x -= (float)((float)nx * L);
    if (x >= L) x = 0.f;
    *n = nx;
    return x;
}

static void put_periodic(azo_state *s, int i)
{
    float *c[3] = { &s->x[i], &s->y[i], &s->z[i] };
    float v[3] = { s->vx[i], s->vy[i], s->vz[i] };
    int t = s->type[i];
    float m = s->mass[t];
    for (int a = 0; a < 3; a++) {
        int nx;
        *c[a] = azo_wrap_coord(*c[a], s->L[a], s->revL[a], &nx);
        if (nx > 0) { s->acPos[t][a]++; s->posMom[a] += (double)(m * v[a]); }
        else if (nx < 0) { s->acNeg[t][a]++; s->negMom[a] += (double)(m * (-v[a])); }
    }
}

static void stage1(azo_state *s)
{
    for (int i = 0; i < s->n; i++) {
        int t = s->type[i];
        float q = s->charge[t], h = s->rMasshdT[t];
        /* nvcc contracts the reference's `vls += rMasshdT * frs` and `xyz += vls * tSt`
         * (src/cuMDfunc.cu:352-368) into FMAs; restated explicitly */
        s->vx[i] = fmaf(h, s->fx[i], s->vx[i]); s->vy[i] = fmaf(h, s->fy[i], s->vy[i]); s->vz[i] = fmaf(h, s->fz[i], s->vz[i]);
        if (!s->frozen[t]) {
            s->x[i] = fmaf(s->vx[i], s->tSt, s->x[i]); s->y[i] = fmaf(s->vy[i], s->tSt, s->y[i]); s->z[i] = fmaf(s->vz[i], s->tSt, s->z[i]);
        }
        put_periodic(s, i);
        s->cell[i] = azo_cell_index(s->x[i], s->y[i], s->z[i], s->g.crev, s->g.cnum);
        s->engElecField += (double)(q * (s->x[i] * s->efield[0] + s->y[i] * s->efield[1] + s->z[i] * s->efield[2]));
        s->fx[i] = -q * s->efield[0]; s->fy[i] = -q * s->efield[1]; s->fz[i] = -q * s->efield[2];
    }
}

/* ---------------------------------------------------------------- sort ---- */
static void sort_atoms(azo_state *s)
{
    int n = s->n, nc = s->g.ncell;
    memset(s->count, 0, sizeof(int) * (nc + 1));
    for (int i = 0; i < n; i++) s->count[s->cell[i]]++;
    int acc = 0;
    for (int c = 0; c < nc; c++) { s->first[c] = acc; acc += s->count[c]; }
    s->first[nc] = acc;
    /* canonical order: by cell, then ascending persistent id (insertion into cell ranges) */
    int *fill = s->t_i[0];
    memset(fill, 0, sizeof(int) * n);
    int *cursor = (int *)xmalloc(sizeof(int) * (nc + 1));
    memcpy(cursor, s->first, sizeof(int) * (nc + 1));
    for (int i = 0; i < n; i++) s->order[cursor[s->cell[i]]++] = i;
    free(cursor);
    for (int c = 0; c < nc; c++) { /* insertion sort by id inside each cell */
        int b = s->first[c], e = b + s->count[c];
        for (int k = b + 1; k < e; k++) {
            int v = s->order[k], kk = k - 1;
            while (kk >= b && s->id[s->order[kk]] > s->id[v]) { s->order[kk + 1] = s->order[kk]; kk--; }
            s->order[kk + 1] = v;
        }
    }
    float *src[11] = { s->x, s->y, s->z, s->vx, s->vy, s->vz, s->fx, s->fy, s->fz, s->eng, s->rad };
    for (int k = 0; k < 11; k++) {
        for (int j = 0; j < n; j++) s->t_f[k][j] = src[k][s->order[j]];
        memcpy(src[k], s->t_f[k], sizeof(float) * n);
    }
    int *isrc[3] = { s->type, s->id, s->cell };
    for (int k = 0; k < 3; k++) {
        for (int j = 0; j < n; j++) s->t_i[0][j] = isrc[k][s->order[j]];
        memcpy(isrc[k], s->t_i[0], sizeof(int) * n);
    }
}

/* -------------------------------------------------------------- forces ---- */
/* pair_1 (src/cuPairs.cu:117-167) evaluated from atom i's side against atom j whose
 * cell image is displaced by `sh` (delta = xi - xj - shift, src/cuPairs.cu:19-24).
 * Full-list evaluation: every unordered pair is seen twice with bitwise-negated delta, so
 * the per-atom sums equal the reference's f_A += f*delta / f_B -= f*delta. */
static inline void pair_accum(azo_state *s, int i, int j, const float sh[3],
                              double *fxi, double *fyi, double *fzi, double *fsc,
                              double *evdw, double *ecoul, long long *npairs)
{
    float dx = s->x[i] - s->x[j] - sh[0];
    float dy = s->y[i] - s->y[j] - sh[1];
    float dz = s->z[i] - s->z[j] - sh[2];
    /* delta.x*delta.x + delta.y*delta.y + delta.z*delta.z (src/cuPairs.cu:19-24) as nvcc -fmad
     * compiles it for sm_100a: FMUL(dy,dy), FFMA(dx,dx,.), FFMA(dz,dz,.) */
    float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
    if (r2 <= s->r2max) {
        float f = 0.f, r = 0.f, ev = 0.f, ec = 0.f;
        int ti = s->type[i], tj = s->type[j];
        if (s->elec.kind) f += coul_feng(&s->elec, r2, &r, s->chprod[ti][tj], &ec);
        const azo_vdw *v = &s->vdw[ti][tj];
        if (v->kind != AZ_VDW_NONE && r2 <= v->r2cut) {
            float ra = s->rad[i], rb = s->rad[j];
            if (v->use_radii && ti == tj) { float lo = ra < rb ? ra : rb, hi = ra < rb ? rb : ra; ra = lo; rb = hi; }
            f += vdw_feng(v, r2, &r, s->elec.kind != 0, ra, rb, &ev);
        }
        if (f > AZ_MX_FRC) f = AZ_MX_FRC; else if (f < -AZ_MX_FRC) f = -AZ_MX_FRC; /* force_autocap */
        *fxi += (double)(f * dx); *fyi += (double)(f * dy); *fzi += (double)(f * dz);
        *fsc += fabs((double)(f * dx)) + fabs((double)(f * dy)) + fabs((double)(f * dz));
        *evdw += 0.5 * (double)ev; *ecoul += 0.5 * (double)ec;
        (*npairs)++;
    }
}

static void forces(azo_state *s)
{
    const azo_grid *g = &s->g;
    double evdw = 0, ecoul = 0;
    long long npairs = 0;
    float rc2 = s->r2max;
    for (int cx = 0; cx < g->cnum[0]; cx++)
    for (int cy = 0; cy < g->cnum[1]; cy++)
    for (int cz = 0; cz < g->cnum[2]; cz++) {
        int c = cx * g->cnYZ + cy * g->cnum[2] + cz;
        int ib = s->first[c], ie = ib + s->count[c];
        if (ib == ie) continue;
        for (int i = ib; i < ie; i++) {
            double fxi = 0, fyi = 0, fzi = 0, fsc = 0;
            for (int ox = -s->sten[0]; ox <= s->sten[0]; ox++) {
                int nx, shx; if (!azo_neighbor(g->cnum[0], cx, ox, &nx, &shx)) continue;
                int aox = ox < 0 ? -ox : ox; float mx = aox ? (float)(aox - 1) * g->csize[0] : 0.f;
                for (int oy = -s->sten[1]; oy <= s->sten[1]; oy++) {
                    int ny, shy; if (!azo_neighbor(g->cnum[1], cy, oy, &ny, &shy)) continue;
                    int aoy = oy < 0 ? -oy : oy; float my = aoy ? (float)(aoy - 1) * g->csize[1] : 0.f;
                    for (int oz = -s->sten[2]; oz <= s->sten[2]; oz++) {
                        int nz, shz; if (!azo_neighbor(g->cnum[2], cz, oz, &nz, &shz)) continue;
                        int aoz = oz < 0 ? -oz : oz; float mz = aoz ? (float)(aoz - 1) * g->csize[2] : 0.f;
                        if ((double)mx * mx + (double)my * my + (double)mz * mz > (double)rc2) continue;
                        int nc = nx * g->cnYZ + ny * g->cnum[2] + nz;
                        float sh[3] = { (float)shx * s->L[0], (float)shy * s->L[1], (float)shz * s->L[2] };
                        int jb = s->first[nc], je = jb + s->count[nc];
                        for (int j = jb; j < je; j++) {
                            if (j == i) continue;
                            pair_accum(s, i, j, sh, &fxi, &fyi, &fzi, &fsc, &evdw, &ecoul, &npairs);
                        }
                    }
                }
            }
            s->fx[i] += (float)fxi; s->fy[i] += (float)fyi; s->fz[i] += (float)fzi;
            s->fscale[i] = (float)fsc;
        }
    }
    s->engVdW += evdw; s->engCoul1 += ecoul;
    s->n_pairs_in_range = npairs / 2;
}

/* ------------------------------------------------ bonds and angles (8f-4) ---- */
/* GROUNDWORK for SURVEY 8f-4 (no kernel exists yet): the arithmetic of the reference's constant bonds and valence angles,
 * pinned on the reference's own device source (tests/golden/ref_device_bonds_angles.npz).
 * Bond potentials, src/cuBonds.cu:1123-1224; kind = the reference's type id (define_bond_potential, :1236-1270):
 * 1 harmonic, 2 Morse, 3 Pedone (Morse + E/r^12), 4 Buckingham, 5 exp-6-8-12.  Returns -dU/dr / r, adds U to *eng.
 * Double-precision literals of the reference (2.0, 12.0, 1.0 / r2) promote their sub-expressions exactly as there. */
float azo_bond_eval(int kind, const float *p, float r2, float r, float *eng)
{
    const float p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3], p4 = p[4];
    switch (kind) {
    case 1: { float x = r - p1; *eng += 0.5f * p0 * x * x; return -p0 / r * x; }
    case 2: {
        float x = r - p2; x = expf(-p1 * x); float y = 1 - x;
        *eng += p0 * y * y - p3;
        return (float)(-2.0 * p0 * p1 * x * y / r);
    }
    case 3: {
        float x = r - p2; x = expf(-p1 * x); float y = 1 - x;
        float ir2 = 1.f / r2, irn = ir2 * ir2; irn = irn * irn * irn;
        *eng += p0 * y * y - p3 - p4 * irn;
        return (float)(-2.0 * p0 * p1 * x * y / r - 12.0 * p4 * irn * ir2);
    }
    case 4: {
        float ir2 = (float)(1.0 / r2), irn = ir2 * ir2;
        *eng += p0 * expf(-r / p1) - p2 * irn * ir2;
        return p0 * expf(-r / p1) / r / p1 - 6.f * p2 * irn * irn;
    }
    case 5: {
        float ir2 = (float)(1.0 / r2), irn = ir2 * ir2;
        *eng += p0 * expf(-r / p1) - p2 * irn * ir2 - p3 * irn * irn - p4 * irn * irn * irn;
        return (float)(p0 * expf(-r / p1) / r / p1 - 6.f * p2 * irn * irn - 8.f * p3 * irn * irn * ir2 - 12.0 * p4 * irn * irn * irn * ir2);
    }
    }
    return 0.f;
}

/* delta_periodic, src/cuMDfunc.cu:178-203: minimum image of one coordinate difference in a rectangular box */
static float delta_periodic1(float d, float L, float halfL) { if (d > halfL) d -= L; else if (d < -halfL) d += L; return d; }

/* angle_hcos, src/cuAngles.cu:230-283: U = k/2 (cos th - cos th0)^2 on (centre, ligand 1, ligand 2) = xyz[0..2];
 * frc[0..2] receive the three forces (the reference adds them with atomicAdd), *eng the energy */
void azo_angle_hcos(float k, float cos0, const float *xyz, const float *L, float *frc, float *eng)
{
    float ij[3], ik[3];
    for (int a = 0; a < 3; a++) {
        ij[a] = delta_periodic1(xyz[3 + a] - xyz[a], L[a], 0.5f * L[a]);
        ik[a] = delta_periodic1(xyz[6 + a] - xyz[a], L[a], 0.5f * L[a]);
    }
    float r2ij = ij[0] * ij[0] + ij[1] * ij[1] + ij[2] * ij[2], rij = sqrtf(r2ij);
    float r2ik = ik[0] * ik[0] + ik[1] * ik[1] + ik[2] * ik[2], rik = sqrtf(r2ik);
    float cos_th = (ij[0] * ik[0] + ij[1] * ik[1] + ij[2] * ik[2]) / rij / rik;
    float dCos = cos_th - cos0;
    float c1 = -k * dCos;
    float c2 = (float)(1.0 / rij / rik);
    for (int a = 0; a < 3; a++) {
        frc[a] = -c1 * (ik[a] * c2 + ij[a] * c2 - cos_th * (ij[a] / r2ij + ik[a] / r2ik));
        frc[3 + a] = c1 * (ik[a] * c2 - cos_th * ij[a] / r2ij);
        frc[6 + a] = c1 * (ij[a] * c2 - cos_th * ik[a] / r2ik);
    }
    *eng += 0.5f * k * dCos * dCos;
}

/* apply_const_bonds, src/cuBonds.cu:709-797 (serial twin: exec_bondlist, src/bonds.cpp:1069-1218, constant bonds): for every
 * bond (id1, id2, type) delta = x1 - x2 through delta_periodic, f = force_eng(r2, r), frc[id1] += f delta, frc[id2] -= f delta.
 * Plain arrays (the boundary for bonds does not exist yet); forces are accumulated in double per atom like forces() does;
 * returns the bond energy. */
double azo_apply_const_bonds(int n_bonds, const int *at1, const int *at2, const int *btype, const int *type_kind, const float *type_par /*5 per type*/,
                             int n, const float *x, const float *y, const float *z, const float *L, double *fx, double *fy, double *fz)
{
    double eng = 0.0;
    (void)n;
    for (int b = 0; b < n_bonds; b++) {
        const int i = at1[b], j = at2[b], t = btype[b];
        if (t == 0) continue; /* empty bond, src/cuBonds.cu:731 */
        float dx = delta_periodic1(x[i] - x[j], L[0], 0.5f * L[0]);
        float dy = delta_periodic1(y[i] - y[j], L[1], 0.5f * L[1]);
        float dz = delta_periodic1(z[i] - z[j], L[2], 0.5f * L[2]);
        float r2 = dx * dx + dy * dy + dz * dz, r = sqrtf(r2), e = 0.f;
        float f = azo_bond_eval(type_kind[t], type_par + 5 * t, r2, r, &e);
        eng += (double)e;
        fx[i] += (double)(f * dx); fy[i] += (double)(f * dy); fz[i] += (double)(f * dz);
        fx[j] -= (double)(f * dx); fy[j] -= (double)(f * dy); fz[j] -= (double)(f * dz);
    }
    return eng;
}

/* apply_angles, src/cuAngles.cu:169-228 (serial twin: exec_anglelist, src/angles.cpp:229-242): for every valence angle
 * (centre, ligand 1, ligand 2, type) the harmonic-cosine term; type_par = (k, cos0) per type, type 0 = empty.  Returns
 * the angle energy. */
double azo_apply_angles(int n_angles, const int *centre, const int *lig1, const int *lig2, const int *atype, const float *type_par /*2 per type*/,
                        const float *x, const float *y, const float *z, const float *L, double *fx, double *fy, double *fz)
{
    double eng = 0.0;
    for (int a = 0; a < n_angles; a++) {
        const int t = atype[a];
        if (t == 0) continue;
        const int id[3] = { centre[a], lig1[a], lig2[a] };
        float xyz[9], f[9], e = 0.f;
        for (int k = 0; k < 3; k++) { xyz[3 * k] = x[id[k]]; xyz[3 * k + 1] = y[id[k]]; xyz[3 * k + 2] = z[id[k]]; }
        azo_angle_hcos(type_par[2 * t], type_par[2 * t + 1], xyz, L, f, &e);
        eng += (double)e;
        for (int k = 0; k < 3; k++) { fx[id[k]] += (double)f[3 * k]; fy[id[k]] += (double)f[3 * k + 1]; fz[id[k]] += (double)f[3 * k + 2]; }
    }
    return eng;
}

/* The bonded terms of one step on the sorted arrays (src/main.cu:307-311 apply_const_bonds, :352-362 apply_angles): the
 * lists name persistent ids, `where` takes them to the current order.  The reference adds them with float atomics; here
 * per-atom sums are formed in double and narrowed once, like forces() does. */
static void bonded(azo_state *s)
{
    s->engBond = 0.0; s->engAngle = 0.0;
    if (!s->n_bonds && !s->n_angles) return;
    const int n = s->n;
    for (int i = 0; i < n; i++) s->where[s->id[i]] = i;
    double *f = (double *)xmalloc(sizeof(double) * 3 * n), *fx = f, *fy = f + n, *fz = f + 2 * n;
    if (s->n_bonds) {
        int *c1 = (int *)xmalloc(sizeof(int) * s->n_bonds), *c2 = (int *)xmalloc(sizeof(int) * s->n_bonds);
        for (int b = 0; b < s->n_bonds; b++) { c1[b] = s->where[s->b_at1[b]]; c2[b] = s->where[s->b_at2[b]]; }
        s->engBond = azo_apply_const_bonds(s->n_bonds, c1, c2, s->b_type, s->bt_kind, s->bt_par, n, s->x, s->y, s->z, s->L, fx, fy, fz);
        free(c1); free(c2);
    }
    if (s->n_angles) {
        int *c0 = (int *)xmalloc(sizeof(int) * s->n_angles), *c1 = (int *)xmalloc(sizeof(int) * s->n_angles), *c2 = (int *)xmalloc(sizeof(int) * s->n_angles);
        for (int a = 0; a < s->n_angles; a++) { c0[a] = s->where[s->a_c[a]]; c1[a] = s->where[s->a_l1[a]]; c2[a] = s->where[s->a_l2[a]]; }
        s->engAngle = azo_apply_angles(s->n_angles, c0, c1, c2, s->a_type, s->at_par, s->x, s->y, s->z, s->L, fx, fy, fz);
        free(c0); free(c1); free(c2);
    }
    for (int i = 0; i < n; i++) {
        s->fx[i] += (float)fx[i]; s->fy[i] += (float)fy[i]; s->fz[i] += (float)fz[i];
        s->fscale[i] += (float)(fabs(fx[i]) + fabs(fy[i]) + fabs(fz[i])); /* the bonded terms join the scale the force tolerance is relative to */
    }
    free(f);
}

/* ------------------------------------------------- reciprocal-space Ewald ---- */
#define FCOUL_SCALE_D 14.3996 /* the engine's Coulomb constant (d_Fcoul_scale, src/main.cu:40); the serial twin's
                               * Fcoul_scale (src/const.h:44) is the same number to 6 digits */
#define PI_D 3.14159265359    /* src/const.h:11 */

/* ewald_const (src/elec.cpp:144-165): the self-energy and the charged-system term, computed once */
double azo_ewald_const(int n, const int *types, const az_species *sp, double alpha, double eps, double vol)
{
    double sq = 0.0, eng = 0.0;
    for (int i = 0; i < n; i++) { double q = sp[types[i]].charge; sq += q; eng += q * q; }
    eng *= -alpha / sqrt(PI_D);
    double q = -0.5 * PI_D * (sq * sq / alpha / alpha) / vol;
    return FCOUL_SCALE_D * (eng + q) / eps;
}

/* The k-vectors of ewald_rec in its own enumeration order (src/elec.cpp:234-324): l in [0,kx), m from 0 (l = 0) or
 * 1-ky, n from 1 (l = m = 0 ... the reference's nmin/mmin switches) or 1-kz, kept if |k|^2 < rkcut2 with
 * rkcut = 1.05 * 2 pi * min(kx/a, ky/b, kz/c) (prepare_elec, src/elec.cpp:383-393).  Returns the count;
 * lmn[3*k..], rk[3*k..] and ak[k] = exp(-|k|^2/(4 alpha^2))/|k|^2 are filled up to cap entries. */
int azo_ewald_kvecs(double la, double lb, double lc, double alpha, int kx, int ky, int kz, int cap, int *lmn, double *rk, double *ak)
{
    const double twopi = 2.0 * PI_D;
    const double ra = 1.0 / la, rb = 1.0 / lb, rc = 1.0 / lc;
    double rkcut = kx * ra;
    if (rkcut > ky * rb) rkcut = ky * rb;
    if (rkcut > kz * rc) rkcut = kz * rc;
    rkcut *= twopi * 1.05;
    const double rkcut2 = rkcut * rkcut, mr4a2 = -0.25 / alpha / alpha;
    int mmin = 0, nmin = 1, ik = 0;
    for (int l = 0; l < kx; l++) {
        double rkx = l * twopi * ra;
        for (int m = mmin; m < ky; m++) {
            double rky = m * twopi * rb;
            for (int n = nmin; n < kz; n++) {
                double rkz = n * twopi * rc;
                double rk2 = rkx * rkx + rky * rky + rkz * rkz;
                if (rk2 < rkcut2) {
                    if (ik < cap) {
                        if (lmn) { lmn[3 * ik] = l; lmn[3 * ik + 1] = m; lmn[3 * ik + 2] = n; }
                        if (rk) { rk[3 * ik] = rkx; rk[3 * ik + 1] = rky; rk[3 * ik + 2] = rkz; }
                        if (ak) ak[ik] = exp(rk2 * mr4a2) / rk2;
                    }
                    ik++;
                }
            }
            nmin = 1 - kz;
        }
        mmin = 1 - ky;
    }
    return ik;
}

/* ewald_rec (src/elec.cpp:167-336) on the current (float) positions, in double like the serial twin: structure factor
 * S(k) = sum_i q_i exp(i k r_i), energy scale * sum_k ak |S|^2 with scale = 4 pi Fcoul / (V eps), force on atom i
 * 2 scale * sum_k ak (Im c_ik Re S - Re c_ik Im S) k, c_ik = q_i exp(i k r_i).  exp(i k r) is built from
 * exp(i 2 pi x/a) by repeated complex multiplication exactly as the reference fills its elc/emc/enc tables. */
static void ewald_rec(azo_state *s)
{
    const int kx = s->ew_k[0], ky = s->ew_k[1], kz = s->ew_k[2], n = s->n;
    if (kx <= 0 && ky <= 0 && kz <= 0) return;
    const double twopi = 2.0 * PI_D;
    int nk = azo_ewald_kvecs(s->boxd[0], s->boxd[1], s->boxd[2], s->ew_alpha, kx, ky, kz, 0, NULL, NULL, NULL);
    int *lmn = (int *)xmalloc(sizeof(int) * 3 * nk);
    double *rk = (double *)xmalloc(sizeof(double) * 3 * nk), *ak = (double *)xmalloc(sizeof(double) * nk);
    azo_ewald_kvecs(s->boxd[0], s->boxd[1], s->boxd[2], s->ew_alpha, kx, ky, kz, nk, lmn, rk, ak);
    const int K[3] = { kx > 1 ? kx : 2, ky > 1 ? ky : 2, kz > 1 ? kz : 2 };
    double *tc[3], *ts[3];
    for (int a = 0; a < 3; a++) { tc[a] = (double *)xmalloc(sizeof(double) * n * K[a]); ts[a] = (double *)xmalloc(sizeof(double) * n * K[a]); }
    const float *pos[3] = { s->x, s->y, s->z };
    for (int a = 0; a < 3; a++)
        for (int i = 0; i < n; i++) {
            double *c = tc[a] + (size_t)i * K[a], *sn = ts[a] + (size_t)i * K[a];
            c[0] = 1.0; sn[0] = 0.0;
            double ang = twopi * (double)pos[a][i] / s->boxd[a];
            c[1] = cos(ang); sn[1] = sin(ang);
            for (int l = 2; l < K[a]; l++) { c[l] = c[l - 1] * c[1] - sn[l - 1] * sn[1]; sn[l] = sn[l - 1] * c[1] + c[l - 1] * sn[1]; }
        }
    const double scale = 2.0 * twopi / (s->boxd[0] * s->boxd[1] * s->boxd[2]) * FCOUL_SCALE_D / s->ew_eps, scale2 = 2.0 * scale;
    double *ckc = (double *)xmalloc(sizeof(double) * n), *cks = (double *)xmalloc(sizeof(double) * n);
    double *fx = (double *)xmalloc(sizeof(double) * n), *fy = (double *)xmalloc(sizeof(double) * n), *fz = (double *)xmalloc(sizeof(double) * n);
    double eng = 0.0;
    for (int k = 0; k < nk; k++) {
        const int l = lmn[3 * k], m = lmn[3 * k + 1], nn = lmn[3 * k + 2];
        const int am = m < 0 ? -m : m, an = nn < 0 ? -nn : nn;
        const double sm = m < 0 ? -1.0 : 1.0, sn_ = nn < 0 ? -1.0 : 1.0; /* negative index: complex conjugate */
        double sumC = 0.0, sumS = 0.0;
        for (int i = 0; i < n; i++) {
            const double ch = (double)s->charge[s->type[i]];
            const double elc = tc[0][(size_t)i * K[0] + l], els = ts[0][(size_t)i * K[0] + l];
            const double emc = tc[1][(size_t)i * K[1] + am], ems = sm * ts[1][(size_t)i * K[1] + am];
            const double enc = tc[2][(size_t)i * K[2] + an], ens = sn_ * ts[2][(size_t)i * K[2] + an];
            const double lmc = elc * emc - els * ems, lms = els * emc + ems * elc;
            ckc[i] = ch * (lmc * enc - lms * ens);
            cks[i] = ch * (lms * enc + lmc * ens);
            sumC += ckc[i]; sumS += cks[i];
        }
        eng += ak[k] * (sumC * sumC + sumS * sumS);
        for (int i = 0; i < n; i++) {
            const double x = ak[k] * (cks[i] * sumC - ckc[i] * sumS) * scale2;
            fx[i] += rk[3 * k] * x; fy[i] += rk[3 * k + 1] * x; fz[i] += rk[3 * k + 2] * x;
        }
    }
    for (int i = 0; i < n; i++) {
        s->frx[i] = (float)fx[i]; s->fry[i] = (float)fy[i]; s->frz[i] = (float)fz[i];
        s->fx[i] += s->frx[i]; s->fy[i] += s->fry[i]; s->fz[i] += s->frz[i];
    }
    s->engCoul2 = scale * eng;
    for (int a = 0; a < 3; a++) { free(tc[a]); free(ts[a]); }
    free(lmn); free(rk); free(ak); free(ckc); free(cks); free(fx); free(fy); free(fz);
}

/* the k-space force of the last force evaluation alone, sorted order */
int azo_ewald_forces(azo_state *s, float *fx, float *fy, float *fz)
{
    if (!s->frx) return AZ_ERR_INVALID;
    memcpy(fx, s->frx, sizeof(float) * s->n); memcpy(fy, s->fry, sizeof(float) * s->n); memcpy(fz, s->frz, sizeof(float) * s->n);
    return 0;
}

/* -------------------------------------------------------------- stage 2 ---- */
static void stage2(azo_state *s)
{
    double ke = 0;
    for (int i = 0; i < s->n; i++) {
        int t = s->type[i]; float h = s->rMasshdT[t];
        s->vx[i] = fmaf(h, s->fx[i], s->vx[i]); s->vy[i] = fmaf(h, s->fy[i], s->vy[i]); s->vz[i] = fmaf(h, s->fz[i], s->vz[i]);
        ke += (double)((s->vx[i] * s->vx[i] + s->vy[i] * s->vy[i] + s->vz[i] * s->vz[i]) * s->mass[t]);
    }
    s->engKin = 0.5 * ke;
    s->engKin_f = (float)s->engKin;
}

/* ----------------------------------------------------------- thermostat ---- */
#define REV_LIGHT 3.33567e-5f /* src/cuTemp.cu:225 */
#define NUM_PI 3.14159f       /* src/cuTemp.cu:228 */

static float f3len(const float *v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static float f3sqr(const float *v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }

/* get_angled_vector, src/cuTemp.cu:395-453 */
void azo_angled_vector(const float *invec, float cos_phi, float theta, float *out)
{
    float v1[3], v2[3], v3[3];
    float l1 = f3len(invec);
    v1[0] = invec[0] / l1; v1[1] = invec[1] / l1; v1[2] = invec[2] / l1;
    if (v1[0] != 0.f) { v2[1] = 1.f; v2[2] = 1.f; v2[0] = -(v1[1] * v2[1] + v1[2] * v2[2]) / v1[0]; }
    else if (v1[1] != 0.f) { v2[0] = 1.f; v2[2] = 1.f; v2[1] = -(v1[2] * v2[2]) / v1[1]; }
    else { v2[0] = 1.f; v2[1] = 0.f; v2[2] = 0.f; }
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1];
    v3[1] = -v1[0] * v2[2] + v1[2] * v2[0];
    v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    float l2 = f3len(v2), l3 = f3len(v3);
    for (int a = 0; a < 3; a++) { v2[a] /= l2; v3[a] /= l3; }
    float sinPhi = sqrtf(1 - cos_phi * cos_phi);
    float sinTh = sinf(theta), cosTh = cosf(theta);
    for (int a = 0; a < 3; a++) out[a] = v1[a] * cos_phi + sinPhi * (cosTh * v2[a] + sinTh * v3[a]);
}

/* One atom through the body of tstat_radi9 (src/cuTemp.cu:736-758): adsorb_rand_photon (:484-507) with the unit vector
 * uv[3 * (rn1 % 3072)], then, if U > 1e-4, radiate_photon3 (:631-685) with the draws rn2, rn3.  Exported so that
 * tests can feed the same draws to the reference's own source compiled on the CPU (oracle/ref_device_harness.cpp). */
void azo_tstat_atom(float mass, float pe, uint32_t rn1, uint32_t rn2, uint32_t rn3, const float *uv, float *v, float *U_io)
{
    float U = *U_io;
    { /* adsorb_rand_photon, src/cuTemp.cu:484-507 */
        float v02 = f3sqr(v);
        const float *u = &uv[3 * (rn1 % AZ_N_UVECT)];
        float ermc = pe * REV_LIGHT / mass;
        v[0] += ermc * u[0]; v[1] += ermc * u[1]; v[2] += ermc * u[2];
        float v12 = f3sqr(v);
        U += pe + 0.5f * mass * (v02 - v12);
    }
    if ((double)U > 1e-4) { /* radiate_photon3, src/cuTemp.cu:631-685 */
        float u0 = U;
        float v02 = f3sqr(v);
        float v0 = sqrtf(v02);
        float ph_eng = 0.9f * u0;
        float ermc = ph_eng * REV_LIGHT / mass;
        float ermcv0 = ermc / v0;
        float rv[3];
        if (!(ermcv0 < 1.f)) { /* fix of the uninitialised vector: anti-parallel to v */
            if (v0 > 0.f) { rv[0] = -v[0] / v0; rv[1] = -v[1] / v0; rv[2] = -v[2] / v0; }
            else { const float *u = &uv[3 * (rn1 % AZ_N_UVECT)]; rv[0] = u[0]; rv[1] = u[1]; rv[2] = u[2]; }
        } else {
            float cos_phi = (float)(rn2 % 2048u) / 1024.f * (1.f - ermcv0);
            cos_phi -= 1.f;
            float theta = (float)(rn3 % 2048u) / 1024.f * NUM_PI;
            azo_angled_vector(v, cos_phi, theta, rv);
        }
        v[0] += ermc * rv[0]; v[1] += ermc * rv[1]; v[2] += ermc * rv[2];
        float v12 = f3sqr(v);
        U -= (ph_eng + 0.5f * mass * (v12 - v02));
    }
    *U_io = U;
}

static void tstat_radi(azo_state *s, int step)
{
    double teng = 0;
    for (int i = 0; i < s->n; i++) {
        int tp = s->type[i];
        float pe = s->photons[i]; /* sorted index, in-order case of src/cuTemp.cu:715-722 */
        uint32_t rn[4];
        philox4x32_10((uint32_t)s->id[i], (uint32_t)step, 0x52414449u /* "RADI" */, 0,
                      (uint32_t)s->seed, (uint32_t)(s->seed >> 32), rn);
        /* rn[0] is the burnt draw of src/cuTemp.cu:738 */
        float v[3] = { s->vx[i], s->vy[i], s->vz[i] };
        float U = s->eng[i];
        azo_tstat_atom(s->mass[tp], pe, rn[1], rn[2], rn[3], s->uv, v, &U);
        s->vx[i] = v[0]; s->vy[i] = v[1]; s->vz[i] = v[2];
        s->eng[i] = U;
        float restrE = U < s->mxEng[tp] ? U : s->mxEng[tp];
        s->rad[i] = s->radA[tp] / (s->radB[tp] - restrE);
        teng += (double)U;
    }
    s->engTemp += teng;
}

/* before_nose + tstat_nose + after_nose(1) (src/cuTemp.cu:115-163): the Nose-Hoover friction is advanced with the
 * current kinetic energy, velocities are scaled, the kinetic energy follows the scaling and the friction is advanced
 * again with the new kinetic energy.  The reference's main applies this block twice per step: before verlet_1stage
 * (src/main.cu:291-299) and inside apply_tstat (src/cuTemp.cu:787-794). */
void azo_nose_scalars(float tSt, float teKin, float rQmass, float qMassTau2, float *engKin, float *chit, float *consInt, float *tscale)
{
    *chit += tSt * (*engKin - teKin) * rQmass;          /* before_nose */
    *tscale = 1.f - tSt * *chit;
    *engKin = *engKin * *tscale * *tscale;              /* after_nose(refresh_kin = 1) */
    *consInt += tSt * *chit * qMassTau2;
    *chit += tSt * (*engKin - teKin) * rQmass;
}

static void nose(azo_state *s)
{
    float ts;
    azo_nose_scalars(s->tSt, s->teKin, s->rQmass, s->qMassTau2, &s->engKin_f, &s->chit, &s->consInt, &ts);
    for (int i = 0; i < s->n; i++) { s->vx[i] *= ts; s->vy[i] *= ts; s->vz[i] *= ts; }
    s->engKin = (double)s->engKin_f;
}

static void apply_tstat(azo_state *s, int step)
{
    if (s->nEq && step <= s->nEq && s->freqEq > 0 && (step % s->freqEq) == 0) {
        if (s->engKin_f != 0.f) { /* temp_scale, src/cuTemp.cu:77-107 */
            float c = (s->tstat == AZ_TSTAT_RADI) ? 0.25f : 1.f;
            float k = sqrtf(c * s->teKin / s->engKin_f);
            for (int i = 0; i < s->n; i++) { s->vx[i] *= k; s->vy[i] *= k; s->vz[i] *= k; }
        }
        s->engKin_f = s->teKin; /* after_tscale, src/cuTemp.cu:109-113 */
        s->engKin = (double)s->teKin;
    }
    if (s->tstat == AZ_TSTAT_RADI) tstat_radi(s, step);
    if (s->tstat == AZ_TSTAT_NOSE) nose(s);
}

/* calc_quantities, src/main.cu:121-162 */
static void calc_quantities(azo_state *s, int step)
{
    float engCoulTot = (float)s->engCoul1 + (float)s->engCoul2 + (float)s->engCoul3; /* src/main.cu:127 */
    s->engPot = engCoulTot + (float)s->engVdW + (float)s->engBond + (float)s->engAngle; /* src/main.cu:128 */
    s->engTot = s->engPot + s->engKin_f;
    float pm[3] = { (float)s->posMom[0], (float)s->posMom[1], (float)s->posMom[2] };
    float nm[3] = { (float)s->negMom[0], (float)s->negMom[1], (float)s->negMom[2] };
    if (step >= AZ_N_MOM - 1) {
        int i = s->iMom;
        float k = (float)(2.f * 1.58e6 / (double)(s->tSt * (float)(AZ_N_MOM - 1)));
        float pp[3], np[3];
        for (int a = 0; a < 3; a++) {
            pp[a] = k * (pm[a] - s->posMomBuf[i][a]) * s->revArea[a];
            np[a] = k * (nm[a] - s->negMomBuf[i][a]) * s->revArea[a];
        }
        i--; if (i < 0) i = AZ_N_MOM - 1;
        for (int a = 0; a < 3; a++) { s->posMomBuf[i][a] = pm[a]; s->negMomBuf[i][a] = nm[a]; }
        s->iMom++; if (s->iMom >= AZ_N_MOM) s->iMom = 0;
        s->pressure = (pp[0] + pp[1] + pp[2] + np[0] + np[1] + np[2]) / 6.f;
    } else {
        for (int a = 0; a < 3; a++) { s->posMomBuf[step][a] = pm[a]; s->negMomBuf[step][a] = nm[a]; }
    }
}

/* One iteration of the loop body src/main.cu:281-383 */
int azo_step(azo_state *s, int first_step, int nsteps)
{
    for (int k = 0; k < nsteps; k++) {
        int step = first_step + k;
        s->engCoul1 = 0; s->engVdW = 0; s->engTemp = 0; s->engElecField = 0; /* reset_quantities */
        if (s->tstat == AZ_TSTAT_NOSE) nose(s);           /* src/main.cu:291-299 */
        stage1(s);
        sort_atoms(s);
        forces(s);
        ewald_rec(s);                                     /* add_elec, src/main.cpp:99 (src/main.cu:330-336) */
        bonded(s);                                        /* apply_const_bonds / apply_angles, src/main.cu:307-311,352-362 */
        if (s->reset_vels && step % s->reset_vels == 0) { /* zero_vel, src/main.cu:372-374, src/cuMDfunc.cu:602-622 */
            for (int i = 0; i < s->n; i++) { s->vx[i] = 0.f; s->vy[i] = 0.f; s->vz[i] = 0.f; }
            s->engKin = 0.0; s->engKin_f = 0.f;
        } else
        stage2(s);
        apply_tstat(s, step);
        calc_quantities(s, step);
        s->last_step = step;
    }
    return 0;
}

/* Force evaluation only, on the current positions (bins + sorts first; no integration).
 * Used to pin the oracle against the reference's serial build and by the parity tests. */
int azo_forces_only(azo_state *s)
{
    s->engCoul1 = 0; s->engVdW = 0;
    for (int i = 0; i < s->n; i++) {
        s->cell[i] = azo_cell_index(s->x[i], s->y[i], s->z[i], s->g.crev, s->g.cnum);
        int t = s->type[i]; float q = s->charge[t];
        s->fx[i] = -q * s->efield[0]; s->fy[i] = -q * s->efield[1]; s->fz[i] = -q * s->efield[2];
    }
    sort_atoms(s);
    forces(s);
    ewald_rec(s);
    bonded(s);
    return 0;
}

int azo_stats(azo_state *s, az_stats *o)
{
    memset(o, 0, sizeof(*o));
    o->step = s->last_step; o->n_atoms_total = s->n;
    o->engTot = s->engTot; o->engKin = s->engKin_f; o->engVdW = (float)s->engVdW;
    o->engCoul1 = (float)s->engCoul1; o->engCoul2 = (float)s->engCoul2; o->engCoul3 = (float)s->engCoul3; o->engCoul2_d = s->engCoul2; o->engTemp = (float)s->engTemp;
    o->engElecField = (float)s->engElecField; o->engPot = s->engPot;
    for (int a = 0; a < 3; a++) { o->posMom[a] = (float)s->posMom[a]; o->negMom[a] = (float)s->negMom[a]; }
    o->pressure = s->pressure;
    o->x_first = s->x[0];
    for (int t = 0; t < s->nspec; t++) for (int a = 0; a < 3; a++) {
        o->specAcBoxPos[t][a] = s->acPos[t][a]; o->specAcBoxNeg[t][a] = s->acNeg[t][a];
    }
    o->engKin_d = s->engKin; o->engVdW_d = s->engVdW; o->engCoul1_d = s->engCoul1; o->engTemp_d = s->engTemp;
    o->engBond = (float)s->engBond; o->engAngle = (float)s->engAngle; o->engBond_d = s->engBond; o->engAngle_d = s->engAngle;
    return 0;
}

int azo_download(azo_state *s, az_atoms_out *o, int order)
{
    if (o->n_atoms < s->n) return AZ_ERR_INVALID;
    for (int i = 0; i < s->n; i++) {
        int k = (order == AZ_ORDER_ORIGINAL) ? s->id[i] : i;
        if (o->xs) { o->xs[k] = s->x[i]; o->ys[k] = s->y[i]; o->zs[k] = s->z[i]; }
        if (o->vxs) { o->vxs[k] = s->vx[i]; o->vys[k] = s->vy[i]; o->vzs[k] = s->vz[i]; }
        if (o->types) o->types[k] = s->type[i];
        if (o->ids) o->ids[k] = s->id[i];
        if (o->engs) o->engs[k] = s->eng[i];
        if (o->radii) o->radii[k] = s->rad[i];
    }
    o->n_written = s->n;
    return 0;
}

int azo_dump(azo_state *s, az_dump *o)
{
    if (o->n_atoms < s->n || o->n_cells < s->g.ncell) return AZ_ERR_INVALID;
    o->n_atoms = s->n; o->n_cells = s->g.ncell;
    for (int a = 0; a < 3; a++) { o->cnum[a] = s->g.cnum[a]; o->csize[a] = s->g.csize[a]; o->crev[a] = s->g.crev[a]; }
    for (int i = 0; i < s->n; i++) {
        if (o->cell_indexes) o->cell_indexes[i] = s->cell[i];
        if (o->ids) o->ids[i] = s->id[i];
        if (o->fx) { o->fx[i] = s->fx[i]; o->fy[i] = s->fy[i]; o->fz[i] = s->fz[i]; }
        if (o->x) { o->x[i] = s->x[i]; o->y[i] = s->y[i]; o->z[i] = s->z[i]; }
        if (o->vx) { o->vx[i] = s->vx[i]; o->vy[i] = s->vy[i]; o->vz[i] = s->vz[i]; }
    }
    for (int c = 0; c < s->g.ncell; c++) {
        if (o->first_atom_in_cell) o->first_atom_in_cell[c] = s->first[c];
        if (o->n_at_in_cell) o->n_at_in_cell[c] = s->count[c];
    }
    return 0;
}

/* per-atom sum of |f*delta| of the last force evaluation (the fp32 tolerance scale) */
int azo_force_scale(azo_state *s, float *out) { memcpy(out, s->fscale, sizeof(float) * s->n); return 0; }
long long azo_pairs_in_range(azo_state *s) { return s->n_pairs_in_range; }
int azo_stencil(azo_state *s, int *out) { out[0] = s->sten[0]; out[1] = s->sten[1]; out[2] = s->sten[2]; return 0; }

/* Minimum-image all-pairs reference (O(N^2), double accumulation) for small systems:
 * an independent check of the cell-stencil traversal above. */
int azo_forces_allpairs(azo_state *s, float *fx, float *fy, float *fz, double *evdw_out, double *ecoul_out)
{
    double evdw = 0, ecoul = 0; long long np = 0;
    for (int i = 0; i < s->n; i++) {
        double ax = 0, ay = 0, az = 0, sc = 0;
        for (int j = 0; j < s->n; j++) {
            if (i == j) continue;
            float sh[3];
            float d[3] = { s->x[i] - s->x[j], s->y[i] - s->y[j], s->z[i] - s->z[j] };
            for (int a = 0; a < 3; a++) {
                float h = 0.5f * s->L[a];
                sh[a] = d[a] > h ? s->L[a] : (d[a] < -h ? -s->L[a] : 0.f);
            }
            pair_accum(s, i, j, sh, &ax, &ay, &az, &sc, &evdw, &ecoul, &np);
        }
        int t = s->type[i]; float q = s->charge[t];
        fx[i] = (float)((double)(-q * s->efield[0]) + ax);
        fy[i] = (float)((double)(-q * s->efield[1]) + ay);
        fz[i] = (float)((double)(-q * s->efield[2]) + az);
    }
    *evdw_out = evdw; *ecoul_out = ecoul;
    return 0;
}


/* brute_rdf (src/cuStat.cu:436-512) restated: every unordered pair, minimum image by delta_periodic
 * (src/cuMDfunc.cu:178-203), r2 < r2max, bin = (int)(sqrt((double)r2) * idRDF), pair index
 * mn * (nSpec - 1) + mn * (1 - mn) / 2 + mx.  Integer counts (the reference adds volume / (N_a N_b) per pair in float;
 * the normalisation of copy_rdf is applied by the caller).  counts[bin * npair + pair] += 1. */
int azo_rdf_counts(azo_state *s, float r_max, float dr, unsigned long long *counts, int nbins)
{
    const float idrdf = (float)(1.0 / (double)dr);
    const float r2max = (float)((double)r_max * (double)r_max);
    const int m = s->nspec - 1, npair = s->nspec * (s->nspec + 1) / 2;
    const float half[3] = { 0.5f * s->L[0], 0.5f * s->L[1], 0.5f * s->L[2] };
    for (int i = 0; i < s->n - 1; i++)
        for (int j = i + 1; j < s->n; j++) {
            float d[3] = { s->x[i] - s->x[j], s->y[i] - s->y[j], s->z[i] - s->z[j] };
            for (int a = 0; a < 3; a++) {
                if (d[a] > half[a]) d[a] -= s->L[a];
                else if (d[a] < -half[a]) d[a] += s->L[a];
            }
            float r2 = fmaf(d[2], d[2], fmaf(d[0], d[0], d[1] * d[1])); /* nvcc contracts dx*dx + dy*dy + dz*dz like this (see the force path) */
            if (r2 < r2max) {
                int iR = (int)(sqrt((double)r2) * (double)idrdf);
                int mn = s->type[i], mx = s->type[j];
                if (mn > mx) { int t = mn; mn = mx; mx = t; }
                int iPair = mn * m + mn * (1 - mn) / 2 + mx;
                if (iR < nbins) counts[(size_t)iR * npair + iPair]++;
            }
        }
    return 0;
}
