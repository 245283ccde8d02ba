// ref_device_harness.cpp -- TEST INFRASTRUCTURE.  The reference's own __device__ functions of the hot path,
// executed on the CPU: oracle/build_ref_device.sh extracts the function bodies VERBATIM from the reference's .cu
// files (into a temporary include file, never into the repository) and compiles them here as host code.  This pins
// the arithmetic that exists only in the reference's CUDA build -- the radius/temperature-dependent pair function and
// the radiative thermostat -- against the reference's source lines, since the CUDA build itself does not compute on
// the B200 box.  tests/golden/make_golden_device.py runs this library and commits the vectors.
//
//   cu_fe_* / cu_fer_* / surk_pot               src/cuVdW.cu:25-257
//   direct_coul, real_ewald, fennel             src/cuElec.cu:27-37,94-134
//   get_angled_vector, rand_uvect,
//   adsorb_rand_photon, radiate_photon3         src/cuTemp.cu:230-234,395-453,484-507,631-685
//   before_nose, after_nose (1-thread kernels)  src/cuTemp.cu:115-123,156-163
//   float3_sqr, float3_length, ...              src/cuUtils.cu:27-75
//   put_periodic                                src/cuMDfunc.cu:35-106
//   count_cell                                  src/cuSort.cu:114-128
//   split_cells, cell_dist_int (host code)      src/cuCellList.cu:9-34,98-132
//   bond_harm/_morse/_pedone/_buck/_e6812       src/cuBonds.cu:1123-1224     (SURVEY 8f-4 groundwork: oracle pinned
//   angle_hcos, delta_periodic                  src/cuAngles.cu:230-283, src/cuMDfunc.cu:178-203   before kernels exist)
//
// The only substitution: rnd_xor128 (src/cuUtils.cu:89-105; racy, seeded from uninitialised memory) is replaced by a
// queue of numbers the caller supplies, so that the same draws can be fed to the oracle.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <cuda_runtime.h> // float3, uint4, make_float3 (host-side vector types only)
#undef __device__
#undef __global__
#undef __host__
#undef __constant__
#define __device__
#define __global__
#define __host__
#define __constant__ static

#include "cuStruct.h" // the reference's cudaMD / cudaVdW (oracle/build_ref_device.sh passes -I to its directory)

static const float d_Fcoul_scale = 14.3996f; // src/main.cu:40
static const float d_sqrtpi = 1.772453f;     // src/main.cu:41
static const int dnUvect = 3072;             // src/cuTemp.h:6

static unsigned int g_rnd[64];
static int g_rnd_n = 0, g_rnd_pos = 0;
unsigned int rnd_xor128(uint4&) { return g_rnd_pos < g_rnd_n ? g_rnd[g_rnd_pos++] : 0u; }

// the serial stand-ins of the device atomics used by put_periodic / count_cell
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p += v; return o; }

// CUDA's float overloads of the math library used by the extracted bodies
static inline void sincos(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }

#include "ref_device_funcs.inc" // verbatim function bodies, generated at build time

extern "C" {

// kind: the reference's vdW type ids (src/vdw.h:15-21); with_r != 0 selects the *_r variant (r known from Coulomb)
int refdev_vdw(int kind, const float* p, float r2, float r_in, int with_r, float rad1, float rad2, float* f_over_r, float* eng)
{
    cudaVdW v;
    memset(&v, 0, sizeof(v));
    v.p0 = p[0]; v.p1 = p[1]; v.p2 = p[2]; v.p3 = p[3]; v.p4 = p[4];
    float e = 0.f, r = r_in, f;
    switch (kind) {
    case 1: f = with_r ? cu_fer_lj(r2, r, &v, e) : cu_fe_lj(r2, &v, e); break;
    case 2: f = with_r ? cu_fer_buck(r2, r, &v, e) : cu_fe_buck(r2, &v, e); break;
    case 4: f = with_r ? cu_fer_bmh(r2, r, &v, e) : cu_fe_bmh(r2, &v, e); break;
    case 5: f = with_r ? cu_fer_elin(r2, r, &v, e) : cu_fe_elin(r2, &v, e); break;
    case 6: f = with_r ? cu_fer_einv(r2, r, &v, e) : cu_fe_einv(r2, &v, e); break;
    case 7: f = surk_pot(r2, rad1, rad2, &v, e); break;
    default: return -1;
    }
    *f_over_r = f; *eng = e;
    return 0;
}

// kind: 1 direct, 2 Ewald real part, 3 Fennell (DSF); returns r through r_out like the reference
int refdev_coul(int kind, float r2, float chprd, float alpha, float daipi2, float elC1, float elC2, float rElec,
                float* f_over_r, float* eng, float* r_out)
{
    static cudaMD md;
    md.alpha = alpha; md.daipi2 = daipi2; md.elC1 = elC1; md.elC2 = elC2; md.rElec = rElec;
    float e = 0.f, r = 0.f, f;
    switch (kind) {
    case 1: f = direct_coul(r2, r, chprd, &md, e); break;
    case 2: f = real_ewald(r2, r, chprd, &md, e); break;
    case 3: f = fennel(r2, r, chprd, &md, e); break;
    default: return -1;
    }
    *f_over_r = f; *eng = e; *r_out = r;
    return 0;
}

void refdev_angled_vector(const float* in, float cos_phi, float theta, float* out)
{
    float3 r = get_angled_vector(make_float3(in[0], in[1], in[2]), cos_phi, theta);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// One atom through the body of tstat_radi9 (src/cuTemp.cu:736-758): absorb a photon of energy pe from the direction
// uvects[rnd[0] % 3072], then, if U > 1e-4, radiate_photon3 with the draws rnd[1], rnd[2]; radius = A / (B - min(U, mx)).
void refdev_tstat_atom(float* vel, float* U, float mass, float pe, const unsigned* rnd, const float* uvects /*3072 x 3*/,
                       float radA, float radB, float mxEng, float* radius)
{
    static cudaMD md;
    static float3 uv[3072];
    for (int k = 0; k < 3072; k++) uv[k] = make_float3(uvects[3 * k], uvects[3 * k + 1], uvects[3 * k + 2]);
    md.uvects = uv;
    g_rnd[0] = rnd[0]; g_rnd[1] = rnd[1]; g_rnd[2] = rnd[2]; g_rnd_n = 3; g_rnd_pos = 0;
    uint4 st = make_uint4(0, 0, 0, 0);
    float3 v = make_float3(vel[0], vel[1], vel[2]);
    float u = *U;
    adsorb_rand_photon(&v, &u, mass, pe, st, &md, 0);
    if (u > 1e-4) radiate_photon3(&v, &u, mass, st, &md, 0);
    vel[0] = v.x; vel[1] = v.y; vel[2] = v.z;
    *U = u;
    float restrE = u < mxEng ? u : mxEng;
    *radius = radA / (radB - restrE);
}

// split_cells with div_type 1 as init_cudaMD calls it (src/cuInit.cu:979)
void refdev_split_cells(const float* L, float r, int nAt, int* cnum, float* csize, float* crev)
{
    static cudaMD md;
    md.leng = make_float3(L[0], L[1], L[2]);
    split_cells(1, r, 0, nAt, &md);
    cnum[0] = md.cNumber.x; cnum[1] = md.cNumber.y; cnum[2] = md.cNumber.z;
    csize[0] = md.cSize.x; csize[1] = md.cSize.y; csize[2] = md.cSize.z;
    crev[0] = md.cRevSize.x; crev[1] = md.cRevSize.y; crev[2] = md.cRevSize.z;
}

int refdev_cell_dist(int xi, int xj, int mx, float csize, float rskip, float* rmin, float* rmax, int* shift)
{
    return cell_dist_int(xi, xj, mx, csize, rskip, *rmin, *rmax, *shift);
}

// put_periodic followed by count_cell for one atom of species 0: wrapped position, cell index, face-crossing counters
// (pos x,y,z then neg x,y,z) and momenta
void refdev_wrap_and_bin(float* xyz, const float* vel, float mass, const float* L, const int* cnum, const float* crev,
                         int* cell, int* crossings, float* mom)
{
    static cudaMD md;
    static int3 acPos[1], acNeg[1];
    static int cellIdx[1], inside[1];
    static int counts[1 << 20];
    md.leng = make_float3(L[0], L[1], L[2]);
    md.revLeng = make_float3(1.f / L[0], 1.f / L[1], 1.f / L[2]); // src/cuInit.cu:937
    md.cNumber = make_int3(cnum[0], cnum[1], cnum[2]);
    md.cRevSize = make_float3(crev[0], crev[1], crev[2]);
    md.cnYZ = cnum[1] * cnum[2];
    md.nCell = cnum[0] * cnum[1] * cnum[2];
    acPos[0] = make_int3(0, 0, 0); acNeg[0] = make_int3(0, 0, 0);
    md.specAcBoxPos = acPos; md.specAcBoxNeg = acNeg;
    md.posMom = make_float3(0.f, 0.f, 0.f); md.negMom = make_float3(0.f, 0.f, 0.f);
    md.cellIndexes = cellIdx; md.insideCellIndex = inside; md.nAtInCell = counts;
    float3 p = make_float3(xyz[0], xyz[1], xyz[2]);
    put_periodic(p, make_float3(vel[0], vel[1], vel[2]), mass, 0, &md);
    xyz[0] = p.x; xyz[1] = p.y; xyz[2] = p.z;
    cellIdx[0] = -1;
    const float3 q = p;
    int c = (int)(floor((double)q.x * (double)md.cRevSize.x) * md.cnYZ + floor((double)q.y * (double)md.cRevSize.y) * md.cNumber.z + floor((double)q.z * (double)md.cRevSize.z));
    if (c >= 0 && c < md.nCell && md.nCell <= (1 << 20)) { counts[c] = 0; count_cell(0, p, &md); *cell = cellIdx[0]; }
    else *cell = c; // out of the grid: the reference would write out of bounds; report the raw index
    crossings[0] = acPos[0].x; crossings[1] = acPos[0].y; crossings[2] = acPos[0].z;
    crossings[3] = acNeg[0].x; crossings[4] = acNeg[0].y; crossings[5] = acNeg[0].z;
    mom[0] = md.posMom.x; mom[1] = md.posMom.y; mom[2] = md.posMom.z; mom[3] = md.negMom.x; mom[4] = md.negMom.y; mom[5] = md.negMom.z;
}

// one Nose-Hoover block as apply_tstat launches it: before_nose, [tstat_nose scales the velocities by md->tscale], after_nose(1)
void refdev_nose(float tSt, float teKin, float rQmass, float qMassTau2, float* engKin, float* chit, float* consInt, float* tscale)
{
    static cudaMD md;
    md.tSt = tSt; md.teKin = teKin; md.rQmass = rQmass; md.qMassTau2 = qMassTau2;
    md.engKin = *engKin; md.chit = *chit; md.consInt = *consInt;
    before_nose(&md);
    *tscale = md.tscale;
    after_nose(1, &md);
    *engKin = md.engKin; *chit = md.chit; *consInt = md.consInt;
}


// bond potentials by the reference's type ids (src/cuBonds.cu:1248-1270: 1 harm, 2 morse, 3 pedone, 4 buck, 5 e6812)
int refdev_bond(int kind, const float* p, float r2, float r, float* f_over_r, float* eng)
{
    cudaBond b;
    memset(&b, 0, sizeof(b));
    b.p0 = p[0]; b.p1 = p[1]; b.p2 = p[2]; b.p3 = p[3]; b.p4 = p[4];
    float e = 0.f, f;
    switch (kind) {
    case 1: f = bond_harm(r2, r, e, &b); break;
    case 2: f = bond_morse(r2, r, e, &b); break;
    case 3: f = bond_pedone(r2, r, e, &b); break;
    case 4: f = bond_buck(r2, r, e, &b); break;
    case 5: f = bond_e6812(r2, r, e, &b); break;
    default: return -1;
    }
    *f_over_r = f; *eng = e;
    return 0;
}

// angle_hcos on three atoms (centre, ligand 1, ligand 2) in a rectangular periodic box: forces (added to zero) and energy
void refdev_angle_hcos(float k, float cos0, const float* xyz /*3 x 3*/, const float* L, float* frc /*3 x 3*/, float* eng)
{
    static cudaMD md;
    float3 pos[3], f[3];
    for (int a = 0; a < 3; a++) { pos[a] = make_float3(xyz[3 * a], xyz[3 * a + 1], xyz[3 * a + 2]); f[a] = make_float3(0.f, 0.f, 0.f); }
    md.xyz = pos; md.frs = f;
    md.leng = make_float3(L[0], L[1], L[2]);
    md.halfLeng = make_float3(0.5f * L[0], 0.5f * L[1], 0.5f * L[2]);   // src/cuInit.cu:937
    cudaAngle t;
    memset(&t, 0, sizeof(t));
    t.p0 = k; t.p1 = cos0;
    int4 ang = make_int4(0, 1, 2, 1);
    float e = 0.f;
    angle_hcos(&ang, &t, &md, e);
    for (int a = 0; a < 3; a++) { frc[3 * a] = f[a].x; frc[3 * a + 1] = f[a].y; frc[3 * a + 2] = f[a].z; }
    *eng = e;
}
} // extern "C"
