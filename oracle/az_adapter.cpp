// TEST INFRASTRUCTURE (like everything under oracle/): the reference-side binding of INTEGRATION.md section B, kept as a file so
// that tests/test_abi.py can compile it against the reference's own headers (src/dataStruct.h, temperature.h, elec.h).
// Nothing in the product builds or links it.  The text below the marker is the code block of INTEGRATION.md, verbatim.
// ---- INTEGRATION.md section B ----
// az_adapter.cpp -- compiled into the reference instead of cuInit.cu / cuPairs.cu / cuMDfunc.cu / ...
#include "aztot_b200.h"
#include "dataStruct.h"      // Atoms, Field, Sim, Box, Elec, Spec, VdW (reference headers)
#include "temperature.h"     // TStat
#include "elec.h"            // tpElecEwald

static az_engine* g_eng;

// replaces init_cudaMD (src/cuInit.cu:756-1199): returns 0 on failure like the original's NULL
int az_init_from_reference(Atoms* atm, Field* fld, Sim* sim, TStat* ts, Box* bx, Elec* el,
                           const double* photons /* TStat::photons or NULL */)
{
    static az_species sp[AZ_MAX_SPECIES];
    static az_vdw_entry tab[AZ_MAX_SPECIES * AZ_MAX_SPECIES];
    az_system_desc d = {};
    d.abi_version = AZ_ABI_VERSION;
    d.n_atoms = atm->nAt;          d.n_species = fld->nSpec;
    d.xs = atm->xs; d.ys = atm->ys; d.zs = atm->zs;             // borrowed for the call only
    d.vxs = atm->vxs; d.vys = atm->vys; d.vzs = atm->vzs;
    d.fxs = atm->fxs; d.fys = atm->fys; d.fzs = atm->fzs;
    d.types = atm->types;
    d.la = bx->la; d.lb = bx->lb; d.lc = bx->lc;                 // rectangular box, src/box.cpp:9-27
    for (int i = 0; i < fld->nSpec; i++) {                       // src/cuInit.cu:843-866
        Spec& s = fld->species[i];
        sp[i].mass = s.mass; sp[i].charge = s.charge; sp[i].frozen = s.frozen;
        sp[i].radA = s.radA; sp[i].radB = s.radB; sp[i].mxEng = s.mxEng;
        for (int j = 0; j < fld->nSpec; j++) {                   // src/cuInit.cu:794-904
            VdW* v = fld->vdws[i][j];
            az_vdw_entry& t = tab[i * fld->nSpec + j];
            if (!v) { t.kind = AZ_VDW_NONE; continue; }
            t.kind = v->type; t.use_radii = v->use_radii;
            t.p0 = v->p0; t.p1 = v->p1; t.p2 = v->p2; t.p3 = v->p3; t.p4 = v->p4; t.r2cut = v->r2cut;
        }
    }
    d.species = sp; d.vdw = tab; d.max_r_vdw = fld->maxRvdw;
    d.elec_kind = el->type; d.r_real = el->rReal; d.alpha = el->alpha; d.daipi2 = el->daipi2;
    d.el_c1 = el->scale; d.el_c2 = el->scale2; d.eps = el->eps;   // prepare_elec, src/elec.cpp:371-406
    if (el->type == tpElecEwald) { d.ewald_k[0] = el->kx; d.ewald_k[1] = el->ky; d.ewald_k[2] = el->kz; }  // 'elec pme': + reciprocal space
    d.tstat_kind = ts->type; d.temperature = ts->Temp; d.t_kin = ts->tKin; d.tstat_tau = ts->tau; d.reset_vels = sim->reset_vels;
    d.photons = photons;                                          // src/temperature.cpp:28-89
    d.timestep = sim->tSt; d.n_equil = sim->nEq; d.eq_freq = sim->freqEq;
    d.r2_max = sim->r2Max; d.cell_size = sim->desired_cell_size;
    d.efield[0] = sim->Ux; d.efield[1] = sim->Uy; d.efield[2] = sim->Uz;
    az_cuda_opts o = {};                                          // device 0, reference schedule (rebuild every step)
    // o.rebuild_every = 10; o.skin = 0.25f;   Verlet list between rebuilds;   o.use_graph = 1;   replay the step loop as a CUDA graph
    // o.pair_variant / o.list_format / o.max_neigh: tested alternatives of the pair path (include/aztot_b200.h); 0 = what was measured fastest
    if (az_engine_create(&d, &o, &g_eng) != AZ_OK) { printf("ERROR: %s\n", az_last_error(NULL)); return 0; }
    return 1;
}

// replaces the loop body of main (src/main.cu:283-383) + stat_iter/copy_stat (src/cuStat.cu:17-71,387-404)
int az_steps_from_reference(int iStep, int n, az_stats* st)
{
    if (az_engine_step(g_eng, iStep, n) != AZ_OK || az_engine_stats(g_eng, st) != AZ_OK) {
        printf("ERROR: %s\n", az_last_error(g_eng));
        return 0;
    }
    return 1;
}

// replaces md_to_host (src/cuInit.cu:1212-1261): sorted order, like the original
int az_to_host(Atoms* atm, float* engs, float* radii)
{
    az_atoms_out out = {};
    out.n_atoms = atm->nAt;
    out.xs = atm->xs; out.ys = atm->ys; out.zs = atm->zs;
    out.vxs = atm->vxs; out.vys = atm->vys; out.vzs = atm->vzs;
    out.types = atm->types; out.engs = engs; out.radii = radii;
    return az_engine_download(g_eng, &out, AZ_ORDER_SORTED) == AZ_OK;
}

void az_free() { az_engine_destroy(g_eng); g_eng = NULL; }      // replaces free_device_md (src/cuInit.cu:1263-1373)
