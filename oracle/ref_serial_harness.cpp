// ref_serial_harness.cpp -- drives the UNMODIFIED serial twin of the reference (aztot_serial,
// /root/reference/src/main.cpp:26-229) from its own object files.  TEST INFRASTRUCTURE ONLY:
// it is compiled into oracle/_ref/ by oracle/build_ref.sh and used (a) to pin oracle/az_oracle.c
// (forces/energies of the reference's pair functions, double precision) and (b) as the
// "reference" CPU baseline of bench.py.  It re-states main.cpp's loop (lines 89-130: integrator1,
// clear_force, add_elec, forcefield, integrate2, calc_chars) without the file output, with a
// monotonic timer around the loop, and dumps forces / energies in a small binary format.
//
// usage (cwd = deck directory): ref_serial_harness <nsteps> <init: full|fast> <dump.bin|->
//   full: init_serial() as the reference does (O(N^2) all_pairs initial forces)
//   fast: the same initialisation calls as init_serial (src/sys_init.cpp:1122-1188) minus the
//         O(N^2) all_pairs; initial forces come from one cell_list() pass instead.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "dataStruct.h"
#include "sys_init.h"
#include "elec.h"
#include "box.h"
#include "ejump.h"
#include "bonds.h"
#include "angles.h"
#include "cell_list.h"
#include "rdf.h"
#include "integrators.h"

static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void dump(FILE* f, const char* tag, Atoms* a, Sim* s)
{
    int n = a->nAt;
    char t[8]; memset(t, 0, 8); strncpy(t, tag, 7);
    fwrite(t, 1, 8, f); fwrite(&n, sizeof(int), 1, f);
    double e[4] = { s->engVdW, s->engElec3, s->engKin, s->engTot };
    fwrite(e, sizeof(double), 4, f);
    fwrite(a->xs, sizeof(double), n, f); fwrite(a->ys, sizeof(double), n, f); fwrite(a->zs, sizeof(double), n, f);
    fwrite(a->vxs, sizeof(double), n, f); fwrite(a->vys, sizeof(double), n, f); fwrite(a->vzs, sizeof(double), n, f);
    fwrite(a->fxs, sizeof(double), n, f); fwrite(a->fys, sizeof(double), n, f); fwrite(a->fzs, sizeof(double), n, f);
}

int main(int argc, char** argv)
{
    int nsteps = argc > 1 ? atoi(argv[1]) : 0;
    int fast = argc > 2 && strcmp(argv[2], "fast") == 0;
    const char* dumpname = argc > 3 ? argv[3] : "-";
    unsigned seed = argc > 4 ? (unsigned)atoi(argv[4]) : 1u;
    srand(seed); // the reference uses time(NULL) (src/main.cpp has none before init; main.cu:242)

    Elec* elec = (Elec*)malloc(sizeof(Elec));
    TStat* tstat = (TStat*)malloc(sizeof(TStat));
    Box* box = (Box*)malloc(sizeof(Box));
    Field* field = (Field*)malloc(sizeof(Field));
    Atoms* atoms = (Atoms*)malloc(sizeof(Atoms));
    Sim* sim = (Sim*)malloc(sizeof(Sim));
    memset(sim, 0, sizeof(Sim)); memset(elec, 0, sizeof(Elec)); memset(tstat, 0, sizeof(TStat));

    double t0 = now();
    if (!init_md(atoms, field, sim, elec, tstat, box)) { printf("{\"error\": \"init_md failed\"}\n"); return 2; }
    if (!fast) {
        if (!init_serial(atoms, field, sim, elec, tstat, box)) { printf("{\"error\": \"init_serial failed\"}\n"); return 2; }
    } else {
        // src/sys_init.cpp:1130-1176 verbatim call sequence, without all_pairs (line 1184)
        if ((sim->flags >> bfSimUseCList) & 1) init_clist(atoms, sim, box, sim->rMax);
        alloc_rdf(sim, field, box);
        init_elec(elec, box, sim, atoms);
        center_box(atoms, box);
        if (!sim->nHead) { printf("{\"error\": \"fast init needs cell_list\"}\n"); return 2; }
        sim->integrator1 = integrate1_clst;
        sim->forcefield = cell_list;
        sim->pair = pair_inter;
        sim->do_jump = jmp_none;
        sim->pair_elec = pair_elecs[elec->type];
        sim->add_elec = add_elecs[elec->type];
        reset_chars(sim);
        // initial forces through the cell list (fill the list with a zero-length drift)
        double dt = sim->tSt; sim->tSt = 0.0;
        for (int i = 0; i < field->nSpec; i++) field->species[i].rMass_hdt = 0.0;
        sim->integrator1(atoms, field->species, sim, box, tstat);
        sim->tSt = dt;
        for (int i = 0; i < field->nSpec; i++) field->species[i].rMass_hdt = 0.5 * sim->tSt / field->species[i].mass;
        clear_force(atoms, field->species, sim, box);
        sim->add_elec(atoms, field, elec, box, sim);
        sim->forcefield(atoms, field, elec, box, sim);
    }
    // SURVEY 8f-4 groundwork: the loop of main.cpp adds the bonded terms after the force field (src/main.cpp:101-104);
    // the initial force evaluation of init_serial does not.  With argv[5] == "bonds" the initial dump carries them too.
    if (argc > 5 && strcmp(argv[5], "bonds") == 0) {
        if (field->nBonds) exec_bondlist(atoms, field, sim, box);
        if (field->nAngles) exec_anglelist(atoms, field, sim, box);
    }
    // a32 pinning: argv[5] == "tstat" writes what read_tstat (src/temperature.cpp:91-259) built for `temperature T radi n` under
    // srand(seed) -- the photon energies of photon_engs and the 3072 preset unit vectors -- to tstat.bin and stops
    if (argc > 5 && strcmp(argv[5], "tstat") == 0) {
        FILE* tf = fopen("tstat.bin", "wb");
        int n = atoms->nAt, nv = 16 * 32 * 2 * 3;
        fwrite(&n, sizeof(int), 1, tf); fwrite(&nv, sizeof(int), 1, tf);
        fwrite(tstat->photons, sizeof(double), n, tf);
        fwrite(tstat->uvectX, sizeof(double), nv, tf); fwrite(tstat->uvectY, sizeof(double), nv, tf); fwrite(tstat->uvectZ, sizeof(double), nv, tf);
        fclose(tf);
        printf("{\"tstat\": %d, \"temperature\": %.6f}\n", n, tstat->Temp);
        return 0;
    }
    double t1 = now();
    FILE* df = strcmp(dumpname, "-") ? fopen(dumpname, "wb") : NULL;
    if (df) dump(df, "init", atoms, sim);

    double tSim = 0.0;
    double t2 = now();
    for (int iSt = 1; iSt <= nsteps; iSt++) {
        reset_chars(sim);
        sim->integrator1(atoms, field->species, sim, box, tstat);
        clear_force(atoms, field->species, sim, box);
        sim->add_elec(atoms, field, elec, box, sim);
        sim->forcefield(atoms, field, elec, box, sim);
        if (field->nBonds) exec_bondlist(atoms, field, sim, box);     // src/main.cpp:101-104
        if (field->nAngles) exec_anglelist(atoms, field, sim, box);
        if (iSt > sim->nEq) integrate2(atoms, field->species, sim, 0, tstat);
        else integrate2(atoms, field->species, sim, (sim->freqEq && (iSt % sim->freqEq) == 0) ? 1 : 0, tstat);
        calc_chars(sim, tSim);
    }
    double t3 = now();
    if (df) { dump(df, "final", atoms, sim); fclose(df); }
    printf("{\"n_atoms\": %d, \"nsteps\": %d, \"init_s\": %.6f, \"loop_s\": %.6f, \"s_per_step\": %.9f, "
           "\"atom_steps_per_s\": %.6e, \"engVdW\": %.10e, \"engCoul\": %.10e, \"engKin\": %.10e, \"n_cells\": %d, "
           "\"engElec1\": %.10e, \"engElec2\": %.10e, \"engBond\": %.10e, \"engAngle\": %.10e, \"n_bonds\": %d}\n",
           atoms->nAt, nsteps, t1 - t0, t3 - t2, nsteps ? (t3 - t2) / nsteps : 0.0,
           nsteps ? (double)atoms->nAt * nsteps / (t3 - t2) : 0.0, sim->engVdW, sim->engElec3, sim->engKin, sim->nHead, sim->engElec1, sim->engElec2, sim->engBond, sim->engAngle, field->nBonds);
    return 0;
}
