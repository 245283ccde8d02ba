// az_main.cpp -- `aztotmd`: the host driver of the drop-in.  It keeps the reference's process
// surface (src/main.cu:239-463): reads atoms.xyz / field.txt / control.txt / cuda.txt from the
// working directory (or the directory given as first argument), runs the step loop on the B200
// engine through the C ABI and writes stat.dat, msd.dat, revcon.xyz, velocities.dat and (radiative
// thermostat) tchars.dat in the reference's formats:
//   stat.dat / msd.dat   start_stat, copy_stat      src/cuStat.cu:40-71,300-358
//   stdout               print_stat                 src/cuMDfunc.cu:318-331
//   revcon.xyz           out_atoms, save_box        src/out_md.cpp:65-86, src/box.cpp:169-176
//   velocities.dat       out_velocities             src/out_md.cpp:126-194
//   tchars.dat           out_thermalchar            src/main.cu:51-118
//   rdf.dat, rdf<N>.dat  rdf_iter, copy_rdf         src/cuStat.cu:514-594 (sampled from the cell grid, not by brute force)
// Statistics rows are produced for iStep % stat == 0 like stat_iter (src/cuStat.cu:387-404).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <string>
#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/aztot_b200.h"

extern "C" const char* az_deck_warnings(const az_deck*);

// copy_rdf (src/cuStat.cu:514-561): normalisation and layout of rdf.dat
static void write_rdf(const char* fname, az_engine* eng, const az_deck* deck, const az_system_desc* d, double dr)
{
    int nb = 0, np = 0, ns = 0;
    if (az_engine_rdf_read(eng, nullptr, 0, &nb, &np, &ns) != AZ_OK || ns <= 0) return;
    std::vector<unsigned long long> cnt((size_t)nb * np);
    if (az_engine_rdf_read(eng, cnt.data(), (int)cnt.size(), nullptr, nullptr, nullptr) != AZ_OK) return;
    FILE* f = fopen(fname, "w");
    if (!f) return;
    const int nspec = d->n_species;
    fprintf(f, "r");
    for (int i = 0; i < nspec; i++)
        for (int j = i; j < nspec; j++) fprintf(f, "\t%s-%s", d->species[i].name, d->species[j].name);
    fprintf(f, "\n");
    const double sphera = 4.0 * 3.14159265359 / 3.0; // src/const.h:11,15
    const double C1 = 2 / (sphera * dr * dr * dr * ns);
    const float volume = (float)(d->la * d->lb * d->lc);
    for (int i = 0; i < nb; i++) {
        fprintf(f, "%f", (i + 0.5) * dr);
        const float C2 = 1.f / (3.f * i * (i + 1.f) + 1.f);
        int k = 0, l = 0;
        for (int j = 0; j < np; j++) {
            const float C3 = (k == l) ? 1.f : 0.5f;
            // what brute_rdf accumulated per pair: volume / N_k / N_l (src/cuStat.cu:497)
            const float acc = (float)cnt[(size_t)i * np + j] * (volume / (float)az_deck_species_count(deck, k) / (float)az_deck_species_count(deck, l));
            fprintf(f, "\t%f", acc * C1 * C2 * C3);
            l++;
            if (l == nspec) { k++; l = k; }
        }
        fprintf(f, "\n");
    }
    fclose(f);
}

// copy_nrdf (src/cuStat.cu:696-777): the same histogram by nuclei.  brute_nrdf adds 2 V / (NN_a NN_b) per pair to the bin of the
// nucleus pair of the two species (src/cuStat.cu:675-676); the integer counts per species pair carry the same information.
static void write_nrdf(const char* fname, az_engine* eng, const az_deck* deck, const az_system_desc* d, double dr)
{
    int nb = 0, np = 0, ns = 0;
    if (az_engine_rdf_read(eng, nullptr, 0, &nb, &np, &ns) != AZ_OK || ns <= 0) return;
    std::vector<unsigned long long> cnt((size_t)nb * np);
    if (az_engine_rdf_read(eng, cnt.data(), (int)cnt.size(), nullptr, nullptr, nullptr) != AZ_OK) return;
    const int nspec = d->n_species;
    std::vector<int> nof(nspec);
    const int nn = az_deck_nuclei(deck, nof.data());
    if (nn <= 0) return;
    FILE* f = fopen(fname, "w");
    if (!f) return;
    fprintf(f, "r");
    for (int i = 0; i < nn; i++)
        for (int j = i; j < nn; j++) fprintf(f, "\t%s-%s", az_deck_nucleus_name(deck, i), az_deck_nucleus_name(deck, j));
    fprintf(f, "\n");
    const double sphera = 4.0 * 3.14159265359 / 3.0;
    const double C1 = 2 / (sphera * dr * dr * dr * ns);
    const float volume = (float)(d->la * d->lb * d->lc);
    const int nnp = nn * (nn + 1) / 2, m = nn - 1;
    std::vector<double> acc(nnp);
    for (int i = 0; i < nb; i++) {
        fprintf(f, "%f", (i + 0.5) * dr);
        const float C2 = 1.f / (3.f * i * (i + 1.f) + 1.f);
        std::fill(acc.begin(), acc.end(), 0.0);
        int k = 0, l = 0;
        for (int j = 0; j < np; j++) { // species pair (k, l), k <= l -> nucleus pair
            int a = nof[k], b = nof[l];
            if (a > b) std::swap(a, b);
            const int ip = a * m + a * (1 - a) / 2 + b; // n_iPair, src/cuStat.cu:668
            acc[ip] += (double)((float)cnt[(size_t)i * np + j] * (volume / (float)az_deck_nucleus_count(deck, a) / (float)az_deck_nucleus_count(deck, b)));
            l++;
            if (l == nspec) { k++; l = k; }
        }
        k = 0; l = 0;
        for (int j = 0; j < nnp; j++) {
            const float C3 = (k == l) ? 1.f : 0.5f;
            fprintf(f, "\t%f", (float)acc[j] * C1 * C2 * C3);
            l++;
            if (l == nn) { k++; l = k; }
        }
        fprintf(f, "\n");
    }
    fclose(f);
}

static void print_stat(int iStep, const az_stats& s, const az_system_desc* d)
{
    printf("%d x1=%.2f ", iStep, s.x_first);
    if (d->elec_kind) printf("C1=%.3G, C2=%.3G ", s.engCoul1, s.engCoul2);
    printf("Kin=%.3G Vdw=%.3G Tot=%.3G P=%.0f ", s.engKin, s.engVdW, s.engTot, s.pressure);
    if (d->tstat_kind == AZ_TSTAT_RADI) printf("U=%.3G", s.engTemp);
    printf("\n");
}

int main(int argc, char** argv)
{
    const char* dir = ".";
    unsigned long long seed = 0;
    int max_steps = -1;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--steps") && i + 1 < argc) max_steps = atoi(argv[++i]);
        else if (argv[i][0] != '-') dir = argv[i];
        else { fprintf(stderr, "usage: aztotmd [deck-dir] [--seed N] [--steps N]\n"); return 2; }
    }
    time_t start_time = time(NULL);
    az_deck* deck = nullptr;
    int rc = az_deck_load(dir, seed, &deck);
    if (deck && az_deck_warnings(deck)[0]) printf("%s", az_deck_warnings(deck));
    if (rc != AZ_OK) {
        printf("%s\n", az_deck_error(deck));
        az_deck_free(deck);
        return 1;
    }
    const az_system_desc* d = az_deck_desc(deck);
    const int nSt = max_steps >= 0 ? max_steps : az_deck_nsteps(deck);
    const int stat = az_deck_stat_period(deck);
    const int nspec = d->n_species;
    if (chdir(dir) != 0) { printf("ERROR: cannot enter %s\n", dir); return 1; }

    az_engine* eng = nullptr;
    rc = az_engine_create(d, az_deck_opts(deck), &eng);
    if (rc != AZ_OK) { printf("ERROR: engine: %s\n", az_last_error(nullptr)); az_deck_free(deck); return 1; }

    // start_stat
    FILE* fs = fopen("stat.dat", "w");
    FILE* fm = fopen("msd.dat", "w");
    if (!fs || !fm) { printf("ERROR: cannot open stat.dat / msd.dat for writing\n"); return 1; }
    const bool radi = d->tstat_kind == AZ_TSTAT_RADI;
    const bool has_bnd = d->n_bond_types > 0, has_ang = d->n_angle_types > 0; // fld->nBdata, fld->nAdata (src/cuStat.cu:311-314)
    fprintf(fs, "time\tstep\tengTot\tengKin\tengVdW\tengCoul1\tengCoul2");
    if (radi) fprintf(fs, "\tengTerm");
    if (has_bnd) fprintf(fs, "\tengBnd");
    if (has_ang) fprintf(fs, "\tengAngle");
    fprintf(fs, "\tmomPx\tmomNx\tmomPy\tmomNy\tmomPz\tmomNz\tpress");
    fprintf(fs, "\ntime, ps\tstep, n\tengTot, eV\tengKin, eV\tengVdW, eV\tengCoul1, eV\tengCoul2, eV");
    if (radi) fprintf(fs, "\tengTerm, eV");
    if (has_bnd) fprintf(fs, "\tengBnd, eV");
    if (has_ang) fprintf(fs, "\tengAngle, eV");
    fprintf(fs, "\tmomPx, eVps/A\tmomNx, eVps/A\tmomPy, eVps/A\tmomNy, eVps/A\tmomPz, eVps/A\tmomNz, eVps/A\tpress, atm\n");
    fprintf(fm, "time\tstep");
    for (int i = 0; i < nspec; i++) fprintf(fm, "\t%s_px\tnx\tpy\tny\tpz\tnz", d->species[i].name);
    fprintf(fm, "\n");

    double rdf_rmax = 0.0, rdf_dr = 0.0;
    int rdf_freq = 0, rdf_outfreq = 0;
    bool rdf_on = az_deck_rdf(deck, &rdf_rmax, &rdf_dr, &rdf_freq, &rdf_outfreq) != 0;
    const bool rdf_nucl = az_deck_rdf_nuclei(deck) != 0;
    if (rdf_on && az_engine_rdf_begin(eng, rdf_rmax, rdf_dr) != AZ_OK) {
        printf("WARNING: no RDF will be written: %s\n", az_last_error(eng));
        rdf_on = false;
    }

    az_stats st;
    memset(&st, 0, sizeof(st));
    int iStep = 0;
    while (iStep < nSt) {
        // advance to the next statistics step (iStep % stat == 0), the next RDF sample, or the end
        int target = (iStep % stat == 0) ? iStep : ((iStep / stat + 1) * stat);
        if (rdf_on) {
            const int tr = (iStep % rdf_freq == 0) ? iStep : ((iStep / rdf_freq + 1) * rdf_freq);
            if (tr < target) target = tr;
        }
        if (target > nSt - 1) target = nSt - 1;
        rc = az_engine_step(eng, iStep, target - iStep + 1);
        if (rc == AZ_OK) rc = az_engine_stats(eng, &st);
        if (rc != AZ_OK) { printf("ERROR: %s\n", az_last_error(eng)); break; }
        iStep = target;
        if (rdf_on && iStep % rdf_freq == 0) { // rdf_iter, src/cuStat.cu:563-594
            if (az_engine_rdf_sample(eng) != AZ_OK) { printf("ERROR: %s\n", az_last_error(eng)); rc = AZ_ERR_CUDA; break; }
            if (rdf_outfreq && iStep % rdf_outfreq == 0) {
                char name[64];
                snprintf(name, sizeof(name), "rdf%d.dat", iStep);
                write_rdf(name, eng, deck, d, rdf_dr);
                if (rdf_nucl) { snprintf(name, sizeof(name), "rdf_n%d.dat", iStep); write_nrdf(name, eng, deck, d, rdf_dr); } // nrdf_iter, src/cuStat.cu:779-797
            }
        }
        if (iStep % stat == 0) {
            print_stat(iStep, st, d);
            double t = iStep * d->timestep;
            fprintf(fs, "%f\t%d\t%f\t%f\t%f\t%f\t%f", t, iStep, st.engTot, st.engKin, st.engVdW, st.engCoul1, st.engCoul2);
            if (radi) fprintf(fs, "\t%f", st.engTemp);
            if (has_bnd) fprintf(fs, "\t%f", st.engBond);
            if (has_ang) fprintf(fs, "\t%f", st.engAngle);
            fprintf(fs, "\t%f\t%f\t%f\t%f\t%f\t%f\t%f\n", st.posMom[0], st.negMom[0], st.posMom[1], st.negMom[1], st.posMom[2], st.negMom[2], st.pressure);
            fprintf(fm, "%f\t%d", t, iStep);
            for (int i = 0; i < nspec; i++)
                fprintf(fm, "\t%d\t%d\t%d\t%d\t%d\t%d", st.specAcBoxPos[i][0], st.specAcBoxNeg[i][0], st.specAcBoxPos[i][1], st.specAcBoxNeg[i][1],
                        st.specAcBoxPos[i][2], st.specAcBoxNeg[i][2]);
            fprintf(fm, "\n");
        }
        iStep++;
    }
    if (rc == AZ_OK) print_stat(iStep, st, d); // final output to screen
    fclose(fs); fclose(fm);
    if (rdf_on) write_rdf("rdf.dat", eng, deck, d, rdf_dr);
    if (rdf_on && rdf_nucl) write_nrdf("rdf_n.dat", eng, deck, d, rdf_dr); // copy_nrdf at the end of the run (src/main.cu:420-424)

    // md_to_host: atoms come back in cell-sorted order, like the reference
    const int n = d->n_atoms;
    std::vector<double> x(n), y(n), z(n), vx(n), vy(n), vz(n);
    std::vector<int> types(n);
    std::vector<float> engs(n), radii(n);
    az_atoms_out out;
    memset(&out, 0, sizeof(out));
    out.n_atoms = n;
    out.xs = x.data(); out.ys = y.data(); out.zs = z.data(); out.vxs = vx.data(); out.vys = vy.data(); out.vzs = vz.data();
    out.types = types.data(); out.engs = engs.data(); out.radii = radii.data();
    int rc2 = az_engine_download(eng, &out, AZ_ORDER_SORTED);
    if (rc2 != AZ_OK) { printf("ERROR: %s\n", az_last_error(eng)); az_engine_destroy(eng); az_deck_free(deck); return 1; }

    std::vector<int> count(nspec, 0);
    for (int i = 0; i < n; i++) count[types[i]]++;
    int mx = 0;
    for (int t = 0; t < nspec; t++) if (count[t] > mx) mx = count[t];
    std::vector<std::vector<int>> by(nspec);
    for (int i = 0; i < n; i++) by[types[i]].push_back(i);

    if (radi) { // out_thermalchar
        FILE* f = fopen("tchars.dat", "w");
        if (f) {
            fprintf(f, "No");
            for (int t = 0; t < nspec; t++) fprintf(f, "\t%s_eng\t%s_rad", d->species[t].name, d->species[t].name);
            fprintf(f, "\n");
            for (int i = 0; i < mx; i++) {
                fprintf(f, "%d", i + 1);
                for (int t = 0; t < nspec; t++)
                    if (i < count[t]) fprintf(f, "\t%f\t%f", (double)engs[by[t][i]], (double)radii[by[t][i]]);
                    else fprintf(f, "\t\t");
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    { // out_atoms
        FILE* f = fopen("revcon.xyz", "w");
        if (f) {
            fprintf(f, "%d\n", n);
            fprintf(f, "%d %f %f %f\n", 1, d->la, d->lb, d->lc);
            for (int i = 0; i < n; i++) fprintf(f, "%s\t%f\t%f\t%f\n", d->species[types[i]].name, x[i], y[i], z[i]);
            fclose(f);
        }
    }
    { // out_velocities
        FILE* f = fopen("velocities.dat", "w");
        if (f) {
            fprintf(f, "No");
            for (int t = 0; t < nspec; t++) fprintf(f, "\t%s\tx\ty\tz", d->species[t].name);
            fprintf(f, "\n");
            for (int i = 0; i < mx; i++) {
                fprintf(f, "%d", i + 1);
                for (int t = 0; t < nspec; t++)
                    if (i < count[t]) {
                        int k = by[t][i];
                        fprintf(f, "\t%f\t%f\t%f\t%f", sqrt(vx[k] * vx[k] + vy[k] * vy[k] + vz[k] * vz[k]), vx[k], vy[k], vz[k]);
                    } else fprintf(f, "\t\t\t\t");
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    az_engine_destroy(eng);
    az_deck_free(deck);
    time_t final_time = time(NULL);
    printf("execution time: %d s, kernel launches on the engine stream only\n", (int)(final_time - start_time));
    return rc == AZ_OK ? 0 : 1;
}
