// az_deck.cpp -- host side of the drop-in: the reference's four input files -> az_system_desc.
//
// Mirrors the parsing layer of the reference for the hot path (host, double precision):
//   token scan         find_int / find_double / find_str     src/utils.cpp:87-195
//   field.txt          read_field, read_spec, read_vdw       src/sys_init.cpp:78-131,174-485, src/vdw.cpp:234-308
//   atoms.xyz          read_atoms_box, read_box              src/sys_init.cpp:487-565, src/box.cpp:9-27
//   control.txt        read_sim, read_tstat, read_elec       src/sys_init.cpp:590-988, src/temperature.cpp:91-259,
//                                                            src/elec.cpp:14-67
//   cuda.txt           read_cuda                             src/cuInit.cu:684-754
//   derived values     init_md, prepare_elec                 src/sys_init.cpp:1036-1119, src/elec.cpp:371-406
//   photon energies    photon_engs                           src/temperature.cpp:28-89
//   'init_vel gaus'    gauss_temp, gauss                     src/temperature.cpp:262-336, src/md_utils.cpp:7-26
//
// The scan semantics are the reference's: every directive is looked up by rewinding the file and
// trying fscanf(template) at each whitespace-separated token, first match wins, unknown words are
// comments.  Directives outside the hot path (bonds, angles, linkage, eJump, red-ox chemistry,
// reciprocal Ewald) are rejected with AZ_ERR_UNSUPPORTED instead of being silently
// dropped.  Random draws (photon energies, Gaussian velocities) come from a seeded counter-based
// generator, not from the time-seeded rand() of src/main.cu:242 -- extra keys: `seed N` in
// control.txt; `rebuild N`, `skin X`, `device N`, `graph 0|1` in cuda.txt (the reference ignores unknown words; graph replay defaults to on below 100 000 atoms).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/aztot_b200.h"

namespace {

// unit factors, src/const.h:16-49
const double kPi = 3.14159265359;
const double kSqrtPi = sqrt(kPi);
const double kESI = 1.60217733E-19, kRSI = 1.0E-10, kTSI = 1.0E-12;
const double kMSI = kESI * kTSI * kTSI / kRSI / kRSI;
const double kMScale = 1.6605402E-27 / kMSI;
const double kKB = 1.3806488E-23 / (1.0 * 1.60217733E-19);

struct Rng { // counter-based stand-in for the reference's rand(): rand01() = rand() % 10000 / 10000
    uint64_t seed; uint64_t ctr = 0; uint32_t buf[4]; int have = 0;
    explicit Rng(uint64_t s) : seed(s) {}
    static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out)
    {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
        for (int r = 0; r < 10; r++) {
            uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += W0; k1 += W1;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
    uint32_t next()
    {
        if (!have) { philox((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x4445434Bu /* "DECK" */, 0, (uint32_t)seed, (uint32_t)(seed >> 32), buf); ctr++; have = 4; }
        return buf[--have];
    }
    double rand01() { return (double)(next() % 10000u) / 10000.0; }
};

// ---- the reference's directive scan (src/utils.cpp:87-195)
template <class T> bool scan_first(FILE* f, const char* templ, T* value)
{
    char skip[128];
    bool missed = true;
    rewind(f);
    while (!feof(f) && (missed = (fscanf(f, templ, value) <= 0)))
        if (fscanf(f, "%127s", skip) <= 0) break;
    return !missed;
}
int scan_number(FILE* f, const char* templ) { int x = 0; return scan_first(f, templ, &x) ? x : 0; } // find_number

struct VdwName { const char* abbr; int kind; int nparam; };
const VdwName kVdw[] = { { "lnjs", AZ_VDW_LJ, 2 }, { "buck", AZ_VDW_BUCK, 3 }, { "p746", AZ_VDW_P746, 3 }, { "bmhs", AZ_VDW_BMH, 5 },
                         { "elin", AZ_VDW_ELIN, 3 }, { "einv", AZ_VDW_EINV, 3 }, { "surk", AZ_VDW_SURK, 4 } };

} // namespace

struct az_deck {
    std::string err;
    std::string warnings;
    az_system_desc desc;
    az_cuda_opts opts;
    int nsteps = 0, stat_period = 1000, nstep_stat = 10, nstep_msd = 10;
    int tstat_step = 0;
    double rdf_rmax = 0.0, rdf_dr = 0.0;   // read_rdf, src/rdf.cpp:14-37
    int rdf_freq = 0, rdf_outfreq = 0, rdf_nuclei = 0;
    std::vector<double> xs, ys, zs, vxs, vys, vzs, photons;
    std::vector<int> types;
    std::vector<az_species> species;
    std::vector<az_vdw_entry> vdw;
    std::vector<int> spec_count;
    // nuclei (second word of a 'spec' line, read_spec src/sys_init.cpp:78-104): Field::nNucl, nnames, nnumbers, Spec::nuclei
    std::vector<std::string> nucl_names;
    std::vector<int> nucl_of_spec, nucl_count;
    // 'bonds' / 'bond_list' + bonds.txt, 'angles' / 'angle_list' + angles.txt
    std::vector<az_bond_type> bond_types;
    std::vector<az_angle_type> angle_types;
    std::vector<int> b_at1, b_at2, b_type, a_c, a_l1, a_l2, a_type;
};

static thread_local std::string g_deck_error;

#define DECK_FAIL(code, ...)                                       \
    do {                                                           \
        char b_[512]; snprintf(b_, sizeof(b_), __VA_ARGS__);       \
        d->err = b_; g_deck_error = b_;                            \
        if (f) fclose(f);                                          \
        if (out) *out = d; else delete d;                          \
        return code;                                               \
    } while (0)
#define DECK_WARN(...) do { char b_[512]; snprintf(b_, sizeof(b_), __VA_ARGS__); d->warnings += b_; d->warnings += "\n"; } while (0)

static int spec_by_name(const az_deck* d, const char* name)
{
    for (size_t i = 0; i < d->species.size(); i++)
        if (strcmp(d->species[i].name, name) == 0) return (int)i;
    return -1;
}

// photon_engs (src/temperature.cpp:28-89) for ONE uniform draw x: the photon energy y in 0..1 eV with
// (1 - x) exp(y / kT) = sum_{k<=4} (y / kT)^k / k!  (the CDF of P ~ e^4 exp(-e / kT)), by the reference's bisection: start at
// y = 0.5, tolerance 1e-3 on the residual, at most 20 halvings, then the previous atom's value.  Returns 0 when the
// reference would draw x again (no sign change on [0, 1]).  The reference feeds it rand01() of the time-seeded rand();
// the deck loader feeds it a seeded counter-based stream.
extern "C" int az_photon_energy(double x, double temperature, double prev, double* e_out)
{
    const double r24 = 1.0 / 24.0, r6 = 1.0 / 6.0; // src/temperature.cpp:18-19
    const double theta = 1.0 / (kKB * temperature), eps = 1e-3;
    auto prob4 = [&](double xx, double y) {
        const double ty = theta * y, ty2 = ty * ty;
        return (1 - xx) * exp(y * theta) - (r24 * ty2 * ty2 + r6 * ty2 * ty + 0.5 * ty * ty + ty + 1);
    };
    const double ra = prob4(x, 0.0), rb = prob4(x, 1.0);
    if (ra * rb > 0) return 0;
    double a = 0.0, b = 1.0, y = 0.5, r = prob4(x, y);
    int k = 0;
    while (r > eps || r < -eps) {
        if (r * ra < 0) { b = y; y = 0.5 * (a + y); } else { a = y; y = 0.5 * (y + b); }
        r = prob4(x, y);
        if (++k >= 20) { y = prev; break; }
    }
    if (e_out) *e_out = (y == y) ? y : prev;
    return 1;
}

extern "C" int az_deck_load(const char* dir, uint64_t seed, az_deck** out)
{
    az_deck* d = new az_deck();
    memset(&d->desc, 0, sizeof(d->desc));
    memset(&d->opts, 0, sizeof(d->opts));
    FILE* f = nullptr;
    if (!dir) DECK_FAIL(AZ_ERR_INVALID, "dir is NULL");
    std::string base = std::string(dir) + "/";

    // ------------------------------------------------------------------ field.txt (read_field)
    f = fopen((base + "field.txt").c_str(), "r");
    if (!f) DECK_FAIL(AZ_ERR_IO, "ERROR[001]! Fatal Error. Can't open Field.txt file");
    int nspec = scan_number(f, " spec %d");
    if (nspec <= 0) DECK_FAIL(AZ_ERR_IO, "ERROR[004]! There is no 'spec' section in the Field.txt file");
    if (nspec > AZ_MAX_SPECIES) DECK_FAIL(AZ_ERR_INVALID, "more than %d species", AZ_MAX_SPECIES);
    d->species.resize(nspec);
    bool charged = false;
    for (int i = 0; i < nspec; i++) { // read_spec: "%s %s %lf %lf %lf" name nucleus mass charge energy
        az_species& s = d->species[i];
        memset(&s, 0, sizeof(s));
        char nm[64], nucl[64]; double mass, q, en;
        if (fscanf(f, "%63s %63s %lf %lf %lf", nm, nucl, &mass, &q, &en) != 5) DECK_FAIL(AZ_ERR_IO, "bad species line %d in field.txt", i + 1);
        strncpy(s.name, nm, 7);
        s.mass = mass * kMScale;
        s.charge = q;
        if (q != 0.0) charged = true;
        int k = -1; // define nucleus (or add new), src/sys_init.cpp:88-103
        for (size_t j = 0; j < d->nucl_names.size(); j++) if (d->nucl_names[j] == nucl) k = (int)j;
        if (k < 0) { k = (int)d->nucl_names.size(); d->nucl_names.push_back(nucl); }
        d->nucl_of_spec.push_back(k);
    }
    { // directives whose physics is outside the hot path must not be dropped silently
        int n;
        if ((n = scan_number(f, " red-ox %d")) > 0) DECK_FAIL(AZ_ERR_UNSUPPORTED, "'red-ox %d': electron transfer chemistry is outside the hot path", n);
        if ((n = scan_number(f, " linkage %d")) > 0) DECK_FAIL(AZ_ERR_UNSUPPORTED, "'linkage %d': reactive bonding is outside the hot path", n);
        if ((n = scan_number(f, " evol_bonds %d")) > 0) DECK_FAIL(AZ_ERR_UNSUPPORTED, "'evol_bonds %d': reactive bonding is outside the hot path", n);
        if ((n = scan_number(f, " h-bonds %d")) > 0) DECK_FAIL(AZ_ERR_UNSUPPORTED, "'h-bonds %d': hydrogen bonds are outside the hot path", n);
        if ((n = scan_number(f, " angle_forming %d")) > 0) DECK_FAIL(AZ_ERR_UNSUPPORTED, "'angle_forming %d': reactive angles are outside the hot path", n);
    }
    { // bond types (read_bond, src/bonds.cpp:125-300): "n spec1 spec2 key params... low high"; constant bonds only ('con con')
        const int nb = scan_number(f, " bonds %d");
        if (nb > 0) {
            d->bond_types.assign(nb + 1, az_bond_type()); // [0] is reserved as 'empty bond' (src/sys_init.cpp:294)
            for (auto& b : d->bond_types) memset(&b, 0, sizeof(b));
            static const struct { const char* key; int kind, np; } kB[] = { { "harm", AZ_BOND_HARM, 2 }, { "mors", AZ_BOND_MORSE, 4 }, { "pdn", AZ_BOND_PEDONE, 5 },
                                                                            { "buck", AZ_BOND_BUCK, 3 }, { "e612", AZ_BOND_E612, 5 } };
            for (int i = 1; i <= nb; i++) {
                int num; char s1[64], s2[64], key[64];
                if (fscanf(f, "%d %63s %63s %63s", &num, s1, s2, key) != 4) DECK_FAIL(AZ_ERR_IO, "bad bond type line %d in field.txt", i);
                const int t1 = spec_by_name(d, s1), t2 = spec_by_name(d, s2);
                if (t1 < 0 || t2 < 0) DECK_FAIL(AZ_ERR_IO, "ERROR[124]: Unknown species in bonds declaration: %d %s %s", num, s1, s2);
                az_bond_type& b = d->bond_types[i];
                b.spec1 = t1; b.spec2 = t2;
                double p[5] = { 0, 0, 0, 0, 0 };
                int np = -1;
                for (const auto& k : kB) if (!strcmp(k.key, key)) { b.kind = k.kind; np = k.np; }
                if (np < 0) DECK_FAIL(AZ_ERR_IO, "ERROR[126]: Unknown potential type in bonds declaration: %d %s", num, key);
                for (int k = 0; k < np; k++) if (fscanf(f, " %lf", &p[k]) != 1) DECK_FAIL(AZ_ERR_IO, "bond type %d: too few parameters", i);
                b.p0 = p[0]; b.p1 = p[1]; b.p2 = p[2]; b.p3 = p[3]; b.p4 = p[4]; // E_scale = r_scale = 1 (src/const.h:39-41)
                for (int lim = 0; lim < 2; lim++) { // lower / upper limit of the bond: only constant bonds are on this path
                    if (fscanf(f, " %63s", key) != 1 || strcmp(key, "con")) DECK_FAIL(AZ_ERR_UNSUPPORTED, "bond type %d: '%s' limit: mutable / breakable bonds are reactive chemistry, outside the hot path", i, key);
                }
            }
        }
        const int na = scan_number(f, " angles %d ");
        if (na > 0) { // read_angle, src/angles.cpp:78-130: "n central key k cos0"
            d->angle_types.assign(na + 1, az_angle_type());
            for (auto& a : d->angle_types) memset(&a, 0, sizeof(a));
            for (int i = 1; i <= na; i++) {
                int num; char s1[64], key[64]; double p0, p1;
                if (fscanf(f, "%d %63s %63s %lf %lf", &num, s1, key, &p0, &p1) != 5) DECK_FAIL(AZ_ERR_IO, "bad angle type line %d in field.txt", i);
                const int t = spec_by_name(d, s1);
                if (t < 0) DECK_FAIL(AZ_ERR_IO, "ERROR[011]: Unknown species in angle declaration: %d %s %s", num, s1, key);
                if (strcmp(key, "hcos")) DECK_FAIL(AZ_ERR_IO, "ERROR[012]: Unknown potential type in angle declaration: %d %s", num, key);
                d->angle_types[i].kind = AZ_ANGLE_HCOS; d->angle_types[i].central = t; d->angle_types[i].p0 = p0; d->angle_types[i].p1 = p1;
            }
        }
    }
    const bool use_bond_list = scan_number(f, " bond_list %d") > 0, use_angle_list = scan_number(f, " angle_list %d") > 0;
    int nfrozen = scan_number(f, " frozensp %d");
    for (int i = 0; i < nfrozen; i++) {
        char nm[64];
        if (fscanf(f, "%63s", nm) != 1) break;
        int j = spec_by_name(d, nm);
        if (j >= 0) d->species[j].frozen = 1; else DECK_WARN("WARNING[b001] Unknown atom type (%s) in 'frozensp' section!", nm);
    }
    d->vdw.assign((size_t)nspec * nspec, az_vdw_entry());
    for (auto& v : d->vdw) memset(&v, 0, sizeof(v));
    double max_r_vdw = 0.0;
    int nvdw = scan_number(f, " vdw %d");
    if (nvdw <= 0) DECK_WARN("WARNING[001] no Van-der-Waals iteractions!");
    for (int i = 0; i < nvdw; i++) { // read_vdw
        char a[64], b[64], c[64]; double rcut, p[5] = { 0, 0, 0, 0, 0 };
        if (fscanf(f, " %63s %63s %63s %lf %lf %lf ", a, b, c, &rcut, &p[0], &p[1]) != 6) DECK_FAIL(AZ_ERR_IO, "bad vdw line %d in field.txt", i + 1);
        const VdwName* vn = nullptr;
        for (const VdwName& k : kVdw) if (strcmp(k.abbr, c) == 0) vn = &k;
        if (!vn) DECK_FAIL(AZ_ERR_IO, "ERROR[006]! Unknown potential type (%s) in %d vdw-line", c, i + 1);
        for (int k = 2; k < vn->nparam; k++) if (fscanf(f, " %lf", &p[k]) != 1) DECK_FAIL(AZ_ERR_IO, "vdw line %d: too few parameters", i + 1);
        int t1 = spec_by_name(d, a), t2 = spec_by_name(d, b);
        if (t1 < 0 || t2 < 0) DECK_FAIL(AZ_ERR_IO, "ERROR[005]! Unknown atom type in vdw-line: %s   %s   %s", a, b, c);
        if (vn->kind == AZ_VDW_P746) DECK_FAIL(AZ_ERR_UNSUPPORTED, "vdw type p746 exists only in the reference's serial build");
        if (rcut > max_r_vdw) max_r_vdw = rcut;
        az_vdw_entry e; memset(&e, 0, sizeof(e));
        e.kind = vn->kind; e.r2cut = rcut * rcut;
        // unit scales are 1 except LJ's factor 4 (vdw_scale0, src/vdw.cpp:215); LJ: p0 = 4e, p1 = s^2, p2 = 24e
        if (vn->kind == AZ_VDW_LJ) { p[0] *= 4.0; p[1] = p[1] * p[1]; p[2] = 6.0 * p[0]; }
        e.p0 = p[0]; e.p1 = p[1]; e.p2 = p[2]; e.p3 = p[3]; e.p4 = p[4];
        e.use_radii = (vn->kind == AZ_VDW_SURK);
        if (d->vdw[t1 * nspec + t2].kind) DECK_WARN("WARNING[002]: Pair potential between %s and %s redeclarated", a, b);
        d->vdw[t1 * nspec + t2] = e;
        if (vn->kind != AZ_VDW_SURK) d->vdw[t2 * nspec + t1] = e;
        else if (t1 != t2) { // mirrored entry (k <-> l) so that a full neighbour list sees the same pair term from both sides
            az_vdw_entry m = e; m.p2 = e.p3; m.p3 = e.p2;
            d->vdw[t2 * nspec + t1] = m;
        }
    }
    { // radii section: nSpec lines "name A B Umax" (src/sys_init.cpp:468-480)
        int j;
        if (scan_first(f, " radii %d", &j))
            for (int i = 0; i < nspec; i++) {
                char nm[64];
                if (fscanf(f, "%63s", nm) != 1) break;
                int t = spec_by_name(d, nm);
                if (t < 0) DECK_FAIL(AZ_ERR_IO, "ERROR[b018] wrong species(%s) in radii section(%d)", nm, i + 1);
                if (fscanf(f, "%lf %lf %lf", &d->species[t].radA, &d->species[t].radB, &d->species[t].mxEng) != 3)
                    DECK_FAIL(AZ_ERR_IO, "bad radii line %d", i + 1);
            }
    }
    fclose(f); f = nullptr;

    // ------------------------------------------------------------------ atoms.xyz (read_atoms_box)
    f = fopen((base + "atoms.xyz").c_str(), "r");
    if (!f) DECK_FAIL(AZ_ERR_IO, "ERROR[007] Can't open file 'atoms.xyz'");
    int n = 0, boxtype = 0;
    double la = 0, lb = 0, lc = 0;
    if (fscanf(f, "%d", &n) != 1 || n <= 0) DECK_FAIL(AZ_ERR_IO, "atoms.xyz: bad atom count");
    if (fscanf(f, "%d", &boxtype) != 1 || boxtype != 1) DECK_FAIL(AZ_ERR_IO, "ERROR[008] Unknown box type!");
    if (fscanf(f, "%lf %lf %lf", &la, &lb, &lc) != 3) DECK_FAIL(AZ_ERR_IO, "atoms.xyz: bad box line");
    d->xs.resize(n); d->ys.resize(n); d->zs.resize(n); d->types.resize(n);
    d->vxs.assign(n, 0.0); d->vys.assign(n, 0.0); d->vzs.assign(n, 0.0);
    d->spec_count.assign(nspec, 0);
    for (int i = 0; i < n; i++) {
        char nm[64];
        if (fscanf(f, "%63s %lf %lf %lf", nm, &d->xs[i], &d->ys[i], &d->zs[i]) != 4) DECK_FAIL(AZ_ERR_IO, "atoms.xyz: bad atom line %d", i + 1);
        int t = spec_by_name(d, nm);
        if (t < 0) DECK_FAIL(AZ_ERR_IO, "ERROR[009]! unknown atom[%d] type=%s in atoms.xyz file", i + 1, nm);
        d->types[i] = t;
        d->spec_count[t]++;
    }
    fclose(f); f = nullptr;
    d->nucl_count.assign(d->nucl_names.size(), 0); // Field::nnumbers, src/sys_init.cpp:539
    for (int i = 0; i < n; i++) d->nucl_count[d->nucl_of_spec[d->types[i]]]++;

    // ------------------------------------------------------------------ bonds.txt / angles.txt (read_bondlist, read_anglelist)
    if (use_bond_list && !d->bond_types.empty()) {
        f = fopen((base + "bonds.txt").c_str(), "r");
        if (!f) DECK_WARN("WARNING[a001] bond list is used, but there is no such file. No bonds are downloaded");
        else {
            int nb = 0;
            if (fscanf(f, "%d", &nb) != 1 || nb < 0) DECK_FAIL(AZ_ERR_IO, "bonds.txt: bad bond count");
            for (int i = 0; i < nb; i++) {
                int a1, a2, k;
                if (fscanf(f, "%d %d %d", &a1, &a2, &k) != 3) DECK_FAIL(AZ_ERR_IO, "bonds.txt: bad line %d", i + 1);
                if (a1 < 0 || a1 >= n || a2 < 0 || a2 >= n || k <= 0 || k >= (int)d->bond_types.size()) DECK_FAIL(AZ_ERR_IO, "bonds.txt: line %d: atom or type out of range", i + 1);
                const az_bond_type& b = d->bond_types[k]; // turn the bond if necessary and verify its type, src/bonds.cpp:50-77
                if (b.spec1 == d->types[a1]) { if (b.spec2 != d->types[a2]) DECK_FAIL(AZ_ERR_IO, "ERROR [121] incorrect type of 2th atom in bond (type: %d, line: %d)", k, i); }
                else if (b.spec1 == d->types[a2] && b.spec2 == d->types[a1]) std::swap(a1, a2);
                else DECK_FAIL(AZ_ERR_IO, "ERROR [123] incorrect type of atoms for bond type(%d) in bond list, line: %d", k, i);
                d->b_at1.push_back(a1); d->b_at2.push_back(a2); d->b_type.push_back(k);
            }
            fclose(f); f = nullptr;
        }
    }
    if (use_angle_list) {
        if (d->angle_types.empty()) DECK_WARN("WARNING[b006] 'anlge_list' directive is ignored, because there are no angle type defintions");
        else {
            f = fopen((base + "angles.txt").c_str(), "r");
            if (!f) DECK_WARN("WARNING[a002] angle list is used, but there is no such file. No anlges are downloaded");
            else {
                int na = 0;
                if (fscanf(f, "%d", &na) != 1 || na < 0) DECK_FAIL(AZ_ERR_IO, "angles.txt: bad angle count");
                for (int i = 0; i < na; i++) {
                    int c, l1, l2, x;
                    if (fscanf(f, "%d %d %d %d", &c, &l1, &l2, &x) != 4) DECK_FAIL(AZ_ERR_IO, "angles.txt: bad line %d", i + 1);
                    if (c < 0 || c >= n || l1 < 0 || l1 >= n || l2 < 0 || l2 >= n) DECK_FAIL(AZ_ERR_IO, "angles.txt: line %d: atom out of range", i + 1);
                    if (x <= 0 || x >= (int)d->angle_types.size()) DECK_FAIL(AZ_ERR_IO, "ERROR[013] wrong atom type number in angles.txt");
                    if (d->types[c] != d->angle_types[x].central) DECK_FAIL(AZ_ERR_IO, "ERROR[014] wrong central atom type in angle list (%d postion)", i);
                    d->a_c.push_back(c); d->a_l1.push_back(l1); d->a_l2.push_back(l2); d->a_type.push_back(x);
                }
                fclose(f); f = nullptr;
            }
        }
    }

    // ------------------------------------------------------------------ control.txt (read_sim)
    f = fopen((base + "control.txt").c_str(), "r");
    if (!f) DECK_FAIL(AZ_ERR_IO, "ERROR[410]: Can't open file 'control.txt'");
    double tSt = 0.0, tSim = 0.0, tEq = 0.0;
    if (!scan_first(f, " timestep %lf ", &tSt)) DECK_FAIL(AZ_ERR_IO, "ERROR[411]: timestep must be declared in control.txt file!");
    int nSt = 0;
    if (!scan_first(f, " timesim %lf ", &tSim)) {
        if (!scan_first(f, " nstep %d", &nSt)) DECK_FAIL(AZ_ERR_IO, "ERROR[412]: no 'nstep' or 'timesim' directives in control.txt file!");
    } else nSt = (int)(tSim / tSt);
    int nEq = 0, freqEq = 0;
    if (!scan_first(f, " timeequil %lf ", &tEq)) {
        if (!(nEq = scan_number(f, " nequil %d "))) DECK_WARN("WARNING[002]: no 'nequil' or 'timeequil' directives in control.txt file - there is no equilibration period!");
    } else nEq = (int)(tEq / tSt);
    if (nEq && !(freqEq = scan_number(f, " eqfreq %d "))) DECK_WARN("WARNING[003]: no t-Scale during equlibration period!");
    uint64_t deck_seed = seed;
    { long long sd; if (scan_first(f, " seed %lld ", &sd)) deck_seed = seed ? seed : (uint64_t)sd; }
    if (!deck_seed) deck_seed = 12345;
    Rng rng(deck_seed);

    // thermostat (read_tstat)
    double temp = 0.0, tstat_tau = 0.0;
    int tstat_kind = AZ_TSTAT_NONE;
    if (!scan_first(f, " temperature %lf ", &temp)) DECK_FAIL(AZ_ERR_IO, "ERROR[404]: temperature is not defined in control.txt file!");
    {
        char ts[64] = "";
        if (fscanf(f, "%63s", ts) != 1) DECK_FAIL(AZ_ERR_IO, "ERROR[405]: unknown thermostat type!");
        if (!strcmp(ts, "none")) tstat_kind = AZ_TSTAT_NONE;
        else if (!strcmp(ts, "nose")) { // read_tstat, src/temperature.cpp:104-110
            if (fscanf(f, " %lf ", &tstat_tau) != 1 || !(tstat_tau > 0.0)) DECK_FAIL(AZ_ERR_IO, "'temperature ... nose': the relaxation time tau is missing");
            tstat_kind = AZ_TSTAT_NOSE;
        }
        else if (!strcmp(ts, "radi")) {
            if (fscanf(f, "%d", &d->tstat_step) != 1) DECK_FAIL(AZ_ERR_IO, "ERROR[a002]: there is no step parameter for radiative thermostat!");
            tstat_kind = AZ_TSTAT_RADI;
        } else DECK_FAIL(AZ_ERR_IO, "ERROR[405]: unknown thermostat type!");
    }
    // electrostatics (read_elec)
    int elec_kind = AZ_ELEC_NONE;
    double rReal = 0.0, alpha = 0.0;
    int ewald_k[3] = { 0, 0, 0 };
    {
        char es[64] = "";
        if (!scan_first(f, " elec %63s", es)) DECK_FAIL(AZ_ERR_IO, "ERROR[401]: electrostatic calculations are not specified in control.txt. Use 'elec' directive!");
        if (!strcmp(es, "none")) {
            if (charged) DECK_WARN("WARNING[b003] The species have charges, but electrostatic directive is none. Charges will be ignored!");
        } else if (!strcmp(es, "dir")) { elec_kind = AZ_ELEC_DIRECT; if (fscanf(f, " %lf ", &rReal) != 1) DECK_FAIL(AZ_ERR_IO, "elec dir: missing cutoff"); }
        else if (!strcmp(es, "pme")) {
            int kx, ky, kz;
            elec_kind = AZ_ELEC_EWALD;
            if (fscanf(f, " %lf %lf %d %d %d", &rReal, &alpha, &kx, &ky, &kz) != 5) DECK_FAIL(AZ_ERR_IO, "elec pme: missing parameters");
            ewald_k[0] = kx; ewald_k[1] = ky; ewald_k[2] = kz;
        } else if (!strcmp(es, "fenn")) { elec_kind = AZ_ELEC_FENNEL; if (fscanf(f, " %lf %lf", &rReal, &alpha) != 2) DECK_FAIL(AZ_ERR_IO, "elec fenn: missing parameters"); }
        else DECK_FAIL(AZ_ERR_IO, "ERROR[404]: Unknown type of electrostatic calculations: %s.", es);
        if (!charged && elec_kind) {
            DECK_WARN("WARNING[b004] The species do not have charges, but electrostatic directive is %s. The electrostatic type is switched to none!", es);
            elec_kind = AZ_ELEC_NONE; // rReal stays set, like the reference (src/elec.cpp:52-57)
        }
    }
    double eps = 1.0;
    if (!scan_first(f, " permittivity %lf ", &eps)) { eps = 1.0; DECK_WARN("WARNING[131]: permittivity was not defined in control.txt file. used value of 1.0!"); }
    // init_vel
    int init_vel = 0; // 0 zero, 1 gaus, 2 const, 3 keng
    {
        char vs[64] = "";
        if (!scan_first(f, " init_vel %63s", vs)) DECK_FAIL(AZ_ERR_IO, "ERROR[406] No init_vel directive in control.txt file!");
        if (!strcmp(vs, "zero")) init_vel = 0;
        else if (!strcmp(vs, "gaus")) init_vel = 1;
        else if (!strcmp(vs, "const")) {
            double vx, vy, vz;
            if (fscanf(f, "%lf %lf %lf", &vx, &vy, &vz) != 3) DECK_FAIL(AZ_ERR_IO, "init_vel const: missing components");
            for (int i = 0; i < n; i++) { d->vxs[i] = vx; d->vys[i] = vy; d->vzs[i] = vz; }
            init_vel = 2;
        } else if (!strcmp(vs, "keng")) {
            double ekin;
            if (fscanf(f, "%lf", &ekin) != 1) DECK_FAIL(AZ_ERR_IO, "init_vel keng: missing energy");
            for (int i = 0; i < n; i++) { // src/sys_init.cpp:767-786
                double vel = sqrt(2.0 * ekin / d->species[d->types[i]].mass);
                double phi = (double)(rng.next() % 32) / 32.0 * 2.0 * kPi, theta = (double)(rng.next() % 32) / 32.0 * 2.0 * kPi;
                double cost = cos(theta);
                d->vzs[i] = sin(theta) * vel; d->vys[i] = sin(phi) * cost * vel; d->vxs[i] = cos(phi) * cost * vel;
            }
            init_vel = 3;
        } else DECK_FAIL(AZ_ERR_IO, "ERROR[407] Unknown value of init_vel directive in control.txt file!");
    }
    if (scan_number(f, " eJump %d ") != 0) DECK_FAIL(AZ_ERR_UNSUPPORTED, "'eJump': electron hopping is outside the hot path");
    double U[3] = { 0, 0, 0 };
    if (scan_first(f, " elecfield %lf ", &U[0])) { if (fscanf(f, " %lf %lf ", &U[1], &U[2]) != 2) { U[1] = U[2] = 0.0; } }
    else U[0] = 0.0;
    const int reset_vels = scan_number(f, " reset_vels %d "); // src/sys_init.cpp:857
    if (reset_vels < 0) DECK_FAIL(AZ_ERR_IO, "'reset_vels' must not be negative");
    { int mn; if (scan_first(f, " max_neigh %d ", &mn) && mn > 0) d->opts.max_neigh = mn; } // rows of the Verlet list per atom
    double cell_size = 0.0;
    if (!scan_first(f, " cell_list %lf ", &cell_size)) DECK_FAIL(AZ_ERR_IO, "'cell_list' is required by the CUDA path (desired_cell_size, src/cuInit.cu:979)");
    { // read_rdf (src/rdf.cpp:14-37): rdf rmax dr freq outfreq [nucl]
        double r;
        if (!scan_first(f, " rdf %lf ", &r)) DECK_WARN("note: no 'rdf' directive (the reference demands one, ERROR[408]): no RDF is sampled");
        else {
            char word[64] = "";
            if (fscanf(f, " %lf %d %d", &d->rdf_dr, &d->rdf_freq, &d->rdf_outfreq) != 3 || !(d->rdf_dr > 0.0) || d->rdf_freq <= 0)
                DECK_FAIL(AZ_ERR_IO, "rdf directive: expected 'rdf rmax dr freq outfreq [nucl]'");
            d->rdf_rmax = r;
            if (fscanf(f, " %63s", word) == 1 && !strcmp(word, "nucl")) {
                d->rdf_nuclei = 1; // nrdf_iter / copy_nrdf instead of rdf_iter / copy_rdf (src/main.cu:392-395): rdf.dat and rdf_n.dat
            }
        }
    }
    d->stat_period = 1000;
    if (!scan_first(f, " stat %d ", &d->stat_period)) { d->stat_period = 1000; DECK_WARN("WARNING[133]: stat directive is not specified, default value of 1000 is used"); }
    if (d->stat_period <= 0) d->stat_period = 1000;
    fclose(f); f = nullptr;

    // ------------------------------------------------------------------ cuda.txt (read_cuda)
    f = fopen((base + "cuda.txt").c_str(), "r");
    if (!f) DECK_FAIL(AZ_ERR_IO, "ERROR[a001]! Fatal Error. Can't open cuda.txt file");
    if (!scan_first(f, " nstep stat %d", &d->nstep_stat)) d->nstep_stat = 10;
    if (!scan_first(f, " nstep msdstat %d", &d->nstep_msd)) d->nstep_msd = 10;
    { int v; float s; // extensions; 'nthread a/b' of the reference are accepted and ignored
        if (scan_first(f, " rebuild %d", &v)) d->opts.rebuild_every = v;
        if (scan_first(f, " skin %f", &s)) d->opts.skin = s;
        // replay the step loop between two 'stat' samples as one CUDA graph.  Default: on for the latency-bound decks (fewer than
        // 100 000 atoms: the case studies' steps are 6-7 kernels of a few microseconds each, so the gaps between launches are a
        // fifth of the step), off for large systems where a step is milliseconds; 'graph 0|1' overrides.  Results are bitwise
        // those of the plain launches (case_graph_replay).
        if (scan_first(f, " graph %d", &v)) d->opts.use_graph = v;
        else d->opts.use_graph = (n < 100000) ? 1 : 0;
        if (scan_first(f, " device %d", &v)) d->opts.device = v;
        if (scan_first(f, " list_format %d", &v) && v >= 0 && v <= 2) d->opts.list_format = v; // 1 = global rows, 2 = tile lists (include/aztot_b200.h)
        if (scan_first(f, " tile_atoms %d", &v) && v > 0) d->opts.tile_atoms = v;
    }
    fclose(f); f = nullptr;

    // ------------------------------------------------------------------ derived values (init_md)
    const bool use_elec = elec_kind != AZ_ELEC_NONE;
    double rMax = use_elec ? rReal : (nvdw > 0 ? max_r_vdw : 0.0); // src/sys_init.cpp:1059-1072
    long long degFree = 3LL * n - (long long)d->b_at1.size() - (tstat_kind ? 1 : 0); // :1099-1102 (Field::nBonds)
    double tKin = 0.5 * temp * kKB * (double)degFree;               // :1106
    if (tstat_kind == AZ_TSTAT_RADI) { // photon_engs: inverse CDF of P ~ e^4 exp(-e/kT) on 0..1 eV by bisection
        d->photons.assign(n, 0.0);
        for (int i = 0; i < n; i++) {
            double y = 0.0;
            int guard = 0;
            while (!az_photon_energy(rng.rand01(), temp, i > 0 ? d->photons[i - 1] : 0.5, &y) && ++guard < 1000) {}
            d->photons[i] = y;
        }
    }
    if (init_vel == 1) { // gauss_temp
        auto gauss = [&](double stdev, double mean) {
            double v1 = 0, v2 = 0, r = 2.0;
            while (r > 1.0 || r == 0.0) { v1 = 2.0 * rng.rand01() - 1.0; v2 = 2.0 * rng.rand01() - 1.0; r = v1 * v1 + v2 * v2; }
            return mean + stdev * v1 * sqrt(-2.0 * log(r) / r);
        };
        double cp[3] = { 0, 0, 0 }, tot = 0.0;
        for (int i = 0; i < n; i++) {
            d->vxs[i] = gauss(0.5, 0.0); d->vys[i] = gauss(0.5, 0.0); d->vzs[i] = gauss(0.5, 0.0);
            double m = d->species[d->types[i]].mass;
            cp[0] += d->vxs[i] * m; cp[1] += d->vys[i] * m; cp[2] += d->vzs[i] * m; tot += m;
        }
        for (int a = 0; a < 3; a++) cp[a] /= tot;
        double kE = 0.0;
        for (int i = 0; i < n; i++) {
            d->vxs[i] -= cp[0]; d->vys[i] -= cp[1]; d->vzs[i] -= cp[2];
            kE += d->species[d->types[i]].mass * (d->vxs[i] * d->vxs[i] + d->vys[i] * d->vys[i] + d->vzs[i] * d->vzs[i]);
        }
        kE *= 0.5;
        double k = kE > 0 ? sqrt(tKin / kE) : 0.0;
        for (int i = 0; i < n; i++) { d->vxs[i] *= k; d->vys[i] *= k; d->vzs[i] *= k; }
    }

    az_system_desc& s = d->desc;
    s.abi_version = AZ_ABI_VERSION;
    s.n_atoms = n; s.n_species = nspec;
    s.xs = d->xs.data(); s.ys = d->ys.data(); s.zs = d->zs.data();
    s.vxs = d->vxs.data(); s.vys = d->vys.data(); s.vzs = d->vzs.data();
    s.types = d->types.data();
    s.la = la; s.lb = lb; s.lc = lc;
    s.species = d->species.data(); s.vdw = d->vdw.data(); s.max_r_vdw = max_r_vdw;
    s.elec_kind = elec_kind; s.r_real = rReal; s.alpha = alpha; s.eps = eps;
    if (elec_kind == AZ_ELEC_EWALD) { // prepare_elec, src/elec.cpp:376-397: the engine derives the k-space constants itself
        s.daipi2 = 2 * alpha / kSqrtPi;
        for (int a = 0; a < 3; a++) s.ewald_k[a] = ewald_k[a];
    }
    if (elec_kind == AZ_ELEC_FENNEL) { // prepare_elec, src/elec.cpp:399-405
        double aRc = alpha * rReal;
        s.daipi2 = 2 * alpha / kSqrtPi;
        s.el_c1 = erfc(aRc) / rReal;
        s.el_c2 = erfc(aRc) / (rReal * rReal) + s.daipi2 * exp(-aRc * aRc) / rReal;
    }
    s.tstat_kind = tstat_kind; s.temperature = temp; s.t_kin = tKin;
    s.tstat_tau = tstat_tau; s.reset_vels = reset_vels;
    s.photons = d->photons.empty() ? nullptr : d->photons.data();
    s.seed = deck_seed;
    s.timestep = tSt; s.n_equil = nEq; s.eq_freq = freqEq;
    s.r2_max = rMax * rMax;
    s.cell_size = cell_size;
    for (int a = 0; a < 3; a++) s.efield[a] = U[a];
    if (!d->b_at1.empty()) {
        s.n_bond_types = (int)d->bond_types.size(); s.bond_types = d->bond_types.data();
        s.n_bonds = (int)d->b_at1.size(); s.bond_at1 = d->b_at1.data(); s.bond_at2 = d->b_at2.data(); s.bond_type = d->b_type.data();
    }
    if (!d->a_c.empty()) {
        s.n_angle_types = (int)d->angle_types.size(); s.angle_types = d->angle_types.data();
        s.n_angles = (int)d->a_c.size(); s.angle_centre = d->a_c.data(); s.angle_lig1 = d->a_l1.data(); s.angle_lig2 = d->a_l2.data(); s.angle_type = d->a_type.data();
    }
    d->nsteps = nSt;
    *out = d;
    return AZ_OK;
}

extern "C" const az_system_desc* az_deck_desc(const az_deck* d) { return d ? &d->desc : nullptr; }
extern "C" const az_cuda_opts* az_deck_opts(const az_deck* d) { return d ? &d->opts : nullptr; }
extern "C" int az_deck_nsteps(const az_deck* d) { return d ? d->nsteps : 0; }
extern "C" int az_deck_stat_period(const az_deck* d) { return d ? d->stat_period : 0; }
extern "C" const char* az_deck_error(const az_deck* d) { return d ? d->err.c_str() : g_deck_error.c_str(); }
extern "C" const char* az_deck_warnings(const az_deck* d) { return d ? d->warnings.c_str() : ""; }
extern "C" int az_deck_nstep_stat(const az_deck* d) { return d ? d->nstep_stat : 0; }
extern "C" int az_deck_species_count(const az_deck* d, int t) { return (d && t >= 0 && t < (int)d->spec_count.size()) ? d->spec_count[t] : 0; }
extern "C" int az_deck_rdf(const az_deck* d, double* r_max, double* dr, int* freq, int* outfreq)
{
    if (!d || d->rdf_freq <= 0) return 0;
    if (r_max) *r_max = d->rdf_rmax;
    if (dr) *dr = d->rdf_dr;
    if (freq) *freq = d->rdf_freq;
    if (outfreq) *outfreq = d->rdf_outfreq;
    return 1;
}
extern "C" int az_deck_nuclei(const az_deck* d, int* nucleus_of_species)
{
    if (!d) return 0;
    if (nucleus_of_species) for (size_t i = 0; i < d->nucl_of_spec.size(); i++) nucleus_of_species[i] = d->nucl_of_spec[i];
    return (int)d->nucl_names.size();
}
extern "C" const char* az_deck_nucleus_name(const az_deck* d, int k) { return (d && k >= 0 && k < (int)d->nucl_names.size()) ? d->nucl_names[k].c_str() : ""; }
extern "C" int az_deck_nucleus_count(const az_deck* d, int k) { return (d && k >= 0 && k < (int)d->nucl_count.size()) ? d->nucl_count[k] : 0; }
extern "C" int az_deck_rdf_nuclei(const az_deck* d) { return d ? d->rdf_nuclei : 0; }
extern "C" void az_deck_free(az_deck* d) { delete d; }
