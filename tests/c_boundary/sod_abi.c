/* sod_abi.c -- the call sequence of INTEGRATION.md section 2 from plain C, with no Python in the way: the closest
 * stand-in for the D shim (extern(C) declarations over the same symbols) this image allows.
 *
 *   init -> block_create -> block_set_geometry -> block_set_bc -> commit -> upload_flow -> compute_dt ->
 *   step x N (with the retry rule of simcore_gasdynamic_step.d:995-999) -> download_flow / download_conserved -> finalize
 *
 * on Sod's shock tube (examples/eilmer/3D/sod-shock-tube/sg/sod.lua: L = 1, p = 1e5 / 1e4 Pa, T = 348.4 / 278.8 K,
 * ideal air, closed ends), one block of 64 x 4 x 4 cells.  The same driver runs the product library (prefix eb200_)
 * and the CPU oracle (prefix orc_), both loaded with dlopen, and compares the conserved quantities: the FMA-free
 * build must be bit-identical, the throughput build within 1e-10.
 *
 * usage: sod_abi <libeb200.so | -> <liboracle.so> [nsteps]      ("-": oracle only, for machines without a GPU)
 * exit code 0 = agreement.  Test infrastructure (tests/test_c_boundary.py). */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/eb200.h"

#define NI 64
#define NJ 4
#define NK 4
#define NG EB200_NGHOST

typedef struct {
    void* h;
    int (*init)(const eb200_config*);
    int (*finalize)(int);
    int (*last_error)(char*, int);
    int (*block_create)(int, int, int, int, int, int);
    int (*block_set_geometry)(int, int, const double*, const double*, const double*, const double*, const double*, const double* const*);
    int (*block_set_bc)(int, int, int, int, const double*, int, int, int, int);
    int (*commit)(int);
    int (*upload_flow)(int, int, const double* const*, int);
    int (*download_conserved)(int, int, double* const*, int);
    int (*compute_dt)(int, double, double, int, double*);
    int (*step)(int, double, double, int*);
} Api;

static void* sym(void* h, const char* prefix, const char* name)
{
    char buf[128];
    snprintf(buf, sizeof buf, "%s%s", prefix, name);
    void* p = dlsym(h, buf);
    if (!p) { fprintf(stderr, "missing symbol %s\n", buf); exit(2); }
    return p;
}

static int load(Api* a, const char* path, const char* prefix)
{
    a->h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!a->h) { fprintf(stderr, "dlopen %s: %s\n", path, dlerror()); return -1; }
    *(void**)&a->init = sym(a->h, prefix, "init");
    *(void**)&a->finalize = sym(a->h, prefix, "finalize");
    *(void**)&a->last_error = sym(a->h, prefix, "last_error");
    *(void**)&a->block_create = sym(a->h, prefix, "block_create");
    *(void**)&a->block_set_geometry = sym(a->h, prefix, "block_set_geometry");
    *(void**)&a->block_set_bc = sym(a->h, prefix, "block_set_bc");
    *(void**)&a->commit = sym(a->h, prefix, "commit");
    *(void**)&a->upload_flow = sym(a->h, prefix, "upload_flow");
    *(void**)&a->download_conserved = sym(a->h, prefix, "download_conserved");
    *(void**)&a->compute_dt = sym(a->h, prefix, "compute_dt");
    *(void**)&a->step = sym(a->h, prefix, "step");
    return 0;
}

#define CHECK(a, call) do { int rc_ = (call); if (rc_ < 0) { char e_[512]; (a)->last_error(e_, 512); \
    fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, e_); return -1; } } while (0)

/* runs the job through one implementation of the ABI; U[q] = conserved quantity q in the padded block layout */
static int run(Api* a, int strict, int nsteps, double** U, double* dt_used, int* steps_done)
{
    const int PI = NI + 2 * NG, PJ = NJ + 2 * NG, PK = NK + 2 * NG;
    const size_t n = (size_t)PI * PJ * PK;
    eb200_config c;
    memset(&c, 0, sizeof c);
    c.dimensions = 3; c.gas_model = EB200_GAS_IDEAL; c.n_species = 1; c.flux_calculator = EB200_FLUX_AUSMDV;
    c.interpolation_order = 2; c.apply_limiter = 1; c.extrema_clipping = 1; c.interpolate_in_local_frame = 1;
    c.apply_entropy_fix = 1; c.update_scheme = EB200_UPDATE_PC; c.max_invalid_cells = 0; c.strict_fp = strict;
    c.thermo_interpolator = EB200_INTERP_RHOU; c.epsilon_van_albada = 1.0e-12; c.M_inf = 0.01;
    c.max_velocity = 30000.0; c.max_temp = 50000.0; c.min_temp = 0.0; c.suggested_low_T_value = 200.0;
    c.ignore_low_T_thermo_update_failure = 1; c.strict_shock_detector = 1;
    c.ideal_mol_mass = 0.02896; c.ideal_gamma = 1.4; c.compression_tolerance = -0.30; c.shear_tolerance = 0.20;
    int sim = a->init(&c);
    if (sim < 0) { char e[512]; a->last_error(e, 512); fprintf(stderr, "init failed: %s\n", e); return -1; }
    CHECK(a, a->block_create(sim, 0, NI, NJ, NK, 0));
    /* geometry of a uniform box (what compute_primary_cell_geometric_data leaves in FVCell / FVInterface):
     * face d has (n, t1, t2) = (e_d, e_d+1, e_d+2), sfluidblock.d:735-850 */
    const double dx = 1.0 / NI, dy = 0.1 / NJ, dz = 0.1 / NK;
    double* vol = malloc(n * 8); double* len[3]; double* face[3];
    const double h[3] = { dx, dy, dz };
    for (int d = 0; d < 3; ++d) { len[d] = malloc(n * 8); face[d] = calloc(10 * n, 8); }
    for (size_t m = 0; m < n; ++m) {
        vol[m] = dx * dy * dz;
        for (int d = 0; d < 3; ++d) {
            len[d][m] = h[d];
            face[d][(size_t)(0 + d) * n + m] = 1.0;                    /* n  = e_d   */
            face[d][(size_t)(3 + (d + 1) % 3) * n + m] = 1.0;          /* t1 = e_d+1 */
            face[d][(size_t)(6 + (d + 2) % 3) * n + m] = 1.0;          /* t2 = e_d+2 */
            face[d][(size_t)9 * n + m] = h[(d + 1) % 3] * h[(d + 2) % 3];
        }
    }
    const double* faces[3] = { face[0], face[1], face[2] };
    CHECK(a, a->block_set_geometry(sim, 0, vol, NULL, len[0], len[1], len[2], faces));
    for (int f = 0; f < 6; ++f) CHECK(a, a->block_set_bc(sim, 0, f, EB200_BC_WALL_WITH_SLIP, NULL, 0, -1, -1, 0));
    CHECK(a, a->commit(sim));
    /* FlowStates: rho, u, p, T, a, velx, vely, velz (ideal_gas.d:89-144) */
    const double R = 8.31451 / 0.02896, g = 1.4, Cv = R / (g - 1.0);
    double* prim[EB200_NPRIM_BASE];
    for (int v = 0; v < EB200_NPRIM_BASE; ++v) prim[v] = calloc(n, 8);
    for (int k = 0; k < PK; ++k) for (int j = 0; j < PJ; ++j) for (int i = 0; i < PI; ++i) {
        const size_t m = ((size_t)k * PJ + j) * PI + i;
        const double x = (i - NG + 0.5) * dx;
        const double p = x < 0.5 ? 1.0e5 : 1.0e4, T = x < 0.5 ? 348.4 : 278.8;
        prim[0][m] = p / (R * T); prim[1][m] = Cv * T; prim[2][m] = p; prim[3][m] = T; prim[4][m] = sqrt(g * R * T);
    }
    const double* cprim[EB200_NPRIM_BASE];
    for (int v = 0; v < EB200_NPRIM_BASE; ++v) cprim[v] = prim[v];
    CHECK(a, a->upload_flow(sim, 0, cprim, EB200_NPRIM_BASE));
    /* determine_time_step_size at step 0 (simcore_gasdynamic_step.d:60-159), then fixed; retry rule :995-999 */
    double out[3];
    CHECK(a, a->compute_dt(sim, 1.0e-6, 0.5, 0, out));
    double dt = out[0] < 1.0e-6 ? out[0] : 1.0e-6, t = 0.0;
    int done = 0;
    for (int s = 0; s < nsteps; ++s) {
        int nbad = 0, attempt = 0, rc;
        while ((rc = a->step(sim, t, dt, &nbad)) == 1 && ++attempt < 3) dt *= 0.2;
        if (rc != 0) { char e[512]; a->last_error(e, 512); fprintf(stderr, "step %d failed (%d): %s\n", s, rc, e); return -1; }
        t += dt; ++done;
    }
    for (int q = 0; q < 5; ++q) U[q] = calloc(n, 8);
    CHECK(a, a->download_conserved(sim, 0, U, 5));
    CHECK(a, a->finalize(sim));
    *dt_used = dt; *steps_done = done;
    free(vol);
    for (int d = 0; d < 3; ++d) { free(len[d]); free(face[d]); }
    for (int v = 0; v < EB200_NPRIM_BASE; ++v) free(prim[v]);
    return 0;
}

static double compare(double** A, double** B, int* identical)
{
    const int PI = NI + 2 * NG, PJ = NJ + 2 * NG;
    double worst = 0.0, scale[5] = { 0, 0, 0, 0, 0 };
    *identical = 1;
    for (int pass = 0; pass < 2; ++pass)
        for (int q = 0; q < 5; ++q)
            for (int k = NG; k < NK + NG; ++k) for (int j = NG; j < NJ + NG; ++j) for (int i = NG; i < NI + NG; ++i) {
                const size_t m = ((size_t)k * PJ + j) * PI + i;
                if (pass == 0) { if (fabs(B[q][m]) > scale[q]) scale[q] = fabs(B[q][m]); continue; }
                if (memcmp(&A[q][m], &B[q][m], 8)) *identical = 0;
                double s = (q >= 1 && q <= 3) ? fmax(scale[1], fmax(scale[2], scale[3])) : scale[q];
                if (s == 0.0) s = 1.0;
                const double e = fabs(A[q][m] - B[q][m]) / s;
                if (e > worst) worst = e;
            }
    return worst;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <libeb200.so | -> <liboracle.so> [nsteps]\n", argv[0]); return 2; }
    const int nsteps = argc > 3 ? atoi(argv[3]) : 40;
    Api orc, prod;
    if (load(&orc, argv[2], "orc_")) return 2;
    double *Uo[5], *Us[5], *Uf[5], dt;
    int steps;
    if (run(&orc, 1, nsteps, Uo, &dt, &steps)) return 1;
    double mass = 0.0;
    for (size_t m = 0; m < (size_t)(NI + 2 * NG) * (NJ + 2 * NG) * (NK + 2 * NG); ++m) mass += Uo[0][m];
    printf("oracle: %d steps, dt = %.6e, sum(rho) over the padded block = %.12e\n", steps, dt, mass);
    if (!strcmp(argv[1], "-")) { printf("oracle only: ok\n"); return 0; }
    if (load(&prod, argv[1], "eb200_")) return 2;
    int same = 0;
    if (run(&prod, 1, nsteps, Us, &dt, &steps)) return 1;
    double e = compare(Us, Uo, &same);
    printf("FMA-free build vs oracle: max rel diff %.3e, bit-identical: %s\n", e, same ? "yes" : "NO");
    if (!same) return 1;
    if (run(&prod, 0, nsteps, Uf, &dt, &steps)) return 1;
    e = compare(Uf, Uo, &same);
    printf("throughput build vs oracle: max rel diff %.3e\n", e);
    if (!(e < 1.0e-10)) return 1;
    printf("C boundary: ok\n");
    return 0;
}
