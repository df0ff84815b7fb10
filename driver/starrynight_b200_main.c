/* starrynight_b200_main.c -- the B200 driver: StarryNight's main() with the Metropolis
 * hot path and the lattice-wide observables running on the GPU.
 *
 * It keeps the reference driver's contract (/root/reference/src/starrynight-main.c):
 *   - reads ./starrynight.cfg with the same libconfig keys and type rules
 *     (starrynight-config.c:98-182), through driver/sn_cfg.h;
 *   - argv[1] overrides T, argv[2] overrides CageStrain (main.c:142-151);
 *   - same flow: initial lattice + solid solution -> initial analysis -> MCEqmSteps
 *     equilibration mega-steps -> MCMegaSteps production mega-steps, each followed by
 *     the mid-point analysis (main.c:180-265);
 *   - same output files and formats: Recombination_T_%04d.log, rdf.dat,
 *     initial_lattice_potential.{xyz,cube}, initial_pot.png, equilib_pot.png and
 *     T_%04d_%d_%03d{-RDF.dat,_potential.xyz,_potential.cube,_potential.png,_MC-PNG.png,_MC-SVG.svg}.
 * What changes is who does the work: every call into montecarlo-core.c / analysis.c
 * becomes a call into libstarrynight_b200.so (include/starrynight_b200.h).  There is
 * no CPU fallback; without a usable GPU the driver stops with an error.
 *
 * Optional extra keys (absent in the stock cfg, so it runs unchanged):
 *   Seed (int)          Philox / MT seed instead of 0xDEADBEEF + T
 *   Device (int)        CUDA device ordinal
 *   Kernel (string)     "auto" | "colour" | "tiled"
 *   Temperatures (int array)   run all these temperatures at once as replicas of one handle (one batch of
 *                       kernels instead of the reference's one process per T, Makefile:49-63).  Every replica
 *                       is initialised and seeded (0xDEADBEEF + T) exactly like a separate run at that T and
 *                       writes that run's T-tagged files, bit for bit.  Overrides T and argv[1].
 *   Checkpoint (string) file rewritten after every production mega-step: lattice, sweeps done (the Philox
 *                       counter), ACCEPT / REJECT and the index of the next mega-step
 *   Restart (string)    continue from such a file: skips the initial analysis, equilibration and hysteresis
 *                       and resumes the production loop where the checkpoint left it, bit for bit
 *   GPUs (int)          Z-slab decomposition over this many GPUs of the box (devices Device, Device+1, ...):
 *                       one slab handle per GPU, boundary updates pushed GPU-to-GPU over NVLink by the sweep
 *                       kernels; the chain is bit-identical to the single-GPU one.  Z must be a multiple of
 *                       GPUs x 4 (GPUs x 32 for the tiled kernel).  The lattice-wide analysis (potential,
 *                       RDF, E-field, recombination: halos of 6-9 sites) runs on a full copy on the first GPU.
 *   Hysteresis : { amplitude = 0.1; steps = 64; cycles = 1; }   triangular Efield.x ramp after
 *                       equilibration (the loop main.c:229-238 only has commented out); prints
 *                       "T: %d Efield: x %f Polar: %f" per field point (main.c:82)
 * Terminal art (outputlattice_dumb_terminal, the ANSI density plots of recombination_calculator) is
 * not reproduced; the numbers those routines print are.
 *
 *   --init-only FILE    build the initial lattice exactly as the run would, write it as raw
 *                       float[X][Y][Z][4] to FILE and exit before touching the GPU (used by tests).
 */
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/starrynight_b200.h"
#include "sn_cfg.h"
#include "sn_lattice_init.h"

/* ---- run parameters, names as in starrynight-config.c:12-93 ------------------ */
static int X = 20, Y = 20, Z = 20, DIM = 3, T = 0;
static double Efield[3] = {0, 0, 0}, K = 1.0, CageStrain = 1.0, MCMegaMultiplier = 1.0;
static int DipoleCutOff = 3, MCMegaSteps = 400, MCEqmSteps = 10, ConstrainToX = 0;
static int DisplayDumbTerminal = 1, CalculateRecombination = 1, CalculateRadialOrderParameter = 0;
static int CalculatePotential = 0, CalculateEfield = 0, SaveDipolesXYZ = 0, SaveDipolesPNG = 0, SaveDipolesSVG = 0, SavePotentialCube = 0;
static const char *InitialLattice = "random";
static float dip_length[10], dip_prevalence[10];
static int dipolecount = 0;
static long long seed_override = -1;
static int device = 0, kernel = SN_KERNEL_AUTO, ngpus = 1;
static const char *checkpoint_path = NULL, *restart_path = NULL;
static int nT = 0, Ts[64];
static double hyst_amplitude = 0.0;
static int hyst_steps = 0, hyst_cycles = 1;

static void die(const char *what)
{
    fprintf(stderr, "starrynight-b200: %s: %s\n", what, sn_last_error());
    exit(EXIT_FAILURE);
}
#define SN(call) do { if ((call) != SN_OK) die(#call); } while (0)

static void load_config(const char *path)
{
    snc_config cfg;
    const snc_node *s;
    double tmp;
    const char *str;
    int i;
    snc_init(&cfg);
    if (!snc_read_file(&cfg, path)) {                                   /* config.c:113-121 */
        fprintf(stderr, "%s:%d - %s\n", cfg.err_file, cfg.err_line, cfg.err_text);
        exit(EXIT_FAILURE);
    }
    snc_lookup_int(&cfg, "T", &T);
    snc_lookup_int(&cfg, "X", &X); snc_lookup_int(&cfg, "Y", &Y); snc_lookup_int(&cfg, "Z", &Z);
    if (snc_lookup_float(&cfg, "Efield.x", &tmp)) Efield[0] = (float)tmp;  /* stored as float, config.c:132 */
    if (snc_lookup_float(&cfg, "Efield.y", &tmp)) Efield[1] = (float)tmp;
    if (snc_lookup_float(&cfg, "Efield.z", &tmp)) Efield[2] = (float)tmp;
    fprintf(stderr, "Efield: x %f y %f z %f\n", Efield[0], Efield[1], Efield[2]);
    snc_lookup_float(&cfg, "K", &K);
    snc_lookup_float(&cfg, "CageStrain", &CageStrain);
    fprintf(stderr, "CageStrain: %f\n", CageStrain);
    s = snc_lookup(&cfg, "Dipoles");
    dipolecount = snc_length(s);
    if (dipolecount > 10) dipolecount = 10;
    for (i = 0; i < dipolecount; i++) dip_length[i] = (float)snc_get_float_elem(s, i);
    s = snc_lookup(&cfg, "Prevalence");
    dipolecount = snc_length(s);                                         /* config.c:148: the second list decides */
    if (dipolecount > 10) dipolecount = 10;
    for (i = 0; i < dipolecount; i++) dip_prevalence[i] = (float)snc_get_float_elem(s, i);
    for (i = 0; i < dipolecount; i++) fprintf(stderr, "Dipole: %d Length: %f Prevalence: %f\n", i, dip_length[i], dip_prevalence[i]);
    snc_lookup_bool(&cfg, "ConstrainToX", &ConstrainToX);
    snc_lookup_int(&cfg, "DipoleCutOff", &DipoleCutOff);
    if (snc_lookup_string(&cfg, "InitialLattice", &str)) InitialLattice = strdup(str);
    snc_lookup_int(&cfg, "MCEqmSteps", &MCEqmSteps);
    snc_lookup_int(&cfg, "MCMegaSteps", &MCMegaSteps);
    snc_lookup_float(&cfg, "MCMoves", &MCMegaMultiplier);
    snc_lookup_bool(&cfg, "DisplayDumbTerminal", &DisplayDumbTerminal);
    snc_lookup_bool(&cfg, "CalculateRecombination", &CalculateRecombination);
    snc_lookup_bool(&cfg, "CalculateRadialOrderParameter", &CalculateRadialOrderParameter);
    snc_lookup_bool(&cfg, "CalculatePotential", &CalculatePotential);
    snc_lookup_bool(&cfg, "CalculateEfield", &CalculateEfield);
    snc_lookup_bool(&cfg, "SaveDipolesSVG", &SaveDipolesSVG);
    snc_lookup_bool(&cfg, "SaveDipolesPNG", &SaveDipolesPNG);
    snc_lookup_bool(&cfg, "SaveDipolesXYZ", &SaveDipolesXYZ);
    snc_lookup_bool(&cfg, "SavePotentialCube", &SavePotentialCube);
    /* optional B200 keys */
    if (snc_lookup_int(&cfg, "Seed", &i)) seed_override = (unsigned int)i;
    snc_lookup_int(&cfg, "Device", &device);
    snc_lookup_int(&cfg, "GPUs", &ngpus);
    if (ngpus < 1 || ngpus > 16) { fprintf(stderr, "GPUs = %d outside 1..16\n", ngpus); exit(EXIT_FAILURE); }
    if (snc_lookup_string(&cfg, "Kernel", &str))
        kernel = !strcmp(str, "colour") ? SN_KERNEL_COLOUR : !strcmp(str, "tiled") ? SN_KERNEL_TILED : SN_KERNEL_AUTO;
    s = snc_lookup(&cfg, "Temperatures");
    nT = s ? snc_length(s) : 0;
    if (nT > 64) nT = 64;
    for (i = 0; i < nT; i++) Ts[i] = (int)snc_get_int_elem(s, i);
    if (snc_lookup_string(&cfg, "Checkpoint", &str)) checkpoint_path = strdup(str);
    if (snc_lookup_string(&cfg, "Restart", &str)) restart_path = strdup(str);
    snc_lookup_float(&cfg, "Hysteresis.amplitude", &hyst_amplitude);
    snc_lookup_int(&cfg, "Hysteresis.steps", &hyst_steps);
    snc_lookup_int(&cfg, "Hysteresis.cycles", &hyst_cycles);
    fprintf(stderr, "Finished loading config file. \n");
    /* strings were strdup'ed; the tree can go */
    snc_destroy(&cfg);
}

/* ---- writers: formats of starrynight-analysis.c --------------------------------- */
static size_t site(int x, int y, int z) { return ((size_t)x * Y + y) * Z + z; }

static void write_potential_xyz(const char *fn, const double *V)          /* analysis.c:264-276 */
{
    FILE *fo = fopen(fn, "w"); int x, y, z;
    if (!fo) { perror(fn); return; }
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++) fprintf(fo, "%d %d %d %f\n", x, y, z, V[site(x, y, z)]);
    fclose(fo);
}

static void write_potential_cube(const char *fn, const double *V)         /* analysis.c:280-308 */
{
    FILE *fo = fopen(fn, "w"); int x, y, z;
    if (!fo) { perror(fn); return; }
    fprintf(fo, "Starrynight Cube file: %d %d %d\n\n", X, Y, Z);
    fprintf(fo, "1 0.0 0.0 0.0\n");
    fprintf(fo, "%d 1.0 0.0 0.0\n", X);
    fprintf(fo, "%d 0.0 1.0 0.0\n", Y);
    fprintf(fo, "%d 0.0 0.0 1.0\n", Z);
    fprintf(fo, "1 0.0 0.0 0.0 0.0 0.0\n");
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) {
        for (z = 0; z < Z; z++) { fprintf(fo, "%g ", V[site(x, y, z)]); if (z % 6 == 5) fprintf(fo, "\n"); }
        fprintf(fo, "\n");
    }
    fclose(fo);
}

static void write_potential_png(const char *fn, const double *V)          /* analysis.c:481-504 (a P2 greymap) */
{
    FILE *fo = fopen(fn, "w"); int i, k, pixel;
    if (!fo) { perror(fn); return; }
    fprintf(fo, "P2\n%d %d\n%d\n", X, Y, SHRT_MAX);
    for (i = 0; i < X; i++) {
        for (k = 0; k < Y; k++) {
            pixel = SHRT_MAX / 2 + (int)(SHRT_MAX * 0.1 * V[site(i, k, 0)]);
            if (pixel < 0) pixel = 0;
            if (pixel > SHRT_MAX) pixel = SHRT_MAX;
            fprintf(fo, "%d ", pixel);
        }
        fprintf(fo, "\n");
    }
    fclose(fo);
}

static void write_rdf(const char *fn, const double *fe, const double *afe, const long long *cnt)   /* analysis.c:582-595 */
{
    FILE *fo = fopen(fn, "a"); int i;                                     /* append, as the reference does */
    if (!fo) { perror(fn); return; }
    fprintf(fo, "# r^2 r orientational_FE_correlation[r^2] orientational_AFE_correlation[r^2] orientational_count[r^2] T\n");
    for (i = 0; i < SN_RDF_BINS; i++)
        if (cnt[i] > 0) fprintf(fo, "%d %f %f %f %lld %d\n", i, sqrt((double)i), fe[i] / (double)cnt[i], afe[i] / (double)cnt[i], cnt[i], T);
    fprintf(fo, "\n");
    fclose(fo);
}

static void write_lattice_xyz(const char *fn, const float *lat)           /* analysis.c:710-728 */
{
    FILE *fo = fopen(fn, "w"); int x, y, z; const float r = 1.6f / 2, d = 4.0f; const double ZSCALE = 5.0;
    if (!fo) { perror(fn); return; }
    fprintf(fo, "%d\n\n", X * Y * Z * 2);
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++) {
        const float *p = lat + site(x, y, z) * 4;
        fprintf(fo, "C %f %f %f\n", d * x + r * p[0], d * y + r * p[1], ZSCALE * (d * z) + r * p[2]);
        fprintf(fo, "N %f %f %f\n", d * x - r * p[0], d * y - r * p[1], ZSCALE * (d * z) - r * p[2]);
    }
    fclose(fo);
}

static void write_lattice_svg(const char *fn, const float *lat)           /* analysis.c:675-706, z = 0 slice */
{
    FILE *fo = fopen(fn, "w"); int x, y;
    if (!fo) { perror(fn); return; }
    fprintf(fo, "<svg xmlns=\"http://www.w3.org/2000/svg\" version=\"1.1\" height=\"%d\" width=\"%d\">\n", X, Y);
    fprintf(fo, " <marker id=\"triangle\" viewBox=\"0 0 10 10\" refX=\"7\" refY=\"5\" markerUnits=\"strokeWidth\" markerWidth=\"2\" markerHeight=\"2\" orient=\"auto\"><path d=\"M 0 0 L 10 5 L 0 10 z\" /></marker>\n");
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) {
        const float *p = lat + site(x, y, 0) * 4; const int g = (int)((-p[2] + 1.0) * 127.0);
        fprintf(fo, " <line x1=\"%f\" y1=\"%f\" x2=\"%f\" y2=\"%f\" style=\"stroke:rgb(%d,%d,%d);stroke-width:0.17\" marker-end=\"url(#triangle)\" />\n",
                y + 0.5 + 0.4 * p[1], x + 0.5 + 0.4 * p[0], y + 0.5 - 0.4 * p[1], x + 0.5 - 0.4 * p[0], g, g, g);
    }
    fprintf(fo, "</svg>\n");
    fclose(fo);
}

static void write_lattice_ppm_hsv(const char *fn, const float *lat)       /* analysis.c:620-671, z = 0 slice */
{
    FILE *fo = fopen(fn, "w"); int i, k;
    if (!fo) { perror(fn); return; }
    fprintf(fo, "P6\n%d %d\n255\n", X, Y);
    for (i = 0; i < X; i++) for (k = 0; k < Y; k++) {
        const float *d = lat + site(i, k, 0) * 4;
        float h = M_PI + atan2(d[1], d[0]), v = 0.5 + 0.4 * d[2], s = 0.6 - 0.6 * fabs(d[2]), r = 0, g = 0, b = 0, f, p, q, t;
        int hp = (int)floor(h / (M_PI / 3.0));
        f = h / (M_PI / 3.0) - (float)hp;
        p = v * (1.0 - s); q = v * (1.0 - f * s); t = v * (1.0 - (1.0 - f) * s);
        switch (hp) { case 0: r = v; g = t; b = p; break; case 1: r = q; g = v; b = p; break; case 2: r = p; g = v; b = t; break;
                      case 3: r = p; g = q; b = v; break; case 4: r = t; g = p; b = v; break; case 5: r = v; g = p; b = q; break; }
        if (d[0] == 0.0 && d[1] == 0.0 && d[2] == 0.0) { r = 0; g = 0; b = 0; }
        fprintf(fo, "%c%c%c", (char)(254.0 * r), (char)(254.0 * g), (char)(254.0 * b));
    }
    fclose(fo);
}

/* ---- analysis hooks, main.c:29-122 ------------------------------------------------ */
static double *Vbuf;
static float *latbuf;
static int cur = 0;                       /* replica the analysis routines look at (temperature batches) */

/* ---- the sweep engine: one handle, or one Z-slab handle per GPU ----------------------------------
 * With GPUs > 1 there is no full-lattice handle: the slabs sweep, and the analysis routines run on the slabs
 * themselves (each over its own sites, planes beyond the slab read from the neighbouring GPUs), merged here. */
static sn_handle *slab[16];
static float *slabbuf;
static double *slabV;
static int slabs_tiled = 1;

static void slab_planes(const float *lat, float *dst, int zfirst, int nplanes)   /* planes zfirst.. (periodic) of lat[X][Y][Z][4] */
{
    int x, y, k;
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (k = 0; k < nplanes; k++)
        memcpy(dst + (((size_t)x * Y + y) * nplanes + k) * 4, lat + site(x, y, ((zfirst + k) % Z + Z) % Z) * 4, 4 * sizeof(float));
}

static void slabs_create(const sn_params *base, const float *lat)
{
    int r, k, ndev = 0; const int nz = Z / ngpus;
    SN(sn_device_count(&ndev));
    if (ndev < 1) { fprintf(stderr, "no CUDA device\n"); exit(EXIT_FAILURE); }
    if (device + ngpus > ndev) fprintf(stderr, "GPUs = %d but %d device(s) visible: several slabs share a device (they split its SMs)\n", ngpus, ndev);
    slabbuf = (float *)malloc((size_t)X * Y * nz * 4 * sizeof(float));
    slabV = (double *)malloc((size_t)X * Y * nz * sizeof(double));
    if (Z % ngpus || !slabbuf || !slabV) { fprintf(stderr, "GPUs = %d does not divide Z = %d\n", ngpus, Z); exit(EXIT_FAILURE); }
    for (r = 0; r < ngpus; r++) {
        sn_params p = *base;
        p.device = (device + r) % ndev; p.z0 = r * nz; p.nz = nz;
        SN(sn_create(&p, &slab[r]));
        slab_planes(lat, slabbuf, r * nz, nz);
        SN(sn_set_lattice(slab[r], 0, slabbuf));
    }
    for (r = 0; r < ngpus; r++) {
        SN(sn_attach_peer(slab[r], 0, slab[(r + ngpus - 1) % ngpus]));
        SN(sn_attach_peer(slab[r], 1, slab[(r + 1) % ngpus]));
    }
    for (r = 0; r < ngpus; r++) SN(sn_pull_ghosts(slab[r]));               /* ghost planes, GPU to GPU */
    SN(sn_kernel_in_use(slab[0], &k));
    slabs_tiled = (k == SN_KERNEL_TILED);
    fprintf(stderr, "Z-slab decomposition: %d GPUs x %d planes, devices %d..%d\n", ngpus, nz, device, device + ngpus - 1);
}

static void engine_sweeps(sn_handle *h, long long n)                       /* MC_moves(MCMinorSteps), main.c:222,248 */
{
    int r; long long done, chunk;
    if (ngpus == 1) { SN(sn_mc_sweeps(h, n)); return; }
    /* The slabs run concurrently and wait for each other on the device.  The tiled kernel is one launch per call.  The
     * colour passes are ~200 launches per sweep and slab, each colour waiting for the neighbours' previous one: queued a
     * whole mega-step at a time from this one thread, slab 0's launches would fill the launch queue before slab 1 has
     * queued anything -- so they go in rounds of one sweep per slab. */
    chunk = slabs_tiled ? n : 1;
    for (done = 0; done < n; done += chunk)
        for (r = 0; r < ngpus; r++) SN(sn_mc_sweeps(slab[r], n - done < chunk ? n - done : chunk));
}

static void engine_sync(sn_handle *h)
{
    int r;
    if (ngpus == 1) { SN(sn_synchronize(h)); return; }
    for (r = 0; r < ngpus; r++) SN(sn_synchronize(slab[r]));
}

static void engine_set_efield(sn_handle *h, const float E[3])
{
    int r;
    for (r = 0; r < (ngpus == 1 ? (nT > 1 ? nT : 1) : 0); r++) SN(sn_set_efield(h, r, E));
    for (r = 0; r < (ngpus > 1 ? ngpus : 0); r++) SN(sn_set_efield(slab[r], 0, E));
}

static void engine_polarisation(sn_handle *h, double P[3])
{
    int r, k; double Q[3];
    if (ngpus == 1) { SN(sn_polarisation(h, cur, P)); return; }
    P[0] = P[1] = P[2] = 0.0;
    for (r = 0; r < ngpus; r++) { SN(sn_polarisation(slab[r], 0, Q)); for (k = 0; k < 3; k++) P[k] += Q[k] / ngpus; }
}

static void engine_counters(sn_handle *h, unsigned long long *acc, unsigned long long *rej, unsigned long long *vac)
{
    int r; unsigned long long a, b, c;
    if (ngpus == 1) { SN(sn_get_counters(h, cur, acc, rej, vac)); return; }
    *acc = *rej = *vac = 0;
    for (r = 0; r < ngpus; r++) { SN(sn_get_counters(slab[r], 0, &a, &b, &c)); *acc += a; *rej += b; *vac += c; }
}

/* the lattice in host order (dumps, checkpoints): one handle, or the slabs' planes put side by side */
static void engine_get_lattice(sn_handle *h, float *lat)
{
    int r, x, y; const int nz = Z / ngpus;
    if (ngpus == 1) { SN(sn_get_lattice(h, cur, lat)); return; }
    for (r = 0; r < ngpus; r++) {
        SN(sn_get_lattice(slab[r], 0, slabbuf));
        for (x = 0; x < X; x++) for (y = 0; y < Y; y++)
            memcpy(lat + site(x, y, r * nz) * 4, slabbuf + ((size_t)x * Y + y) * nz * 4, (size_t)nz * 4 * sizeof(float));
    }
}

static void slab_map_into(double *dst, int r)                              /* slabV[X][Y][nz] -> dst[X][Y][Z] at z0 = r nz */
{
    int x, y; const int nz = Z / ngpus;
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++)
        memcpy(dst + site(x, y, r * nz), slabV + ((size_t)x * Y + y) * nz, (size_t)nz * sizeof(double));
}

static void refresh_potential(sn_handle *h)
{
    int r;
    if (ngpus == 1) { SN(sn_potential_map(h, cur, Vbuf)); return; }
    for (r = 0; r < ngpus; r++) { SN(sn_potential_map(slab[r], 0, slabV)); slab_map_into(Vbuf, r); }
}

static void do_rdf(sn_handle *h, const char *fn)
{
    double fe[SN_RDF_BINS], afe[SN_RDF_BINS]; long long cnt[SN_RDF_BINS];
    if (ngpus == 1) SN(sn_rdf(h, cur, fe, afe, cnt));
    else {
        double f1[SN_RDF_BINS], a1[SN_RDF_BINS]; long long c1[SN_RDF_BINS]; int r, b;
        for (b = 0; b < SN_RDF_BINS; b++) { fe[b] = afe[b] = 0.0; cnt[b] = 0; }
        for (r = 0; r < ngpus; r++) {                                       /* sums and counts add up over the slabs */
            SN(sn_rdf(slab[r], 0, f1, a1, c1));
            for (b = 0; b < SN_RDF_BINS; b++) { fe[b] += f1[b]; afe[b] += a1[b]; cnt[b] += c1[b]; }
        }
    }
    write_rdf(fn, fe, afe, cnt);
}

static void write_efield_xyz(sn_handle *h, const char *fn, int cutoff, int half_offset)   /* analysis.c:379-389, 468-479 */
{
    FILE *fo; int x, y, z, r;
    if (ngpus == 1) SN(sn_efield_map(h, cur, cutoff, half_offset, Vbuf));
    else for (r = 0; r < ngpus; r++) { SN(sn_efield_map(slab[r], 0, cutoff, half_offset, slabV)); slab_map_into(Vbuf, r); }
    fo = fopen(fn, "w");
    if (!fo) { perror(fn); return; }
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++) fprintf(fo, "%d %d %d %f\n", x, y, z, Vbuf[site(x, y, z)]);
    fclose(fo);
}

static void do_recombination(sn_handle *h, FILE *log)                      /* analysis.c:96-228 */
{
    double r[SN_RECOMB_N];
    if (ngpus == 1) SN(sn_recombination(h, cur, r));
    else {
        double parts[16 * SN_RECOMB_PARTIAL_N]; int k;
        for (k = 0; k < ngpus; k++) SN(sn_recombination_partial(slab[k], 0, parts + k * SN_RECOMB_PARTIAL_N));
        SN(sn_recombination_finish(ngpus, parts, r));
    }
    if (log) {
        fprintf(log, "T: %d ZBe: %e ZBh: %e ZFDe: %e ZFDh: %e R_Boltz: %e ", T, r[0], r[1], r[2], r[3], r[4]);
        fprintf(log, "R_FD: %e FD-Total-electron: %e FD-Total-hole: %e\n", r[5], r[6], r[7]);
        fflush(log);
    }
    {                                                                      /* the echo below the density plots, :224-227 */
        fprintf(stderr, "Density eMAX: %f hMAX: %f\nRMAX: %e\n", r[8], r[9], r[10]);
        fprintf(stderr, "T: %d ZBe: %e ZBh: %e ZFDe: %e ZFDh: %e R_Boltz: %e \n", T, r[0], r[1], r[2], r[3], r[4]);
        fprintf(stderr, "R_FD: %e FD-Total-electron: %e FD-Total-hole: %e\n", r[5], r[6], r[7]);
    }
}

static void terminal_summary(void)
{
    /* the last line outputlattice_dumb_terminal prints (analysis.c:923-925), from the z = 0 slice of V */
    int x, y; double mean = 0, var = 0, dmax = 0;
    for (y = 0; y < Y; y++) for (x = 0; x < X; x++) { double p = Vbuf[site(x, y, 0)]; mean += p; var += p * p; if (fabs(p) > dmax) dmax = fabs(p); }
    fprintf(stderr, "T: %d DMAX: %f new_DMAX: %f variance: %f mean: %f\n", T, dmax, dmax, var / (X * Y), mean / (X * Y));
}

static void analysis_initial(sn_handle *h)                                 /* main.c:29-48 */
{
    int need_v = CalculatePotential || SavePotentialCube || DisplayDumbTerminal;
    if (CalculateEfield) write_efield_xyz(h, "initial_lattice_efield.xyz", 4, 0);          /* main.c:31-32 */
    if (CalculateEfield) write_efield_xyz(h, "initial_lattice_efieldoffset.xyz", 2, 1);
    if (need_v) refresh_potential(h);
    if (CalculatePotential) write_potential_xyz("initial_lattice_potential.xyz", Vbuf);
    if (SavePotentialCube) write_potential_cube("initial_lattice_potential.cube", Vbuf);
    if (SaveDipolesSVG || SaveDipolesPNG || SaveDipolesXYZ) engine_get_lattice(h, latbuf);
    if (SaveDipolesSVG) write_lattice_svg("initial-SVG.svg", latbuf);
    if (CalculatePotential) write_potential_png("initial_pot.png", Vbuf);
    if (SaveDipolesXYZ) write_lattice_xyz("initial_dipoles.xyz", latbuf);
    if (CalculateRadialOrderParameter) do_rdf(h, "rdf.dat");
    if (SaveDipolesPNG) write_lattice_ppm_hsv("initial.png", latbuf);
    if (DisplayDumbTerminal) terminal_summary();
    if (CalculateRecombination) do_recombination(h, stderr);                                /* main.c:47 */
}

static void analysis_midpoint(sn_handle *h, int MCstep, FILE *log)         /* main.c:51-103 */
{
    char name[160], prefix[100];
    int need_v = CalculatePotential || SavePotentialCube || DisplayDumbTerminal;
    sprintf(prefix, "T_%04d_%d_%03d", T, (int)CageStrain, MCstep);       /* main.c:58 */
    if (need_v) refresh_potential(h);
    if (DisplayDumbTerminal) terminal_summary();
    if (CalculateRecombination) do_recombination(h, log);                                   /* main.c:73 */
    sprintf(name, "%s-RDF.dat", prefix);
    if (CalculateRadialOrderParameter) do_rdf(h, name);
    sprintf(name, "%s_efield.xyz", prefix);
    if (CalculateEfield) write_efield_xyz(h, name, 4, 0);                                   /* main.c:85-86; reuses Vbuf */
    if (CalculateEfield && need_v) refresh_potential(h);
    sprintf(name, "%s_potential.xyz", prefix);
    if (CalculatePotential) write_potential_xyz(name, Vbuf);
    sprintf(name, "%s_potential.cube", prefix);
    if (SavePotentialCube) write_potential_cube(name, Vbuf);
    sprintf(name, "%s_potential.png", prefix);
    if (CalculatePotential) write_potential_png(name, Vbuf);
    if (SaveDipolesPNG || SaveDipolesSVG) engine_get_lattice(h, latbuf);
    sprintf(name, "%s_MC-PNG.png", prefix);
    if (SaveDipolesPNG) write_lattice_ppm_hsv(name, latbuf);
    sprintf(name, "%s_MC-SVG.svg", prefix);
    if (SaveDipolesSVG) write_lattice_svg(name, latbuf);
}

/* ---- checkpoint / restart ------------------------------------------------------------- */
typedef struct {
    char magic[8];                       /* "SNB200C1" */
    int X, Y, Z, T;
    unsigned long long seed, sweeps, accept, reject, vacant;
    int next_megastep, pad;
} sn_checkpoint_header;

static void engine_counters(sn_handle *h, unsigned long long *acc, unsigned long long *rej, unsigned long long *vac);

static void checkpoint_write(sn_handle *h, unsigned long long seed, int next_megastep)
{
    sn_checkpoint_header hd; char tmp[512]; FILE *f; const size_t n = (size_t)X * Y * Z * 4;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, "SNB200C1", 8);
    hd.X = X; hd.Y = Y; hd.Z = Z; hd.T = T; hd.seed = seed; hd.next_megastep = next_megastep;
    engine_get_lattice(h, latbuf);
    SN(sn_get_sweep_count(ngpus > 1 ? slab[0] : h, &hd.sweeps));
    engine_counters(h, &hd.accept, &hd.reject, &hd.vacant);
    snprintf(tmp, sizeof tmp, "%s.tmp", checkpoint_path);
    f = fopen(tmp, "wb");
    if (!f || fwrite(&hd, sizeof hd, 1, f) != 1 || fwrite(latbuf, sizeof(float), n, f) != n) { perror(tmp); exit(EXIT_FAILURE); }
    fclose(f);
    if (rename(tmp, checkpoint_path)) { perror(checkpoint_path); exit(EXIT_FAILURE); }   /* never a half-written checkpoint */
}

static void checkpoint_read(sn_checkpoint_header *hd, unsigned long long seed)
{
    FILE *f = fopen(restart_path, "rb"); const size_t n = (size_t)X * Y * Z * 4;
    if (!f || fread(hd, sizeof *hd, 1, f) != 1 || memcmp(hd->magic, "SNB200C1", 8)) { fprintf(stderr, "Restart: cannot read '%s'\n", restart_path); exit(EXIT_FAILURE); }
    if (hd->X != X || hd->Y != Y || hd->Z != Z || hd->T != T || hd->seed != seed) {
        fprintf(stderr, "Restart: '%s' holds a %dx%dx%d lattice at T=%d, seed %llX; the configuration asks for %dx%dx%d, T=%d, seed %llX\n",
                restart_path, hd->X, hd->Y, hd->Z, hd->T, hd->seed, X, Y, Z, T, seed);
        exit(EXIT_FAILURE);
    }
    if (fread(latbuf, sizeof(float), n, f) != n) { fprintf(stderr, "Restart: '%s' is truncated\n", restart_path); exit(EXIT_FAILURE); }
    fclose(f);
    fprintf(stderr, "Restart from '%s': %llu sweeps done, next mega-step %d\n", restart_path, hd->sweeps, hd->next_megastep);
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char *argv[])
{
    const char *init_only = NULL, *cfgpath = "starrynight.cfg";
    int i, npos = 0;
    char name[160];
    FILE *log, *logs[64];
    int r;
    sn_mt19937 mt;
    sn_params p;
    sn_handle *h = NULL;
    long long sweeps_per_megastep;
    unsigned long long acc = 0, rej = 0, vac = 0;
    size_t nsites;
    int histo[10], first_megastep = 0;
    unsigned long long run_seed = 0;
    sn_checkpoint_header ck;

    fprintf(stderr, "Starry Night - Monte Carlo brushstrokes (B200 build: %s).\n", sn_version());
    for (i = 1; i < argc; i++)
        if (!strcmp(argv[i], "--init-only") && i + 1 < argc) { init_only = argv[++i]; }
        else if (!strcmp(argv[i], "--config") && i + 1 < argc) { cfgpath = argv[++i]; }

    fprintf(stderr, "Loading config...\n");
    load_config(cfgpath);
    for (i = 1; i < argc; i++) {                                           /* main.c:142-151 */
        if (!strcmp(argv[i], "--init-only") || !strcmp(argv[i], "--config")) { i++; continue; }
        if (npos == 0) { sscanf(argv[i], "%d", &T); fprintf(stderr, "Command line temperature: T = %d\n", T); }
        if (npos == 1) { sscanf(argv[i], "%lf", &CageStrain); fprintf(stderr, "Command Line CageStrain: CageStrain = %lf\n", CageStrain); }
        npos++;
    }
    if (nT > 0) T = Ts[0]; else { nT = 1; Ts[0] = T; }
    if (nT > 1 && (ngpus > 1 || checkpoint_path || restart_path || seed_override >= 0)) {
        fprintf(stderr, "Temperatures cannot be combined with GPUs > 1, Checkpoint, Restart or Seed\n");
        return EXIT_FAILURE;
    }
    nsites = (size_t)X * Y * Z;
    fprintf(stderr, "Memory allocation for lattice with X=%d Y=%d Z=%d\n", X, Y, Z);
    latbuf = (float *)calloc(nsites * 4, sizeof(float));
    Vbuf = (double *)calloc(nsites, sizeof(double));
    if (!latbuf || !Vbuf) { fprintf(stderr, "out of host memory\n"); return EXIT_FAILURE; }

    {   /* main.c:172-176: seed the twister with 0xDEADBEEF + T (int arithmetic wraps as there) */
        unsigned int SEED = seed_override >= 0 ? (unsigned int)seed_override : (unsigned int)(0xDEADBEEFu + (unsigned int)T);
        sn_mt_seed(&mt, SEED);
        fprintf(stderr, "Mersenne Twister initialised... seed: %X\t", SEED);
        if (!sn_init_lattice(latbuf, X, Y, Z, DIM, InitialLattice, &mt)) {
            fprintf(stderr, "unknown InitialLattice '%s', using random (main.c:185)\n", InitialLattice);
            sn_init_lattice(latbuf, X, Y, Z, DIM, "random", &mt);
        }
        fprintf(stderr, "Lattice initialised...");
        sn_init_solid_solution(latbuf, X, Y, Z, dipolecount, dip_length, dip_prevalence, &mt, histo);
        fprintf(stderr, "\nSolid Solution: ");
        for (i = 0; i < dipolecount; i++) fprintf(stderr, "    Dipole %d: Length: %f Count: %d", i, dip_length[i], histo[i]);
        fprintf(stderr, "\nSolid solution formed...\n");
        run_seed = SEED;
        if (restart_path && !init_only) checkpoint_read(&ck, SEED);       /* replaces the freshly built lattice */
        if (init_only) {
            FILE *f = fopen(init_only, "wb");
            if (!f || fwrite(latbuf, sizeof(float), nsites * 4, f) != nsites * 4) { perror(init_only); return EXIT_FAILURE; }
            fclose(f);
            return 0;
        }
        sprintf(name, "Recombination_T_%04d.log", T);                      /* main.c:165-178 */
        log = fopen(name, restart_path ? "a" : "w");                       /* a resumed run keeps the lines of the mega-steps before the checkpoint */
        fprintf(stderr, "Log file '%s' opened. ", name);
        if (log && !restart_path) fprintf(log, "# Starrynight - simulation run on time(NULL)= %ld\n# Mersenne Twister Seed: %X\n", (long)time(NULL), SEED);
        if (log && restart_path) fprintf(log, "# restarted from '%s' at mega-step %d on time(NULL)= %ld\n", restart_path, ck.next_megastep, (long)time(NULL));

        SN(sn_default_params(&p));
        p.X = X; p.Y = Y; p.Z = Z; p.cutoff = DipoleCutOff; p.CageStrain = CageStrain; p.K = K;
        p.Efield[0] = (float)Efield[0]; p.Efield[1] = (float)Efield[1]; p.Efield[2] = (float)Efield[2];
        p.beta = 1 / ((float)T / 300.0);                                   /* main.c:215 */
        p.ConstrainToX = ConstrainToX; p.DIM = DIM; p.nreplicas = nT; p.seed = SEED; p.device = device; p.kernel = kernel;
        logs[0] = log;
    }
    if (ngpus > 1) slabs_create(&p, latbuf);                               /* one Z-slab handle per GPU, no full-lattice handle */
    else { SN(sn_create(&p, &h)); SN(sn_set_lattice(h, 0, latbuf)); }      /* lattice malloc + gen_neighbour, main.c:155-180 */
    { int nnb = 0; SN(sn_neighbour_table(ngpus > 1 ? slab[0] : h, &nnb, NULL, NULL));
      fprintf(stderr, "\nNeighbour list generated: %d neighbours found with DipoleCutOff=%d.\n", nnb, DipoleCutOff); }
    for (r = 1; r < nT; r++) {
        /* the other temperatures of the batch: initial state, seed and log of a separate run at Ts[r] (main.c:165-215) */
        const unsigned int SEED = (unsigned int)(0xDEADBEEFu + (unsigned int)Ts[r]);
        sn_mt_seed(&mt, SEED);
        if (!sn_init_lattice(latbuf, X, Y, Z, DIM, InitialLattice, &mt)) sn_init_lattice(latbuf, X, Y, Z, DIM, "random", &mt);
        sn_init_solid_solution(latbuf, X, Y, Z, dipolecount, dip_length, dip_prevalence, &mt, histo);
        SN(sn_set_lattice(h, r, latbuf));
        SN(sn_set_beta(h, r, 1 / ((float)Ts[r] / 300.0)));
        SN(sn_set_replica_seed(h, r, SEED));
        sprintf(name, "Recombination_T_%04d.log", Ts[r]);
        logs[r] = fopen(name, "w");
        if (logs[r]) fprintf(logs[r], "# Starrynight - simulation run on time(NULL)= %ld\n# Mersenne Twister Seed: %X\n", (long)time(NULL), SEED);
    }
    if (nT > 1) fprintf(stderr, "Temperature batch: %d replicas, T = %d .. %d\n", nT, Ts[0], Ts[nT - 1]);
    if (restart_path) {
        if (ngpus == 1) SN(sn_set_sweep_count(h, ck.sweeps));
        SN(sn_set_counters(ngpus > 1 ? slab[0] : h, 0, ck.accept, ck.reject, ck.vacant));
        for (r = 0; r < (ngpus > 1 ? ngpus : 0); r++) SN(sn_set_sweep_count(slab[r], ck.sweeps));
        first_megastep = ck.next_megastep;
    } else for (cur = 0; cur < nT; cur++) { T = Ts[cur]; analysis_initial(h); }
    cur = 0; T = Ts[0];

    sweeps_per_megastep = (long long)(MCMegaMultiplier + 0.5);             /* MCMinorSteps = X*Y*Z*MCMoves attempts, config.c:166 */
    if (sweeps_per_megastep < 1 && MCMegaMultiplier > 0) sweeps_per_megastep = 1;
    fprintf(stderr, "\n\tMC startup. 'Do I dare disturb the universe?'\n");
    fprintf(stderr, "'.' is %e MC moves attempted.\n", (double)sweeps_per_megastep * (double)nsites);
    fprintf(stderr, "Equilibriation MC moves... %e\n", (double)sweeps_per_megastep * (double)nsites * (double)MCEqmSteps);
    if (restart_path) goto production;
    for (i = 0; i < MCEqmSteps; i++) { fprintf(stderr, ","); engine_sweeps(h, sweeps_per_megastep); }   /* main.c:219-223 */
    engine_sync(h);
    for (cur = 0; cur < nT; cur++) {                                                        /* untagged files: the last T wins */
        if (CalculateEfield) write_efield_xyz(h, "equilib_lattice_efield.xyz", 4, 0);       /* main.c:225 */
        if (CalculatePotential) { refresh_potential(h); write_potential_png("equilib_pot.png", Vbuf); }
        if (SaveDipolesSVG) { engine_get_lattice(h, latbuf); write_lattice_svg("equilib-SVG.svg", latbuf); }
    }
    cur = 0;

    if (hyst_steps > 0 && hyst_amplitude != 0.0) {
        /* triangular ramp 0 -> +A -> -A -> 0 of Efield.x, one mega-step of sweeps per field point */
        int c, s; const int n = 4 * hyst_steps;
        for (c = 0; c < hyst_cycles; c++) for (s = 0; s < n; s++) {
            const double ph = (double)s / hyst_steps;                      /* 0..4 */
            const double e = hyst_amplitude * (ph < 1 ? ph : ph < 3 ? 2 - ph : ph - 4);
            float E[3] = {(float)e, (float)Efield[1], (float)Efield[2]}; double P[3];
            engine_set_efield(h, E);
            engine_sweeps(h, sweeps_per_megastep);
            for (cur = 0; cur < nT; cur++) { engine_polarisation(h, P); fprintf(stdout, "T: %d Efield: x %f Polar: %f\n", Ts[cur], e, P[0]); }
            cur = 0;
        }
        { float E[3] = {(float)Efield[0], (float)Efield[1], (float)Efield[2]}; engine_set_efield(h, E); }
        fflush(stdout);
    }

production:
    for (i = first_megastep; i < MCMegaSteps; i++) {                       /* main.c:244-265, the hot loop */
        double tic = now_s(), toc, tac;
        engine_sweeps(h, sweeps_per_megastep);
        engine_sync(h);
        toc = now_s();
        for (cur = 0; cur < nT; cur++) { T = Ts[cur]; analysis_midpoint(h, i, logs[cur]); }
        cur = 0; T = Ts[0];
        fflush(stdout);
        tac = now_s();
        fprintf(stderr, "MC Moves (per second): %f MHz\n", 1e-6 * (double)sweeps_per_megastep * (double)nsites * nT / (toc - tic));
        fprintf(stderr, "Output routines: %f s ; Efficiency of MC moves vs. analysis %.2f%%\n", tac - toc, 100.0 * (toc - tic) / (tac - tic));
        if (checkpoint_path) checkpoint_write(h, run_seed, i + 1);
    }
    fprintf(stderr, "\n");
    for (cur = 0; cur < nT; cur++) {
        engine_counters(h, &acc, &rej, &vac);
        if (nT > 1) fprintf(stderr, "T: %d ", Ts[cur]);
        fprintf(stderr, "Monte Carlo moves - ACCEPT: %llu REJECT: %llu ratio: %f\n", acc, rej, (float)acc / (float)(rej + acc));
    }
    fprintf(stderr, " For us, there is only the trying. The rest is not our business. ~T.S.Eliot\n\n");
    for (r = 0; r < nT; r++) if (logs[r]) fclose(logs[r]);
    if (h) SN(sn_destroy(h));
    for (i = 0; i < (ngpus > 1 ? ngpus : 0); i++) SN(sn_destroy(slab[i]));
    free(latbuf); free(Vbuf); free(slabbuf); free(slabV);
    return 0;
}
