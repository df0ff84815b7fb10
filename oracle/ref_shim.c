/* TEST INFRASTRUCTURE -- not part of the product; never linked into
 * libstarrynight_b200.so.
 *
 * Builds the reference's UNMODIFIED sources, where they lie under
 * /root/reference/src, into a shared library that tests can call through
 * ctypes.  Nothing from the reference is copied: the one translation unit
 * /root/reference/src/starrynight-main.c (which itself #includes the other
 * .c files, lines 18-26) is #included below, with
 *     -Dmain=starrynight_main   so the library has no main(), and optionally
 *     -DREF_F64                 which maps `float` to `double` AFTER the system
 *                               headers are in, giving the all-FP64 build of
 *                               the same source that the 1e-12 energy bar is
 *                               asserted against.
 * The wrappers only move data between flat arrays and the reference's globals
 * (`lattice`, `X`, `Y`, `Z`, ... starrynight-config.c:12-93) and call its
 * file-static functions; all arithmetic runs in the reference's own code.
 *
 * Build: see oracle/Makefile (outputs go to oracle/_ref/, git-ignored).
 */
#include <math.h>
#include <limits.h>
#include <time.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <stdarg.h>
#include <libconfig.h>   /* the stub in oracle/stub */

/* The reference narrates every neighbour and lattice site on stderr
 * (montecarlo-core.c:63).  Route its fprintf through a switchable sink. */
static int ref_quiet = 1;
static int ref_fprintf(FILE *f, const char *fmt, ...)
{
    va_list ap; int r;
    if (ref_quiet && f == stderr) return 0;
    va_start(ap, fmt); r = vfprintf(f, fmt, ap); va_end(ap);
    return r;
}
#define fprintf ref_fprintf

#ifdef REF_F64
#define float double
#endif

#include "starrynight-main.c"   /* found via -I/root/reference/src */

#ifdef REF_F64
#undef float
#endif
#undef fprintf

#define REF_API __attribute__((visibility("default")))

static int ref_alloc_x = 0, ref_alloc_y = 0;

static void ref_free_lattice(void)
{
    int x, y;
    if (!lattice) return;
    for (x = 0; x < ref_alloc_x; x++) {
        for (y = 0; y < ref_alloc_y; y++) free(lattice[x][y]);
        free(lattice[x]);
    }
    free(lattice); lattice = NULL;
}

REF_API int ref_is_f64(void)
{
#ifdef REF_F64
    return 1;
#else
    return 0;
#endif
}

REF_API void ref_set_quiet(int q) { ref_quiet = q; }

/* Set the globals the hot path reads (montecarlo-core.c / config.c), allocate
 * the lattice the way main() does (main.c:155-161) and build the neighbour
 * list (main.c:180). */
REF_API void ref_configure(int x, int y, int z, int cutoff, double cagestrain, double k,
                           double ex, double ey, double ez, double beta_, int constrain, int dim, int temperature)
{
    int i, j;
    ref_free_lattice();
    X = x; Y = y; Z = z; DipoleCutOff = cutoff; CageStrain = cagestrain; K = k;
    Efield.x = ex; Efield.y = ey; Efield.z = ez; Efield.length = 0;
    beta = beta_; ConstrainToX = constrain; DIM = dim; T = temperature;
    ACCEPT = 0; REJECT = 0;
    lattice = (struct dipole ***)malloc(sizeof(struct dipole **) * X);
    for (i = 0; i < X; i++) {
        lattice[i] = (struct dipole **)malloc(sizeof(struct dipole *) * Y);
        for (j = 0; j < Y; j++) lattice[i][j] = (struct dipole *)calloc(Z, sizeof(struct dipole));
    }
    ref_alloc_x = X; ref_alloc_y = Y;
    neighbour = 0;
    gen_neighbour();
}

REF_API void ref_set_beta(double b) { beta = b; }
REF_API void ref_set_efield(double ex, double ey, double ez) { Efield.x = ex; Efield.y = ey; Efield.z = ez; }
REF_API void ref_set_cagestrain(double c) { CageStrain = c; }
REF_API void ref_set_K(double k) { K = k; }

/* flat [x][y][z][4] = (x,y,z,length), z fastest: the reference's own order */
REF_API void ref_set_lattice(const double *a)
{
    int x, y, z; size_t i = 0;
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++, i += 4) {
        lattice[x][y][z].x = a[i]; lattice[x][y][z].y = a[i + 1];
        lattice[x][y][z].z = a[i + 2]; lattice[x][y][z].length = a[i + 3];
    }
}

REF_API void ref_get_lattice(double *a)
{
    int x, y, z; size_t i = 0;
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++, i += 4) {
        a[i] = lattice[x][y][z].x; a[i + 1] = lattice[x][y][z].y;
        a[i + 2] = lattice[x][y][z].z; a[i + 3] = lattice[x][y][z].length;
    }
}

REF_API int ref_neighbour_count(void) { return neighbour; }
REF_API void ref_get_neighbours(int *dxyz, double *d)
{
    int i;
    for (i = 0; i < neighbour; i++) {
        dxyz[3 * i] = neighbours[i].dx; dxyz[3 * i + 1] = neighbours[i].dy; dxyz[3 * i + 2] = neighbours[i].dz;
        d[i] = neighbours[i].d;
    }
}

/* dE of rotating site (x,y,z) to (nx,ny,nz): site_energy, montecarlo-core.c:76 */
REF_API void ref_site_energy_batch(int n, const int *sites, const double *newdip, double *dE)
{
    int i;
    for (i = 0; i < n; i++) {
        struct dipole nd; int x = sites[3 * i], y = sites[3 * i + 1], z = sites[3 * i + 2];
        nd.x = newdip[3 * i]; nd.y = newdip[3 * i + 1]; nd.z = newdip[3 * i + 2];
        nd.length = lattice[x][y][z].length;      /* montecarlo-core.c:173 */
        dE[i] = site_energy(x, y, z, &nd, &lattice[x][y][z]);
    }
}

/* Interaction energy of every site through the reference's own site_energy:
 * e[i] = site_energy(x,y,z, new=lattice[x][y][z], old={0,0,0,len_i})
 * (SURVEY.md section 8a row A7).  Uses whatever Efield/K/CageStrain are set. */
REF_API void ref_site_interaction_map(double *e)
{
    int x, y, z; size_t i = 0;
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++, i++) {
        struct dipole zero; zero.x = 0; zero.y = 0; zero.z = 0; zero.length = lattice[x][y][z].length;
        e[i] = site_energy(x, y, z, &lattice[x][y][z], &zero);
    }
}

REF_API void ref_seed(unsigned long s) { init_genrand(s); }
REF_API void ref_mc_moves(int n) { MC_moves(n); }
REF_API void ref_get_counters(unsigned long long *acc, unsigned long long *rej) { *acc = ACCEPT; *rej = REJECT; }
REF_API void ref_reset_counters(void) { ACCEPT = 0; REJECT = 0; }
REF_API unsigned long ref_genrand_int32(void) { return genrand_int32(); }
REF_API double ref_genrand_real1(void) { return genrand_real1(); }
REF_API double ref_genrand_real2(void) { return genrand_real2(); }
REF_API void ref_random_sphere_point(double *p)
{ struct dipole d; random_sphere_point(&d); p[0] = d.x; p[1] = d.y; p[2] = d.z; }
REF_API void ref_random_X_point(double *p)
{ struct dipole d; random_X_point(&d); p[0] = d.x; p[1] = d.y; p[2] = d.z; }

REF_API double ref_polarisation(void) { return polarisation(); }
REF_API double ref_landau_order(void) { return landau_order(); }
REF_API double ref_dipole_potential(int x, int y, int z) { return dipole_potential(x, y, z); }
REF_API void ref_potential_map(double *v)
{
    int x, y, z; size_t i = 0;
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++, i++) v[i] = dipole_potential(x, y, z);
}
/* dipole_electricfield / dipole_electricfieldoffset (analysis.c:310-465).  The offset variant
 * prints every term to stderr (analysis.c:361-368): the caller silences stderr around it. */
REF_API double ref_dipole_electricfield(int cutoff, int x, int y, int z) { return dipole_electricfield(cutoff, x, y, z); }
REF_API void ref_efield_map(int cutoff, int half_offset, double *v)
{
    int x, y, z; size_t i = 0;
    FILE *saved = stderr;
    if (half_offset) stderr = fopen("/dev/null", "w");
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++, i++)
        v[i] = half_offset ? dipole_electricfieldoffset(cutoff, x, y, z) : dipole_electricfield(cutoff, x, y, z);
    if (half_offset) { fclose(stderr); stderr = saved; }
}
/* recombination_calculator (analysis.c:96-228): the two log lines go to `logfile`, the terminal art to /dev/null */
REF_API void ref_recombination(const char *logfile)
{
    FILE *log = fopen(logfile, "w"), *saved = stderr;
    stderr = fopen("/dev/null", "w");
    recombination_calculator(log);
    fclose(stderr); stderr = saved;
    fclose(log);
}
/* appends one block to `filename` exactly as the reference does (analysis.c:528) */
REF_API void ref_radial_order_parameter(const char *filename) { radial_order_parameter((char *)filename); }
REF_API void ref_lattice_potential_XYZ(const char *filename) { lattice_potential_XYZ((char *)filename); }
REF_API void ref_lattice_potential_cube(const char *filename) { lattice_potential_cube((char *)filename); }
REF_API void ref_outputpotential_png(const char *filename) { outputpotential_png((char *)filename); }

/* initial lattices (lattice.c:25-137); returns 0 if the name is unknown */
REF_API int ref_initialise_lattice(const char *name)
{
    if (!strcmp(name, "random")) initialise_lattice_random();
    else if (!strcmp(name, "ferroelectric")) initialise_lattice_ferroelectric();
    else if (!strcmp(name, "buckled")) initialise_lattice_buckled();
    else if (!strcmp(name, "antiferro_wall")) initialise_lattice_antiferro_wall();
    else if (!strcmp(name, "ferro_wall")) initialise_lattice_ferro_wall();
    else if (!strcmp(name, "antiferro_slip")) initialise_lattice_antiferro_slip();
    else if (!strcmp(name, "spectrum")) initialise_lattice_spectrum();
    else if (!strcmp(name, "slab_delete")) initialise_lattice_slab_delete();
    else return 0;
    return 1;
}

REF_API void ref_solid_solution(int n, const double *length, const double *prevalence)
{
    int i;
    dipolecount = n;
    for (i = 0; i < n; i++) { dipoles[i].length = length[i]; dipoles[i].prevalence = prevalence[i]; }
    solid_solution();
}

/* run the reference's whole program (cwd must hold starrynight.cfg) */
REF_API int ref_main(int argc, char **argv) { return starrynight_main(argc, argv); }
