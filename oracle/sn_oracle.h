/* TEST INFRASTRUCTURE -- CPU oracle for the StarryNight Metropolis hot path.
 *
 * A plain-C restatement of the reference algorithm, each function citing the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (libstarrynight_b200.so) never links or calls it.
 *
 * Pinning: validated bit-for-bit against the reference's own unmodified
 * sources built by oracle/Makefile into oracle/_ref/ (tests/test_oracle_vs_ref.py)
 * and against golden vectors generated from that build (tests/golden/).  The
 * reference's own test-suite asserts nothing for this path,
 * so those two are the pin.
 *
 * Every routine exists twice: *_f32 follows the native build (float terms,
 * double accumulation -- montecarlo-core.c:99-106) and *_f64 follows the same
 * source compiled with float->double (the build the 1e-12 FP64 bar refers to).
 * Lattices are flat arrays [x][y][z][4] = (x, y, z, length), z fastest, which is
 * the memory order of the reference's lattice[x][y][z] (config.c:32-36).
 */
#ifndef SN_ORACLE_H
#define SN_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int X, Y, Z;          /* config.c:12-14 */
    int cutoff;           /* DipoleCutOff, config.c:70 */
    double CageStrain;    /* config.c:68 */
    double K;             /* config.c:66 */
    double Efield[3];     /* config.c:64 (stored as float in the f32 build) */
    double beta;          /* config.c:62; main.c:215 */
    int ConstrainToX;     /* config.c:85 */
    int DIM;              /* config.c:16 */
    int T;                /* config.c:24, only echoed in output files */
} sno_params;

typedef struct { unsigned long mt[624]; int left; int next; } sno_mt;

#define SNO_MAXNB 10000   /* montecarlo-core.c:29 */
#define SNO_RECOMB_N 11   /* ZBe ZBh ZFDe ZFDh R_Boltz R_FD e_total h_total eMAX hMAX RMAX (analysis.c:96-228) */
#define SNO_RDF_BINS 81   /* analysis.c:540-550: bins 0..80 are zeroed and printed */

/* MT19937 (mt19937ar-cok.c:63-196): published Matsumoto-Nishimura algorithm */
void sno_mt_seed(sno_mt *s, unsigned long seed);
unsigned long sno_mt_int32(sno_mt *s);
double sno_mt_real1(sno_mt *s);
double sno_mt_real2(sno_mt *s);

/* montecarlo-core.c:38-72; returns the neighbour count */
int sno_gen_neighbours(const sno_params *p, int *dxyz, double *d);

#define SNO_DECL(SUF, REAL)                                                                        \
    void sno_site_energy_batch_##SUF(const sno_params *p, const REAL *lat, int n, const int *sites, \
                                     const REAL *newdip, double *dE);                               \
    void sno_site_interaction_map_##SUF(const sno_params *p, const REAL *lat, double *e);           \
    void sno_total_energy_##SUF(const sno_params *p, const REAL *lat, double out[4]);               \
    void sno_random_sphere_point_##SUF(const sno_params *p, sno_mt *s, REAL out[3]);                \
    void sno_random_X_point_##SUF(sno_mt *s, REAL out[3]);                                          \
    void sno_mc_moves_##SUF(const sno_params *p, REAL *lat, sno_mt *s, long long moves,             \
                            unsigned long long *accept, unsigned long long *reject);                \
    int sno_initialise_lattice_##SUF(const sno_params *p, REAL *lat, sno_mt *s, const char *name);  \
    void sno_solid_solution_##SUF(const sno_params *p, REAL *lat, sno_mt *s, int n,                 \
                                  const double *length, const double *prevalence, int *histo);     \
    double sno_polarisation_##SUF(const sno_params *p, const REAL *lat);                            \
    double sno_landau_order_##SUF(const sno_params *p, const REAL *lat);                            \
    double sno_dipole_potential_##SUF(const sno_params *p, const REAL *lat, int x, int y, int z);   \
    void sno_potential_map_##SUF(const sno_params *p, const REAL *lat, double *v);                  \
    void sno_rdf_##SUF(const sno_params *p, const REAL *lat, REAL *fe, REAL *afe, int *count);      \
    double sno_dipole_electricfield_##SUF(const sno_params *p, const REAL *lat, int cutoff,         \
                                          int half_offset, int x, int y, int z);                    \
    void sno_efield_map_##SUF(const sno_params *p, const REAL *lat, int cutoff, int half_offset,    \
                              double *v);                                                           \
    void sno_recombination_##SUF(const sno_params *p, const REAL *lat, double out[SNO_RECOMB_N]);

SNO_DECL(f32, float)
SNO_DECL(f64, double)

#ifdef __cplusplus
}
#endif
#endif
