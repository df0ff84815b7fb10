/* TEST INFRASTRUCTURE -- CPU oracle for the StarryNight Metropolis hot path.
 * See sn_oracle.h for the contract and the pinning statement.  Built by
 * oracle/Makefile with -O2 -ffp-contract=off (the reference's stock x86-64
 * build has no FMA contraction either).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "sn_oracle.h"

/* ---- MT19937, mt19937ar-cok.c:63-196 (Matsumoto & Nishimura 2002) ---------
 * The reference ships the Cokus-optimised variant; it emits the same stream as
 * the textbook generator restated here: init by the Knuth LCG 1812433253, a
 * 624-word twist with matrix 0x9908b0df, and the 11/7/15/18 tempering. */
void sno_mt_seed(sno_mt *s, unsigned long seed)
{
    int j;
    s->mt[0] = seed & 0xffffffffUL;                                         /* :66 */
    for (j = 1; j < 624; j++)
        s->mt[j] = (1812433253UL * (s->mt[j - 1] ^ (s->mt[j - 1] >> 30)) + (unsigned long)j) & 0xffffffffUL; /* :68-73 */
    s->left = 1; s->next = 624;
}

static void sno_mt_twist(sno_mt *s)
{
    int k; unsigned long y;
    for (k = 0; k < 624; k++) {                                              /* :108-128 */
        y = (s->mt[k] & 0x80000000UL) | (s->mt[(k + 1) % 624] & 0x7fffffffUL);
        s->mt[k] = s->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    s->next = 0;
}

unsigned long sno_mt_int32(sno_mt *s)
{
    unsigned long y;
    if (s->next >= 624) sno_mt_twist(s);
    y = s->mt[s->next++];
    y ^= (y >> 11);                                                          /* :138-142 */
    y ^= (y << 7) & 0x9d2c5680UL;
    y ^= (y << 15) & 0xefc60000UL;
    y ^= (y >> 18);
    return y & 0xffffffffUL;
}

double sno_mt_real1(sno_mt *s) { return (double)sno_mt_int32(s) * (1.0 / 4294967295.0); }   /* :164-179 */
double sno_mt_real2(sno_mt *s) { return (double)sno_mt_int32(s) * (1.0 / 4294967296.0); }   /* :182-196 */

/* montecarlo-core.c:18-21 */
static int sno_rand_int(sno_mt *s, int span)
{
    return (int)(sno_mt_int32(s) % (unsigned long)span);
}

/* ---- neighbour list, montecarlo-core.c:38-72 ------------------------------ */
typedef struct { int dx, dy, dz; double d; } sno_nb;   /* d holds a float value in the f32 build */

static int sno_build_nb_prec(const sno_params *p, sno_nb *nb, int f64)
{
    int dx, dy, dz, n = 0;
    int zcut = p->cutoff;
    if (p->Z == 1) zcut = 0;                                                /* :44-45 */
    for (dx = -p->cutoff; dx <= p->cutoff; dx++)
        for (dy = -p->cutoff; dy <= p->cutoff; dy++)
            for (dz = -zcut; dz <= zcut; dz++) {
                double d;
                if (dx == 0 && dy == 0 && dz == 0) continue;
                if (f64) d = sqrt((double)dx * dx + dy * dy + dz * dz);     /* :54 under float->double */
                else d = (float)sqrt((float)dx * dx + dy * dy + dz * dz);   /* :54 */
                if (d > (double)p->cutoff) continue;                        /* :56 */
                if (n >= SNO_MAXNB) return n;
                nb[n].dx = dx; nb[n].dy = dy; nb[n].dz = dz; nb[n].d = d; n++;
            }
    return n;
}

int sno_gen_neighbours(const sno_params *p, int *dxyz, double *d)
{
    sno_nb *nb = (sno_nb *)malloc(sizeof(sno_nb) * SNO_MAXNB);
    int n = sno_build_nb_prec(p, nb, 0), i;
    for (i = 0; i < n; i++) {
        dxyz[3 * i] = nb[i].dx; dxyz[3 * i + 1] = nb[i].dy; dxyz[3 * i + 2] = nb[i].dz; d[i] = nb[i].d;
    }
    free(nb);
    return n;
}

#define REAL float
#define FN(name) name##_f32
#define sno_build_nb(p, nb) sno_build_nb_prec(p, nb, 0)
#include "sn_oracle_impl.h"
#undef REAL
#undef FN

#undef sno_build_nb
#define REAL double
#define FN(name) name##_f64
#define sno_build_nb(p, nb) sno_build_nb_prec(p, nb, 1)
#include "sn_oracle_impl.h"
#undef REAL
#undef FN
