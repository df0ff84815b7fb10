/* sn_lattice_init.h -- host-side initial lattices and solid solution for the driver.
 *
 * Same states as /root/reference/src/starrynight-lattice.c:25-171, written for a flat
 * float[X][Y][Z][4] array (the layout sn_set_lattice takes).  `random` and the species
 * map consume the MT19937 stream in the reference's order, so a given seed gives the
 * reference's own starting configuration.
 */
#ifndef SN_LATTICE_INIT_H
#define SN_LATTICE_INIT_H

#include <math.h>
#include <string.h>
#include "sn_mt19937.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* Marsaglia (1972) point on the unit sphere, circle when dim < 3 (config.c:203-227) */
static void sn_init_sphere_point(sn_mt19937 *mt, int dim, float *p)
{
    float x1, x2;
    do {
        x1 = 2.0 * sn_mt_real1(mt) - 1.0;
        x2 = 2.0 * sn_mt_real1(mt) - 1.0;
    } while (x1 * x1 + x2 * x2 > 1.0);
    if (dim < 3) {
        p[0] = (x1 * x1 - x2 * x2) / (x1 * x1 + x2 * x2);
        p[1] = 2 * x1 * x2 / (x1 * x1 + x2 * x2);
        p[2] = 0.0;
    } else {
        p[0] = 2 * x1 * sqrt(1 - x1 * x1 - x2 * x2);
        p[1] = 2 * x2 * sqrt(1 - x1 * x1 - x2 * x2);
        p[2] = 1.0 - 2.0 * (x1 * x1 + x2 * x2);
    }
}

/* returns 0 for an unknown name (the reference then keeps `random`, main.c:185) */
static int sn_init_lattice(float *lat, int X, int Y, int Z, int dim, const char *name, sn_mt19937 *mt)
{
    int x, y, z;
    const char *kinds[] = {"random", "ferroelectric", "buckled", "antiferro_wall", "ferro_wall", "antiferro_slip", "spectrum", "slab_delete"};
    int kind = -1, i;
    for (i = 0; i < 8; i++) if (!strcmp(name, kinds[i])) kind = i;
    if (kind < 0) return 0;
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++) {
        float *d = lat + (((size_t)x * Y + y) * Z + z) * 4;
        switch (kind) {
        case 0: sn_init_sphere_point(mt, dim, d); break;
        case 1: d[0] = 1.0f; d[1] = 0.0f; d[2] = 0.0f; break;
        case 2: d[0] = (float)(x % 2); d[1] = (float)(y % 2); d[2] = (float)(z % 2); break;
        case 3:                                       /* two antiferroelectric domains at right angles */
            if ((y < Y / 2) ^ (x > X / 2)) { d[0] = (float)((2. * ((z + y) % 2)) - 1.0); d[1] = 0.0f; }
            else { d[0] = 0.0f; d[1] = (float)((2. * ((x + z) % 2)) - 1.0); }
            d[2] = 0.0f; break;
        case 4: d[0] = 0.0f; d[1] = (x < X / 2) ? -1.0f : 1.0f; d[2] = 0.0f; break;
        case 5:
            d[0] = (float)((2. * ((z + y + (x < X / 2 ? 0 : 1)) % 2)) - 1.0); d[1] = 0.0f; d[2] = 0.0f; break;
        case 6: {
            float angle = 2 * M_PI * (x * X + y) / ((float)X * Y);
            d[0] = sin(angle); d[1] = cos(angle); d[2] = 0.0f; break; }
        case 7: if (x < 6) { d[0] = 0.0f; d[1] = 0.0f; d[2] = 0.0f; } break;
        }
    }
    return 1;
}

/* lattice.c:139-171: species length per site drawn from the prevalence table */
static void sn_init_solid_solution(float *lat, int X, int Y, int Z, int n, const float *length, const float *prevalence,
                                   sn_mt19937 *mt, int *histo)
{
    int x, y, z, i;
    float len[10] = {0}, prev[10] = {0};
    for (i = 0; i < n && i < 10; i++) { len[i] = length[i]; prev[i] = prevalence[i]; if (histo) histo[i] = 0; }
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++) {
        float sample = sn_mt_real1(mt);
        for (i = 0; sample > prev[i] && i < 9; sample -= prev[i], i++);
        lat[(((size_t)x * Y + y) * Z + z) * 4 + 3] = len[i];
        if (histo) histo[i]++;
    }
}

#endif
