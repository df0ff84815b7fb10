/* sn_mt19937.h -- MT19937 (Matsumoto & Nishimura 2002) for the B200 driver.
 *
 * The device kernels use counter-based Philox; the HOST still needs the
 * reference's generator so that initial lattices and the solid-solution species
 * map come out exactly as /root/reference/src/starrynight-lattice.c would build
 * them from init_genrand(0xDEADBEEF + T) (main.c:172-176).  Textbook algorithm,
 * same stream as the reference's mt19937ar-cok.c.
 */
#ifndef SN_MT19937_H
#define SN_MT19937_H

typedef struct { unsigned long mt[624]; int idx; } sn_mt19937;

static void sn_mt_seed(sn_mt19937 *s, unsigned long seed)
{
    int j;
    s->mt[0] = seed & 0xffffffffUL;
    for (j = 1; j < 624; j++) s->mt[j] = (1812433253UL * (s->mt[j - 1] ^ (s->mt[j - 1] >> 30)) + (unsigned long)j) & 0xffffffffUL;
    s->idx = 624;
}

static unsigned long sn_mt_u32(sn_mt19937 *s)
{
    unsigned long y;
    if (s->idx >= 624) {
        int k;
        for (k = 0; k < 624; k++) {
            y = (s->mt[k] & 0x80000000UL) | (s->mt[(k + 1) % 624] & 0x7fffffffUL);
            s->mt[k] = s->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        s->idx = 0;
    }
    y = s->mt[s->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680UL;
    y ^= (y << 15) & 0xefc60000UL;
    y ^= (y >> 18);
    return y & 0xffffffffUL;
}

/* genrand_real1: uniform on [0,1] */
static double sn_mt_real1(sn_mt19937 *s) { return (double)sn_mt_u32(s) * (1.0 / 4294967295.0); }

#endif
