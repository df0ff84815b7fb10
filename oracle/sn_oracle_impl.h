/* TEST INFRASTRUCTURE -- included twice by sn_oracle.c with
 *   REAL = float,  FN(x) = x##_f32   (native reference build)
 *   REAL = double, FN(x) = x##_f64   (reference built with float->double)
 * so one text states both precisions, exactly like the reference's single
 * source does under -Dfloat=double.  Arithmetic order and types follow the
 * reference statement by statement; do not "simplify" expressions here.
 */

#define IDX(p, x, y, z) ((((size_t)(x) * (p)->Y + (y)) * (p)->Z + (z)) * 4)

/* config.c:189-199 -- sum=0; sum+=ax*bx; sum+=ay*by; sum+=az*bz in REAL */
static REAL FN(dot3)(const REAL *a, const REAL *b)
{
    REAL sum = 0.0;
    sum += a[0] * b[0];
    sum += a[1] * b[1];
    sum += a[2] * b[2];
    return sum;
}

/* montecarlo-core.c:76-141 */
static double FN(site_energy)(const sno_params *p, const REAL *lat, const sno_nb *nb, int nnb,
                              int x, int y, int z, const REAL *newd, const REAL *oldd)
{
    double dE = 0.0;
    int i;
    const int X = p->X, Y = p->Y, Z = p->Z;
    for (i = 0; i < nnb; i++) {                                            /* :91 */
        int dx = nb[i].dx, dy = nb[i].dy, dz = nb[i].dz;                   /* :94 */
        REAL d = (REAL)nb[i].d;                                            /* :95 */
        const REAL *test = lat + IDX(p, (X + x + dx) % X, (Y + y + dy) % Y, (Z + z + dz) % Z); /* :97 */
        REAL n[3];
        n[0] = (REAL)dx / d; n[1] = (REAL)dy / d; n[2] = (REAL)dz / d;     /* :99 */
        dE += (oldd[3] * test[3]) *                                        /* :102-106, all REAL */
              ((FN(dot3)(newd, test) - 3 * FN(dot3)(n, newd) * FN(dot3)(n, test)) -
               (FN(dot3)(oldd, test) - 3 * FN(dot3)(n, oldd) * FN(dot3)(n, test))) / (d * d * d);
        if ((dx * dx + dy * dy + dz * dz) == 1)                            /* :113-115, double */
            dE += -p->CageStrain * FN(dot3)(newd, test) + p->CageStrain * FN(dot3)(oldd, test);
    }
    {   /* :120-121 -- REAL subtraction, then added to the double */
        REAL ef[3]; ef[0] = (REAL)p->Efield[0]; ef[1] = (REAL)p->Efield[1]; ef[2] = (REAL)p->Efield[2];
        dE += +FN(dot3)(newd, ef) - FN(dot3)(oldd, ef);
    }
    if (p->K > 0.0) {                                                       /* :124-134 */
        REAL n[3];
        n[0] = 1.0; n[1] = 0.0; n[2] = 0.0;
        dE += -p->K * fabs(FN(dot3)(newd, n)) + p->K * fabs(FN(dot3)(oldd, n));
        n[0] = 0.0; n[1] = 1.0; n[2] = 0.0;
        dE += -p->K * fabs(FN(dot3)(newd, n)) + p->K * fabs(FN(dot3)(oldd, n));
    }
    return dE;
}

void FN(sno_site_energy_batch)(const sno_params *p, const REAL *lat, int n, const int *sites,
                               const REAL *newdip, double *dE)
{
    sno_nb *nb = (sno_nb *)malloc(sizeof(sno_nb) * SNO_MAXNB);
    int nnb = sno_build_nb(p, nb), i;
    for (i = 0; i < n; i++) {
        const REAL *old = lat + IDX(p, sites[3 * i], sites[3 * i + 1], sites[3 * i + 2]);
        REAL nd[4];
        nd[0] = newdip[3 * i]; nd[1] = newdip[3 * i + 1]; nd[2] = newdip[3 * i + 2];
        nd[3] = old[3];                                                     /* :173 */
        dE[i] = FN(site_energy)(p, lat, nb, nnb, sites[3 * i], sites[3 * i + 1], sites[3 * i + 2], nd, old);
    }
    free(nb);
}

/* e_i = site_energy(new = p_i, old = {0,0,0,len_i}) */
void FN(sno_site_interaction_map)(const sno_params *p, const REAL *lat, double *e)
{
    sno_nb *nb = (sno_nb *)malloc(sizeof(sno_nb) * SNO_MAXNB);
    int nnb = sno_build_nb(p, nb), x, y, z; size_t i = 0;
    for (x = 0; x < p->X; x++) for (y = 0; y < p->Y; y++) for (z = 0; z < p->Z; z++, i++) {
        const REAL *me = lat + IDX(p, x, y, z);
        REAL zero[4]; zero[0] = 0; zero[1] = 0; zero[2] = 0; zero[3] = me[3];
        e[i] = FN(site_energy)(p, lat, nb, nnb, x, y, z, me, zero);
    }
    free(nb);
}

/* Total lattice energy.  The reference has no working routine (main.c:63), so H
 * is DEFINED as the function whose single-site difference is site_energy:
 *   H = 1/2 sum_i (dd_i + cage_i) + sum_i p_i.E - K sum_i (|p_ix| + |p_iy|) [K>0]
 * out = { E_dd, E_cage, E_field, E_K }, each obtained by running the
 * reference-order site_energy with the other couplings switched off. */
void FN(sno_total_energy)(const sno_params *p, const REAL *lat, double out[4])
{
    size_t n = (size_t)p->X * p->Y * p->Z, i;
    double *e = (double *)malloc(sizeof(double) * n);
    sno_params q = *p; double s;
    q.CageStrain = 0; q.K = 0; q.Efield[0] = q.Efield[1] = q.Efield[2] = 0;
    FN(sno_site_interaction_map)(&q, lat, e); for (s = 0, i = 0; i < n; i++) s += e[i]; out[0] = 0.5 * s;
    q.CageStrain = p->CageStrain; q.cutoff = 1;        /* cage only lives on |r|=1 */
    {   /* subtract the |r|=1 dipole part: run cutoff 1 with and without cage */
        double s0, s1; sno_params q0 = q; q0.CageStrain = 0;
        FN(sno_site_interaction_map)(&q0, lat, e); for (s0 = 0, i = 0; i < n; i++) s0 += e[i];
        FN(sno_site_interaction_map)(&q, lat, e);  for (s1 = 0, i = 0; i < n; i++) s1 += e[i];
        out[1] = 0.5 * (s1 - s0);
    }
    q = *p; q.cutoff = 0; q.CageStrain = 0; q.K = 0;   /* cutoff 0 => empty neighbour list */
    FN(sno_site_interaction_map)(&q, lat, e); for (s = 0, i = 0; i < n; i++) s += e[i]; out[2] = s;
    q = *p; q.cutoff = 0; q.CageStrain = 0; q.Efield[0] = q.Efield[1] = q.Efield[2] = 0;
    FN(sno_site_interaction_map)(&q, lat, e); for (s = 0, i = 0; i < n; i++) s += e[i]; out[3] = s;
    free(e);
}

/* config.c:203-227, Marsaglia 1972 */
void FN(sno_random_sphere_point)(const sno_params *p, sno_mt *s, REAL out[3])
{
    REAL x1, x2;
    do {
        x1 = 2.0 * sno_mt_real1(s) - 1.0;
        x2 = 2.0 * sno_mt_real1(s) - 1.0;
    } while (x1 * x1 + x2 * x2 > 1.0);
    if (p->DIM < 3) {
        out[0] = (x1 * x1 - x2 * x2) / (x1 * x1 + x2 * x2);
        out[1] = 2 * x1 * x2 / (x1 * x1 + x2 * x2);
        out[2] = 0.0;
    } else {
        out[0] = 2 * x1 * sqrt(1 - x1 * x1 - x2 * x2);
        out[1] = 2 * x2 * sqrt(1 - x1 * x1 - x2 * x2);
        out[2] = 1.0 - 2.0 * (x1 * x1 + x2 * x2);
    }
}

/* config.c:230-263 */
void FN(sno_random_X_point)(sno_mt *s, REAL out[3])
{
    int i = sno_rand_int(s, 6), x = 0, y = 0, z = 0;
    switch (i) { case 0: x = 1; break; case 1: x = -1; break; case 2: y = 1; break;
                 case 3: y = -1; break; case 4: z = 1; break; case 5: z = -1; break; }
    out[0] = (REAL)x; out[1] = (REAL)y; out[2] = (REAL)z;
}

/* montecarlo-core.c:143-191 */
void FN(sno_mc_moves)(const sno_params *p, REAL *lat, sno_mt *s, long long moves,
                      unsigned long long *accept, unsigned long long *reject)
{
    sno_nb *nb = (sno_nb *)malloc(sizeof(sno_nb) * SNO_MAXNB);
    int nnb = sno_build_nb(p, nb);
    long long m;
    for (m = 0; m < moves; m++) {
        int x = sno_rand_int(s, p->X), y = sno_rand_int(s, p->Y), z = sno_rand_int(s, p->Z); /* :159-161 */
        REAL *old = lat + IDX(p, x, y, z), nd[4];
        REAL dE;                                                            /* :154 `float dE` */
        if (old[3] == 0.0) continue;                                        /* :163 */
        if (p->ConstrainToX) FN(sno_random_X_point)(s, nd);                 /* :168-171 */
        else FN(sno_random_sphere_point)(p, s, nd);
        nd[3] = old[3];                                                     /* :173 */
        dE = FN(site_energy)(p, lat, nb, nnb, x, y, z, nd, old);            /* :177 narrowing */
        if (dE < 0.0 || exp(-dE * p->beta) > sno_mt_real2(s)) {             /* :179 */
            old[0] = nd[0]; old[1] = nd[1]; old[2] = nd[2];
            (*accept)++;
        } else (*reject)++;
    }
    free(nb);
}

/* lattice.c:25-137; `random` consumes the MT stream like the reference */
int FN(sno_initialise_lattice)(const sno_params *p, REAL *lat, sno_mt *s, const char *name)
{
    int x, y, z; const int X = p->X, Y = p->Y, Z = p->Z;
    int kind = !strcmp(name, "random") ? 0 : !strcmp(name, "ferroelectric") ? 1 : !strcmp(name, "buckled") ? 2 :
               !strcmp(name, "antiferro_wall") ? 3 : !strcmp(name, "ferro_wall") ? 4 :
               !strcmp(name, "antiferro_slip") ? 5 : !strcmp(name, "spectrum") ? 6 :
               !strcmp(name, "slab_delete") ? 7 : -1;
    if (kind < 0) return 0;
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++) {
        REAL *d = lat + IDX(p, x, y, z);
        switch (kind) {
        case 0: FN(sno_random_sphere_point)(p, s, d); break;                /* :25-37 */
        case 1: d[0] = 1.0; d[1] = 0.0; d[2] = 0.0; break;                  /* :39-47 */
        case 2: d[0] = x % 2; d[1] = y % 2; d[2] = z % 2; break;            /* :49-57 */
        case 3:                                                             /* :59-74 */
            if ((y < Y / 2) ^ (x > X / 2)) { d[0] = (2. * ((z + y) % 2)) - 1.0; d[1] = 0.0; }
            else { d[0] = 0.0; d[1] = (2. * ((x + z) % 2)) - 1.0; }
            d[2] = 0.0; break;
        case 4:                                                             /* :76-88 */
            d[0] = 0.0; d[1] = (x < X / 2) ? -1.0 : 1.0; d[2] = 0.0; break;
        case 5:                                                             /* :90-105 */
            if (x < X / 2) { d[0] = (2. * ((z + y) % 2)) - 1.0; d[1] = 0.0; }
            else { d[0] = (2. * ((z + y + 1) % 2)) - 1.0; d[1] = 0.0; }
            d[2] = 0.0; break;
        case 6: {                                                           /* :107-124 */
            REAL angle = 2 * M_PI * (x * X + y) / ((REAL)X * Y);
            d[0] = sin(angle); d[1] = cos(angle); d[2] = 0.0; break; }
        case 7:                                                             /* :126-137, only x<6 touched */
            if (x < 6) { d[0] = 0.0; d[1] = 0.0; d[2] = 0.0; }
            break;
        }
    }
    return 1;
}

/* lattice.c:139-171 */
void FN(sno_solid_solution)(const sno_params *p, REAL *lat, sno_mt *s, int n,
                            const double *length, const double *prevalence, int *histo)
{
    int x, y, z, i;
    REAL len[10] = {0}, prev[10] = {0};   /* struct mixture holds floats (config.c:39-43); globals are zeroed */
    for (i = 0; i < n && i < 10; i++) { len[i] = (REAL)length[i]; prev[i] = (REAL)prevalence[i]; if (histo) histo[i] = 0; }
    for (x = 0; x < p->X; x++) for (y = 0; y < p->Y; y++) for (z = 0; z < p->Z; z++) {
        REAL sample = sno_mt_real1(s);                                      /* :154 */
        for (i = 0; sample > prev[i] && i < 9; sample -= prev[i], i++);     /* :159 (the reference runs off the table if prevalences sum < 1) */
        lat[IDX(p, x, y, z) + 3] = len[i];                                  /* :162 */
        if (histo) histo[i]++;
    }
}

/* analysis.c:48-62 */
double FN(sno_polarisation)(const sno_params *p, const REAL *lat)
{
    double P = 0.0; int x, y, z; REAL n[3];
    n[0] = 1.0; n[1] = 0.0; n[2] = 0.0;
    for (x = 0; x < p->X; x++) for (y = 0; y < p->Y; y++) for (z = 0; z < p->Z; z++)
        P += FN(dot3)(lat + IDX(p, x, y, z), n);
    return P / (double)(p->X * p->Y * p->Z);
}

/* analysis.c:506-526 -- REAL accumulators, and `/ (double)N * (double)N` as written */
double FN(sno_landau_order)(const sno_params *p, const REAL *lat)
{
    REAL o[3]; int x, y, z; double landau;
    o[0] = 0.0; o[1] = 0.0; o[2] = 0.0;
    for (x = 0; x < p->X; x++) for (y = 0; y < p->Y; y++) for (z = 0; z < p->Z; z++) {
        const REAL *d = lat + IDX(p, x, y, z);
        o[0] += d[0]; o[1] += d[1]; o[2] += d[2];
    }
    landau = FN(dot3)(o, o) / (double)(p->X * p->Y * p->Z) * (double)(p->X * p->Y * p->Z);
    return landau;
}

/* analysis.c:65-94 */
double FN(sno_dipole_potential)(const sno_params *p, const REAL *lat, int x, int y, int z)
{
    int dx, dy, dz; const int MAX = 6; double pot = 0.0; REAL d, r[3];
    const int X = p->X, Y = p->Y, Z = p->Z;
    for (dx = -MAX; dx <= MAX; dx++) for (dy = -MAX; dy <= MAX; dy++) for (dz = -MAX; dz <= MAX; dz++) {
        const REAL *t;
        if (dx == 0 && dy == 0 && dz == 0) continue;
        r[0] = (REAL)dx; r[1] = (REAL)dy; r[2] = (REAL)dz;
        d = sqrt((REAL)r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);            /* :84 */
        if (d > (REAL)MAX) continue;
        t = lat + IDX(p, (X + x + dx) % X, (Y + y + dy) % Y, (Z + z + dz) % Z);
        pot += t[3] * FN(dot3)(t, r) / (d * d * d);                         /* :90-91 */
    }
    return pot;
}

void FN(sno_potential_map)(const sno_params *p, const REAL *lat, double *v)
{
    int x, y, z; size_t i = 0;
    for (x = 0; x < p->X; x++) for (y = 0; y < p->Y; y++) for (z = 0; z < p->Z; z++, i++)
        v[i] = FN(sno_dipole_potential)(p, lat, x, y, z);
}

/* analysis.c:528-598.  fe/afe/count have SNO_RDF_BINS (=81) entries and hold the
 * accumulated sums and counts BEFORE the division at :587-588; bin r^2=81 is left
 * out because the reference neither zeroes nor prints it (:549, :583). */
void FN(sno_rdf)(const sno_params *p, const REAL *lat, REAL *fe, REAL *afe, int *count)
{
    const int CUTOFF = 9; int x, y, z, dx, dy, dz, i;
    const int X = p->X, Y = p->Y, Z = p->Z;
    for (i = 0; i < SNO_RDF_BINS; i++) { fe[i] = 0.0; afe[i] = 0.0; count[i] = 0; }
    for (x = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++)
        for (dx = -CUTOFF; dx <= CUTOFF; dx++) for (dy = -CUTOFF; dy <= CUTOFF; dy++) for (dz = -CUTOFF; dz <= CUTOFF; dz++) {
            int d2 = dx * dx + dy * dy + dz * dz; REAL FE, AFE, d, n[3];
            const REAL *a, *b;
            if (d2 > CUTOFF * CUTOFF) continue;
            a = lat + IDX(p, x, y, z);
            b = lat + IDX(p, (x + dx + X) % X, (y + dy + Y) % Y, (z + dz + Z) % Z);
            FE = FN(dot3)(a, b);
            d = sqrt((REAL)dx * dx + dy * dy + dz * dz);
            if (d == 0) d = 1;
            n[0] = (REAL)dx / d; n[1] = (REAL)dy / d; n[2] = (REAL)dz / d;
            AFE = FE - 3 * FN(dot3)(n, a) * FN(dot3)(n, b);
            if (d2 < SNO_RDF_BINS) { fe[d2] += FE; afe[d2] += AFE; count[d2]++; }
        }
}

/* analysis.c:393-465 (dipole_electricfield: integer offsets, self term excluded, -p_i/3 added at the end)
 * and :310-376 (dipole_electricfieldoffset: the field half a lattice step off the sites, offsets
 * dx + 0.5 for dx in [-CUTOFF-1, CUTOFF-1], nothing excluded).  Species lengths are NOT applied there.
 * Statement order and types as written: r, n, the contribution and the running field are `struct dipole`
 * members (REAL), `radial` is a double, `3*n.x*radial - p.x` is evaluated in double and narrowed. */
double FN(sno_dipole_electricfield)(const sno_params *p, const REAL *lat, int CUTOFF, int half, int x, int y, int z)
{
    int dx, dy, dz; const int X = p->X, Y = p->Y, Z = p->Z;
    REAL E[3], c[3], r[3], n[3], d; double radial;
    const int lo = half ? -CUTOFF - 1 : -CUTOFF, hi = half ? CUTOFF - 1 : CUTOFF;
    E[0] = 0.0; E[1] = 0.0; E[2] = 0.0;
    for (dx = lo; dx <= hi; dx++) for (dy = lo; dy <= hi; dy++) for (dz = lo; dz <= hi; dz++) {
        const REAL *t;
        if (!half && dx == 0 && dy == 0 && dz == 0) continue;                /* :411 */
        if (half) { r[0] = (REAL)(dx) + 0.5; r[1] = (REAL)(dy) + 0.5; r[2] = (REAL)(dz) + 0.5; }   /* :330 */
        else { r[0] = (REAL)(dx); r[1] = (REAL)(dy); r[2] = (REAL)(dz); }    /* :414 */
        d = sqrt((REAL)r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);             /* :416 */
        if (d > (REAL)CUTOFF) continue;                                       /* :418 */
        n[0] = r[0] / d; n[1] = r[1] / d; n[2] = r[2] / d;                    /* :426 */
        t = lat + IDX(p, (X + x + dx) % X, (Y + y + dy) % Y, (Z + z + dz) % Z);
        radial = FN(dot3)(n, t);                                              /* :429 */
        c[0] = 3 * n[0] * radial - t[0];                                      /* :432-434 */
        c[1] = 3 * n[1] * radial - t[1];
        c[2] = 3 * n[2] * radial - t[2];
        c[0] /= d * d * d; c[1] /= d * d * d; c[2] /= d * d * d;              /* :437-439 */
        E[0] += c[0]; E[1] += c[1]; E[2] += c[2];                             /* :442-444 */
    }
    if (!half) {                                                              /* :457-459 */
        const REAL *s = lat + IDX(p, x, y, z);
        E[0] -= 1 / 3.0 * s[0]; E[1] -= 1 / 3.0 * s[1]; E[2] -= 1 / 3.0 * s[2];
    }
    return sqrt(FN(dot3)(E, E));                                              /* :464 */
}

void FN(sno_efield_map)(const sno_params *p, const REAL *lat, int cutoff, int half, double *v)
{
    int x, y, z; size_t i = 0;
    for (x = 0; x < p->X; x++) for (y = 0; y < p->Y; y++) for (z = 0; z < p->Z; z++, i++)
        v[i] = FN(sno_dipole_electricfield)(p, lat, cutoff, half, x, y, z);
}

/* analysis.c:96-170, the physics of recombination_calculator: Boltzmann and Fermi-Dirac partition sums
 * of the screened dipole potential, then the e/h densities and their overlap.  out = ZBe ZBh ZFDe ZFDh
 * R_Boltz R_FD electron_total hole_total eMAX hMAX RMAX (the maxima are over the z = 0 plane, :157-159).
 * R_Boltz is (X*Y*Z)*(X*Y*Z)/(ZBe*ZBh) with the product taken in int there (:131; wraps beyond 215^2
 * sites); here it is taken in double. */
void FN(sno_recombination)(const sno_params *p, const REAL *lat, double out[SNO_RECOMB_N])
{
    int x, y, z; const int X = p->X, Y = p->Y, Z = p->Z; const double N = (double)X * Y * Z;
    double BETA = 1 / (0.025), potentialeV = 0.165, pot;                      /* :104-106 */
    double ZBe = 0.0, ZBh = 0.0, ZFDe = 0.0, ZFDh = 0.0, et = 0.0, ht = 0.0, rt = 0.0, eMAX = 0.0, hMAX = 0.0, RMAX = 0.0;
    double *V = (double *)malloc(sizeof(double) * (size_t)N);
    size_t i = 0;
    potentialeV /= 5;
    FN(sno_potential_map)(p, lat, V);
    for (i = 0; i < (size_t)N; i++) {                                         /* :112-127 */
        pot = potentialeV * V[i];
        ZBe += exp(-pot * BETA); ZBh += exp(pot * BETA);
        ZFDe += 1.0 / (exp(pot * BETA) + 1.0); ZFDh += 1.0 / (exp(-pot * BETA) + 1.0);
    }
    for (x = 0, i = 0; x < X; x++) for (y = 0; y < Y; y++) for (z = 0; z < Z; z++, i++) {   /* :140-166 */
        double e, h, e0, h0, pot0;
        pot = potentialeV * V[i];
        e = 1.0 / (exp(pot * BETA) + 1.0) / ZFDe; h = 1.0 / (exp(-pot * BETA) + 1.0) / ZFDh;
        pot0 = potentialeV * V[i - (size_t)z];
        e0 = 1.0 / (exp(pot0 * BETA) + 1.0) / ZFDe; h0 = 1.0 / (exp(-pot0 * BETA) + 1.0) / ZFDh;
        if (e0 > eMAX) eMAX = e0;
        if (h0 > hMAX) hMAX = h0;
        if (e0 * h0 > RMAX) RMAX = e0 * h0;
        et += e; ht += h; rt += e * h;
    }
    out[0] = ZBe; out[1] = ZBh; out[2] = ZFDe; out[3] = ZFDh; out[4] = N * N / (ZBe * ZBh);
    out[5] = N * rt; out[6] = et; out[7] = ht; out[8] = eMAX; out[9] = hMAX; out[10] = RMAX;
    free(V);
}

#undef IDX
