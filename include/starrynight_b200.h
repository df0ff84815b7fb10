/* starrynight_b200.h -- C ABI of libstarrynight_b200.so
 *
 * B200-native (sm_100a) replacement for StarryNight's Metropolis Monte Carlo
 * hot path and the lattice-wide observables that read the same state.
 *
 * The reference has no FFI: the path is reached by direct calls to file-static
 * functions over globals inside one translation unit.
 * Each entry point below names the reference call site / global it replaces.
 * A driver written like /root/reference/src/starrynight-main.c calls
 *     sn_create            where main() mallocs `lattice` and calls gen_neighbour()   (main.c:155-161,180)
 *     sn_set_lattice       after initialise_lattice()/solid_solution()                (main.c:203-205)
 *     sn_set_beta          where main() sets beta = 1/((float)T/300.0)                (main.c:215,239)
 *     sn_mc_sweeps         where main() calls MC_moves(MCMinorSteps)                  (main.c:222,248)
 *     sn_get_lattice / observables   where analysis_*() read `lattice`               (main.c:29-103)
 *     sn_get_counters      where main() prints ACCEPT / REJECT                        (main.c:273)
 *
 * Conventions: every function returns 0 on success and a non-zero sn_status on
 * failure; sn_last_error() returns a message for the calling thread's last
 * failure.  The caller owns all host buffers; the handle owns all device
 * memory.  One host thread per handle.  Calls are synchronous with respect to
 * the host unless stated otherwise.  There is no CPU fallback: if no CUDA
 * device is usable sn_create fails.
 *
 * Host lattice layout (everywhere): float[X][Y][nz][4] = (x, y, z, length),
 * z fastest -- the memory order of the reference's `struct dipole
 * lattice[x][y][z]` (config.c:32-36; main.c:155-161).  `nz` is the handle's own
 * Z-slab (nz == Z unless the lattice is slab-decomposed across GPUs).
 */
#ifndef STARRYNIGHT_B200_H
#define STARRYNIGHT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SN_API __attribute__((visibility("default")))
#else
#define SN_API
#endif

typedef struct sn_handle sn_handle;

typedef enum {
    SN_OK = 0,
    SN_ERR_INVALID = 1,      /* bad argument */
    SN_ERR_CUDA = 2,         /* CUDA runtime / driver error */
    SN_ERR_UNSUPPORTED = 3,  /* configuration the library cannot run */
    SN_ERR_NOMEM = 4
} sn_status;

/* arithmetic of the energy-audit entry points */
typedef enum {
    SN_PREC_F32 = 0,      /* FP32 local-field arithmetic, the colour-pass / resident sweep kernels' own (1e-5 bar); the tiled
                             kernel's register arithmetic is audited attempt by attempt with sn_mc_sweep_audit */
    SN_PREC_F64 = 1,      /* all-FP64, reference statement order (1e-12 bar vs the float->double reference build) */
    SN_PREC_REPLICA = 2   /* float terms + double accumulation exactly as montecarlo-core.c:99-121 (bit-equal to the native reference) */
} sn_precision;

typedef enum {
    SN_KERNEL_AUTO = 0,    /* tiled kernel when the lattice allows it, the resident kernel when the lattice fits in one
                              CTA's shared memory, else colour passes */
    SN_KERNEL_COLOUR = 1,  /* one launch per colour sublattice, neighbours read from global memory */
    SN_KERNEL_TILED = 2,   /* TMA-staged shared-memory tiles (needs cutoff 2 or 3, X, Y, nz >= 20, nz a multiple of 4; Z-slabs: multiples of 32 planes) */
    SN_KERNEL_RESIDENT = 4, /* lattice resident in shared memory, one CTA per replica, all sweeps of a call in one launch
                              (needs X*Y*Z*16 B <= 227 KB, no Z-slabs); bit-identical to SN_KERNEL_COLOUR */
    SN_KERNEL_TILED_PHASED = 3  /* the same kernel, one launch per tile-parity phase instead of one dataflow
                                   launch per call: validation only, bit-identical results */
} sn_kernel;

typedef struct {
    int X, Y, Z;              /* global lattice, config.c:12-14 / cfg keys X Y Z */
    int cutoff;               /* DipoleCutOff, config.c:70 */
    double CageStrain;        /* config.c:68 */
    double K;                 /* config.c:66 */
    float Efield[3];          /* config.c:64, initial value for every replica */
    double beta;              /* config.c:62, initial value for every replica */
    int ConstrainToX;         /* config.c:85 */
    int DIM;                  /* config.c:16 (2 => proposals on the XY circle) */
    int nreplicas;            /* independent lattices advanced by one call (T / field sweeps, seeds) */
    unsigned long long seed;  /* Philox key; replaces init_genrand(0xDEADBEEF+T), main.c:172-176 */
    int device;               /* CUDA device ordinal */
    int z0, nz;               /* this handle's Z-slab [z0, z0+nz) of the global lattice; nz = 0 means the whole Z */
    int kernel;               /* sn_kernel */
} sn_params;

/* number of r^2 bins sn_rdf fills: r^2 = 0..80 (analysis.c:540-550) */
#define SN_RDF_BINS 81

SN_API const char *sn_last_error(void);
SN_API const char *sn_version(void);

/* CUDA devices this process can use (0 without a GPU; sn_create then fails: there is no CPU path) */
SN_API int sn_device_count(int *n);

/* fills *p with the reference's defaults (config.c:12-93): 20^3, cutoff 3, ... */
SN_API int sn_default_params(sn_params *p);

/* replaces: lattice malloc + gen_neighbour()  (main.c:155-161,180; montecarlo-core.c:38-72) */
SN_API int sn_create(const sn_params *p, sn_handle **out);
SN_API int sn_destroy(sn_handle *h);

/* the neighbour list the handle built; dxyz has 3*n ints, d has n floats, in the
 * reference's order (montecarlo-core.c:47-62).  Pass NULL arrays to query n. */
SN_API int sn_neighbour_table(sn_handle *h, int *n, int *dxyz, float *d);

/* replaces direct writes/reads of `lattice` (config.c:32-36).  host buffer:
 * float[X][Y][nz][4].  H2D / D2H copies happen inside the call. */
SN_API int sn_set_lattice(sn_handle *h, int replica, const float *xyzlen);
SN_API int sn_get_lattice(sn_handle *h, int replica, float *xyzlen);
/* The same, queued on the handle's stream without waiting: the buffer (pinned host memory, or the copy is not
 * asynchronous) must stay untouched until sn_synchronize (or any synchronous call on the handle) returns.  With two
 * handles used alternately, one lattice travels over PCIe while the other is being swept (see sn_order_after). */
SN_API int sn_set_lattice_async(sn_handle *h, int replica, const float *xyzlen);
SN_API int sn_get_lattice_async(sn_handle *h, int replica, float *xyzlen);
/* `count` consecutive replicas at once, host block float[count][X][Y][nz][4]: one copy and one kernel for the whole batch
 * (a T x CageStrain grid of small lattices would otherwise pay the per-call overhead a thousand times per step) */
SN_API int sn_set_lattices_async(sn_handle *h, int first, int count, const float *xyzlen);
SN_API int sn_get_lattices_async(sn_handle *h, int first, int count, float *xyzlen);
/* h's later work waits (on the device) for the sweeps queued so far on `other`: fixes the order of the two handles'
 * persistent sweep kernels when they are used as a double buffer.  Z-slab handles that share a GPU normally split its
 * SMs (their kernels may have to be resident together); handles ordered this way never run at once and are given
 * the whole share from the first call on -- the caller must then keep ordering every launch, in the same order on
 * every GPU of the decomposition. */
SN_API int sn_order_after(sn_handle *h, sn_handle *other);

/* replaces assignments to the globals beta (main.c:215,239), Efield (config.c:132-134),
 * CageStrain (main.c:149) between MC_moves calls.  Stream-ordered: the new value applies to the sweeps queued after the
 * call (and to the audit / energy entry points called after it); the calls do not wait for earlier sweeps to finish. */
SN_API int sn_set_beta(sn_handle *h, int replica, double beta);
SN_API int sn_set_efield(sn_handle *h, int replica, const float E[3]);
SN_API int sn_set_cagestrain(sn_handle *h, double cagestrain);           /* every replica */
/* one replica: the reference's sweep mode is a T x CageStrain grid of independent runs (Makefile:52-54
 * `superparallel`; argv[2] -> CageStrain, main.c:147-151), here one replica batch */
SN_API int sn_set_replica_cagestrain(sn_handle *h, int replica, double cagestrain);

/* replaces MC_moves(X*Y*Z*nsweeps) (montecarlo-core.c:143-149; main.c:222,248).
 * One sweep attempts one Metropolis update at every site of every replica, in
 * colour-sublattice order; each attempt is a fresh full dE over the cut-off
 * sphere (site_energy, montecarlo-core.c:76-141) followed by the accept test
 * (montecarlo-core.c:179).  Returns after the work is queued on the handle's
 * stream; any later call that reads results synchronises. */
SN_API int sn_mc_sweeps(sn_handle *h, long long nsweeps);

/* same, bracketed by CUDA events on the handle's stream: *ms = device time of
 * the whole call, *launches = kernels launched.  Used by bench.py. */
SN_API int sn_mc_sweeps_timed(sn_handle *h, long long nsweeps, double *ms, long long *launches);

SN_API int sn_synchronize(sn_handle *h);

/* Audit of MC_move (montecarlo-core.c:151-191) as the sweep kernel executes it: runs ONE sweep (it counts like
 * sn_mc_sweeps(h, 1): same chain, same counters) and returns one record of SN_AUDIT_WORDS floats per attempt,
 * records[replica][x][y][zlocal][8] =
 *   { trial dipole x, y, z (config.c:203-263),  accept uniform u (:179),  dE the kernel used (:154),
 *     decision: 1 accepted, 0 rejected, 2 vacant site skipped (:163),
 *     group: ordinal of the set of mutually independent sites the attempt was made in -- attempts of a lower
 *            group come before it in the chain, attempts of one group never read each other's sites,
 *     0 }
 * A host replay in group order on a copy of the lattice reproduces every dE with the reference's own
 * site_energy (tests/test_gpu_audit.py).  Tiled handles run the tiled kernel itself (audit instantiation). */
#define SN_AUDIT_WORDS 8
SN_API int sn_mc_sweep_audit(sn_handle *h, float *records);

/* replaces the globals ACCEPT / REJECT (config.c:26-27; montecarlo-core.c:187-190).
 * vacant = attempts that hit a length==0 site, which the reference neither
 * accepts nor rejects (montecarlo-core.c:163). */
SN_API int sn_get_counters(sn_handle *h, int replica, unsigned long long *accept,
                    unsigned long long *reject, unsigned long long *vacant);
SN_API int sn_reset_counters(sn_handle *h);
SN_API int sn_set_counters(sn_handle *h, int replica, unsigned long long accept, unsigned long long reject,
                    unsigned long long vacant);

/* Checkpoint / restart (the reference has none; long runs on large lattices need it).  The state of the
 * chain is the lattice (sn_get_lattice / sn_set_lattice), the counters above and the number of sweeps done,
 * which is the counter word of the per-site Philox streams: a handle created with the same parameters and
 * seed, given the saved lattice and sweep count, continues the chain bit for bit.  On Z-slab handles call
 * sn_set_sweep_count on every slab before the next sn_mc_sweeps. */
/* Give one replica its own Philox key: it then draws exactly the numbers replica 0 of a handle created with
 * `seed` would draw, so a batch of replicas (a temperature or field sweep) reproduces as many independent
 * runs -- the reference's only parallel mode (Makefile:49-63, one process per T, seed 0xDEADBEEF + T). */
SN_API int sn_set_replica_seed(sn_handle *h, int replica, unsigned long long seed);
SN_API int sn_get_sweep_count(sn_handle *h, unsigned long long *sweeps_done);
SN_API int sn_set_sweep_count(sn_handle *h, unsigned long long sweeps_done);

/* audit of site_energy (montecarlo-core.c:76-141): dE[i] of rotating site
 * sites[3i..3i+2] = (x, y, zlocal) to newdip[3i..3i+2], lattice unchanged. */
SN_API int sn_site_energy(sn_handle *h, int replica, int precision, int n, const int *sites,
                   const float *newdip, double *dE);

/* total lattice energy, defined so that site_energy is its exact single-site
 * difference (the reference has none: main.c:63).  out = {E_dd, E_cage, E_field, E_K}
 *   E_dd   = 1/2 sum_i sum_j l_i l_j [p_i.p_j - 3 (n.p_i)(n.p_j)] / d^3
 *   E_cage = -1/2 CageStrain sum_i sum_nn p_i.p_j
 *   E_field= sum_i p_i.E            E_K = -K sum_i (|p_ix| + |p_iy|)  [K > 0]
 * For a slab handle the sums run over the handle's own sites. */
SN_API int sn_total_energy(sn_handle *h, int replica, int precision, double out[4]);

/* replaces polarisation() (analysis.c:48-62): P = (1/N) sum_i p_i, all three components */
SN_API int sn_polarisation(sn_handle *h, int replica, double P[3]);
/* 64-bit content hash of the handle's own sites (position-keyed, summed mod 2^64): the hashes of the Z-slabs of a
 * decomposed lattice add up to the hash of the same lattice on one GPU iff every site holds the same bits.  The
 * reference has nothing like it; bench.py and the multi-GPU tests use it to show N-GPU chain == 1-GPU chain. */
SN_API int sn_state_hash(sn_handle *h, int replica, unsigned long long *hash);
/* which sweep kernel sn_mc_sweeps runs on this handle (an sn_kernel value other than SN_KERNEL_AUTO) */
SN_API int sn_kernel_in_use(sn_handle *h, int *kernel);

/* Diagnostic (host only, no GPU needed): the work order of the tiled kernel for one sweep of a periodic X x Y x Z lattice --
 * items[n][5] = {replica, tile x, tile y, tile z, phase} in the order the persistent CTAs take them; *n_items = tiles x
 * replicas.  Adjacent tiles (26-neighbourhood, periodic) never share a phase; an item starts once its neighbours have
 * finished the items that precede it in this order.  items may be NULL to query the count. */
SN_API int sn_tile_schedule(int X, int Y, int Z, int nreplicas, unsigned long long sweep, int *n_items, int *items, int max_items);
/* replaces landau_order() (analysis.c:506-526): |sum_i p_i|^2 / N * N as written there */
SN_API int sn_landau_order(sn_handle *h, int replica, double *landau);
/* Observables on a Z-slab handle cover the handle's own sites; planes beyond the slab are read from the neighbouring
 * GPUs over NVLink (the slab must be at least as high as the stencil radius: 9 for the RDF, 6 for the potential).
 * Sums (sn_rdf, sn_total_energy, sn_polarisation x sites, sn_recombination_partial) add up over the slabs; maps are
 * per slab.  Call them on every slab between the same two sweeps.
 *
 * replaces radial_order_parameter() (analysis.c:528-598): accumulated sums and
 * counts per r^2 = 0..80 BEFORE the division at :587-588, in FP64 / int64 */
SN_API int sn_rdf(sn_handle *h, int replica, double *fe_sum, double *afe_sum, long long *count);
/* replaces dipole_potential() over the lattice (analysis.c:65-94,264-308): V[X][Y][nz] */
SN_API int sn_potential_map(sn_handle *h, int replica, double *V);
/* replaces dipole_electricfield(cutoff, x, y, z) over the lattice (analysis.c:393-479, lattice_Efield_XYZ uses
 * cutoff 4) and, with half_offset != 0, dipole_electricfieldoffset (analysis.c:310-389, cutoff 2): |E| per site,
 * Emag[X][Y][nz] */
SN_API int sn_efield_map(sn_handle *h, int replica, int cutoff, int half_offset, double *Emag);
/* replaces the physics of recombination_calculator() (analysis.c:96-170): out = ZBe ZBh ZFDe ZFDh R_Boltz R_FD
 * FD-Total-electron FD-Total-hole eMAX hMAX RMAX */
#define SN_RECOMB_N 11
SN_API int sn_recombination(sn_handle *h, int replica, double out[SN_RECOMB_N]);
/* the same in two halves for a Z-slab decomposed lattice: the partial sums over one handle's own sites (5 partition /
 * occupation sums, 3 maxima over the global z = 0 plane, the site count), then the normalisations over all slabs */
#define SN_RECOMB_PARTIAL_N 9
SN_API int sn_recombination_partial(sn_handle *h, int replica, double part[SN_RECOMB_PARTIAL_N]);
SN_API int sn_recombination_finish(int nparts, const double *parts, double out[SN_RECOMB_N]);

/* Test aid: Philox4x32-10 (Salmon et al., SC'11), the counter-based generator that replaces the reference's global
 * MT19937 stream (mt19937ar-cok.c; montecarlo-core.c:159-161,179), evaluated for n inputs of 6 words
 * (counter[4], key[2]) by the host build of the function and by the device.  out_* receive 4 words per input;
 * out_device may be NULL (then no GPU is touched).  Checked against the Random123 known-answer vectors. */
SN_API int sn_philox_kat(int n, const unsigned int *counter_key, unsigned int *out_host, unsigned int *out_device);

/* measurement aid for bench.py: sustained FFMA throughput of `device` in TFLOP/s, the
 * denominator of the FP32 CUDA-core roofline (MEASURED_PEAKS.json carries no FP32 figure) */
SN_API int sn_bench_fp32_peak(int device, double *tflops);
/* the same for DFMA: the denominator of the FP64 roofline of the observable kernels (analysis.c sums in double) */
SN_API int sn_bench_fp64_peak(int device, double *tflops);

/* ---- Z-slab decomposition across GPUs (one handle per GPU) ------------------
 * A slab handle keeps `cutoff` ghost planes below and above its own planes.
 * sn_get_boundary / sn_set_ghost move them through host memory (bootstrap and
 * CPU tests); sn_ipc_* wire the device-to-device path used by sn_mc_sweeps. */
/* side 0 = lowest `cutoff` own planes, 1 = highest; buffer float[X][Y][cutoff][4] */
SN_API int sn_get_boundary(sn_handle *h, int replica, int side, float *planes);
/* side 0 = ghost planes below z0 (the lower neighbour's top planes), 1 = above */
SN_API int sn_set_ghost(sn_handle *h, int replica, int side, const float *planes);
/* device-to-device alternative to sn_get_boundary + sn_set_ghost: once every slab has its lattice (sn_set_lattice)
 * and its neighbours (sn_ipc_attach / sn_attach_peer), each slab copies the neighbours' boundary planes into its
 * ghost planes over NVLink, bracketed by the device-side handshake.  Stream-ordered; call it on every slab. */
SN_API int sn_pull_ghosts(sn_handle *h);
/* 64-byte CUDA IPC handles of the lattice buffer and the phase-flag buffer */
SN_API int sn_ipc_export(sn_handle *h, void *lattice_handle64, void *flags_handle64);
/* attach the neighbour that owns the planes below (side 0) / above (side 1) */
SN_API int sn_ipc_attach(sn_handle *h, int side, const void *lattice_handle64, const void *flags_handle64);
/* same wiring for two handles living in one process (peer access is enabled).  Slabs of one lattice wait for each
 * other from inside their persistent sweep kernels, so they must be able to run at the same time: one slab per GPU
 * is the intended layout; several slab handles of one process on the same device share its SMs equally (handles in
 * different processes cannot see each other: keep to one slab per GPU there). */
SN_API int sn_attach_peer(sn_handle *h, int side, sn_handle *peer);

#ifdef __cplusplus
}
#endif
#endif /* STARRYNIGHT_B200_H */
