/* libint_b200.h -- C ABI of the B200 (sm_100a) two-electron Coulomb integral + direct Fock path.
 *
 * Drop-in boundary. In the reference (evaleev/libint @ 7a1a9d8) the hot path sits behind
 *   (1) the generated C interface   Libint_t + libint2_build_eri[la][lb][lc][ld](Libint_t*)
 *       (src/bin/libint/iface.cc:114-185, used at include/libint2/engine.impl.h:623-635,1898-1899)
 *   (2) the C++ API                 libint2::Engine::compute2<coulomb, xx_xx|xs_xx, 0>
 *       (include/libint2/engine.h:787-791, engine.impl.h:1151-2113)
 *   (3) its consumer                compute_2body_fock (tests/hartree-fock/hartree-fock++.cc:1574-1772)
 * Those interfaces are one-quartet-at-a-time.  This library keeps their *data contracts*
 * (Shell normalization, ShellPair primitive screening, Cartesian / solid-harmonic orderings,
 * row-major n1*n2*n3*n4 result layout, G = J - K/2 digestion convention) and exposes the
 * batched entry points a binding needs; INTEGRATION.md shows the Engine-side glue.
 *
 * Conventions
 *   - plain C: pointers, sizes, error codes (0 = success, negative = error; no exceptions
 *     cross this boundary).  lb200_last_error() returns a message for the calling context.
 *   - every "on_device" flag says whether the corresponding buffer is device memory of the
 *     context's GPU (e.g. a torch CUDA tensor) or host memory.
 *   - shells carry normalization-embedded coefficients, i.e. what libint2::Shell::contr[0].coeff
 *     holds after Shell::renorm() (include/libint2/shell.h:958-999); lb200_shell_renorm does that.
 *   - no CPU fallback: every compute entry point fails with LB200_ERR_CUDA if no GPU is usable.
 */
#ifndef LIBINT_B200_H
#define LIBINT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB200_OK 0
#define LB200_ERR_INVALID (-1)     /* bad argument / unsupported class                         */
#define LB200_ERR_CUDA (-2)        /* CUDA runtime error (see lb200_last_error)                  */
#define LB200_ERR_LMAX (-3)        /* angular momentum beyond the built kernels
                                      (mirrors Engine::lmax_exceeded, engine.h:893-916)         */
#define LB200_ERR_NOMEM (-4)

/* ScreeningMethod values, identical to include/libint2/shell.h:1041-1059 */
#define LB200_SCREEN_ORIGINAL 0x0001
#define LB200_SCREEN_CONSERVATIVE 0x0010
#define LB200_SCREEN_SCHWARZ 0x0100
#define LB200_SCREEN_SCHWARZ_INF 0x1000

#define LB200_MAX_AM 4             /* s..g per shell (LIBINT2_MAX_AM_eri analogue)              */

typedef struct lb200_context lb200_context;
typedef struct lb200_basis lb200_basis;
typedef struct lb200_pairs lb200_pairs;
typedef struct lb200_fock lb200_fock;
typedef struct lb200_comm lb200_comm;
typedef struct lb200_df3c lb200_df3c;

/* ---- library / context life cycle: replaces libint2::initialize()/finalize()
 *      (include/libint2/initialize.h:76-136) and Engine construction (engine.h:503-526). */
int lb200_version(void);
int lb200_device_count(void);
int lb200_context_create(int device, lb200_context** out);
int lb200_context_destroy(lb200_context* ctx);
const char* lb200_last_error(const lb200_context* ctx);
/* use an externally owned CUDA stream (cudaStream_t cast to void*), e.g. torch's current stream */
int lb200_context_set_stream(lb200_context* ctx, void* cuda_stream);
int lb200_context_synchronize(lb200_context* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long lb200_context_launch_count(const lb200_context* ctx);
/* FP64 FMA throughput of the context's GPU, measured by a register-resident FMA loop
 * (the denominator of the FP64 roofline; no reference counterpart).  Returns TFLOP/s. */
int lb200_fp64_peak_probe(lb200_context* ctx, int iters, double* tflops, double* ms);

/* ---- Shell normalization: Shell::renorm(), shell.h:958-999. coeff is updated in place;
 *      max_ln_coeff (shell.h:1001-1011) is optional. */
int lb200_shell_renorm(int l, int nprim, const double* alpha, double* coeff,
                       int enforce_unit_normalization, double* max_ln_coeff);

/* ---- basis = std::vector<libint2::Shell> flattened (include/libint2/basis.h.in, shell.h:720):
 *      l[nshell], pure[nshell], nprim[nshell], origin[3*nshell], then alpha/coeff concatenated
 *      (sum nprim entries); one contraction per shell, as Engine::compute2 requires
 *      (engine.impl.h:1167-1171). */
int lb200_basis_create(lb200_context* ctx, int nshell, const int* l, const int* pure,
                       const int* nprim, const double* origin, const double* alpha,
                       const double* coeff, lb200_basis** out);
/* the unit shell of Shell::unit() (shell.h:906-909,949-953): bra2 of the 3-centre integrals */
int lb200_basis_create_unit(lb200_context* ctx, lb200_basis** out);
int lb200_basis_destroy(lb200_basis* bs);
int lb200_basis_nbf(const lb200_basis* bs);
int lb200_basis_nshell(const lb200_basis* bs);
/* first basis function of every shell: BasisSet::shell2bf() */
int lb200_basis_shell2bf(const lb200_basis* bs, int* out);

/* ---- shell-pair data: ShellPair::init (shell.h:1138-1328), built once and kept on the GPU.
 *      All pairs of one call must belong to one class (l(s1) >= l(s2), same l's and purity).
 *      screening: ORIGINAL / CONSERVATIVE use ln_prec as in shell.h:1162-1232;
 *      SCHWARZ / SCHWARZ_INF need prim_schwarz (one factor per primitive pair, ordered
 *      [pair][p1][p2], = schwarz_factor_evaluator of hartree-fock++.cc:1390-1412); for SCHWARZ_INF
 *      NULL makes the library compute them on the GPU (sqrt of max |(ab|ab)| per primitive pair);
 *      SCHWARZ (Frobenius norm) with NULL is rejected with LB200_ERR_INVALID.
 *      All pairs of one block share one kind of second shell (ordinary or Shell::unit()). */
int lb200_pairs_create(lb200_context* ctx, const lb200_basis* bs1, const lb200_basis* bs2,
                       int npair, const int* s1, const int* s2, int screening, double ln_prec,
                       const double* prim_schwarz, lb200_pairs** out);
int lb200_pairs_destroy(lb200_pairs* p);
/* info[0..5] = la, lb, npair, total primitive pairs kept, pure_a, pure_b */
int lb200_pairs_info(const lb200_pairs* p, long long* info);
/* copies the primitive-pair records of pair i: 9 doubles each
 * {P[3], K, one_over_gamma, nonsph_screen_fac, ln_scr, p1, p2} (PrimPairData, shell.h:1084-1092);
 * returns the count */
int lb200_pairs_get(const lb200_pairs* p, int i, double* out, int cap);

/* ---- batched Engine::compute2<coulomb, *, 0>: one shell set per task.
 *      tasks = ntasks x {bra pair index, ket pair index}.  Output: ntasks blocks of
 *      n_a*n_b*n_c*n_d doubles, row-major in the order (bra.first, bra.second, ket.first,
 *      ket.second); Cartesian (`pure_out` = 0) or, with `pure_out` = 1, transformed for the
 *      shells flagged pure (engine.impl.h:1965-1985).  A task whose primitives are all
 *      screened out (results()[0] == nullptr in the reference, engine.impl.h:1781-1784)
 *      yields zeros.  precision <= 0 disables primitive screening (Engine::set_precision). */
int lb200_eri_batch(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket,
                    long long ntasks, const int* tasks, int tasks_on_device, int screening,
                    double precision, int pure_out, double* out, int out_on_device);
/* 1 if a kernel exists for (la lb|lc ld) with la >= lb, lc >= ld (either bra/ket order), else 0:
 * the analogue of a null libint2_build_eri[la][lb][lc][ld] entry (engine.impl.h:1898). */
int lb200_eri_class_supported(int la, int lb, int lc, int ld);
/* doubles per task written by lb200_eri_batch */
long long lb200_eri_block_size(const lb200_pairs* bra, const lb200_pairs* ket, int pure_out);

/* ---- the reference's innermost plugin call, batched: libint2_build_eri[la][lb][lc][ld](Libint_t*)
 *      (src/bin/libint/iface.cc:114-185, used at include/libint2/engine.impl.h:1898-1899).  The
 *      caller supplies the per-primitive prerequisites Engine::compute2 writes into Libint_t
 *      (engine.impl.h:1514-1641), in its own bra/ket orientation; VRR, contraction and HRR run on
 *      the GPU.  Class (la lb|lc ld) with la >= lb, lc >= ld, either bra/ket order.
 *        prim_off[ntasks+1]  records of shell set t are [prim_off[t], prim_off[t+1])  (contrdepth)
 *        recs                LB200_PREREQ_DOUBLES doubles per primitive quartet:
 *                            (ss|ss)^(m) m = 0..24 | PA[3] | QC[3] | WP[3] | WQ[3] |
 *                            oo2z oo2e oo2ze roz roe      (PA = 0 for a unit-shell bra2: 3eri/2eri)
 *        geom[ntasks][6]     AB[3] = A - B, CD[3] = C - D
 *        out[ntasks][ncart(la) ncart(lb) ncart(lc) ncart(ld)]   row-major, Cartesian
 *      Host buffers; synchronous.  include/libint2_b200_iface (liblibint_b200_iface.so) wraps this
 *      into the Libint_t / libint2_build_* ABI itself. */
#define LB200_PREREQ_DOUBLES 42
int lb200_eri_prereq_batch(lb200_context* ctx, int la, int lb, int lc, int ld, long long ntasks,
                           const int* prim_off, const double* recs, const double* geom,
                           double* out);

/* ---- direct Fock build: compute_2body_fock (hartree-fock++.cc:1574-1772).
 *      pairs (s1 >= s2) = the significant shell-pair list obs_shellpair_list
 *      (hartree-fock++.cc:1305-1381); lb200_significant_pairs computes it.
 *      create() evaluates the Schwarz matrix (hartree-fock++.cc:1230-1298) and the
 *      SchwarzInf primitive-pair data (:1383-1431) on the GPU. */
int lb200_significant_pairs(const lb200_basis* bs, double threshold, int* s1, int* s2,
                            long long cap, long long* count);
/* the same list evaluated on the GPU (one thread per shell pair); cap >= nshell(nshell+1)/2 returns the
 * whole list in one call */
int lb200_significant_pairs_device(lb200_context* ctx, const lb200_basis* bs, double threshold, int* s1,
                                   int* s2, long long cap, long long* count);
int lb200_fock_create(lb200_context* ctx, const lb200_basis* obs, long long npair, const int* s1,
                      const int* s2, lb200_fock** out);
int lb200_fock_destroy(lb200_fock* f);
int lb200_fock_schwarz(const lb200_fock* f, double* K /* nshell*nshell, host */);
/* rank that owns the quartet (bra pair | ket pair), pairs named by their canonical index
 * s1*(s1+1)/2 + s2 (s1 >= s2); host-callable copy of the rule the screening kernel applies.
 * Ownership goes by the BRA pair alone (bra = the pair of the class with the larger
 * angular-momentum key, or the larger canonical index inside one class), so every rank
 * enumerates only its own rows of the task matrix.
 * Replaces the reference's thread round-robin s1234 % nthreads (hartree-fock++.cc:1665). */
int lb200_fock_task_owner(int bra_pair_index, int ket_pair_index, int nranks);
/* G = 1/2 (g + g^T), g accumulated as in hartree-fock++.cc:1721-1743 from density D (nbf x nbf,
 * row-major).  Only the quartets with (task id % nranks) == rank are processed, so that N
 * processes each produce a partial G to be summed (ncclAllReduce by the caller); pass 0, 1
 * for the whole build.  stats (optional, 4 doubles): shell quartets computed, kernel
 * launches, device milliseconds, candidate quartets screened. */
int lb200_fock_build(lb200_fock* f, const double* D, int D_on_device, double precision,
                     int use_schwarz, int rank, int nranks, double* G, int G_on_device,
                     double* stats);

/* ---- first geometric derivatives: batched Engine::compute2<Operator::coulomb, BraKet::xx_xx, 1>
 *      (engine.impl.h:1151-2113 with deriv_order 1; hartree-fock++.cc:1978).  Same arguments as lb200_eri_batch;
 *      output: ntasks x 12 blocks of n_a*n_b*n_c*n_d doubles, block d = 3 * centre + xyz with the centres in
 *      the order (bra.first, bra.second, ket.first, ket.second) -- Engine::results()[0..11].  Built from the
 *      ordinary class kernels with one angular momentum shifted (d/dA_x (ab|cd) = 2 alpha_a (a+1_x b|cd) -
 *      a_x (a-1_x b|cd), src/bin/test_eri/eri.h:383-460), fourth centre by translational invariance; a class
 *      whose raised twins have no kernel (l = 3 next to l >= 1) returns LB200_ERR_LMAX -- the analogue of
 *      LIBINT2_MAX_AM_eri1 < LIBINT2_MAX_AM_eri.  Four-centre blocks only (no Shell::unit()). */
/* the six shifted shell sets a class (la lb|lc ld), la >= lb, lc >= ld, is differentiated from -- A+, A-, B+, B-,
 * C+, C- -- as 5 integers each: doubles per task (0: the lowered shell does not exist) and the strides of the
 * component indices (a, b, c, d) inside a block.  Host-only; LB200_ERR_LMAX as lb200_eri_deriv1_batch. */
int lb200_eri_deriv1_plan(int la, int lb, int lc, int ld, long long* plan /* 30 */);
int lb200_eri_deriv1_batch(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket,
                           long long ntasks, const int* tasks, int tasks_on_device, int screening,
                           double precision, int pure_out, double* out, int out_on_device);

/* ---- two-body forces of the direct SCF driver: F2[3 * atom + xyz] = sum_ij G1[3 atom + xyz]_ij D_ij with
 *      G1 = compute_2body_fock_deriv<1>(obs, atoms, D) (hartree-fock++.cc:1775-2055, used at :648-656),
 *      evaluated on the GPU without forming the 3 * natoms matrices G1: same quartets, screening and rank
 *      ownership as lb200_fock_build; shell2atom[nshell] = BasisSet::shell2atom(atoms); grad = 3 * natoms
 *      doubles (host).  With nranks > 1 every rank returns a partial gradient to be summed.
 *      stats (optional, 3 doubles): shell quartets, kernel launches, device milliseconds. */
int lb200_fock_grad(lb200_fock* f, const double* D, int D_on_device, double precision, int use_schwarz,
                    int rank, int nranks, int natoms, const int* shell2atom, double* grad, double* stats);

/* ---- batched Engine::compute2 over an implicit Cartesian product of pair ranges: task t =
 *      (bra pair b0 + t / nk, ket pair k0 + t % nk), t < nb * nk -- no task list is built or read
 *      (the three-centre sweep and the two-centre metric are such products).  Device output only. */
int lb200_eri_product(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket, int b0, int nb,
                      int k0, int nk, int screening, double precision, int pure_out, double* out_device);

/* ---- one-body integrals on the GPU: overlap S, kinetic T, nuclear attraction V of a basis, the three
 *      compute_1body_ints<Operator::overlap|kinetic|nuclear> calls of tests/hartree-fock/hartree-fock++.cc:
 *      267-275 (Engine::compute1, engine.impl.h:181-561).  charges = natom x {Z, x, y, z} (host;
 *      make_point_charges, :1064); S, T, V = nbf x nbf row-major, device (on_device = 1) or host. */
int lb200_onebody(lb200_context* ctx, const lb200_basis* bs, int natom, const double* charges, double* S,
                  double* T, double* V, int on_device);

/* ---- one-body contributions to the nuclear forces, the first block of the reference driver's force section
 *      (tests/hartree-fock/hartree-fock++.cc:601-627):  F1[3 atom + xyz] = 2 sum_ij (T1 + V1)_ij D_ij  and
 *      FPulay[3 atom + xyz] = -2 sum_ij S1_ij W_ij  with S1 / T1 / V1 = compute_1body_ints_deriv<overlap | kinetic |
 *      nuclear>(1, obs, atoms) (:1154-1228; Engine::compute1 with deriv_order 1) and W = C_occ eps_occ C_occ^T.  The
 *      3 natom derivative matrices are never formed: each shell pair's derivative integrals are contracted with its
 *      density block on the GPU.  charges = natom x {Z, x, y, z} (host); shell2atom[nshell] (host;
 *      BasisSet::shell2atom); D, W = nbf x nbf row-major, device (on_device = 1) or host; F1, FPulay: host,
 *      3 natom doubles each.  LB200_ERR_LMAX if a shell has l > 4. */
int lb200_onebody_forces(lb200_context* ctx, const lb200_basis* bs, int natom, const double* charges,
                         const int* shell2atom, const double* D, const double* W, int on_device, double* F1,
                         double* FPulay);

/* ---- density fitting: three-centre integrals (P|mu nu) as dense slabs and the two-centre metric (P|Q):
 *      the DF set-up of tests/hartree-fock/hartree-fock++.cc:2215-2262 (Zxy[ndf][n][n] via
 *      Engine::compute2<coulomb, xs_xx>(dfbs[s1], Shell::unit(), obs[s2], obs[s3])) and
 *      compute_2body_2index_ints (:1517-1571).
 *      create: obs pairs (s1 >= s2) = the significant orbital shell pairs (lb200_significant_pairs).
 *      slab:   Z[P - first][mu][nu] (device, row-major, pure where flagged, both (mu,nu) and (nu,mu)) for the
 *              functions of DF shells [P0, P0 + nP); zeroed first.  threshold > 0 drops triplets whose
 *              Schwarz-type bound sqrt|(P|P)| sqrt|(mu nu|mu nu)| is below it (the reference computes all);
 *              precision as in lb200_eri_batch.  stats (optional, 2 doubles): triplets computed, triplets total.
 *              Slabs of disjoint shell ranges are independent: ranks shard by DF shell, no collective.
 *      metric: V[ndf][ndf] (device).   info[0..5] = nbf, ndf, DF shells, orbital pairs, shell triplets, groups. */
int lb200_df3c_create(lb200_context* ctx, const lb200_basis* obs, const lb200_basis* dfbs, long long npair,
                      const int* s1, const int* s2, lb200_df3c** out);
int lb200_df3c_destroy(lb200_df3c* f);
int lb200_df3c_slab(lb200_df3c* f, int P0, int nP, double threshold, double precision, double* Z_device,
                    double* stats);
int lb200_df3c_metric(lb200_df3c* f, double* V_device);
int lb200_df3c_info(const lb200_df3c* f, long long* info);

/* ---- multi-GPU: one process per GPU, each builds the partial G of the bra rows it owns
 *      (lb200_fock_build with rank / nranks), then one in-place ncclAllReduce(sum, f64) over NVLink
 *      combines them -- the GPU form of the reference's sum over thread-private G's
 *      (hartree-fock++.cc:1753-1755).  NCCL is bound at run time (the library does not link it).
 *        id = lb200_comm_unique_id() on rank 0 (128 bytes, ncclUniqueId), distributed by the host's own
 *        means (MPI_Bcast, a file, torch.distributed ...), then lb200_comm_create on every rank; or wrap a
 *        communicator the host already has (ncclComm_t cast to void*) with lb200_comm_from_nccl.
 *      lb200_fock_allreduce runs on the context's stream (asynchronous, like every device-buffer call). */
int lb200_comm_unique_id(char* id, int cap /* >= 128 */);
int lb200_comm_create(lb200_context* ctx, int nranks, int rank, const char* id, lb200_comm** out);
int lb200_comm_from_nccl(lb200_context* ctx, void* nccl_comm, lb200_comm** out);
int lb200_comm_destroy(lb200_comm* c);
int lb200_comm_rank(const lb200_comm* c);
int lb200_comm_size(const lb200_comm* c);
int lb200_fock_allreduce(lb200_comm* c, double* G_device, long long count);

/* Profiling of the build (no reference counterpart): with profiling on, every (bra class, ket class,
 * contraction buckets) launch of the next builds is timed with CUDA events (one stream sync per
 * launch: diagnostics, not for timed runs) and its surviving primitive quartets are counted in the
 * kernel.  get_profile copies 9 doubles per row -- la, lb, lc, ld, bra bucket, ket bucket, device ms,
 * shell quartets, surviving primitive quartets -- sorted by time; returns the row count (rows = NULL:
 * count only).  bench.py derives the per-class FP64 roofline from it. */
int lb200_fock_set_profile(lb200_fock* f, int on);
long long lb200_fock_get_profile(const lb200_fock* f, double* rows, long long cap);

#ifdef __cplusplus
}
#endif
#endif /* LIBINT_B200_H */
