// Host-side objects behind the C ABI (include/libint_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/libint_b200.h"
#include "eri_kernel.cuh"

struct lb200_context {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  double* d_boys = nullptr;
  lb200::RowInfo* d_rows = nullptr;
  // sparse cart->pure tables for the generic transform kernel: per l, CSR over pure index
  int* d_sph_rowptr = nullptr;   // [(kMaxShellL+1)][2*kMaxShellL+2]
  int* d_sph_col = nullptr;
  double* d_sph_val = nullptr;
  int* d_sph_base = nullptr;     // [(kMaxShellL+1)] offset of each l into col/val
  long long launches = 0;
  // scratch of the host-buffer lb200_eri_batch path, kept between calls (grown on demand):
  // task list, two Cartesian / two transformed chunk buffers, copy stream and its events
  // (slots 5..7: the derivative path, deriv.cu)
  static constexpr int kScratchSlots = 8;
  void* d_scratch[kScratchSlots] = {};
  size_t scratch_bytes[kScratchSlots] = {};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  mutable std::string err;
};

struct lb200_basis {
  lb200_context* ctx = nullptr;
  int nshell = 0, nbf = 0;
  std::vector<int> l, pure, nprim, off, shell2bf;
  std::vector<double> O, alpha, coeff, max_ln_coeff;
  bool is_unit(int s) const { return nprim[s] == 1 && alpha[off[s]] == 0.0 && l[s] == 0; }
  int size(int s) const { return pure[s] ? 2 * l[s] + 1 : (l[s] + 1) * (l[s] + 2) / 2; }
};

struct lb200_deriv_blocks;

struct lb200_pairs {
  lb200_context* ctx = nullptr;
  lb200::PairBlock dev{};  // device view
  // host copies (used by lb200_pairs_get, the Fock driver and the derivative path).  Blocks whose
  // primitive records were made on the GPU (pairs_device.cu) fill prim / Kraw / p1p2 lazily:
  // lb200::pairs_host_mirror() downloads them on first use.
  std::vector<int> prim_off, shell, bf, p1p2;
  std::vector<lb200::PrimPair> prim;
  std::vector<double> AB, Kraw;   // Kraw: K of shell.h:1241-1243 without the coefficient product
  std::vector<double> A;          // centre of the first shell of every pair
  long long nprim_total = 0;      // primitive pairs kept (= prim.size() once mirrored)
  bool host_valid = true;
  const double* d_Kraw = nullptr; // device-built blocks: K and (p1, p2) of every record, inside d_block
  const int2* d_p1p2 = nullptr;
  // exponents of the two bases (copies; the derivative path scales K*c_a*c_b by 2 alpha)
  std::vector<double> alpha1, alpha2;
  std::vector<int> off1, off2;
  lb200_deriv_blocks* deriv = nullptr;   // shifted-angular-momentum twins (deriv.cu), built on demand
  void* d_block = nullptr;  // single allocation backing all device arrays
};

namespace lb200 {

int set_error(const lb200_context* ctx, int code, const std::string& msg);
int check_cuda(const lb200_context* ctx, cudaError_t e, const char* what);

// class dispatch (dispatch.cu): launches kernel-oriented class (la lb|lc ld)
bool class_supported(int la, int lb, int lc, int ld);
int order_key(int la, int lb);
cudaError_t launch_eri(int la, int lb, int lc, int ld, const EriParams& p, const RowInfo* rows,
                       int mode, int num_sms, cudaStream_t stream);

// implicit task list of run_store: every (bra pair b0 + i, ket pair k0 + j), i < nb, j < nk, j fastest
struct ProductTasks { int b0, nb, k0, nk; };
// store-mode launch for device-resident tasks (explicit list, or `prod`) into a device buffer (Cartesian)
int run_store(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket, long long ntasks,
              const int2* d_tasks, int screening, double precision, double* d_out,
              const ProductTasks* prod = nullptr);

// generic cart -> pure transform of a batch of shell sets (transform.cu)
cudaError_t launch_pure_transform(const lb200_context* ctx, const double* in, double* out,
                                  long long ntasks, const int l[4], const int pure[4],
                                  cudaStream_t stream);
// per-task max |x| over blocks of n doubles
cudaError_t launch_block_absmax(const double* in, double* out, long long ntasks, long long n,
                                cudaStream_t stream);

// pair-block assembly shared by lb200_pairs_create and the Fock driver
int build_pairs(lb200_context* ctx, const lb200_basis* bs1, const lb200_basis* bs2, int npair,
                const int* s1, const int* s2, int screening, double ln_prec,
                const double* prim_schwarz, const double* pair_schwarz, lb200_pairs** out);
// ShellPair::init for a block of pairs on the GPU (pairs_device.cu): count pass, then -- after the host's
// prefix sum -- the fill pass straight into the block's device arrays
struct DevicePrimBuilder {
  lb200_context* ctx = nullptr;
  int npair = 0, screening = 0;
  double ln_prec = 0;
  char* d_tmp = nullptr;
  void* params = nullptr;
  ~DevicePrimBuilder();
  int init(lb200_context* ctx, const lb200_basis* bs1, const lb200_basis* bs2, int npair, const int* s1,
           const int* s2, int screening, double ln_prec, const double* prim_schwarz);
  int count(std::vector<int>& counts);
  int fill(const int* d_prim_off, PrimPair* d_prim, double* d_Kraw, int2* d_p1p2);
};
// fills the host copies of a block whose primitive records were made on the GPU
int pairs_host_mirror(const lb200_pairs* P);
// uploads prim_off / geometry / shell / gidx / schwarz (+ the host-made primitive records when
// `with_prims`) into one device allocation and points P->dev at it; extra_prim_bytes reserves room per
// primitive record behind the block (device-built K and p1p2 arrays)
int upload_pairs(lb200_context* ctx, lb200_pairs* P, const std::vector<PairGeom>& geom,
                 const double* pair_schwarz, bool with_prims);
void free_deriv_blocks(lb200_pairs* P);
// grow-only device scratch of the context (synchronises the context's streams when it has to grow)
int ctx_scratch(lb200_context* ctx, int slot, size_t bytes, void** out);

// ---- first geometric derivatives (deriv.cu) --------------------------------------------------------
// The twelve derivative shell sets of a quartet are assembled from six ordinary shell sets with one
// angular momentum shifted: d/dA_x (ab|cd) = 2 alpha_a (a+1_x b|cd) - a_x (a-1_x b|cd)  (the relation the
// reference's closed-form check uses, src/bin/test_eri/eri.h:383-460), centres A, B, C explicitly and D by
// translational invariance.  DerivSets names the six scratch buffers of one chunk of tasks.
struct DerivBuf {
  const double* p;      // [ntasks][blk], null when the lowered shell does not exist (l = 0)
  long long blk;
  int s[4];             // strides of the (a, b, c, d) component indices inside a block
};
struct DerivSets {
  int l[4], n[4];       // original class, caller's order (bra.first, bra.second, ket.first, ket.second)
  DerivBuf buf[6];      // A+, A-, B+, B-, C+, C-
  long long doubles_per_task;
};
// describes the six sets of (bra | ket) and checks that every shifted class has a kernel
int deriv_plan(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket, DerivSets& ds);
int deriv_plan_class(lb200_context* ctx, int la, int lb, int lc, int ld, DerivSets& ds);
// evaluates the six sets of `ntasks` tasks into `scratch` (doubles_per_task * ntasks doubles) on the
// context's stream and fills ds.buf[*].p
int deriv_eval(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket, long long ntasks,
               const int2* d_tasks, int screening, double precision, double* scratch, DerivSets& ds);
// out[t][12][n0 n1 n2 n3], Cartesian
cudaError_t launch_deriv_store(const DerivSets& ds, long long ntasks, double* out, cudaStream_t st);
// grad[3 * atom + xyz] += deg * sum_abcd dI * (2 D_ab D_cd - 1/2 D_ac D_bd - 1/2 D_ad D_bc), Cartesian-ised D
struct DerivGradParams {
  const int2* tasks;        // (bra pair, ket pair)
  const int4* ftasks;       // the screening kernel's records: degeneracy code in bits 30-31 of .y
  const int* bra_shell;     // [npair][2]
  const int* ket_shell;
  const int* shell2cbf;     // first Cartesian function of every shell
  const int* shell2atom;
  const double* Dc;         // [nbfc][nbfc]
  int nbfc;
  double* grad;             // [3 * natoms]
};
cudaError_t launch_deriv_grad(const DerivSets& ds, long long ntasks, const DerivGradParams& gp, cudaStream_t st);
// Dc = C^T D C: the density in the Cartesian functions of every shell (pure shells back-transformed)
cudaError_t launch_cartesianize_density(const lb200_context* ctx, const double* D, int nbf, double* Dc, int nbfc,
                                        int nshell, const int* d_l, const int* d_pure, const int* d_shell2bf,
                                        const int* d_shell2cbf, const int* d_cbf2shell, cudaStream_t st);
cudaError_t launch_unpack_tasks(const int4* ftasks, const unsigned* count, int2* tasks, cudaStream_t st,
                                long long cap);
// SchwarzInf primitive factors, one per (pair, p1, p2) (hartree-fock++.cc:1390-1412)
int compute_prim_schwarz(lb200_context* ctx, const lb200_basis* bs1, const lb200_basis* bs2,
                         int npair, const int* s1, const int* s2, std::vector<double>& out);

}  // namespace lb200
