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
  void* d_scratch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[5] = {0, 0, 0, 0, 0};
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

struct lb200_pairs {
  lb200_context* ctx = nullptr;
  lb200::PairBlock dev{};  // device view
  // host copies (small; used by lb200_pairs_get and the Fock driver)
  std::vector<int> prim_off, shell, bf, p1p2;
  std::vector<lb200::PrimPair> prim;
  std::vector<double> AB, Kraw;   // Kraw: K of shell.h:1241-1243 without the coefficient product
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
// SchwarzInf primitive factors, one per (pair, p1, p2) (hartree-fock++.cc:1390-1412)
int compute_prim_schwarz(lb200_context* ctx, const lb200_basis* bs1, const lb200_basis* bs2,
                         int npair, const int* s1, const int* s2, std::vector<double>& out);

}  // namespace lb200
