// lb200_onebody_forces: the one-body and Pulay force contributions of the reference's direct-SCF driver
// (tests/hartree-fock/hartree-fock++.cc:601-627) on the GPU.  One thread per shell pair (s1 >= s2) runs
// ob1::pair_forces (onebody_deriv.cuh) on the Cartesian-ised densities C^T D C and C^T W C; the 3 * natoms sums of a
// CTA are collected in shared memory and flushed with one FP64 atomic per entry.  A set-up-sized step (O(N^2 natoms)
// once per geometry), written for correctness like onebody.cu, not tuned.
#include <algorithm>
#include <vector>

#include "internal.h"
#include "onebody_deriv.cuh"

using namespace lb200;

namespace {

struct ForceParams {
  int nshell, natom, nbfc;
  long long npairs;
  const int *l, *nprim, *off, *shell2cbf, *shell2atom;
  const double *O, *alpha, *coeff, *charges;
  const double *Dc, *Wc;      // [nbfc][nbfc]
  double* F;                  // [2][3 * natom]: F1, F_Pulay
  int use_shared;
};

struct DeviceAcc {
  double* buf;   // shared ([2][3 natom]) or global
  int n3;
  __device__ void add(int which, int idx, double v) {
    if (v != 0.0) atomicAdd(buf + which * n3 + idx, v);
  }
};

__global__ void __launch_bounds__(64) onebody_forces_kernel(const ForceParams p) {
  extern __shared__ double sF[];
  const int n3 = 3 * p.natom;
  if (p.use_shared) {
    for (int i = threadIdx.x; i < 2 * n3; i += blockDim.x) sF[i] = 0.0;
    __syncthreads();
  }
  DeviceAcc acc{p.use_shared ? sF : p.F, n3};
  for (long long tix = blockIdx.x * (long long)blockDim.x + threadIdx.x; tix < p.npairs;
       tix += (long long)gridDim.x * blockDim.x) {
    int a = (int)((sqrt(8.0 * (double)tix + 1.0) - 1.0) * 0.5);
    while ((long long)a * (a + 1) / 2 > tix) --a;
    while ((long long)(a + 1) * (a + 2) / 2 <= tix) ++a;
    const int b = (int)(tix - (long long)a * (a + 1) / 2);
    const int la = p.l[a], lb = p.l[b];
    const int na = nc(la), nb = nc(lb);
    const int oa = p.shell2cbf[a], ob = p.shell2cbf[b];
    double wD[ob1::kNC * ob1::kNC], wW[ob1::kNC * ob1::kNC];
    for (int i = 0; i < na; ++i)
      for (int j = 0; j < nb; ++j) {
        const size_t ab = (size_t)(oa + i) * p.nbfc + ob + j, ba = (size_t)(ob + j) * p.nbfc + oa + i;
        wD[i * nb + j] = a == b ? p.Dc[ab] : p.Dc[ab] + p.Dc[ba];
        wW[i * nb + j] = a == b ? p.Wc[ab] : p.Wc[ab] + p.Wc[ba];
      }
    ob1::pair_forces(la, lb, p.O + 3 * a, p.O + 3 * b, p.nprim[a], p.alpha + p.off[a], p.coeff + p.off[a],
                     p.nprim[b], p.alpha + p.off[b], p.coeff + p.off[b], wD, wW, p.shell2atom[a],
                     p.shell2atom[b], p.natom, p.charges, acc);
  }
  if (p.use_shared) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * n3; i += blockDim.x)
      if (sF[i] != 0.0) atomicAdd(p.F + i, sF[i]);
  }
}

}  // namespace

// F1[3 natom] = 2 sum (T1 + V1) o D, FPulay[3 natom] = -2 sum S1 o W (host); D, W = nbf x nbf row-major,
// device (on_device = 1) or host; charges = natom x {Z, x, y, z} (host); shell2atom[nshell] (host)
extern "C" int lb200_onebody_forces(lb200_context* ctx, const lb200_basis* bs, int natom, const double* charges,
                                    const int* shell2atom, const double* D, const double* W, int on_device,
                                    double* F1, double* FPulay) {
  if (!ctx || !bs || natom <= 0 || !charges || !shell2atom || !D || !W || !F1 || !FPulay) return LB200_ERR_INVALID;
  const int ns = bs->nshell, n = bs->nbf;
  for (int s = 0; s < ns; ++s) {
    if (shell2atom[s] < 0 || shell2atom[s] >= natom)
      return set_error(ctx, LB200_ERR_INVALID, "lb200_onebody_forces: shell2atom entry out of range");
    if (bs->l[s] > ob1::kL) return set_error(ctx, LB200_ERR_LMAX, "lb200_onebody_forces: shell l exceeds the supported maximum");
  }
  cudaSetDevice(ctx->device);
  const int n3 = 3 * natom;
  std::fill(F1, F1 + n3, 0.0);
  std::fill(FPulay, FPulay + n3, 0.0);
  const long long npairs = (long long)ns * (ns + 1) / 2;
  if (npairs == 0) return LB200_OK;
  std::vector<int> shell2cbf(ns), cbf2shell;
  int nbfc = 0;
  for (int s = 0; s < ns; ++s) {
    shell2cbf[s] = nbfc;
    const int k = (bs->l[s] + 1) * (bs->l[s] + 2) / 2;
    cbf2shell.insert(cbf2shell.end(), k, s);
    nbfc += k;
  }
  const size_t nprimtot = bs->alpha.size(), n2 = (size_t)n * n, nc2 = (size_t)nbfc * nbfc;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_l = 0, o_pu = al(o_l + ns * 4), o_np = al(o_pu + ns * 4), o_off = al(o_np + ns * 4);
  const size_t o_s2b = al(o_off + (ns + 1) * 4), o_s2c = al(o_s2b + ns * 4), o_c2s = al(o_s2c + ns * 4);
  const size_t o_s2a = al(o_c2s + (size_t)nbfc * 4), o_O = al(o_s2a + ns * 4), o_al = al(o_O + 3 * (size_t)ns * 8);
  const size_t o_co = al(o_al + nprimtot * 8), o_ch = al(o_co + nprimtot * 8), o_F = al(o_ch + 4 * (size_t)natom * 8);
  const size_t o_Dc = al(o_F + 2 * (size_t)n3 * 8), o_Wc = al(o_Dc + nc2 * 8), o_D = al(o_Wc + nc2 * 8);
  const size_t total = o_D + (on_device ? 0 : 2 * al(n2 * 8));
  char* d = nullptr;
  int rc = check_cuda(ctx, cudaMalloc(&d, total), "cudaMalloc(onebody_forces)");
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  auto up = [&](size_t o, const void* src, size_t bytes) { cudaMemcpyAsync(d + o, src, bytes, cudaMemcpyHostToDevice, st); };
  up(o_l, bs->l.data(), ns * 4);
  up(o_pu, bs->pure.data(), ns * 4);
  up(o_np, bs->nprim.data(), ns * 4);
  up(o_off, bs->off.data(), (ns + 1) * 4);
  up(o_s2b, bs->shell2bf.data(), ns * 4);
  up(o_s2c, shell2cbf.data(), ns * 4);
  up(o_c2s, cbf2shell.data(), (size_t)nbfc * 4);
  up(o_s2a, shell2atom, ns * 4);
  up(o_O, bs->O.data(), 3 * (size_t)ns * 8);
  up(o_al, bs->alpha.data(), nprimtot * 8);
  up(o_co, bs->coeff.data(), nprimtot * 8);
  up(o_ch, charges, 4 * (size_t)natom * 8);
  cudaMemsetAsync(d + o_F, 0, 2 * (size_t)n3 * 8, st);
  const double *dD = D, *dW = W;
  if (!on_device) {
    up(o_D, D, n2 * 8);
    up(o_D + al(n2 * 8), W, n2 * 8);
    dD = reinterpret_cast<const double*>(d + o_D);
    dW = reinterpret_cast<const double*>(d + o_D + al(n2 * 8));
  }
  auto ip = [&](size_t o) { return reinterpret_cast<const int*>(d + o); };
  auto dp = [&](size_t o) { return reinterpret_cast<double*>(d + o); };
  cudaError_t e = launch_cartesianize_density(ctx, dD, n, dp(o_Dc), nbfc, ns, ip(o_l), ip(o_pu), ip(o_s2b), ip(o_s2c),
                                              ip(o_c2s), st);
  if (e == cudaSuccess)
    e = launch_cartesianize_density(ctx, dW, n, dp(o_Wc), nbfc, ns, ip(o_l), ip(o_pu), ip(o_s2b), ip(o_s2c), ip(o_c2s), st);
  ctx->launches += 2;
  if (e == cudaSuccess) {
    ForceParams p;
    p.nshell = ns; p.natom = natom; p.nbfc = nbfc; p.npairs = npairs;
    p.l = ip(o_l); p.nprim = ip(o_np); p.off = ip(o_off); p.shell2cbf = ip(o_s2c); p.shell2atom = ip(o_s2a);
    p.O = dp(o_O); p.alpha = dp(o_al); p.coeff = dp(o_co); p.charges = dp(o_ch);
    p.Dc = dp(o_Dc); p.Wc = dp(o_Wc); p.F = dp(o_F);
    const size_t shmem = 2 * (size_t)n3 * 8;
    p.use_shared = shmem <= 40 * 1024;
    const int grid = (int)std::min<long long>((npairs + 63) / 64, (long long)ctx->num_sms * 16);
    onebody_forces_kernel<<<grid, 64, p.use_shared ? shmem : 0, st>>>(p);
    ++ctx->launches;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    cudaMemcpyAsync(F1, d + o_F, (size_t)n3 * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(FPulay, d + o_F + (size_t)n3 * 8, (size_t)n3 * 8, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
  }
  rc = check_cuda(ctx, e, "onebody_forces");
  cudaFree(d);
  return rc;
}
