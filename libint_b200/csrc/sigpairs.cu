// Significant shell-pair list on the GPU: compute_shellpairs of the reference's direct-SCF driver
// (tests/hartree-fock/hartree-fock++.cc:1305-1381): a pair (s1 >= s2) is kept if the shells share a
// centre or the Frobenius norm of their overlap block (solid harmonics where flagged) reaches the
// threshold.  One thread per pair: primitive loop with the Obara-Saika 1-d overlap recursion, Cartesian
// block in local memory, sparse cart->pure on both indices, norm.  The host version
// (lb200_significant_pairs, host_ints.cc) is the same arithmetic in a serial O(N^2) loop -- 11 s of
// set-up for (H2O)_256 / cc-pVTZ (15.9 M pairs); this kernel takes milliseconds.
#include <vector>

#include "internal.h"

using namespace lb200;

namespace {

struct SigParams {
  int nshell;
  long long npairs;
  const int *l, *pure, *nprim, *off;
  const double *O, *alpha, *coeff;
  const int *sph_rowptr, *sph_col, *sph_base;
  const double* sph_val;
  double threshold;
  unsigned char* flag;
};

constexpr int kNC = (kMaxShellL + 1) * (kMaxShellL + 2) / 2;   // 15

__global__ void sigpair_kernel(const SigParams p) {
  constexpr int RP = 2 * kMaxShellL + 2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < p.npairs;
       t += (long long)gridDim.x * blockDim.x) {
    int a = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long long)a * (a + 1) / 2 > t) --a;
    while ((long long)(a + 1) * (a + 2) / 2 <= t) ++a;
    const int b = (int)(t - (long long)a * (a + 1) / 2);
    const double Ax = p.O[3 * a], Ay = p.O[3 * a + 1], Az = p.O[3 * a + 2];
    const double Bx = p.O[3 * b], By = p.O[3 * b + 1], Bz = p.O[3 * b + 2];
    if (Ax == Bx && Ay == By && Az == Bz) { p.flag[t] = 1; continue; }
    const int l1 = p.l[a], l2 = p.l[b];
    const int n1 = nc(l1), n2 = nc(l2);
    const double A[3] = {Ax, Ay, Az}, B[3] = {Bx, By, Bz};
    double AB2 = 0;
    for (int k = 0; k < 3; ++k) AB2 += (A[k] - B[k]) * (A[k] - B[k]);
    double S[kNC * kNC];
    for (int i = 0; i < n1 * n2; ++i) S[i] = 0.0;
    double I[3][kMaxShellL + 1][kMaxShellL + 1];
    for (int p1 = 0; p1 < p.nprim[a]; ++p1)
      for (int p2 = 0; p2 < p.nprim[b]; ++p2) {
        const double a1 = p.alpha[p.off[a] + p1], a2 = p.alpha[p.off[b] + p2];
        const double c = p.coeff[p.off[a] + p1] * p.coeff[p.off[b] + p2];
        const double g = a1 + a2, oog = 1.0 / g, rho = a1 * a2 * oog;
        const double pio = 3.14159265358979323846 * oog;
        const double pref = c * exp(-rho * AB2) * pio * sqrt(pio);
        for (int k = 0; k < 3; ++k) {
          const double P = (a1 * A[k] + a2 * B[k]) * oog;
          const double PA = P - A[k], PB = P - B[k];
          I[k][0][0] = 1.0;
          for (int i = 0; i <= l1; ++i) {
            if (i > 0) I[k][i][0] = PA * I[k][i - 1][0] + (i > 1 ? (i - 1) * 0.5 * oog * I[k][i - 2][0] : 0.0);
            for (int j = 1; j <= l2; ++j)
              I[k][i][j] = PB * I[k][i][j - 1] + (j > 1 ? (j - 1) * 0.5 * oog * I[k][i][j - 2] : 0.0) +
                           (i > 0 ? i * 0.5 * oog * I[k][i - 1][j - 1] : 0.0);
          }
        }
        int i = 0;
        for (int x1 = l1; x1 >= 0; --x1)
          for (int y1 = l1 - x1; y1 >= 0; --y1, ++i) {
            const int z1 = l1 - x1 - y1;
            int j = 0;
            for (int x2 = l2; x2 >= 0; --x2)
              for (int y2 = l2 - x2; y2 >= 0; --y2, ++j)
                S[i * n2 + j] += pref * I[0][x1][x2] * I[1][y1][y2] * I[2][z1][l2 - x2 - y2];
          }
      }
    // Frobenius norm of C1 S C2^T (C = sparse cart->pure rows where the shell is pure, else identity)
    const bool pu1 = p.pure[a] != 0 && l1 > 0, pu2 = p.pure[b] != 0 && l2 > 0;
    const int m1 = pu1 ? 2 * l1 + 1 : n1, m2 = pu2 ? 2 * l2 + 1 : n2;
    double nrm = 0.0;
    for (int i = 0; i < m1; ++i) {
      double row[kNC];   // row i of C1 S
      for (int j = 0; j < n2; ++j) {
        double v = 0.0;
        if (pu1) {
          for (int k = p.sph_rowptr[l1 * RP + i]; k < p.sph_rowptr[l1 * RP + i + 1]; ++k)
            v += p.sph_val[p.sph_base[l1] + k] * S[p.sph_col[p.sph_base[l1] + k] * n2 + j];
        } else {
          v = S[i * n2 + j];
        }
        row[j] = v;
      }
      for (int m = 0; m < m2; ++m) {
        double v = 0.0;
        if (pu2) {
          for (int k = p.sph_rowptr[l2 * RP + m]; k < p.sph_rowptr[l2 * RP + m + 1]; ++k)
            v += p.sph_val[p.sph_base[l2] + k] * row[p.sph_col[p.sph_base[l2] + k]];
        } else {
          v = row[m];
        }
        nrm += v * v;
      }
    }
    p.flag[t] = sqrt(nrm) >= p.threshold ? 1 : 0;
  }
}

}  // namespace

extern "C" int lb200_significant_pairs_device(lb200_context* ctx, const lb200_basis* bs, double threshold,
                                              int* s1, int* s2, long long cap, long long* count) {
  if (!ctx || !bs || !count) return LB200_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const int ns = bs->nshell;
  const long long T = (long long)ns * (ns + 1) / 2;
  *count = 0;
  if (T == 0) return LB200_OK;
  const size_t nprimtot = bs->alpha.size();
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_l = 0, o_pu = al(o_l + ns * 4), o_np = al(o_pu + ns * 4), o_off = al(o_np + ns * 4);
  const size_t o_O = al(o_off + (ns + 1) * 4), o_al = al(o_O + 3 * (size_t)ns * 8), o_co = al(o_al + nprimtot * 8);
  const size_t o_fl = al(o_co + nprimtot * 8), total = o_fl + (size_t)T;
  char* d = nullptr;
  int rc = check_cuda(ctx, cudaMalloc(&d, total), "cudaMalloc(significant pairs)");
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  cudaMemcpyAsync(d + o_l, bs->l.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_pu, bs->pure.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_np, bs->nprim.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_off, bs->off.data(), (ns + 1) * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_O, bs->O.data(), 3 * (size_t)ns * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_al, bs->alpha.data(), nprimtot * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_co, bs->coeff.data(), nprimtot * 8, cudaMemcpyHostToDevice, st);
  SigParams p;
  p.nshell = ns; p.npairs = T;
  p.l = reinterpret_cast<const int*>(d + o_l); p.pure = reinterpret_cast<const int*>(d + o_pu);
  p.nprim = reinterpret_cast<const int*>(d + o_np); p.off = reinterpret_cast<const int*>(d + o_off);
  p.O = reinterpret_cast<const double*>(d + o_O); p.alpha = reinterpret_cast<const double*>(d + o_al);
  p.coeff = reinterpret_cast<const double*>(d + o_co);
  p.sph_rowptr = ctx->d_sph_rowptr; p.sph_col = ctx->d_sph_col; p.sph_base = ctx->d_sph_base; p.sph_val = ctx->d_sph_val;
  p.threshold = threshold;
  p.flag = reinterpret_cast<unsigned char*>(d + o_fl);
  const int grid = (int)std::min<long long>((T + 127) / 128, (long long)ctx->num_sms * 64);
  sigpair_kernel<<<grid, 128, 0, st>>>(p);
  ++ctx->launches;
  std::vector<unsigned char> flag((size_t)T);
  cudaMemcpyAsync(flag.data(), p.flag, (size_t)T, cudaMemcpyDeviceToHost, st);
  rc = check_cuda(ctx, cudaStreamSynchronize(st), "significant pairs");
  cudaFree(d);
  if (rc) return rc;
  long long n = 0, t = 0;
  for (int a = 0; a < ns; ++a)
    for (int b = 0; b <= a; ++b, ++t)
      if (flag[(size_t)t]) {
        if (n < cap && s1 && s2) { s1[n] = a; s2[n] = b; }
        ++n;
      }
  *count = n;
  return (s1 && n > cap) ? LB200_ERR_NOMEM : LB200_OK;
}
