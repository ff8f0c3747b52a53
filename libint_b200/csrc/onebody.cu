// One-body integrals on the GPU: overlap S, kinetic T and nuclear attraction V over a basis,
// the three compute_1body_ints<Operator::overlap | kinetic | nuclear> calls of the reference's direct-SCF
// driver (tests/hartree-fock/hartree-fock++.cc:267-275, :1064-1152; Engine::compute1,
// include/libint2/engine.impl.h:181-561) that precede the Fock builds.
// One thread per shell pair (s1 >= s2): primitive loop, McMurchie-Davidson Hermite expansion
// (E coefficients per dimension, R_{tuv} auxiliaries per point charge with the Boys function from a
// convergent series / asymptotic form), Cartesian blocks in local memory, sparse cart->pure, both
// triangles of the dense matrices written.  Same basis-function convention as the two-electron path:
// coefficients as renormalized by Shell::renorm (shell.h:958-999), every Cartesian component of a shell
// carrying the normalization of x^l, STANDARD component order, solid harmonics of solidharmonics.h.
// An O(N^2 natoms) set-up step: written for correctness and to keep the SCF loop on the device, not
// tuned.
#include <vector>

#include "internal.h"

using namespace lb200;

namespace {

constexpr int kL1 = kMaxShellL;          // bra l
constexpr int kL2 = kMaxShellL + 2;      // ket l + 2 (kinetic-energy relation)
constexpr int kT = kL1 + kL2 + 1;        // Hermite index range of E
constexpr int kLV = 2 * kMaxShellL;      // total l of the nuclear-attraction auxiliaries
constexpr int kNC1 = (kMaxShellL + 1) * (kMaxShellL + 2) / 2;

struct OneBodyParams {
  int nshell, nbf, natom;
  long long npairs;
  const int *l, *pure, *nprim, *off, *shell2bf;
  const double *O, *alpha, *coeff;
  const double* charges;   // [natom][4]: Z, x, y, z
  const int *sph_rowptr, *sph_col, *sph_base;
  const double* sph_val;
  double *S, *T, *V;
};

// F_m(U), m = 0..mmax: series + downward recursion below 35, asymptotic value + upward recursion above
__device__ void boys_series(int mmax, double U, double* F) {
  const double eU = exp(-U);
  if (U < 35.0) {
    double term = 1.0 / (2 * mmax + 1), s = term;
    for (int k = 1; k < 400; ++k) {
      term *= 2.0 * U / (2 * mmax + 2 * k + 1);
      s += term;
      if (term < 1e-17 * s) break;
    }
    F[mmax] = eU * s;
    for (int m = mmax; m > 0; --m) F[m - 1] = (2.0 * U * F[m] + eU) / (2 * m - 1);
  } else {
    F[0] = 0.5 * sqrt(3.14159265358979323846 / U) * erf(sqrt(U));
    for (int m = 0; m < mmax; ++m) F[m + 1] = ((2 * m + 1) * F[m] - eU) / (2.0 * U);
  }
}

__device__ inline int tuv_index(int t, int u, int v) {   // t + u + v <= kLV, dense cube index
  return (t * (kLV + 1) + u) * (kLV + 1) + v;
}

__global__ void __launch_bounds__(64) onebody_kernel(const OneBodyParams p) {
  constexpr int RP = 2 * kMaxShellL + 2;
  for (long long tix = blockIdx.x * (long long)blockDim.x + threadIdx.x; tix < p.npairs;
       tix += (long long)gridDim.x * blockDim.x) {
    int a = (int)((sqrt(8.0 * (double)tix + 1.0) - 1.0) * 0.5);
    while ((long long)a * (a + 1) / 2 > tix) --a;
    while ((long long)(a + 1) * (a + 2) / 2 <= tix) ++a;
    const int b = (int)(tix - (long long)a * (a + 1) / 2);
    const int la = p.l[a], lb = p.l[b];
    const int na = nc(la), nb = nc(lb);
    const double A[3] = {p.O[3 * a], p.O[3 * a + 1], p.O[3 * a + 2]};
    const double B[3] = {p.O[3 * b], p.O[3 * b + 1], p.O[3 * b + 2]};
    double AB2 = 0;
    for (int k = 0; k < 3; ++k) AB2 += (A[k] - B[k]) * (A[k] - B[k]);
    double S[kNC1 * kNC1], T[kNC1 * kNC1], V[kNC1 * kNC1];
    for (int i = 0; i < na * nb; ++i) S[i] = T[i] = V[i] = 0.0;
    double E[3][kL1 + 1][kL2 + 1][kT + 1];
    double R[(kLV + 1) * (kLV + 1) * (kLV + 1)];        // R^0_{tuv}
    double Rn[kLV + 1][kLV + 1];                         // scratch: R^n along one recursion chain
    const int L = la + lb;
    for (int p1 = 0; p1 < p.nprim[a]; ++p1)
      for (int p2 = 0; p2 < p.nprim[b]; ++p2) {
        const double ea = p.alpha[p.off[a] + p1], eb = p.alpha[p.off[b] + p2];
        const double w = p.coeff[p.off[a] + p1] * p.coeff[p.off[b] + p2];
        const double g = ea + eb, o2p = 0.5 / g;
        const double pref = w * exp(-ea * eb / g * AB2);
        double P[3];
        for (int k = 0; k < 3; ++k) P[k] = (ea * A[k] + eb * B[k]) / g;
        // Hermite expansion coefficients E^{ij}_t per dimension, ket up to lb + 2
        for (int d = 0; d < 3; ++d) {
          const double PA = P[d] - A[d], PB = P[d] - B[d];
          for (int i = 0; i <= la; ++i)
            for (int j = 0; j <= lb + 2; ++j)
              for (int t = 0; t <= kT; ++t) E[d][i][j][t] = 0.0;
          E[d][0][0][0] = 1.0;
          for (int i = 0; i < la; ++i)
            for (int t = 0; t <= i + 1; ++t)
              E[d][i + 1][0][t] = (t > 0 ? o2p * E[d][i][0][t - 1] : 0.0) + PA * E[d][i][0][t] +
                                  (t + 1) * E[d][i][0][t + 1];
          for (int j = 0; j < lb + 2; ++j)
            for (int i = 0; i <= la; ++i)
              for (int t = 0; t <= i + j + 1; ++t)
                E[d][i][j + 1][t] = (t > 0 ? o2p * E[d][i][j][t - 1] : 0.0) + PB * E[d][i][j][t] +
                                    (t + 1) * E[d][i][j][t + 1];
        }
        const double s1 = sqrt(3.14159265358979323846 / g);
        auto S1 = [&](int d, int i, int j) { return j >= 0 ? E[d][i][j][0] * s1 : 0.0; };
        auto T1 = [&](int d, int i, int j) {
          return -2.0 * eb * eb * S1(d, i, j + 2) + eb * (2 * j + 1) * S1(d, i, j) -
                 0.5 * j * (j - 1) * S1(d, i, j - 2);
        };
        // overlap and kinetic energy
        {
          int ia = 0;
          for (int ax = la; ax >= 0; --ax)
            for (int ay = la - ax; ay >= 0; --ay, ++ia) {
              const int az = la - ax - ay;
              int ib = 0;
              for (int bx = lb; bx >= 0; --bx)
                for (int by = lb - bx; by >= 0; --by, ++ib) {
                  const int bz = lb - bx - by;
                  const double sx = S1(0, ax, bx), sy = S1(1, ay, by), sz = S1(2, az, bz);
                  S[ia * nb + ib] += pref * sx * sy * sz;
                  T[ia * nb + ib] += pref * (T1(0, ax, bx) * sy * sz + sx * T1(1, ay, by) * sz + sx * sy * T1(2, az, bz));
                }
            }
        }
        // nuclear attraction: sum over the point charges
        for (int c = 0; c < p.natom; ++c) {
          const double Z = p.charges[4 * c];
          const double PC[3] = {P[0] - p.charges[4 * c + 1], P[1] - p.charges[4 * c + 2], P[2] - p.charges[4 * c + 3]};
          double Fm[kLV + 1];
          boys_series(L, g * (PC[0] * PC[0] + PC[1] * PC[1] + PC[2] * PC[2]), Fm);
          // R^n_{tuv}: R^n_{000} = (-2g)^n F_n; R^n_{t+1,u,v} = t R^{n+1}_{t-1,u,v} + X R^{n+1}_{t,u,v}, etc.
          // Built with a (t,u,v)-outer loop and an n-chain in Rn: for every (u, v) first the v chain, ...
          // Simple dense scheme: Raux[n][tuv] needs too much local memory, so R^0_{tuv} is obtained by
          // recursing each target down to s-type auxiliaries along z, then y, then x (three nested chains).
          for (int t = 0; t <= L; ++t)
            for (int u = 0; u <= L - t; ++u)
              for (int v = 0; v <= L - t - u; ++v) R[tuv_index(t, u, v)] = 0.0;
          // z chains: Rz[n][v] = R^n_{00v}
          double Rz[kLV + 1][kLV + 1];
          {
            double f = 1.0;
            for (int n = 0; n <= L; ++n) { Rz[n][0] = f * Fm[n]; f *= -2.0 * g; }
            for (int v = 1; v <= L; ++v)
              for (int n = 0; n <= L - v; ++n)
                Rz[n][v] = (v > 1 ? (v - 1) * Rz[n + 1][v - 2] : 0.0) + PC[2] * Rz[n + 1][v - 1];
          }
          for (int v = 0; v <= L; ++v) {
            // y chains on top of R^n_{00v}: Rn[n][u] = R^n_{0uv}
            for (int n = 0; n <= L - v; ++n) Rn[n][0] = Rz[n][v];
            for (int u = 1; u <= L - v; ++u)
              for (int n = 0; n <= L - v - u; ++n)
                Rn[n][u] = (u > 1 ? (u - 1) * Rn[n + 1][u - 2] : 0.0) + PC[1] * Rn[n + 1][u - 1];
            for (int u = 0; u <= L - v; ++u) {
              // x chain on top of R^n_{0uv}
              double Rx[kLV + 1][kLV + 1];
              for (int n = 0; n <= L - v - u; ++n) Rx[n][0] = Rn[n][u];
              for (int t = 1; t <= L - v - u; ++t)
                for (int n = 0; n <= L - v - u - t; ++n)
                  Rx[n][t] = (t > 1 ? (t - 1) * Rx[n + 1][t - 2] : 0.0) + PC[0] * Rx[n + 1][t - 1];
              for (int t = 0; t <= L - v - u; ++t) R[tuv_index(t, u, v)] = Rx[0][t];
            }
          }
          const double fac = -Z * 2.0 * 3.14159265358979323846 / g * pref;
          int ia = 0;
          for (int ax = la; ax >= 0; --ax)
            for (int ay = la - ax; ay >= 0; --ay, ++ia) {
              const int az = la - ax - ay;
              int ib = 0;
              for (int bx = lb; bx >= 0; --bx)
                for (int by = lb - bx; by >= 0; --by, ++ib) {
                  const int bz = lb - bx - by;
                  double v = 0.0;
                  for (int t = 0; t <= ax + bx; ++t)
                    for (int u = 0; u <= ay + by; ++u) {
                      const double eu = E[0][ax][bx][t] * E[1][ay][by][u];
                      for (int vv = 0; vv <= az + bz; ++vv) v += eu * E[2][az][bz][vv] * R[tuv_index(t, u, vv)];
                    }
                  V[ia * nb + ib] += fac * v;
                }
            }
        }
      }
    // cart -> pure on both indices, then both triangles of the three matrices
    const bool pu1 = p.pure[a] != 0 && la > 0, pu2 = p.pure[b] != 0 && lb > 0;
    const int ma = pu1 ? 2 * la + 1 : na, mb = pu2 ? 2 * lb + 1 : nb;
    const int bfa = p.shell2bf[a], bfb = p.shell2bf[b];
    for (int which = 0; which < 3; ++which) {
      const double* M = which == 0 ? S : (which == 1 ? T : V);
      double* out = which == 0 ? p.S : (which == 1 ? p.T : p.V);
      for (int i = 0; i < ma; ++i) {
        double row[kNC1];
        for (int j = 0; j < nb; ++j) {
          double v = 0.0;
          if (pu1) {
            for (int k = p.sph_rowptr[la * RP + i]; k < p.sph_rowptr[la * RP + i + 1]; ++k)
              v += p.sph_val[p.sph_base[la] + k] * M[p.sph_col[p.sph_base[la] + k] * nb + j];
          } else {
            v = M[i * nb + j];
          }
          row[j] = v;
        }
        for (int m = 0; m < mb; ++m) {
          double v = 0.0;
          if (pu2) {
            for (int k = p.sph_rowptr[lb * RP + m]; k < p.sph_rowptr[lb * RP + m + 1]; ++k)
              v += p.sph_val[p.sph_base[lb] + k] * row[p.sph_col[p.sph_base[lb] + k]];
          } else {
            v = row[m];
          }
          out[(size_t)(bfa + i) * p.nbf + bfb + m] = v;
          out[(size_t)(bfb + m) * p.nbf + bfa + i] = v;
        }
      }
    }
  }
}

}  // namespace

// S, T, V (device or host, nbf x nbf each, row-major); charges = natom x {Z, x, y, z} (host)
extern "C" int lb200_onebody(lb200_context* ctx, const lb200_basis* bs, int natom, const double* charges,
                             double* S, double* T, double* V, int on_device) {
  if (!ctx || !bs || natom < 0 || (natom > 0 && !charges) || !S || !T || !V) return LB200_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const int ns = bs->nshell, n = bs->nbf;
  const long long npairs = (long long)ns * (ns + 1) / 2;
  if (npairs == 0) return LB200_OK;
  const size_t nprimtot = bs->alpha.size(), n2 = (size_t)n * n;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_l = 0, o_pu = al(o_l + ns * 4), o_np = al(o_pu + ns * 4), o_off = al(o_np + ns * 4);
  const size_t o_s2b = al(o_off + (ns + 1) * 4), o_O = al(o_s2b + ns * 4), o_al = al(o_O + 3 * (size_t)ns * 8);
  const size_t o_co = al(o_al + nprimtot * 8), o_ch = al(o_co + nprimtot * 8);
  const size_t o_S = al(o_ch + 4 * (size_t)natom * 8 + 8), total = o_S + (on_device ? 0 : 3 * al(n2 * 8));
  char* d = nullptr;
  int rc = check_cuda(ctx, cudaMalloc(&d, total), "cudaMalloc(onebody)");
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  cudaMemcpyAsync(d + o_l, bs->l.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_pu, bs->pure.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_np, bs->nprim.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_off, bs->off.data(), (ns + 1) * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_s2b, bs->shell2bf.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_O, bs->O.data(), 3 * (size_t)ns * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_al, bs->alpha.data(), nprimtot * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d + o_co, bs->coeff.data(), nprimtot * 8, cudaMemcpyHostToDevice, st);
  if (natom > 0) cudaMemcpyAsync(d + o_ch, charges, 4 * (size_t)natom * 8, cudaMemcpyHostToDevice, st);
  OneBodyParams p;
  p.nshell = ns; p.nbf = n; p.natom = natom; p.npairs = npairs;
  p.l = reinterpret_cast<const int*>(d + o_l); p.pure = reinterpret_cast<const int*>(d + o_pu);
  p.nprim = reinterpret_cast<const int*>(d + o_np); p.off = reinterpret_cast<const int*>(d + o_off);
  p.shell2bf = reinterpret_cast<const int*>(d + o_s2b);
  p.O = reinterpret_cast<const double*>(d + o_O); p.alpha = reinterpret_cast<const double*>(d + o_al);
  p.coeff = reinterpret_cast<const double*>(d + o_co); p.charges = reinterpret_cast<const double*>(d + o_ch);
  p.sph_rowptr = ctx->d_sph_rowptr; p.sph_col = ctx->d_sph_col; p.sph_base = ctx->d_sph_base; p.sph_val = ctx->d_sph_val;
  if (on_device) {
    p.S = S; p.T = T; p.V = V;
  } else {
    p.S = reinterpret_cast<double*>(d + o_S);
    p.T = reinterpret_cast<double*>(d + o_S + al(n2 * 8));
    p.V = reinterpret_cast<double*>(d + o_S + 2 * al(n2 * 8));
  }
  const int grid = (int)std::min<long long>((npairs + 63) / 64, (long long)ctx->num_sms * 32);
  onebody_kernel<<<grid, 64, 0, st>>>(p);
  ++ctx->launches;
  if (!on_device) {
    cudaMemcpyAsync(S, p.S, n2 * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(T, p.T, n2 * 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(V, p.V, n2 * 8, cudaMemcpyDeviceToHost, st);
  }
  rc = check_cuda(ctx, cudaStreamSynchronize(st), "onebody");
  cudaFree(d);
  return rc;
}
