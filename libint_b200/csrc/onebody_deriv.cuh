// One-body contributions to the nuclear forces of the reference's direct-SCF driver
// (tests/hartree-fock/hartree-fock++.cc:601-627): F1 = 2 sum_ij (T1 + V1)_ij D_ij and F_Pulay = -2 sum_ij S1_ij W_ij
// with S1 / T1 / V1 = compute_1body_ints_deriv<overlap | kinetic | nuclear>(1, obs, atoms) (:1154-1228;
// Engine::compute1 with deriv_order 1, engine.impl.h:181-561).  The 3 * natoms derivative matrices are never
// formed: one shell pair's derivative integrals are contracted with its density block on the fly.
//
// Scheme (McMurchie-Davidson, same conventions as onebody.cu): d/dA_x of a Cartesian Gaussian is
// 2 alpha (a + 1_x| - a_x (a - 1_x|, applied to the Hermite expansion coefficients E^{ij}_t of one dimension.
//  * overlap, kinetic energy: only d/dA is evaluated, d/dB = -d/dA (they depend on A - B only);
//  * nuclear attraction: the density-weighted derivative coefficients are gathered once per primitive pair into
//    six "Hermite densities" Lambda^{A|B, x|y|z}_{tuv}; every point charge then costs one R_{tuv} table and six dot
//    products, and its own centre receives -(d/dA + d/dB) of its term (translational invariance).
// Everything here is __host__ __device__ and works on the density in the CARTESIAN functions of every shell
// (launch_cartesianize_density), so that the same source is compiled by g++ into the CPU test harness
// (tests/cxx/onebody_forces_host.cc: test infrastructure, checked against the numpy derivative integrals that
// reproduce the reference's golden forces) and by nvcc into onebody_forces_kernel (onebody_deriv.cu).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LB200_OB_HD __host__ __device__
#else
#define LB200_OB_HD
#endif

namespace lb200 {
namespace ob1 {

constexpr int kL = 4;                    // largest shell l (kMaxShellL)
constexpr int kNI = kL + 2;              // first index 0 .. la + 1
constexpr int kNJ = kL + 3;              // second index 0 .. lb + 2 (kinetic-energy relation; lb + 1 for d/dB)
constexpr int kNT = 2 * kL + 5;          // Hermite index 0 .. la + lb + 3, one more read by the recursion
constexpr int kNR = 2 * kL + 2;          // orders 0 .. la + lb + 1 of the nuclear-attraction auxiliaries
constexpr int kNC = (kL + 1) * (kL + 2) / 2;
constexpr int kNTet = kNR * (kNR + 1) * (kNR + 2) / 6;   // (t, u, v) with t + u + v <= kNR - 1
constexpr double kPi = 3.14159265358979323846;

// F_m(U), m = 0..mmax: series + downward recursion below 35, asymptotic value + upward recursion above
LB200_OB_HD inline void boys_series(int mmax, double U, double* F) {
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
    F[0] = 0.5 * sqrt(kPi / U) * erf(sqrt(U));
    for (int m = 0; m < mmax; ++m) F[m + 1] = ((2 * m + 1) * F[m] - eU) / (2.0 * U);
  }
}

// compact index of (t, u, v), t + u + v <= N
struct Tet {
  int N, off[kNR + 1];
  LB200_OB_HD explicit Tet(int n) : N(n) {
    off[0] = 0;
    for (int t = 0; t <= N; ++t) off[t + 1] = off[t] + (N - t + 1) * (N - t + 2) / 2;
  }
  LB200_OB_HD int size() const { return off[N + 1]; }
  LB200_OB_HD int operator()(int t, int u, int v) const {
    const int M = N - t;
    return off[t] + u * (M + 1) - u * (u - 1) / 2 + v;
  }
};

// R^0_{tuv}(p, P - C), t + u + v <= N, into R[tet(t, u, v)]: R^n_{000} = (-2p)^n F_n(p |PC|^2),
// R^n_{t+1,u,v} = t R^{n+1}_{t-1,u,v} + X_PC R^{n+1}_{t,u,v} -- a z chain, y chains on top of it, x chains on top of those
LB200_OB_HD inline void hermite_R(int N, double p, const double* PC, const Tet& tet, double* R) {
  double Fm[kNR];
  boys_series(N, p * (PC[0] * PC[0] + PC[1] * PC[1] + PC[2] * PC[2]), Fm);
  double Rz[kNR][kNR], Ry[kNR][kNR], Rx[kNR][kNR];
  double f = 1.0;
  for (int n = 0; n <= N; ++n) { Rz[n][0] = f * Fm[n]; f *= -2.0 * p; }
  for (int v = 1; v <= N; ++v)
    for (int n = 0; n <= N - v; ++n)
      Rz[n][v] = (v > 1 ? (v - 1) * Rz[n + 1][v - 2] : 0.0) + PC[2] * Rz[n + 1][v - 1];
  for (int v = 0; v <= N; ++v) {
    for (int n = 0; n <= N - v; ++n) Ry[n][0] = Rz[n][v];
    for (int u = 1; u <= N - v; ++u)
      for (int n = 0; n <= N - v - u; ++n)
        Ry[n][u] = (u > 1 ? (u - 1) * Ry[n + 1][u - 2] : 0.0) + PC[1] * Ry[n + 1][u - 1];
    for (int u = 0; u <= N - v; ++u) {
      for (int n = 0; n <= N - v - u; ++n) Rx[n][0] = Ry[n][u];
      for (int t = 1; t <= N - v - u; ++t)
        for (int n = 0; n <= N - v - u - t; ++n)
          Rx[n][t] = (t > 1 ? (t - 1) * Rx[n + 1][t - 2] : 0.0) + PC[0] * Rx[n + 1][t - 1];
      for (int t = 0; t <= N - v - u; ++t) R[tet(t, u, v)] = Rx[0][t];
    }
  }
}

// One shell pair (a, b), a >= b.  ea/ca, eb/cb: exponents and normalization-embedded coefficients of the
// two shells; wD / wW: the pair's Cartesian density blocks [ncart(la)][ncart(lb)] -- D_ab (+ D_ba^T when the
// shells differ) for the T + V term, the same of the energy-weighted density W for the overlap term;
// charges = natom x {Z, x, y, z}.  acc.add(0, 3 * atom + xyz, v) accumulates F1, acc.add(1, ...) F_Pulay.
template <class Acc>
LB200_OB_HD void pair_forces(int la, int lb, const double* A, const double* B, int npa, const double* ea,
                             const double* ca, int npb, const double* eb, const double* cb, const double* wD,
                             const double* wW, int atomA, int atomB, int natom, const double* charges, Acc& acc) {
  const int nb = (lb + 1) * (lb + 2) / 2;
  double wmax = 0.0;
  for (int i = 0; i < (la + 1) * (la + 2) / 2 * nb; ++i) wmax = fmax(wmax, fmax(fabs(wD[i]), fabs(wW[i])));
  if (wmax == 0.0) return;
  double AB2 = 0.0;
  for (int k = 0; k < 3; ++k) AB2 += (A[k] - B[k]) * (A[k] - B[k]);
  const int N = la + lb + 1;
  const Tet tet(N);
  const int ntet = tet.size();
  double E[3][kNI][kNJ][kNT];
  double Lam[6][kNTet];            // A x, A y, A z, B x, B y, B z
  double R[kNTet];
  double gS[3] = {0, 0, 0}, gT[3] = {0, 0, 0};      // d/dA of the overlap / kinetic terms, density-weighted
  double gVA[3] = {0, 0, 0}, gVB[3] = {0, 0, 0};    // d/dA, d/dB of the nuclear-attraction term, all charges
  for (int p1 = 0; p1 < npa; ++p1)
    for (int p2 = 0; p2 < npb; ++p2) {
      const double al = ea[p1], be = eb[p2];
      const double g = al + be, o2p = 0.5 / g;
      const double pref = ca[p1] * cb[p2] * exp(-al * be / g * AB2);
      if (fabs(pref) * wmax < 1e-22) continue;   // below 1e-10 of the parity tolerance with any E coefficient of an l <= 4 pair
      double P[3];
      for (int k = 0; k < 3; ++k) P[k] = (al * A[k] + be * B[k]) / g;
      for (int d = 0; d < 3; ++d) {
        const double PA = P[d] - A[d], PB = P[d] - B[d];
        for (int i = 0; i <= la + 1; ++i)
          for (int j = 0; j <= lb + 2; ++j)
            for (int t = 0; t < kNT; ++t) E[d][i][j][t] = 0.0;
        E[d][0][0][0] = 1.0;
        for (int i = 0; i <= la; ++i)
          for (int t = 0; t <= i + 1; ++t)
            E[d][i + 1][0][t] = (t > 0 ? o2p * E[d][i][0][t - 1] : 0.0) + PA * E[d][i][0][t] + (t + 1) * E[d][i][0][t + 1];
        for (int j = 0; j <= lb + 1; ++j)
          for (int i = 0; i <= la + 1; ++i)
            for (int t = 0; t <= i + j + 1; ++t)
              E[d][i][j + 1][t] = (t > 0 ? o2p * E[d][i][j][t - 1] : 0.0) + PB * E[d][i][j][t] + (t + 1) * E[d][i][j][t + 1];
      }
      const double s1 = sqrt(kPi / g);
      auto s = [&](int d, int i, int j) { return (i >= 0 && j >= 0) ? E[d][i][j][0] * s1 : 0.0; };
      auto k1 = [&](int d, int i, int j) {   // 1-d kinetic-energy integral, operator on the second function
        return i < 0 ? 0.0
                     : -2.0 * be * be * s(d, i, j + 2) + be * (2 * j + 1) * s(d, i, j) - 0.5 * j * (j - 1) * s(d, i, j - 2);
      };
      for (int k = 0; k < 6; ++k)
        for (int i = 0; i < ntet; ++i) Lam[k][i] = 0.0;
      int ia = 0;
      for (int ax = la; ax >= 0; --ax)
        for (int ay = la - ax; ay >= 0; --ay, ++ia) {
          const int a3[3] = {ax, ay, la - ax - ay};
          int ib = 0;
          for (int bx = lb; bx >= 0; --bx)
            for (int by = lb - bx; by >= 0; --by, ++ib) {
              const int b3[3] = {bx, by, lb - bx - by};
              const double wd = pref * wD[ia * nb + ib], ww = pref * wW[ia * nb + ib];
              if (wd == 0.0 && ww == 0.0) continue;
              double sv[3], kv[3], sA[3], kA[3];
              for (int d = 0; d < 3; ++d) {
                const int i = a3[d], j = b3[d];
                sv[d] = s(d, i, j);
                kv[d] = k1(d, i, j);
                sA[d] = 2.0 * al * s(d, i + 1, j) - i * s(d, i - 1, j);
                kA[d] = 2.0 * al * k1(d, i + 1, j) - i * k1(d, i - 1, j);
              }
              gS[0] += ww * sA[0] * sv[1] * sv[2];
              gS[1] += ww * sv[0] * sA[1] * sv[2];
              gS[2] += ww * sv[0] * sv[1] * sA[2];
              gT[0] += wd * (kA[0] * sv[1] * sv[2] + sA[0] * kv[1] * sv[2] + sA[0] * sv[1] * kv[2]);
              gT[1] += wd * (kv[0] * sA[1] * sv[2] + sv[0] * kA[1] * sv[2] + sv[0] * sA[1] * kv[2]);
              gT[2] += wd * (kv[0] * sv[1] * sA[2] + sv[0] * kv[1] * sA[2] + sv[0] * sv[1] * kA[2]);
              if (wd == 0.0 || natom == 0) continue;
              // Hermite densities of the six derivatives: dimension d carries the derivative coefficients
              for (int side = 0; side < 2; ++side)
                for (int d = 0; d < 3; ++d) {
                  double c[3][kNT];
                  int n[3];
                  for (int e = 0; e < 3; ++e) {
                    const int i = a3[e], j = b3[e];
                    if (e != d) {
                      n[e] = i + j;
                      for (int t = 0; t <= n[e]; ++t) c[e][t] = E[e][i][j][t];
                    } else if (side == 0) {
                      n[e] = i + j + 1;
                      for (int t = 0; t <= n[e]; ++t)
                        c[e][t] = 2.0 * al * E[e][i + 1][j][t] - (i > 0 ? i * E[e][i - 1][j][t] : 0.0);
                    } else {
                      n[e] = i + j + 1;
                      for (int t = 0; t <= n[e]; ++t)
                        c[e][t] = 2.0 * be * E[e][i][j + 1][t] - (j > 0 ? j * E[e][i][j - 1][t] : 0.0);
                    }
                  }
                  double* L = Lam[3 * side + d];
                  for (int t = 0; t <= n[0]; ++t)
                    for (int u = 0; u <= n[1]; ++u) {
                      const double cu = wd * c[0][t] * c[1][u];
                      for (int v = 0; v <= n[2]; ++v) L[tet(t, u, v)] += cu * c[2][v];
                    }
                }
            }
        }
      for (int c = 0; c < natom; ++c) {
        const double Z = charges[4 * c];
        if (Z == 0.0) continue;
        const double PC[3] = {P[0] - charges[4 * c + 1], P[1] - charges[4 * c + 2], P[2] - charges[4 * c + 3]};
        hermite_R(N, g, PC, tet, R);
        const double fac = -Z * 2.0 * kPi / g;
        for (int d = 0; d < 3; ++d) {
          double va = 0.0, vb = 0.0;
          for (int i = 0; i < ntet; ++i) {
            va += Lam[d][i] * R[i];
            vb += Lam[3 + d][i] * R[i];
          }
          va *= fac;
          vb *= fac;
          gVA[d] += va;
          gVB[d] += vb;
          acc.add(0, 3 * c + d, -2.0 * (va + vb));   // the operator's own centre
        }
      }
    }
  for (int d = 0; d < 3; ++d) {
    acc.add(0, 3 * atomA + d, 2.0 * gVA[d]);
    acc.add(0, 3 * atomB + d, 2.0 * gVB[d]);
    if (atomA != atomB) {
      acc.add(0, 3 * atomA + d, 2.0 * gT[d]);
      acc.add(0, 3 * atomB + d, -2.0 * gT[d]);
      acc.add(1, 3 * atomA + d, -2.0 * gS[d]);
      acc.add(1, 3 * atomB + d, 2.0 * gS[d]);
    }
  }
}

}  // namespace ob1
}  // namespace lb200
