// oracle/truth.cc -- TEST INFRASTRUCTURE (parity arbiter), not product code.
//
// Extended-precision evaluation of contracted Cartesian Coulomb shell sets (ab|cd): the same
// Obara-Saika / Head-Gordon-Pople scheme the reference's generated kernels implement
// (src/bin/libint/vrr_11_twoprep_11.h:144-463, hrr.h:193-330, dg.cc:1128-1188; prerequisites
// as include/libint2/engine.impl.h:1310-1767, pair data as shell.h:1138-1256), carried out
// entirely in `long double` (64-bit mantissa, eps 1.1e-19) or `__float128` (113-bit, eps 1.9e-34)
// from the same double-precision shell data, with a Boys function that does not share the
// interpolation table of boys.h (convergent series + downward recursion / asymptotic form).
//
// Purpose: the GPU kernels and the CPU oracle (reference Engine + restated kernels) order their
// floating-point operations differently, and the HRR cancels large intermediates when |AB|, |CD|
// are large (tests/eri/test.cc:77-83 notes the loss for (dp|dd), (dd|dd)), so the two cannot agree
// to 1e-14 absolute on every element.  This file is the arbiter: both are compared with a result
// whose own rounding error is 1e3 (long double) to 1e18 (quad) times smaller.
//
// Output: hi = (double)x and lo = (double)(x - hi) for every integral, so that the comparison
// err = |(got - hi) - lo| is not limited by rounding the truth to double.
//
// Only tests/ and bench.py's parity leg may load the library built from this file.
#include <quadmath.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }
// STANDARD ordering (include/libint2/cgshell_ordering.h)
inline int cart_index(int l, int x, int y) { return ((l - x + 1) * (l - x)) / 2 + l - x - y; }

template <class R> R r_exp(R x);
template <class R> R r_sqrt(R x);
template <> inline long double r_exp(long double x) { return expl(x); }
template <> inline long double r_sqrt(long double x) { return sqrtl(x); }
template <> inline __float128 r_exp(__float128 x) { return expq(x); }
template <> inline __float128 r_sqrt(__float128 x) { return sqrtq(x); }
template <class R> R r_eps();
template <> inline long double r_eps() { return 1.0e-21L; }
template <> inline __float128 r_eps() { return 1.0e-36Q; }
template <class R> R r_pi();
template <> inline long double r_pi() { return 3.14159265358979323846264338327950288L; }
template <> inline __float128 r_pi() { return M_PIq; }

// Boys function F_m(T), m = 0..mmax.  T <= 120: F_mmax from the series
// exp(-T) sum_k (2T)^k / ((2m+1)(2m+3)...(2m+2k+1)) (all terms positive, no cancellation), then
// the stable downward recursion F_{m-1} = (2T F_m + exp(-T)) / (2m-1).  T > 120:
// F_0 = sqrt(pi/T)/2 (erfc(sqrt 120) ~ 1e-54) and the upward recursion, stable for T >> m.
template <class R>
void boys(R T, int mmax, R* F) {
  const R eT = r_exp<R>(-T);
  if (T > R(120)) {
    F[0] = r_sqrt<R>(r_pi<R>() / T) / R(2);
    for (int m = 0; m < mmax; ++m) F[m + 1] = (R(2 * m + 1) * F[m] - eT) / (R(2) * T);
    return;
  }
  R term = R(1) / R(2 * mmax + 1), sum = term;
  for (int k = 1; k < 2000; ++k) {
    term *= R(2) * T / R(2 * mmax + 2 * k + 1);
    sum += term;
    if (term < r_eps<R>() * sum) break;
  }
  F[mmax] = eT * sum;
  for (int m = mmax; m > 0; --m) F[m - 1] = (R(2) * T * F[m] + eT) / R(2 * m - 1);
}

struct Cart {
  std::vector<std::vector<int>> x, y, z;
  Cart() {
    x.resize(20); y.resize(20); z.resize(20);
    for (int l = 0; l < 20; ++l) {
      x[l].resize(ncart(l)); y[l].resize(ncart(l)); z[l].resize(ncart(l));
      for (int a = l; a >= 0; --a)
        for (int b = l - a; b >= 0; --b) {
          const int i = cart_index(l, a, b);
          x[l][i] = a; y[l][i] = b; z[l][i] = l - a - b;
        }
    }
  }
  int q(int l, int i, int d) const { return d == 0 ? x[l][i] : (d == 1 ? y[l][i] : z[l][i]); }
  int dir(int l, int i) const { return x[l][i] ? 0 : (y[l][i] ? 1 : 2); }   // OSVRR_xs_xs.h:74-78
  int shifted(int l, int i, int d, int s) const {
    int a = x[l][i] + (d == 0 ? s : 0), b = y[l][i] + (d == 1 ? s : 0);
    return cart_index(l + s, a, b);
  }
};
const Cart& cart() {
  static Cart c;
  return c;
}

// one contracted Cartesian shell set; shells: l[4], nprim[4], O[4][3], alpha/coeff per shell
// (coefficients carry the normalization, Shell::renorm, shell.h:958-999)
template <class R>
void shell_set(const int* l, const int* nprim, const double* const* O, const double* const* alpha,
               const double* const* coeff, R* out) {
  const Cart& ct = cart();
  const int la = l[0], lb = l[1], lc = l[2], ld = l[3];
  const int emax = la + lb, fmax = lc + ld, L = emax + fmax;
  R A[3], B[3], C[3], D[3], AB[3], CD[3], AB2 = 0, CD2 = 0;
  for (int k = 0; k < 3; ++k) {
    A[k] = O[0][k]; B[k] = O[1][k]; C[k] = O[2][k]; D[k] = O[3][k];
    AB[k] = A[k] - B[k]; CD[k] = C[k] - D[k];
    AB2 += AB[k] * AB[k]; CD2 += CD[k] * CD[k];
  }
  // V[e][f]: ncart(e) x ncart(f) x (L-e-f+1)
  std::vector<std::vector<std::vector<R>>> V(emax + 1, std::vector<std::vector<R>>(fmax + 1));
  for (int e = 0; e <= emax; ++e)
    for (int f = 0; f <= fmax; ++f) V[e][f].assign((size_t)ncart(e) * ncart(f) * (L - e - f + 1), R(0));
  auto nm = [&](int e, int f) { return L - e - f + 1; };
  std::vector<std::vector<std::vector<R>>> Ct(emax + 1, std::vector<std::vector<R>>(fmax + 1));
  for (int e = la; e <= emax; ++e)
    for (int f = lc; f <= fmax; ++f) Ct[e][f].assign((size_t)ncart(e) * ncart(f), R(0));
  std::vector<R> Fm(L + 1);
  const R sqrt2pi54 = r_sqrt<R>(R(2)) * r_sqrt<R>(r_pi<R>() * r_pi<R>() * r_sqrt<R>(r_pi<R>()));  // sqrt(2) pi^(5/4)
  for (int pa = 0; pa < nprim[0]; ++pa)
    for (int pb = 0; pb < nprim[1]; ++pb) {
      const R a1 = alpha[0][pa], a2 = alpha[1][pb];
      const R gp = a1 + a2, oogp = R(1) / gp;
      R P[3], PA[3];
      for (int k = 0; k < 3; ++k) { P[k] = (a1 * A[k] + a2 * B[k]) * oogp; PA[k] = P[k] - A[k]; }
      const R Kab = sqrt2pi54 * r_exp<R>(-a1 * a2 * oogp * AB2) * oogp * R(coeff[0][pa]) * R(coeff[1][pb]);
      for (int pc = 0; pc < nprim[2]; ++pc)
        for (int pd = 0; pd < nprim[3]; ++pd) {
          const R a3 = alpha[2][pc], a4 = alpha[3][pd];
          const R gq = a3 + a4, oogq = R(1) / gq;
          R Q[3], QC[3], W[3], WP[3], WQ[3], PQ2 = 0;
          for (int k = 0; k < 3; ++k) { Q[k] = (a3 * C[k] + a4 * D[k]) * oogq; QC[k] = Q[k] - C[k]; }
          const R Kcd = sqrt2pi54 * r_exp<R>(-a3 * a4 * oogq * CD2) * oogq * R(coeff[2][pc]) * R(coeff[3][pd]);
          const R gpq = gp + gq, oogpq = R(1) / gpq, rho = gp * gq * oogpq;
          for (int k = 0; k < 3; ++k) {
            W[k] = (gp * P[k] + gq * Q[k]) * oogpq;
            WP[k] = W[k] - P[k]; WQ[k] = W[k] - Q[k];
            PQ2 += (P[k] - Q[k]) * (P[k] - Q[k]);
          }
          const R pfac = Kab * Kcd * r_sqrt<R>(oogpq);
          boys<R>(PQ2 * rho, L, Fm.data());
          const R oo2z = R(0.5) * oogp, oo2e = R(0.5) * oogq, oo2ze = R(0.5) * oogpq;
          const R roz = rho * oogp, roe = rho * oogq;
          for (int m = 0; m <= L; ++m) V[0][0][m] = Fm[m] * pfac;
          // build on A (vrr_11_twoprep_11.h:154-222, c = 0)
          for (int e = 1; e <= emax; ++e)
            for (int ie = 0; ie < ncart(e); ++ie) {
              const int d = ct.dir(e, ie), im1 = ct.shifted(e, ie, d, -1);
              const int qd1 = ct.q(e, ie, d) - 1;
              const R* s1 = &V[e - 1][0][(size_t)im1 * nm(e - 1, 0)];
              R* t = &V[e][0][(size_t)ie * nm(e, 0)];
              for (int m = 0; m < nm(e, 0); ++m) t[m] = PA[d] * s1[m] + WP[d] * s1[m + 1];
              if (qd1 > 0) {
                const int im2 = ct.shifted(e - 1, im1, d, -1);
                const R* s2 = &V[e - 2][0][(size_t)im2 * nm(e - 2, 0)];
                for (int m = 0; m < nm(e, 0); ++m) t[m] += R(qd1) * oo2z * (s2[m] - roz * s2[m + 1]);
              }
            }
          // build on C (vrr_11_twoprep_11.h:305-383)
          for (int f = 1; f <= fmax; ++f)
            for (int e = 0; e <= emax; ++e)
              for (int ie = 0; ie < ncart(e); ++ie)
                for (int jf = 0; jf < ncart(f); ++jf) {
                  const int d = ct.dir(f, jf), jm1 = ct.shifted(f, jf, d, -1);
                  const int qd1 = ct.q(f, jf, d) - 1;
                  const int n = nm(e, f);
                  const R* s1 = &V[e][f - 1][((size_t)ie * ncart(f - 1) + jm1) * nm(e, f - 1)];
                  R* t = &V[e][f][((size_t)ie * ncart(f) + jf) * n];
                  for (int m = 0; m < n; ++m) t[m] = QC[d] * s1[m] + WQ[d] * s1[m + 1];
                  if (qd1 > 0) {
                    const int jm2 = ct.shifted(f - 1, jm1, d, -1);
                    const R* s2 = &V[e][f - 2][((size_t)ie * ncart(f - 2) + jm2) * nm(e, f - 2)];
                    for (int m = 0; m < n; ++m) t[m] += R(qd1) * oo2e * (s2[m] - roe * s2[m + 1]);
                  }
                  const int qe = ct.q(e, ie, d);
                  if (qe > 0) {
                    const int iem1 = ct.shifted(e, ie, d, -1);
                    const R* s4 = &V[e - 1][f - 1][((size_t)iem1 * ncart(f - 1) + jm1) * nm(e - 1, f - 1)];
                    for (int m = 0; m < n; ++m) t[m] += R(qe) * oo2ze * s4[m + 1];
                  }
                }
          for (int e = la; e <= emax; ++e)
            for (int f = lc; f <= fmax; ++f) {
              const int n = ncart(e) * ncart(f), nmm = nm(e, f);
              for (int i = 0; i < n; ++i) Ct[e][f][i] += V[e][f][(size_t)i * nmm];
            }
        }
    }
  // ket HRR (hrr.h:324): K[e] = (e0|lc ld) as [ie][ic][id]
  std::vector<std::vector<R>> K(emax + 1);
  for (int e = la; e <= emax; ++e) {
    const int ne = ncart(e);
    std::vector<std::vector<R>> cur(fmax - lc + 1);
    for (int c = lc; c <= fmax; ++c) cur[c - lc] = Ct[e][c];
    for (int dd = 1; dd <= ld; ++dd) {
      std::vector<std::vector<R>> nxt(fmax - dd - lc + 1);
      for (int c = lc; c <= fmax - dd; ++c) {
        auto& o = nxt[c - lc];
        o.resize((size_t)ne * ncart(c) * ncart(dd));
        const auto& lo = cur[c - lc];
        const auto& hi = cur[c + 1 - lc];
        for (int ie = 0; ie < ne; ++ie)
          for (int ic = 0; ic < ncart(c); ++ic)
            for (int id = 0; id < ncart(dd); ++id) {
              const int dir = ct.dir(dd, id);
              const int idm1 = ct.shifted(dd, id, dir, -1), icp1 = ct.shifted(c, ic, dir, +1);
              o[((size_t)ie * ncart(c) + ic) * ncart(dd) + id] =
                  hi[((size_t)ie * ncart(c + 1) + icp1) * ncart(dd - 1) + idm1] +
                  CD[dir] * lo[((size_t)ie * ncart(c) + ic) * ncart(dd - 1) + idm1];
            }
      }
      cur.swap(nxt);
    }
    K[e] = std::move(cur[0]);
  }
  // bra HRR (hrr.h:246)
  const int ncd = ncart(lc) * ncart(ld);
  std::vector<std::vector<R>> cur(emax - la + 1);
  for (int a = la; a <= emax; ++a) cur[a - la] = std::move(K[a]);
  for (int bb = 1; bb <= lb; ++bb) {
    std::vector<std::vector<R>> nxt(emax - bb - la + 1);
    for (int a = la; a <= emax - bb; ++a) {
      auto& o = nxt[a - la];
      o.resize((size_t)ncart(a) * ncart(bb) * ncd);
      const auto& lo = cur[a - la];
      const auto& hi = cur[a + 1 - la];
      for (int ia = 0; ia < ncart(a); ++ia)
        for (int ib = 0; ib < ncart(bb); ++ib) {
          const int dir = ct.dir(bb, ib);
          const int ibm1 = ct.shifted(bb, ib, dir, -1), iap1 = ct.shifted(a, ia, dir, +1);
          const R* h = &hi[((size_t)iap1 * ncart(bb - 1) + ibm1) * ncd];
          const R* lw = &lo[((size_t)ia * ncart(bb - 1) + ibm1) * ncd];
          R* oo = &o[((size_t)ia * ncart(bb) + ib) * ncd];
          for (int k = 0; k < ncd; ++k) oo[k] = h[k] + AB[dir] * lw[k];
        }
    }
    cur.swap(nxt);
  }
  const size_t n = (size_t)ncart(la) * ncart(lb) * ncd;
  for (size_t i = 0; i < n; ++i) out[i] = cur[0][i];
}

template <class R>
void run_batch(int nshell, const int* l, const int* nprim, const double* O, const double* alpha,
               const double* coeff, long nq, const int* q4, int nthreads, double* hi, double* lo) {
  std::vector<long> off(nshell + 1, 0);
  for (int s = 0; s < nshell; ++s) off[s + 1] = off[s] + nprim[s];
  if (nq == 0) return;
  const int* q0 = q4;
  const long blk = (long)ncart(l[q0[0]]) * ncart(l[q0[1]]) * ncart(l[q0[2]]) * ncart(l[q0[3]]);
  auto work = [&](int tid, int nthr) {
    std::vector<R> buf(blk);
    for (long q = tid; q < nq; q += nthr) {
      const int* s = q4 + 4 * q;
      int ll[4], np[4];
      const double *Op[4], *al[4], *co[4];
      for (int k = 0; k < 4; ++k) {
        ll[k] = l[s[k]]; np[k] = nprim[s[k]];
        Op[k] = O + 3 * s[k]; al[k] = alpha + off[s[k]]; co[k] = coeff + off[s[k]];
      }
      shell_set<R>(ll, np, Op, al, co, buf.data());
      for (long i = 0; i < blk; ++i) {
        const double h = (double)buf[i];
        hi[q * blk + i] = h;
        if (lo) lo[q * blk + i] = (double)(buf[i] - R(h));
      }
    }
  };
  const int nthr = std::max(1, nthreads);
  std::vector<std::thread> th;
  for (int t = 1; t < nthr; ++t) th.emplace_back(work, t, nthr);
  work(0, nthr);
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

// All quartets of one call must belong to one class (same four l's): q4 = nq x 4 shell indices.
// quad = 0: long double, 1: __float128.  hi/lo: nq * ncart^4 doubles each (lo may be NULL).
int lbt_eri_batch(int quad, int nshell, const int* l, const int* nprim, const double* O,
                  const double* alpha, const double* coeff, long nq, const int* q4, int nthreads,
                  double* hi, double* lo) {
  if (quad)
    run_batch<__float128>(nshell, l, nprim, O, alpha, coeff, nq, q4, nthreads, hi, lo);
  else
    run_batch<long double>(nshell, l, nprim, O, alpha, coeff, nq, q4, nthreads, hi, lo);
  return 0;
}

// Boys function in extended precision (for pinning the arbiter itself)
void lbt_boys(int quad, double T, int mmax, double* out) {
  if (quad) {
    std::vector<__float128> F(mmax + 1);
    boys<__float128>(T, mmax, F.data());
    for (int m = 0; m <= mmax; ++m) out[m] = (double)F[m];
  } else {
    std::vector<long double> F(mmax + 1);
    boys<long double>(T, mmax, F.data());
    for (int m = 0; m <= mmax; ++m) out[m] = (double)F[m];
  }
}
}
