// Host-side helper: significant shell-pair list of the direct SCF driver,
// compute_shellpairs of tests/hartree-fock/hartree-fock++.cc:1305-1381: a pair (s1 >= s2) is
// kept if the shells share a centre or the Frobenius norm of their overlap block is
// >= threshold.  The overlap block is evaluated with the textbook Obara-Saika 1-d
// recursion (this O(N^2) set-up step is not part of the GPU hot path).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/libint_b200.h"
#include "cart.cuh"
#include "internal_host.h"
#include "sph_coefs.cuh"

namespace {

using namespace lb200;

template <int L>
void sph_rows(std::vector<std::vector<std::pair<int, double>>>& rows) {
  rows.assign(2 * L + 1, {});
  for (int k = 0; k < Sph<L>::nnz; ++k) rows[Sph<L>::e[k].m].push_back({Sph<L>::e[k].c, Sph<L>::e[k].v});
}
const std::vector<std::vector<std::pair<int, double>>>& sph(int l) {
  static std::vector<std::vector<std::vector<std::pair<int, double>>>> tab = [] {
    std::vector<std::vector<std::vector<std::pair<int, double>>>> t(7);
    sph_rows<0>(t[0]); sph_rows<1>(t[1]); sph_rows<2>(t[2]); sph_rows<3>(t[3]);
    sph_rows<4>(t[4]); sph_rows<5>(t[5]); sph_rows<6>(t[6]);
    return t;
  }();
  return tab[l];
}

double overlap_block_norm(const lb200_basis_view& bs, int s1, int s2) {
  const int l1 = bs.l[s1], l2 = bs.l[s2];
  const int n1 = nc(l1), n2 = nc(l2);
  std::vector<double> S((size_t)n1 * n2, 0.0);
  const double* A = bs.O + 3 * s1;
  const double* B = bs.O + 3 * s2;
  double AB2 = 0, AB[3];
  for (int k = 0; k < 3; ++k) { AB[k] = A[k] - B[k]; AB2 += AB[k] * AB[k]; }
  double I[3][12][12];
  for (int p1 = 0; p1 < bs.nprim[s1]; ++p1)
    for (int p2 = 0; p2 < bs.nprim[s2]; ++p2) {
      const double a1 = bs.alpha[bs.off[s1] + p1], a2 = bs.alpha[bs.off[s2] + p2];
      const double c = bs.coeff[bs.off[s1] + p1] * bs.coeff[bs.off[s2] + p2];
      const double g = a1 + a2, oog = 1 / g, rho = a1 * a2 * oog;
      const double pref = c * std::exp(-rho * AB2) * std::pow(M_PI * oog, 1.5);
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
      for (int i = 0; i < n1; ++i) {
        const C3 q1 = cxyz(l1, i);
        for (int j = 0; j < n2; ++j) {
          const C3 q2 = cxyz(l2, j);
          S[(size_t)i * n2 + j] += pref * I[0][q1.x][q2.x] * I[1][q1.y][q2.y] * I[2][q1.z][q2.z];
        }
      }
    }
  // to pure where flagged
  std::vector<double> T;
  int m1 = n1, m2 = n2;
  if (bs.pure[s1]) {
    const auto& r = sph(l1);
    m1 = 2 * l1 + 1;
    T.assign((size_t)m1 * n2, 0.0);
    for (int m = 0; m < m1; ++m)
      for (auto& e : r[m])
        for (int j = 0; j < n2; ++j) T[(size_t)m * n2 + j] += e.second * S[(size_t)e.first * n2 + j];
    S.swap(T);
  }
  if (bs.pure[s2]) {
    const auto& r = sph(l2);
    m2 = 2 * l2 + 1;
    T.assign((size_t)m1 * m2, 0.0);
    for (int i = 0; i < m1; ++i)
      for (int m = 0; m < m2; ++m)
        for (auto& e : r[m]) T[(size_t)i * m2 + m] += e.second * S[(size_t)i * n2 + e.first];
    S.swap(T);
  }
  double nrm = 0;
  for (double v : S) nrm += v * v;
  return std::sqrt(nrm);
}

}  // namespace

extern "C" int lb200_significant_pairs(const lb200_basis* bs, double threshold, int* s1, int* s2,
                                       long long cap, long long* count) {
  if (!bs || !count) return LB200_ERR_INVALID;
  const lb200_basis_view v = lb200_view(bs);
  long long n = 0;
  for (int a = 0; a < v.nshell; ++a)
    for (int b = 0; b <= a; ++b) {
      const bool same = v.O[3 * a] == v.O[3 * b] && v.O[3 * a + 1] == v.O[3 * b + 1] &&
                        v.O[3 * a + 2] == v.O[3 * b + 2];
      bool sig = same;
      if (!same) {
        sig = overlap_block_norm(v, a, b) >= threshold;
      }
      if (sig) {
        if (n < cap && s1 && s2) { s1[n] = a; s2[n] = b; }
        ++n;
      }
    }
  *count = n;
  return (s1 && n > cap) ? LB200_ERR_NOMEM : LB200_OK;
}
