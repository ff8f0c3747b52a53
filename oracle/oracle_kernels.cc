// oracle/oracle_kernels.cc -- TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// The reference repository ships the *generator* of the integral kernels, not
// the kernels (libint2_build_eri[la][lb][lc][ld] are emitted by build_libint,
// /root/reference/src/bin/libint/build_libint.cc:984-1169, which cannot be
// built in this image: needs Boost.MPL + GMP C++ headers).  This file is a
// plain run-time-loop CPU restatement of what those generated kernels compute,
// written against the reference's own specification of the recurrences:
//
//   * kernel driver (zero contracted block, loop over contrdepth primitives
//     accumulating the VRR output, then HRR):     src/bin/libint/dg.cc:1128-1188
//   * Obara-Saika VRR for (a0|c0)^(m):     src/bin/libint/vrr_11_twoprep_11.h:144-463
//                                          src/lib/libint/OSVRR_xs_xs.h:49-185
//   * HRR (a b| = (a+1 b-1| + AB (a b-1|:             src/bin/libint/hrr.h:193-330
//   * canonical class rule la>=lb, lc>=ld, la+lb<=lc+ld: build_libint.cc:78-83
//   * Cartesian component order (STANDARD):        include/libint2/cgshell_ordering.h
//   * output layout ((a*nb+b)*nc+c)*nd+d, borrowed pointer into stack:
//                                              doc/progman/progman.tex:440-474
//   * evaluator life-cycle functions:               src/bin/libint/iface.cc:302-418
//
// It exports exactly the C symbols the generated library would
// (libint2_build_eri / _3eri / _2eri tables, libint2_static_init/cleanup,
// libint2_{init,need_memory,cleanup}_<task>), so the reference's *unmodified*
// header-only libint2::Engine links against it; "reference Engine + these
// kernels" is the CPU oracle every CUDA parity test is checked against.  The
// restatement itself is pinned against the reference's independent closed-form
// evaluator eri() (src/bin/test_eri/eri.h:121-380) in tests/test_oracle.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the library built from this file.

#include <libint2/util/generated/libint2_iface.h>

#include <algorithm>
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

namespace {

constexpr int kMaxL = 2 * LIBINT2_MAX_AM;  // highest AM reached by VRR on one centre pair

inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }

// STANDARD ordering (cgshell_ordering.h): x exponent runs l..0, then y runs (l-x)..0
inline int cart_index(int l, int x, int y) { return ((l - x + 1) * (l - x)) / 2 + l - x - y; }

// Index algebra of the Cartesian components, tabulated once: the recurrences below look their
// operands up here instead of recomputing quantum numbers per element.
constexpr int kTabL = kMaxL + 2;                       // shells 0 .. kMaxL+1
constexpr int kTabN = (kTabL + 1) * (kTabL + 2) / 2;   // components of the largest tabulated shell
struct CartTable {
  // xyz[l][idx][0..2]
  std::vector<std::vector<std::vector<int>>> xyz;
  short q[kTabL + 1][kTabN][3];    // quantum numbers
  short dir[kTabL + 1][kTabN];     // build direction: first of x,y,z with a nonzero quantum number
  short dec[kTabL + 1][kTabN][3];  // index (in shell l-1) of the component with q[d]-1, -1 if q[d] == 0
  short inc[kTabL + 1][kTabN][3];  // index (in shell l+1) of the component with q[d]+1
  CartTable() {
    xyz.resize(kMaxL + 2);
    for (int l = 0; l <= kMaxL + 1; ++l) {
      xyz[l].resize(ncart(l));
      for (int x = l; x >= 0; --x)
        for (int y = l - x; y >= 0; --y) {
          const int z = l - x - y;
          const int i = cart_index(l, x, y);
          xyz[l][i] = {x, y, z};
          const int qq[3] = {x, y, z};
          for (int d = 0; d < 3; ++d) {
            q[l][i][d] = (short)qq[d];
            int m[3] = {x, y, z}, p[3] = {x, y, z};
            --m[d];
            ++p[d];
            dec[l][i][d] = qq[d] > 0 ? (short)cart_index(l - 1, m[0], m[1]) : (short)-1;
            inc[l][i][d] = (short)cart_index(l + 1, p[0], p[1]);
          }
          dir[l][i] = (short)(x ? 0 : (y ? 1 : 2));
        }
    }
  }
};
const CartTable& cart() {
  static CartTable t;
  return t;
}

// direction along which a Cartesian component is built/decremented: first of
// x,y,z with a nonzero quantum number (OSVRR_xs_xs.h:74-78)
inline int build_dir(const std::vector<int>& q) { return q[0] ? 0 : (q[1] ? 1 : 2); }

struct Scratch {
  // [e][f] -> block of ncart(e)*ncart(f)*(nm) doubles, element ((ie*nf+jf)*nm + m)
  std::vector<double> v;
  std::vector<double> contr;  // contracted (e0|f0), e in [la,la+lb], f in [lc,lc+ld]
  std::vector<double> hrr1, hrr2, ket;  // HRR arenas, grown on demand, never freed
};

inline const double* fm_ptr(const Libint_t* p) {
  // the (ss|ss)^(m) members are laid out contiguously, m ascending, VECLEN = 1
  return &p->_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_0[0];
}

// -----------------------------------------------------------------------------
// generic contracted (la lb|lc ld) shell set; unit_b / unit_d mark the unit
// shell s-functions of the 3- and 2-centre integrals (no PA / QC terms, no HRR)
// -----------------------------------------------------------------------------
void build_generic(const Libint_t* inteval, int la, int lb, int lc, int ld, bool unit_b,
                   bool unit_d) {
  const CartTable& ct = cart();
  const int emax = la + lb, fmax = lc + ld, L = emax + fmax;
  const int contrdepth = inteval[0].contrdepth;

  // offsets of the primitive VRR blocks
  int off[kMaxL + 1][kMaxL + 1];
  int nmv[kMaxL + 1][kMaxL + 1];
  int total = 0;
  for (int e = 0; e <= emax; ++e)
    for (int f = 0; f <= fmax; ++f) {
      nmv[e][f] = L - e - f + 1;
      off[e][f] = total;
      total += ncart(e) * ncart(f) * nmv[e][f];
    }
  // offsets of the contracted targets
  int coff[kMaxL + 1][kMaxL + 1];
  int ctotal = 0;
  for (int e = la; e <= emax; ++e)
    for (int f = lc; f <= fmax; ++f) {
      coff[e][f] = ctotal;
      ctotal += ncart(e) * ncart(f);
    }

  static thread_local Scratch s;
  if ((int)s.v.size() < total) s.v.resize(total);
  s.contr.assign(ctotal, 0.0);  // dg.cc:1128-1140 (zero out the contracted block)
  double* V = s.v.data();

  for (int p = 0; p < contrdepth; ++p) {  // dg.cc:1142-1188
    const Libint_t* pd = inteval + p;
    const double* Fm = fm_ptr(pd);
    double PA[3] = {0, 0, 0}, QC[3] = {0, 0, 0}, WP[3] = {0, 0, 0}, WQ[3] = {0, 0, 0};
    if (!unit_b && L > 0) { PA[0] = pd->PA_x[0]; PA[1] = pd->PA_y[0]; PA[2] = pd->PA_z[0]; }
    if (!unit_d && L > 0) { QC[0] = pd->QC_x[0]; QC[1] = pd->QC_y[0]; QC[2] = pd->QC_z[0]; }
    if (emax > 0) { WP[0] = pd->WP_x[0]; WP[1] = pd->WP_y[0]; WP[2] = pd->WP_z[0]; }
    if (fmax > 0) { WQ[0] = pd->WQ_x[0]; WQ[1] = pd->WQ_y[0]; WQ[2] = pd->WQ_z[0]; }
    const double oo2z = L > 0 ? pd->oo2z[0] : 0, oo2e = L > 0 ? pd->oo2e[0] : 0,
                 oo2ze = L > 0 ? pd->oo2ze[0] : 0, roz = L > 0 ? pd->roz[0] : 0,
                 roe = L > 0 ? pd->roe[0] : 0;

    // [00|00]^(m)
    for (int m = 0; m <= L; ++m) V[off[0][0] + m] = Fm[m];

    // build on A: [e+1 0|00]^(m), vrr_11_twoprep_11.h:154-222 with c = 0
    for (int e = 1; e <= emax; ++e) {
      const int nm = nmv[e][0];
      for (int ie = 0; ie < ncart(e); ++ie) {
        const int d = ct.dir[e][ie];
        const int qm1d = ct.q[e][ie][d] - 1;
        const int im1 = ct.dec[e][ie][d];
        const double* s1 = V + off[e - 1][0] + im1 * nmv[e - 1][0];
        double* t = V + off[e][0] + ie * nm;
        if (qm1d > 0) {
          const int im2 = ct.dec[e - 1][im1][d];
          const double* s2 = V + off[e - 2][0] + im2 * nmv[e - 2][0];
          const double fac = qm1d * oo2z;
          for (int m = 0; m < nm; ++m) {
            double val = WP[d] * s1[m + 1] + fac * (s2[m] - roz * s2[m + 1]);
            if (!unit_b) val += PA[d] * s1[m];
            t[m] = val;
          }
        } else {
          for (int m = 0; m < nm; ++m) {
            double val = WP[d] * s1[m + 1];
            if (!unit_b) val += PA[d] * s1[m];
            t[m] = val;
          }
        }
      }
    }

    // build on C: [e0|f+1 0]^(m), vrr_11_twoprep_11.h:305-383 (mirror image)
    for (int f = 1; f <= fmax; ++f) {
      const int nf = ncart(f), nfm1 = ncart(f - 1), nfm2 = f >= 2 ? ncart(f - 2) : 0;
      for (int e = 0; e <= emax; ++e) {
        const int nm = nmv[e][f];
        for (int ie = 0; ie < ncart(e); ++ie) {
          for (int jf = 0; jf < nf; ++jf) {
            const int d = ct.dir[f][jf];
            const int qm1d = ct.q[f][jf][d] - 1;
            const int jm1 = ct.dec[f][jf][d];
            const double* s1 = V + off[e][f - 1] + (ie * nfm1 + jm1) * nmv[e][f - 1];
            double* t = V + off[e][f] + (ie * nf + jf) * nm;
            for (int m = 0; m < nm; ++m) {
              double val = WQ[d] * s1[m + 1];
              if (!unit_d) val += QC[d] * s1[m];
              t[m] = val;
            }
            if (qm1d > 0) {
              const int jm2 = ct.dec[f - 1][jm1][d];
              const double* s2 = V + off[e][f - 2] + (ie * nfm2 + jm2) * nmv[e][f - 2];
              const double fac = qm1d * oo2e;
              for (int m = 0; m < nm; ++m) t[m] += fac * (s2[m] - roe * s2[m + 1]);
            }
            const int qed = ct.q[e][ie][d];
            if (qed > 0) {
              const int iem1 = ct.dec[e][ie][d];
              const double* s4 =
                  V + off[e - 1][f - 1] + (iem1 * nfm1 + jm1) * nmv[e - 1][f - 1];
              const double fac = qed * oo2ze;
              for (int m = 0; m < nm; ++m) t[m] += fac * s4[m + 1];
            }
          }
        }
      }
    }

    // accumulate primitive [e0|f0]^(0) into the contracted block
    for (int e = la; e <= emax; ++e)
      for (int f = lc; f <= fmax; ++f) {
        const int n = ncart(e) * ncart(f), nm = nmv[e][f];
        const double* src = V + off[e][f];
        double* dst = s.contr.data() + coff[e][f];
        for (int i = 0; i < n; ++i) dst[i] += src[i * nm];
      }
  }

  // ---------------- HRR, ket first: (e0|c d) from (e0|f0), hrr.h:324 -----------
  // after this step K[e] holds (e0|lc ld) as [ie][ic][id].  All intermediates live in two
  // thread-local arenas that are only ever grown (no allocation per call).
  const int nc = ncart(lc), nd = ncart(ld), na = ncart(la), nb = ncart(lb);
  const int ncd = nc * nd;
  size_t kreq = 0, koff[kMaxL + 1];
  for (int e = la; e <= emax; ++e) { koff[e] = kreq; kreq += (size_t)ncart(e) * ncd; }
  size_t levmax = 0;
  for (int e = la; e <= emax; ++e) {
    for (int dd = 0; dd <= ld; ++dd) {
      size_t n = 0;
      for (int c = lc; c <= fmax - dd; ++c) n += (size_t)ncart(e) * ncart(c) * ncart(dd);
      levmax = std::max(levmax, n);
    }
  }
  for (int bb = 0; bb <= lb; ++bb) {
    size_t n = 0;
    for (int a = la; a <= emax - bb; ++a) n += (size_t)ncart(a) * ncart(bb) * ncd;
    levmax = std::max(levmax, n);
  }
  if (s.hrr1.size() < levmax) s.hrr1.resize(levmax);
  if (s.hrr2.size() < levmax) s.hrr2.resize(levmax);
  if (s.ket.size() < kreq) s.ket.resize(kreq);
  {
    double CD[3] = {0, 0, 0};
    if (ld > 0) { CD[0] = inteval[0].CD_x[0]; CD[1] = inteval[0].CD_y[0]; CD[2] = inteval[0].CD_z[0]; }
    for (int e = la; e <= emax; ++e) {
      const int ne = ncart(e);
      // level dd: blocks c = lc .. fmax-dd, block c stored [ie][ic][id] at loff[c - lc]
      double* cur = s.hrr1.data();
      double* nxt = s.hrr2.data();
      size_t curoff[kMaxL + 2], nxtoff[kMaxL + 2];
      {
        size_t o = 0;
        for (int c = lc; c <= fmax; ++c) {
          curoff[c - lc] = o;
          std::memcpy(cur + o, s.contr.data() + coff[e][c], sizeof(double) * ne * ncart(c));
          o += (size_t)ne * ncart(c);
        }
      }
      for (int dd = 1; dd <= ld; ++dd) {
        const int ndd = ncart(dd), ndm1 = ncart(dd - 1);
        size_t o = 0;
        for (int c = lc; c <= fmax - dd; ++c) {
          const int ncc = ncart(c), ncp1 = ncart(c + 1);
          nxtoff[c - lc] = o;
          double* out = nxt + o;
          o += (size_t)ne * ncc * ndd;
          const double* lo = cur + curoff[c - lc];      // (e0| c,   dd-1)
          const double* hi = cur + curoff[c + 1 - lc];  // (e0| c+1, dd-1)
          for (int ie = 0; ie < ne; ++ie)
            for (int ic = 0; ic < ncc; ++ic) {
              for (int id = 0; id < ndd; ++id) {
                const int dir = ct.dir[dd][id];
                const int idm1 = ct.dec[dd][id][dir];
                const int icp1 = ct.inc[c][ic][dir];
                out[((size_t)ie * ncc + ic) * ndd + id] =
                    hi[((size_t)ie * ncp1 + icp1) * ndm1 + idm1] +
                    CD[dir] * lo[((size_t)ie * ncc + ic) * ndm1 + idm1];
              }
            }
        }
        std::swap(cur, nxt);
        for (int c = lc; c <= fmax - dd; ++c) curoff[c - lc] = nxtoff[c - lc];
      }
      std::memcpy(s.ket.data() + koff[e], cur + curoff[0], sizeof(double) * ne * ncd);
    }
  }

  // ---------------- HRR, bra: (a b|cd) from (e0|cd), hrr.h:246 -----------------
  const double* result = nullptr;
  {
    double AB[3] = {0, 0, 0};
    if (lb > 0) { AB[0] = inteval[0].AB_x[0]; AB[1] = inteval[0].AB_y[0]; AB[2] = inteval[0].AB_z[0]; }
    // level bb: blocks a = la .. emax-bb, block a stored [ia][ibcur][cd]
    const double* cur = s.ket.data();
    size_t curoff[kMaxL + 2], nxtoff[kMaxL + 2];
    for (int a = la; a <= emax; ++a) curoff[a - la] = koff[a];
    double* bufs[2] = {s.hrr1.data(), s.hrr2.data()};
    for (int bb = 1; bb <= lb; ++bb) {
      const int nbb = ncart(bb), nbm1 = ncart(bb - 1);
      double* nxt = bufs[bb & 1];
      size_t o = 0;
      for (int a = la; a <= emax - bb; ++a) {
        const int naa = ncart(a);
        nxtoff[a - la] = o;
        double* out = nxt + o;
        o += (size_t)naa * nbb * ncd;
        const double* lo = cur + curoff[a - la];
        const double* hi = cur + curoff[a + 1 - la];
        for (int ia = 0; ia < naa; ++ia) {
          for (int ib = 0; ib < nbb; ++ib) {
            const int dir = ct.dir[bb][ib];
            const int ibm1 = ct.dec[bb][ib][dir];
            const int iap1 = ct.inc[a][ia][dir];
            const double* h = hi + ((size_t)iap1 * nbm1 + ibm1) * ncd;
            const double* l = lo + ((size_t)ia * nbm1 + ibm1) * ncd;
            double* ov = out + ((size_t)ia * nbb + ib) * ncd;
            for (int k = 0; k < ncd; ++k) ov[k] = h[k] + AB[dir] * l[k];
          }
        }
      }
      cur = nxt;
      for (int a = la; a <= emax - bb; ++a) curoff[a - la] = nxtoff[a - la];
    }
    result = cur + curoff[0];
  }

  std::memcpy(inteval[0].stack, result, sizeof(double) * na * nb * ncd);
  inteval[0].targets[0] = inteval[0].stack;
}

template <int la, int lb, int lc, int ld>
void build4(const Libint_t* p) {
  build_generic(p, la, lb, lc, ld, false, false);
}
template <int l, int lc, int ld>
void build3(const Libint_t* p) {
  build_generic(p, l, 0, lc, ld, true, false);
}
template <int l1, int l2>
void build2(const Libint_t* p) {
  build_generic(p, l1, 0, l2, 0, true, true);
}

constexpr int N = LIBINT2_MAX_AM + 1;

template <int I>
void fill4_one() {
  constexpr int la = I / (N * N * N), lb = (I / (N * N)) % N, lc = (I / N) % N, ld = I % N;
  // canonical classes only (build_libint.cc:78-83); others stay null
  if (la >= lb && lc >= ld && la + lb <= lc + ld && (la + lb + lc + ld) > 0)
    libint2_build_eri[la][lb][lc][ld] = &build4<la, lb, lc, ld>;
}
template <int I>
void fill3_one() {
  constexpr int l = I / (N * N), lc = (I / N) % N, ld = I % N;
  // 3-centre: bra = (l s|, ket canonical lc >= ld (build_libint.cc:1212-1375)
  if (lc >= ld && (l + lc + ld) > 0) libint2_build_3eri[l][lc][ld] = &build3<l, lc, ld>;
}
template <int I>
void fill2_one() {
  constexpr int l1 = I / N, l2 = I % N;
  if ((l1 + l2) > 0) libint2_build_2eri[l1][l2] = &build2<l1, l2>;
}
template <int... I>
void fill4(std::integer_sequence<int, I...>) { (fill4_one<I>(), ...); }
template <int... I>
void fill3(std::integer_sequence<int, I...>) { (fill3_one<I>(), ...); }
template <int... I>
void fill2(std::integer_sequence<int, I...>) { (fill2_one<I>(), ...); }

size_t need_memory(int max_am) {
  // contracted Cartesian target of the largest class
  const size_t n = ncart(max_am);
  return n * n * n * n + 16;
}

void init_eval(Libint_t* inteval, int max_am, void* buf) {
  // iface.cc:302-418: buf == 0 => library allocates the stack
  double* stack = buf ? static_cast<double*>(buf)
                      : static_cast<double*>(std::malloc(need_memory(max_am) * sizeof(double)));
  inteval[0].stack = stack;
  inteval[0].vstack = stack;
  inteval[0].targets[0] = nullptr;
  inteval[0].veclen = 1;
  inteval[0].contrdepth = 0;
}

}  // namespace

extern "C" {

void (*libint2_build_default[LIBINT2_MAX_AM_default + 1][LIBINT2_MAX_AM_default + 1])(const Libint_t*);
void (*libint2_build_eri[N][N][N][N])(const Libint_t*);
void (*libint2_build_3eri[N][N][N])(const Libint_t*);
void (*libint2_build_2eri[N][N])(const Libint_t*);

void libint2_static_init() {
  std::memset(libint2_build_default, 0, sizeof(libint2_build_default));
  std::memset(libint2_build_eri, 0, sizeof(libint2_build_eri));
  std::memset(libint2_build_3eri, 0, sizeof(libint2_build_3eri));
  std::memset(libint2_build_2eri, 0, sizeof(libint2_build_2eri));
  fill4(std::make_integer_sequence<int, N * N * N * N>{});
  fill3(std::make_integer_sequence<int, N * N * N>{});
  fill2(std::make_integer_sequence<int, N * N>{});
}
void libint2_static_cleanup() {}

size_t libint2_need_memory_default(int max_am) { return need_memory(max_am); }
size_t libint2_need_memory_eri(int max_am) { return need_memory(max_am); }
size_t libint2_need_memory_3eri(int max_am) { return need_memory(max_am); }
size_t libint2_need_memory_2eri(int max_am) { return need_memory(max_am); }
void libint2_init_default(Libint_t* e, int max_am, void* buf) { init_eval(e, max_am, buf); }
void libint2_init_eri(Libint_t* e, int max_am, void* buf) { init_eval(e, max_am, buf); }
void libint2_init_3eri(Libint_t* e, int max_am, void* buf) { init_eval(e, max_am, buf); }
void libint2_init_2eri(Libint_t* e, int max_am, void* buf) { init_eval(e, max_am, buf); }
void libint2_cleanup_default(Libint_t* e) {
  std::free(e[0].stack);
  e[0].stack = nullptr;
}
void libint2_cleanup_eri(Libint_t* e) { libint2_cleanup_default(e); }
void libint2_cleanup_3eri(Libint_t* e) { libint2_cleanup_default(e); }
void libint2_cleanup_2eri(Libint_t* e) { libint2_cleanup_default(e); }
}
