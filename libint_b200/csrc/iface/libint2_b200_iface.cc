// liblibint_b200_iface.so -- the reference's C plugin boundary on top of the B200 library.
//
// Exports exactly what a consumer of the generated libint library binds
// (src/bin/libint/iface.cc:114-185,302-418; used by include/libint2/engine.impl.h:623-635,
// 1898-1899 and tests/unit/c-api.c:45-170):
//   Libint_t (include/libint2/util/generated/libint2_types.h),
//   libint2_build_eri[la][lb][lc][ld], libint2_build_3eri[l][lc][ld], libint2_build_2eri[l1][l2],
//   libint2_build_default, libint2_static_init/cleanup,
//   libint2_{need_memory,init,cleanup}_{default,eri,3eri,2eri},
//   and for first derivatives libint2_build_eri1[la][lb][lc][ld] + libint2_{need_memory,init,cleanup}_eri1
//   (LIBINT2_MAX_DERIV_ORDER 1: Engine(Operator::coulomb, max_nprim, max_l, 1), engine.impl.h:1704-1751).
// so that the reference's unmodified header-only libint2::Engine (and any C caller that fills
// Libint_t itself) links against this library instead of libint2.a.
//
// Each build function packs the caller's Libint_t[contrdepth] prerequisites, runs VRR +
// contraction + HRR of that class on the GPU (lb200_eri_prereq_batch: the production class
// kernels fed with caller-made prerequisites) and copies the contracted Cartesian shell set
// into the evaluator's host `stack`; targets[0] = stack (a borrowed pointer valid until the
// next build call on that evaluator, doc/progman/progman.tex:440-474).  One shell set per call
// means one PCIe round trip per call: this is the correctness drop-in; throughput lives behind
// the batched entry points of include/libint_b200.h.
//
// No CPU fallback: without a usable GPU the first build call aborts with a message (the
// reference interface has no error return; its preconditions are asserts).
#include <libint2/util/generated/libint2_iface.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "../../../include/libint_b200.h"

namespace {

inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }

std::mutex g_mutex;
std::vector<lb200_context*> g_contexts;   // one per calling thread, destroyed by static_cleanup
unsigned long g_generation = 0;

struct ThreadState {
  lb200_context* ctx = nullptr;
  unsigned long generation = 0;
  std::vector<double> recs;
};
thread_local ThreadState tls;

lb200_context* thread_context() {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (tls.ctx && tls.generation == g_generation) return tls.ctx;
  int device = 0;
  if (const char* e = std::getenv("LB200_DEVICE")) device = std::atoi(e);
  lb200_context* c = nullptr;
  const int rc = lb200_context_create(device, &c);
  if (rc != LB200_OK) {
    std::fprintf(stderr,
                 "libint2 (B200): no usable CUDA device %d (lb200_context_create = %d); this library has "
                 "no CPU fallback\n", device, rc);
    std::abort();
  }
  g_contexts.push_back(c);
  tls.ctx = c;
  tls.generation = g_generation;
  return c;
}

// (ss|ss)^(m) members are laid out contiguously, m ascending, VECLEN = 1
inline const double* fm_ptr(const Libint_t* p) {
  return &p->_aB_s___0__s___1___TwoPRep_s___0__s___1___Ab__up_0[0];
}

// ncenter 4: (la lb|lc ld); 3: (la s|lc ld) with the unit shell as bra2 (engine.impl.h:1836-1873);
// 2: (la s|lc s) (:1875-1886).  Members Engine::compute2 leaves unset for a braket (PA/AB for
// xs_*, QC/CD for xs_xs, WP / WQ for an s-only side, engine.impl.h:1514-1641) are taken as 0.
void build_any(const Libint_t* inteval, int la, int lb, int lc, int ld, int ncenter) {
  lb200_context* ctx = thread_context();
  const int n = inteval[0].contrdepth;
  const int L = la + lb + lc + ld;
  std::vector<double>& r = tls.recs;
  r.assign((size_t)LB200_PREREQ_DOUBLES * (n > 0 ? n : 1), 0.0);
  for (int i = 0; i < n; ++i) {
    const Libint_t* p = inteval + i;
    double* o = r.data() + (size_t)LB200_PREREQ_DOUBLES * i;
    const double* F = fm_ptr(p);
    for (int m = 0; m <= L; ++m) o[m] = F[m];
    double* g = o + 25;
    if (ncenter == 4) { g[0] = p->PA_x[0]; g[1] = p->PA_y[0]; g[2] = p->PA_z[0]; }
    if (ncenter != 2) { g[3] = p->QC_x[0]; g[4] = p->QC_y[0]; g[5] = p->QC_z[0]; }
    if (la + lb > 0) { g[6] = p->WP_x[0]; g[7] = p->WP_y[0]; g[8] = p->WP_z[0]; }
    if (lc + ld > 0) { g[9] = p->WQ_x[0]; g[10] = p->WQ_y[0]; g[11] = p->WQ_z[0]; }
    g[12] = p->oo2z[0]; g[13] = p->oo2e[0]; g[14] = p->oo2ze[0]; g[15] = p->roz[0]; g[16] = p->roe[0];
  }
  double geom[6] = {0, 0, 0, 0, 0, 0};
  if (ncenter == 4 && lb > 0) { geom[0] = inteval[0].AB_x[0]; geom[1] = inteval[0].AB_y[0]; geom[2] = inteval[0].AB_z[0]; }
  if (ncenter != 2 && ld > 0) { geom[3] = inteval[0].CD_x[0]; geom[4] = inteval[0].CD_y[0]; geom[5] = inteval[0].CD_z[0]; }
  const int off[2] = {0, n};
  const int rc = lb200_eri_prereq_batch(ctx, la, lb, lc, ld, 1, off, r.data(), geom, inteval[0].stack);
  if (rc != LB200_OK) {
    std::fprintf(stderr, "libint2 (B200): build (%d %d|%d %d) failed (%d): %s\n", la, lb, lc, ld, rc,
                 lb200_last_error(ctx));
    std::abort();
  }
  inteval[0].targets[0] = inteval[0].stack;
}

// ---- first derivatives: libint2_build_eri1[la][lb][lc][ld] -------------------------------------------------
// Twelve targets, index 3 * centre + xyz.  Same construction as lb200_eri_deriv1_batch (csrc/deriv.cu):
// d/dA_x (ab|cd) = (a+1_x b|cd)[2 alpha_a] - a_x (a-1_x b|cd) from six ordinary shell sets with one angular
// momentum shifted, the fourth centre by translational invariance.  The factor 2 alpha of the raised sets is what
// Engine::compute2 stores per primitive in two_alpha{0,1}_{bra,ket} (engine.impl.h:1739-1750); it scales that
// primitive's (ss|ss)^(m).  A pair whose second shell becomes the larger one is handed over swapped (PA -> PB,
// AB -> BA; the members exist, engine.impl.h:1525,1567).
struct Cart { int x, y, z; };
inline Cart cart_of(int l, int i) {
  int k = 0;
  for (int x = l; x >= 0; --x)
    for (int y = l - x; y >= 0; --y, ++k)
      if (k == i) return Cart{x, y, l - x - y};
  return Cart{0, 0, 0};
}
inline int cart_index(int x, int y, int z) {
  const int l = x + y + z;
  return ((l - x + 1) * (l - x)) / 2 + l - x - y;
}

void build_deriv_any(const Libint_t* inteval, int la, int lb, int lc, int ld) {
  lb200_context* ctx = thread_context();
  const int n = inteval[0].contrdepth;
  const int l[4] = {la, lb, lc, ld};
  const int nn[4] = {ncart(la), ncart(lb), ncart(lc), ncart(ld)};
  const size_t blk = (size_t)nn[0] * nn[1] * nn[2] * nn[3];
  static thread_local std::vector<double> sets[6];
  std::vector<double>& r = tls.recs;
  int stride[6][4];
  for (int k = 0; k < 6; ++k) {
    const int c = k / 2, sgn = (k & 1) ? -1 : +1;
    sets[k].clear();
    if (l[c] + sgn < 0) continue;
    int m[4] = {l[0], l[1], l[2], l[3]};
    m[c] += sgn;
    const int Lk = m[0] + m[1] + m[2] + m[3];
    const bool sw_bra = m[0] < m[1], sw_ket = m[2] < m[3];
    r.assign((size_t)LB200_PREREQ_DOUBLES * (n > 0 ? n : 1), 0.0);
    for (int i = 0; i < n; ++i) {
      const Libint_t* p = inteval + i;
      double* o = r.data() + (size_t)LB200_PREREQ_DOUBLES * i;
      const double* F = fm_ptr(p);
      double scale = 1.0;
      if (sgn > 0)
        scale = c == 0 ? p->two_alpha0_bra[0] : (c == 1 ? p->two_alpha0_ket[0] : p->two_alpha1_bra[0]);
      for (int q = 0; q <= Lk; ++q) o[q] = F[q] * scale;
      double* g = o + 25;
      if (sw_bra) { g[0] = p->PB_x[0]; g[1] = p->PB_y[0]; g[2] = p->PB_z[0]; }
      else { g[0] = p->PA_x[0]; g[1] = p->PA_y[0]; g[2] = p->PA_z[0]; }
      if (sw_ket) { g[3] = p->QD_x[0]; g[4] = p->QD_y[0]; g[5] = p->QD_z[0]; }
      else { g[3] = p->QC_x[0]; g[4] = p->QC_y[0]; g[5] = p->QC_z[0]; }
      g[6] = p->WP_x[0]; g[7] = p->WP_y[0]; g[8] = p->WP_z[0];
      g[9] = p->WQ_x[0]; g[10] = p->WQ_y[0]; g[11] = p->WQ_z[0];
      g[12] = p->oo2z[0]; g[13] = p->oo2e[0]; g[14] = p->oo2ze[0]; g[15] = p->roz[0]; g[16] = p->roe[0];
    }
    const double sb = sw_bra ? -1.0 : 1.0, sk = sw_ket ? -1.0 : 1.0;
    const double geom[6] = {sb * inteval[0].AB_x[0], sb * inteval[0].AB_y[0], sb * inteval[0].AB_z[0],
                            sk * inteval[0].CD_x[0], sk * inteval[0].CD_y[0], sk * inteval[0].CD_z[0]};
    const int pa = sw_bra ? m[1] : m[0], pb = sw_bra ? m[0] : m[1], pc = sw_ket ? m[3] : m[2], pd = sw_ket ? m[2] : m[3];
    const int mm[4] = {ncart(m[0]), ncart(m[1]), ncart(m[2]), ncart(m[3])};
    sets[k].assign((size_t)mm[0] * mm[1] * mm[2] * mm[3], 0.0);
    const int off[2] = {0, n};
    const int rc = lb200_eri_prereq_batch(ctx, pa, pb, pc, pd, 1, off, r.data(), geom, sets[k].data());
    if (rc != LB200_OK) {
      std::fprintf(stderr, "libint2 (B200): derivative build (%d %d|%d %d), shifted set %d failed (%d): %s\n", la, lb,
                   lc, ld, k, rc, lb200_last_error(ctx));
      std::abort();
    }
    // strides of the (a, b, c, d) component indices in the layout [pa][pb][pc][pd] just written
    const int i0 = sw_bra ? 1 : 0, i1 = sw_bra ? 0 : 1, i2 = sw_ket ? 3 : 2, i3 = sw_ket ? 2 : 3;
    stride[k][i3] = 1;
    stride[k][i2] = mm[i3];
    stride[k][i1] = mm[i2] * mm[i3];
    stride[k][i0] = mm[i1] * mm[i2] * mm[i3];
  }
  double* out = inteval[0].stack;
  for (int ia = 0, e = 0; ia < nn[0]; ++ia)
    for (int ib = 0; ib < nn[1]; ++ib)
      for (int ic = 0; ic < nn[2]; ++ic)
        for (int id = 0; id < nn[3]; ++id, ++e) {
          const int idx[4] = {ia, ib, ic, id};
          double dsum[3] = {0, 0, 0};
          for (int c = 0; c < 3; ++c) {
            const Cart q = cart_of(l[c], idx[c]);
            const int qv[3] = {q.x, q.y, q.z};
            size_t rest_p = 0, rest_m = 0;
            for (int x = 0; x < 4; ++x)
              if (x != c) {
                rest_p += (size_t)idx[x] * stride[2 * c][x];
                if (l[c] > 0) rest_m += (size_t)idx[x] * stride[2 * c + 1][x];
              }
            for (int d = 0; d < 3; ++d) {
              int up[3] = {q.x, q.y, q.z};
              ++up[d];
              double v = sets[2 * c][rest_p + (size_t)cart_index(up[0], up[1], up[2]) * stride[2 * c][c]];
              if (qv[d] > 0) {
                int dn[3] = {q.x, q.y, q.z};
                --dn[d];
                v -= qv[d] * sets[2 * c + 1][rest_m + (size_t)cart_index(dn[0], dn[1], dn[2]) * stride[2 * c + 1][c]];
              }
              out[(size_t)(3 * c + d) * blk + e] = v;
              dsum[d] += v;
            }
          }
          for (int d = 0; d < 3; ++d) out[(size_t)(9 + d) * blk + e] = -dsum[d];   // translational invariance
        }
  for (int s = 0; s < 12; ++s) inteval[0].targets[s] = out + (size_t)s * blk;
}

template <int la, int lb, int lc, int ld>
void build4d(const Libint_t* p) { build_deriv_any(p, la, lb, lc, ld); }

constexpr int N4 = LIBINT2_MAX_AM_eri + 1, N3 = LIBINT2_MAX_AM_3eri + 1, N2 = LIBINT2_MAX_AM_2eri + 1;

template <int la, int lb, int lc, int ld>
void build4(const Libint_t* p) { build_any(p, la, lb, lc, ld, 4); }
template <int l, int lc, int ld>
void build3(const Libint_t* p) { build_any(p, l, 0, lc, ld, 3); }
template <int l1, int l2>
void build2(const Libint_t* p) { build_any(p, l1, 0, l2, 0, 2); }

size_t need_memory(int max_am) {
  const size_t n = ncart(max_am);
  return n * n * n * n + 16;   // the contracted Cartesian target of the largest class
}

void init_eval(Libint_t* inteval, int max_am, void* buf) {   // iface.cc:357-393
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
void (*libint2_build_eri[N4][N4][N4][N4])(const Libint_t*);
void (*libint2_build_3eri[N3][N3][N3])(const Libint_t*);
void (*libint2_build_2eri[N2][N2])(const Libint_t*);
void (*libint2_build_eri1[LIBINT2_MAX_AM_eri1 + 1][LIBINT2_MAX_AM_eri1 + 1][LIBINT2_MAX_AM_eri1 + 1][LIBINT2_MAX_AM_eri1 + 1])(const Libint_t*);

}

namespace {

// canonical classes only (build_libint.cc:78-83: la >= lb, lc >= ld, la+lb <= lc+ld); an entry stays
// null when the GPU library has no kernel for the class (engine.impl.h:1898 asserts on it)
template <int I>
void fill4_one() {
  constexpr int la = I / (N4 * N4 * N4), lb = (I / (N4 * N4)) % N4, lc = (I / N4) % N4, ld = I % N4;
  if constexpr (la >= lb && lc >= ld && la + lb <= lc + ld && (la + lb + lc + ld) > 0) {
    if (lb200_eri_class_supported(la, lb, lc, ld)) libint2_build_eri[la][lb][lc][ld] = &build4<la, lb, lc, ld>;
  }
}
template <int I>
void fill3_one() {
  constexpr int l = I / (N3 * N3), lc = (I / N3) % N3, ld = I % N3;
  if constexpr (lc >= ld && (l + lc + ld) > 0) {
    if (lb200_eri_class_supported(l, 0, lc, ld)) libint2_build_3eri[l][lc][ld] = &build3<l, lc, ld>;
  }
}
template <int I>
void fill2_one() {
  constexpr int l1 = I / N2, l2 = I % N2;
  if constexpr ((l1 + l2) > 0) {
    if (lb200_eri_class_supported(l1, 0, l2, 0)) libint2_build_2eri[l1][l2] = &build2<l1, l2>;
  }
}
constexpr int N4D = LIBINT2_MAX_AM_eri1 + 1;
template <int I>
void fill4d_one() {
  constexpr int la = I / (N4D * N4D * N4D), lb = (I / (N4D * N4D)) % N4D, lc = (I / N4D) % N4D, ld = I % N4D;
  // same canonical rule; (ss|ss) included: its derivatives are (ps|ss)-type sets (tests/eri/test.cc:238-249)
  if constexpr (la >= lb && lc >= ld && la + lb <= lc + ld) {
    long long plan[30];
    if (lb200_eri_deriv1_plan(la, lb, lc, ld, plan) == LB200_OK) libint2_build_eri1[la][lb][lc][ld] = &build4d<la, lb, lc, ld>;
  }
}
template <int... I>
void fill4d(std::integer_sequence<int, I...>) { (fill4d_one<I>(), ...); }
template <int... I>
void fill4(std::integer_sequence<int, I...>) { (fill4_one<I>(), ...); }
template <int... I>
void fill3(std::integer_sequence<int, I...>) { (fill3_one<I>(), ...); }
template <int... I>
void fill2(std::integer_sequence<int, I...>) { (fill2_one<I>(), ...); }

}  // namespace

extern "C" {

void libint2_static_init() {
  std::memset(libint2_build_default, 0, sizeof(libint2_build_default));
  std::memset(libint2_build_eri, 0, sizeof(libint2_build_eri));
  std::memset(libint2_build_3eri, 0, sizeof(libint2_build_3eri));
  std::memset(libint2_build_2eri, 0, sizeof(libint2_build_2eri));
  fill4(std::make_integer_sequence<int, N4 * N4 * N4 * N4>{});
  fill3(std::make_integer_sequence<int, N3 * N3 * N3>{});
  fill2(std::make_integer_sequence<int, N2 * N2>{});
  std::memset(libint2_build_eri1, 0, sizeof(libint2_build_eri1));
  fill4d(std::make_integer_sequence<int, N4D * N4D * N4D * N4D>{});
}

void libint2_static_cleanup() {
  std::lock_guard<std::mutex> lock(g_mutex);
  for (lb200_context* c : g_contexts) lb200_context_destroy(c);
  g_contexts.clear();
  ++g_generation;   // thread-local handles of this generation are dead
}

size_t libint2_need_memory_default(int max_am) { return need_memory(max_am); }
size_t libint2_need_memory_eri(int max_am) { return need_memory(max_am); }
size_t libint2_need_memory_3eri(int max_am) { return need_memory(max_am); }
size_t libint2_need_memory_2eri(int max_am) { return need_memory(max_am); }
void libint2_init_default(Libint_t* e, int max_am, void* buf) { init_eval(e, max_am, buf); }
void libint2_init_eri(Libint_t* e, int max_am, void* buf) { init_eval(e, max_am, buf); }
void libint2_init_3eri(Libint_t* e, int max_am, void* buf) { init_eval(e, max_am, buf); }
void libint2_init_2eri(Libint_t* e, int max_am, void* buf) { init_eval(e, max_am, buf); }
void libint2_cleanup_default(Libint_t* e) {   // iface.cc:395-413
  std::free(e[0].stack);
  e[0].stack = nullptr;
  e[0].vstack = nullptr;
}
void libint2_cleanup_eri(Libint_t* e) { libint2_cleanup_default(e); }
size_t libint2_need_memory_eri1(int max_am) { return 12 * need_memory(max_am); }
void libint2_init_eri1(Libint_t* e, int max_am, void* buf) {
  double* stack = buf ? static_cast<double*>(buf)
                      : static_cast<double*>(std::malloc(libint2_need_memory_eri1(max_am) * sizeof(double)));
  e[0].stack = stack;
  e[0].vstack = stack;
  for (int s = 0; s < 12; ++s) e[0].targets[s] = nullptr;
  e[0].veclen = 1;
  e[0].contrdepth = 0;
}
void libint2_cleanup_eri1(Libint_t* e) { libint2_cleanup_default(e); }
void libint2_cleanup_3eri(Libint_t* e) { libint2_cleanup_default(e); }
void libint2_cleanup_2eri(Libint_t* e) { libint2_cleanup_default(e); }

}  // extern "C"
