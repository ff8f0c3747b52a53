// liblibint_b200_iface.so -- the reference's C plugin boundary on top of the B200 library.
//
// Exports exactly what a consumer of the generated libint library binds
// (src/bin/libint/iface.cc:114-185,302-418; used by include/libint2/engine.impl.h:623-635,
// 1898-1899 and tests/unit/c-api.c:45-170):
//   Libint_t (include/libint2/util/generated/libint2_types.h),
//   libint2_build_eri[la][lb][lc][ld], libint2_build_3eri[l][lc][ld], libint2_build_2eri[l1][l2],
//   libint2_build_default, libint2_static_init/cleanup,
//   libint2_{need_memory,init,cleanup}_{default,eri,3eri,2eri}.
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
void libint2_cleanup_3eri(Libint_t* e) { libint2_cleanup_default(e); }
void libint2_cleanup_2eri(Libint_t* e) { libint2_cleanup_default(e); }

}  // extern "C"
