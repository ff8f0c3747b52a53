// Compile-time Cartesian-shell index algebra shared by all kernels.
//
// Component order inside a Cartesian shell is libint's STANDARD ordering
// (reference: include/libint2/cgshell_ordering.h, section "STANDARD ordering":
// x exponent runs l..0, then y runs (l-x)..0; INT_CARTINDEX(l,x,y) =
// ((l-x+1)(l-x))/2 + l-x-y) so that results are laid out exactly as the
// reference's generated kernels lay them out.
#pragma once
#include <utility>

#ifdef __CUDACC__
#define LB_HD __host__ __device__
#else
#define LB_HD
#endif

namespace lb200 {

LB_HD constexpr int nc(int l) { return (l + 1) * (l + 2) / 2; }
// number of Cartesian components in shells 0..l  (0 for l < 0)
LB_HD constexpr int nc_upto(int l) { return l < 0 ? 0 : (l + 1) * (l + 2) * (l + 3) / 6; }
LB_HD constexpr int npure(int l) { return 2 * l + 1; }

struct C3 {
  int x, y, z;
};
LB_HD constexpr int cidx(int l, int x, int y) { return ((l - x + 1) * (l - x)) / 2 + l - x - y; }
LB_HD constexpr int cidx(C3 q) { return cidx(q.x + q.y + q.z, q.x, q.y); }
LB_HD constexpr C3 cxyz(int l, int i) {
  int ii = 0;
  while ((ii + 1) * (ii + 2) / 2 <= i) ++ii;
  const int k = i - ii * (ii + 1) / 2;
  return C3{l - ii, ii - k, k};
}
// direction along which a component is built / decremented: first of x,y,z with a
// nonzero quantum number (same rule as the reference's generic VRR,
// src/lib/libint/OSVRR_xs_xs.h:74-78)
LB_HD constexpr int cdir(C3 q) { return q.x ? 0 : (q.y ? 1 : 2); }
LB_HD constexpr int cget(C3 q, int d) { return d == 0 ? q.x : (d == 1 ? q.y : q.z); }
LB_HD constexpr C3 cadd(C3 q, int d, int s) {
  return C3{q.x + (d == 0 ? s : 0), q.y + (d == 1 ? s : 0), q.z + (d == 2 ? s : 0)};
}

// compile-time loop: f(std::integral_constant<int, 0>{}), ..., f(<N-1>)
template <int... I, class F>
LB_HD inline __attribute__((always_inline)) void static_for_impl(std::integer_sequence<int, I...>,
                                                                 F&& f) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
LB_HD inline __attribute__((always_inline)) void static_for(F&& f) {
  static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int cmin(int a, int b) { return a < b ? a : b; }

}  // namespace lb200
