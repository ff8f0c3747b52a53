// Class dispatch: (la lb|lc ld) in kernel orientation -> explicit instantiation (gen/*.cu).
#include "internal.h"
#include "launch.cuh"

namespace lb200 {

#define LB200_CLASS(a, b, c, d)                                                              \
  extern template cudaError_t launch_class_any<a, b, c, d>(const EriParams&, const RowInfo*, \
                                                           int, int, cudaStream_t);
#include "gen/dispatch_table.inc"
#undef LB200_CLASS

int order_key(int la, int lb) { return (la + lb) * 100 + la * 10 + lb; }

bool class_supported(int la, int lb, int lc, int ld) {
#define LB200_CLASS(a, b, c, d) \
  if (la == a && lb == b && lc == c && ld == d) return true;
#include "gen/dispatch_table.inc"
#undef LB200_CLASS
  return false;
}

cudaError_t launch_eri(int la, int lb, int lc, int ld, const EriParams& p, const RowInfo* rows,
                       int mode, int num_sms, cudaStream_t stream) {
#define LB200_CLASS(a, b, c, d)                 \
  if (la == a && lb == b && lc == c && ld == d) \
    return launch_class_any<a, b, c, d>(p, rows, mode, num_sms, stream);
#include "gen/dispatch_table.inc"
#undef LB200_CLASS
  return cudaErrorInvalidValue;
}

}  // namespace lb200
