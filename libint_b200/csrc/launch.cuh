// Host-side launchers of one class (explicitly instantiated in gen/eri_inst_*.cu).
//
// A class (X|Y) (X = the pair class with the larger order key) is served by
// eri_rowreg_kernel (eri_rowreg.cuh) with rows = the smaller pair and the other pair unrolled in
// registers whenever X has l1+l2 <= 4; for the (fd|, (ff| bras the roles are exchanged.
#pragma once
#include "eri_rowreg_prim.cuh"

namespace lb200 {

constexpr int kSmemLimit = 227 * 1024;  // usable dynamic shared memory per CTA on sm_100
constexpr int kMaxDevices = 64;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute and the occupancy depends
// on the device: both are set / queried once per (kernel, current device), not once per process
template <class Kern>
cudaError_t kernel_ctas_per_sm(Kern kern, int threads, int smem, int* cache, int* out) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  if (cache[dev] == 0) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    int nb = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem);
    if (err != cudaSuccess) return err;
    cache[dev] = nb < 1 ? 1 : nb;
  }
  *out = cache[dev];
  return cudaSuccess;
}

template <int LA, int LB, int LC, int LD, int MODE>
cudaError_t launch_rowreg(const EriParams& p, const RowInfo* rows, int num_sms,
                          cudaStream_t stream) {
  using K = RR<LA, LB, LC, LD>;
  constexpr int SMEM = K::QPC * K::qsize(MODE == kModeFock) * 8;
  static_assert(SMEM <= kSmemLimit, "row-register kernel exceeds shared memory");
  auto kern = eri_rowreg_kernel<LA, LB, LC, LD, MODE>;
  // the shared-memory opt-in and the occupancy are per device: cache them per device ordinal
  static int cache[kMaxDevices] = {};
  int ctas_per_sm = 0;
  if (cudaError_t err = kernel_ctas_per_sm(kern, K::THREADS, SMEM, cache, &ctas_per_sm)) return err;
  long long grid = (long long)num_sms * ctas_per_sm;
  if (!p.ntasks_dev) {
    const long long need = ((long long)p.ntasks + K::QPC - 1) / K::QPC;
    if (need < grid) grid = need;
  }
  if (grid < 1) return cudaSuccess;
  kern<<<(unsigned)grid, K::THREADS, SMEM, stream>>>(p, rows);
  return cudaGetLastError();
}

// uncontracted pair blocks, store mode, more than one row per quartet: pipelined kernel
template <int LA, int LB, int LC, int LD, bool TR, bool FOCK = false>
cudaError_t launch_rowreg_prim_tr(const EriParams& p, const RowInfo* rows, int num_sms,
                                  cudaStream_t stream) {
  using K = RRP<LA, LB, LC, LD, TR, FOCK>;
  constexpr int SMEM = K::QPC * K::QSIZE * 8;
  static_assert(SMEM <= kSmemLimit, "pipelined row-register kernel exceeds shared memory");
  auto kern = eri_rowreg_prim_kernel<LA, LB, LC, LD, TR, FOCK>;
  // the shared-memory opt-in and the occupancy are per device: cache them per device ordinal
  static int cache[kMaxDevices] = {};
  int ctas_per_sm = 0;
  if (cudaError_t err = kernel_ctas_per_sm(kern, K::THREADS, SMEM, cache, &ctas_per_sm)) return err;
  long long grid = (long long)num_sms * ctas_per_sm;
  if (!p.ntasks_dev) {
    const long long need = ((long long)p.ntasks + K::QPC - 1) / K::QPC;
    if (need < grid) grid = need;
  }
  if (grid < 1) return cudaSuccess;
  kern<<<(unsigned)grid, K::THREADS, SMEM, stream>>>(p, rows);
  return cudaGetLastError();
}
template <int LA, int LB, int LC, int LD>
cudaError_t launch_rowreg_prim(const EriParams& p, const RowInfo* rows, int num_sms,
                               cudaStream_t stream) {
  if constexpr (LA == LC && LB == LD) {
    // a diagonal class is always launched in its natural orientation (launch_class)
    return launch_rowreg_prim_tr<LA, LB, LC, LD, false>(p, rows, num_sms, stream);
  } else {
    return p.transpose_out ? launch_rowreg_prim_tr<LA, LB, LC, LD, true>(p, rows, num_sms, stream)
                           : launch_rowreg_prim_tr<LA, LB, LC, LD, false>(p, rows, num_sms, stream);
  }
}

template <int LA, int LB, int LC, int LD, int MODE>
cudaError_t launch_class(const EriParams& p, const RowInfo* rows, int num_sms,
                         cudaStream_t stream) {
  if constexpr (MODE != kModeFock && LA == LC && LB == LD && LA + LB <= 4) {
    // both pairs of one class: either may be the row side.  Pick the orientation whose output
    // is in the caller's order, so that the bra-HRR lanes store straight to global memory
    // instead of going through the transposing copy-out.
    if (!p.transpose_out) {
      if constexpr (RR<LA, LB, LC, LD>::NEC > 1) {
        if (p.uncontracted) return launch_rowreg_prim<LA, LB, LC, LD>(p, rows, num_sms, stream);
      }
      return launch_rowreg<LA, LB, LC, LD, MODE>(p, rows, num_sms, stream);
    }
  }
  // Fock mode, (fd| / (ff| against an (ss| or (ps| pair: the small pair still provides the rows and the
  // big pair is unrolled (register pyramid of depth 5 / 6, partly in local memory).  With rows = the
  // big pair these classes ran 56 / 84 lanes per quartet through CTA-wide barriers at < 1 % of the
  // FP64 peak (profiles/r02a_ncu_fock_3210.txt, _3300.txt).
#ifndef LB200_ROWS_SMALL_MAXE
#define LB200_ROWS_SMALL_MAXE 1
#endif
  constexpr bool ROWS_SMALL = LA + LB <= 4 || (MODE == kModeFock && LC + LD <= LB200_ROWS_SMALL_MAXE);
  if constexpr (ROWS_SMALL) {
    // rows = (LC LD| (the smaller pair), (LA LB) unrolled: the kernel sees bra and ket swapped
    EriParams q = p;
    q.bra = p.ket;
    q.ket = p.bra;
    q.swap_tasks = p.swap_tasks ^ 1;
    q.transpose_out = p.transpose_out ^ 1;
    if constexpr (RR<LC, LD, LA, LB>::NEC > 1) {
      if (p.uncontracted) {
        // Fock mode: the pipelined kernel up to 35 rows per quartet.  Round 1 measured a gain only for the
        // (ps| row classes; with 64-byte records and the Boys recursion it now wins (a little) everywhere:
        // (H2O)_64 def2-TZVP 2.288 -> 2.245 s, cc-pVTZ 5.538 -> 5.387 s for MAXNEC 4 -> 35 (56 / 84: no change)
        if constexpr (MODE == kModeFock) {
#ifndef LB200_FOCK_PRIM_MAXNEC
#define LB200_FOCK_PRIM_MAXNEC 35
#endif
          if constexpr (RR<LC, LD, LA, LB>::NEC <= LB200_FOCK_PRIM_MAXNEC)
            return launch_rowreg_prim_tr<LC, LD, LA, LB, false, true>(q, rows, num_sms, stream);
        } else {
          return launch_rowreg_prim<LC, LD, LA, LB>(q, rows, num_sms, stream);
        }
      }
    }
    return launch_rowreg<LC, LD, LA, LB, MODE>(q, rows, num_sms, stream);
  } else {
    // (fd|, (ff| bras: rows = the big pair, the other pair unrolled (up to l1+l2 = 6; beyond 4
    // the register pyramid spills to local memory -- these classes are rare)
    if constexpr (MODE != kModeFock) {
      if (p.uncontracted) return launch_rowreg_prim<LA, LB, LC, LD>(p, rows, num_sms, stream);
    }
    return launch_rowreg<LA, LB, LC, LD, MODE>(p, rows, num_sms, stream);
  }
}

// one entry point per class; mode selects the instantiation
template <int LA, int LB, int LC, int LD>
cudaError_t launch_class_any(const EriParams& p, const RowInfo* rows, int mode, int num_sms,
                             cudaStream_t stream) {
  if (mode == kModeFock) return launch_class<LA, LB, LC, LD, kModeFock>(p, rows, num_sms, stream);
  if (mode == kModePrereq) return launch_class<LA, LB, LC, LD, kModePrereq>(p, rows, num_sms, stream);
  return launch_class<LA, LB, LC, LD, kModeStoreCart>(p, rows, num_sms, stream);
}

}  // namespace lb200
