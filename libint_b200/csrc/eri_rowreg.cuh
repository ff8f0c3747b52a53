// Row-parallel, register-resident Obara-Saika / Head-Gordon-Pople kernel for one class
// (LA LB|LC LD).
//
// What it replaces: the generated libint2_build_eri[..] kernel (spec: src/bin/libint/dg.cc:
// 1128-1188) and the per-primitive prerequisite set-up of Engine::compute2
// (include/libint2/engine.impl.h:1310-1767) for a batch of shell quartets; recurrences as in
// vrr_11_twoprep_11.h:154-222 (build on A), :305-383 (build on C) and hrr.h:246,:324.
//
// Layout.  The pair with the smaller Cartesian footprint is the "row" side (LA LB|: every
// (e 0| component with e <= LA+LB is one ROW and one lane owns one row of one quartet, so a
// CTA of THREADS lanes works on THREADS / NEC quartets at a time (NEC = 1 degenerates to a
// plain thread-per-quartet kernel).  Everything on the other side |LC LD) is unrolled at
// compile time and lives in the lane's registers:
//
//   per surviving primitive quartet
//     prerequisites                      every lane, registers
//     Boys F_m(T)*pfac                   lanes m of the quartet -> shared -> all lanes
//     [row 0|00]^(m), m <= LC+LD         private chain x..xy..yz..z in registers, no exchange
//     [row 0|f 0]^(m), f = 1..LC+LD      level by level in registers; only the cross term
//                                        [row-1_d 0|f-1_d 0]^(m+1) comes from another lane,
//                                        through one shared-memory read per value
//     accumulate (e0|f0), e>=LA, f>=LC   registers
//   ket HRR -> (row 0|c d)               registers, per row
//   transpose through shared memory, bra HRR -> (a b|c d) by lanes over (quartet, cd), registers
//   store, or cart->pure + J/K digestion (Fock mode)
//
// FP64 operands therefore come from registers; shared memory carries ~1.5 doubles per value
// instead of the ~7 of the all-shared team kernel (eri_kernel.cuh), which the FP64:LDS
// throughput ratio of the SM (64 DFMA/clk vs 16 doubles/clk) makes the difference between a
// shared-memory-bound and an FP64-bound kernel.
#pragma once
#include "eri_kernel.cuh"
#include "fock_digest.cuh"

namespace lb200 {

template <int LA, int LB, int LC, int LD>
struct RR {
  static_assert(LA >= LB && LC >= LD, "class must be canonical within each pair");
  static constexpr int EMAX = LA + LB, FMAX = LC + LD, L = EMAX + FMAX;
  static constexpr int NEC = nc_upto(EMAX);        // rows per quartet
  static constexpr int NECX = nc_upto(EMAX - 1);   // rows that feed cross terms (e < EMAX)
  static constexpr int ROW0 = nc_upto(LA - 1), NRT = NEC - ROW0, RTP = NRT | 1;
  static constexpr int F0 = nc_upto(LC - 1), NFT = nc_upto(FMAX) - F0;
  static constexpr int NA = nc(LA), NB = nc(LB), NC = nc(LC), ND = nc(LD);
  static constexpr int NAB = NA * NB, NCD = NC * ND, CS = NCD | 1;
  // cross-term slots: level f < FMAX, component j, m = 1..FMAX-f
  static constexpr int xbase(int f) {
    int s = 0;
    for (int g = 0; g < f; ++g) s += nc(g) * (FMAX - g);
    return s;
  }
  static constexpr int XSLOTS = xbase(FMAX);
  static constexpr int xslot(int f, int j, int m) { return xbase(f) + j * (FMAX - f) + (m - 1); }
  // per-quartet shared-memory region (doubles); the phases alias each other
  static constexpr int HDR = 8;                       // AB[3], CD[3], flags
  static constexpr int OFF_F = HDR;                   // Boys values
  static constexpr int OFF_X = OFF_F + ((L + 2) & ~1);
  static constexpr int PRIM_DOUBLES = OFF_X + (EMAX > 0 ? XSLOTS * NECX : 0);
  static constexpr int TB_DOUBLES = HDR + (LB > 0 ? NCD * RTP : 0);
  // Fock mode: final integrals [NAB][CS] at HDR, then a second buffer used first as the
  // row->column transpose buffer and afterwards by the cart->pure passes
  static constexpr int OFF_B2 = HDR + NAB * CS;
  // staged density blocks: written (cp.async) while the K loop runs, so clear of its area too
  static constexpr int OFF_D = cmax(OFF_B2 + cmax(NAB * NCD, LB > 0 ? NCD * RTP : 0), (PRIM_DOUBLES + 1) & ~1);
  static constexpr int FOCK_DOUBLES = OFF_D + fock_dblock_doubles<LA, LB, LC, LD>();
  static constexpr int STORE_DOUBLES = OFF_B2 + (LB > 0 ? NCD * RTP : 0);
  // Stride between the regions of consecutive quartets.  The lanes of a warp that belong to
  // different quartets touch the same offset of their regions at the same time, so the stride
  // decides the bank pattern: with stride = NEC (mod 16 eight-byte banks) a half-warp's
  // quartets tile the banks (row-indexed accesses) or land in distinct banks (broadcast
  // accesses); a stride that is a multiple of 16 would serialise every access QPG ways.
  // CTA-wide groups (NEC > 16): stride = 2 (mod 16 banks) measured best of the eight even residues on the
  // sweep ((dd|dd) 37.7 -> 37.2 ms, (dp|dd) 18.6 -> 18.1, (dp|dp) 11.8 -> 11.4 per 10^7 quartets)
#ifndef LB200_PAD_WIDE
#define LB200_PAD_WIDE 2
#endif
  static constexpr int pad_stride(int s) {
    s = (s + 1) & ~1;
    if (NEC > 1 && NEC <= 16) {
      const int want = (NEC + (NEC & 1)) % 16;
      while (s % 16 != want) s += 2;
    } else if (NEC > 16 && LB200_PAD_WIDE >= 0) {
      while (s % 16 != LB200_PAD_WIDE) s += 2;
    }
    return s;
  }
  static constexpr int qsize(bool fock) {
    int s = cmax(PRIM_DOUBLES, STORE_DOUBLES);
    if (fock) s = cmax(s, FOCK_DOUBLES);
    return pad_stride(s);
  }
  static constexpr int threads() {
    if (NEC <= 32) return 128;
    if (NEC == 35) return 128;   // 3 quartets, 105 lanes
    if (NEC == 56) return 128;   // 2 quartets, 112 lanes
    return 96;                   // NEC = 84: 1 quartet
  }
  static constexpr int THREADS = threads();
  // sync group: for NEC <= 16 the quartets of a warp never leave it, every barrier is a
  // __syncwarp and the warps of a CTA run independent rounds; larger NEC: the whole CTA
  static constexpr bool WL = NEC <= 16;
  static constexpr int GROUP = WL ? 32 : THREADS;
  static constexpr int QPG = GROUP / NEC;             // quartets per group round
  static constexpr int NG = THREADS / GROUP;
  // CTAs per SM the register allocation is bounded for (measured per pyramid depth: the F = 4
  // pyramid wants the full 255 registers, shallower ones gain more from a third / fourth CTA)
  // Thread-per-quartet classes (NEC = 1) are latency-bound in the Fock build (long scoreboard
  // ~23 cycles per issue for uncontracted (ps|ss)): more resident warps pay there.
#ifndef LB200_TQ_MINB_LO
#define LB200_TQ_MINB_LO 4
#endif
#ifndef LB200_TQ_MINB_MID
#define LB200_TQ_MINB_MID 3
#endif
  static constexpr int MINB = NEC == 1 ? (FMAX >= 4 ? 2 : (FMAX >= 2 ? LB200_TQ_MINB_MID : LB200_TQ_MINB_LO))
                                       : (FMAX >= 4 ? 2 : (FMAX >= 2 ? 3 : 4));
  static constexpr int QPC = NG * QPG;                // quartets in flight per CTA
};

// one level of the register pyramid: [row 0|f 0]^(m), component j, m = 0..FMAX-F
template <int FMAX, int F>
struct Lvl {
  static constexpr int NM = FMAX - F + 1;
  double v[nc(F) * NM];
};

struct RowMeta {   // per lane, fixed for the whole kernel
  int row, e;
  int rm[3];       // row index of the component with q[d]-1 (0 if q[d] == 0)
  double q[3];     // x,y,z quantum numbers of the row
};

// ACCUM: targets are added to the contraction accumulators (general kernel) or assigned
// (uncontracted kernel, eri_rowreg_prim.cuh)
template <class K, int F, bool ACCUM = true, class P1, class P2>
__device__ __forceinline__ void rr_build_level(Lvl<K::FMAX, F>& out, const P1& p1, const P2& p2,
                                               const double (&QC)[3], const double (&WQ)[3],
                                               const double (&koo2e)[6], double roe,
                                               const double (&ce)[3], double* __restrict__ Xq,
                                               const RowMeta& rmeta, double* __restrict__ acc) {
  constexpr int NM = K::FMAX - F + 1;
#ifndef LB200_X_XPREF
#define LB200_X_XPREF 1
#endif
  // Shared-memory latency in front of the last FMA of every value is the largest stall of the d-class kernels
  // (short scoreboard, profiles/r02_ncu_full_2222.txt).  LB200_X_XPREF = 1 (default): the cross terms of
  // component j+1 are loaded before component j is computed; 2: all cross terms of the level first; 0: at
  // use.  Measured per 2^20 (dd|dd) / (dp|dd) / (dp|dp) quartets: 3.89 / 1.91 / 1.21 ms (0) -> 3.81 / 1.83 /
  // 1.19 (1), 3.80 / 1.84 / 1.20 (2); whole sweep with the H-zeroing change below 125.5 -> 121.9 ms, Fock
  // build unchanged (profiles/r03_variants.txt).
  [[maybe_unused]] double xall[(LB200_X_XPREF == 2 && K::EMAX > 0) ? nc(F) * NM : 1];
  if constexpr (LB200_X_XPREF == 2 && K::EMAX > 0) {
    static_for<nc(F)>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      constexpr C3 q = cxyz(F, j);
      constexpr int d = cdir(q);
      constexpr int jm1 = cidx(cadd(q, d, -1));
      static_for<NM>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        xall[j * NM + m] = Xq[K::xslot(F - 1, jm1, m + 1) * K::NECX + rmeta.rm[d]];
      });
    });
  }
  [[maybe_unused]] double xcur[NM], xnxt[NM];
  if constexpr (LB200_X_XPREF == 1 && K::EMAX > 0) {
    constexpr C3 q0 = cxyz(F, 0);
    constexpr int d0 = cdir(q0);
    constexpr int j0m1 = cidx(cadd(q0, d0, -1));
    static_for<NM>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      xcur[m] = Xq[K::xslot(F - 1, j0m1, m + 1) * K::NECX + rmeta.rm[d0]];
    });
  }
  static_for<nc(F)>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr C3 q = cxyz(F, j);
    constexpr int d = cdir(q);
    constexpr int qd = cget(q, d);
    constexpr int jm1 = cidx(cadd(q, d, -1));
    if constexpr (LB200_X_XPREF == 1 && K::EMAX > 0 && j + 1 < nc(F)) {
      constexpr C3 qn = cxyz(F, j + 1);
      constexpr int dn = cdir(qn);
      constexpr int jn1 = cidx(cadd(qn, dn, -1));
      static_for<NM>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        xnxt[m] = Xq[K::xslot(F - 1, jn1, m + 1) * K::NECX + rmeta.rm[dn]];
      });
    }
    static_for<NM>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      double v = QC[d] * p1.v[jm1 * (NM + 1) + m] + WQ[d] * p1.v[jm1 * (NM + 1) + m + 1];
      if constexpr (qd > 1) {
        constexpr int jm2 = cidx(cadd(q, d, -2));
        v += koo2e[qd - 1] * (p2.v[jm2 * (NM + 2) + m] - roe * p2.v[jm2 * (NM + 2) + m + 1]);
      }
      if constexpr (K::EMAX > 0) {
        if constexpr (LB200_X_XPREF == 2) v += ce[d] * xall[j * NM + m];
        else if constexpr (LB200_X_XPREF == 1) v += ce[d] * xcur[m];
        else v += ce[d] * Xq[K::xslot(F - 1, jm1, m + 1) * K::NECX + rmeta.rm[d]];
      }
      out.v[j * NM + m] = v;
      if constexpr (F < K::FMAX && m >= 1 && K::EMAX > 0) {
        if (rmeta.row < K::NECX) Xq[K::xslot(F, j, m) * K::NECX + rmeta.row] = v;
      }
      if constexpr (m == 0 && F >= K::LCv) {
        if constexpr (ACCUM) acc[nc_upto(F - 1) - K::F0 + j] += v;
        else acc[nc_upto(F - 1) - K::F0 + j] = v;
      }
    });
    if constexpr (LB200_X_XPREF == 1 && K::EMAX > 0 && j + 1 < nc(F))
      static_for<NM>([&](auto mc) { xcur[decltype(mc)::value] = xnxt[decltype(mc)::value]; });
  });
}

// The same level with the group barrier INSIDE it: first everything that lives in this lane's registers (the
// two-term and four-term parts of every value, ~3/4 of the level's FP64 work), then the barrier that makes the
// cross terms of level F-1 visible, then the cross-term FMAs, the hand-over of this level's values and the targets.
// A warp that finished level F-1 early computes instead of waiting at the barrier.  (K::EMAX > 0 only.)
template <class K, int F, bool ACCUM, class P1, class P2, class SyncFn>
__device__ __forceinline__ void rr_build_level_split(Lvl<K::FMAX, F>& out, const P1& p1, const P2& p2,
                                                     const double (&QC)[3], const double (&WQ)[3],
                                                     const double (&koo2e)[6], double roe, const double (&ce)[3],
                                                     double* __restrict__ Xq, const RowMeta& rmeta,
                                                     double* __restrict__ acc, SyncFn&& sync) {
  constexpr int NM = K::FMAX - F + 1;
  static_for<nc(F)>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr C3 q = cxyz(F, j);
    constexpr int d = cdir(q);
    constexpr int qd = cget(q, d);
    constexpr int jm1 = cidx(cadd(q, d, -1));
    static_for<NM>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      double v = QC[d] * p1.v[jm1 * (NM + 1) + m] + WQ[d] * p1.v[jm1 * (NM + 1) + m + 1];
      if constexpr (qd > 1) {
        constexpr int jm2 = cidx(cadd(q, d, -2));
        v += koo2e[qd - 1] * (p2.v[jm2 * (NM + 2) + m] - roe * p2.v[jm2 * (NM + 2) + m + 1]);
      }
      out.v[j * NM + m] = v;
    });
  });
  sync();
  static_for<nc(F)>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr C3 q = cxyz(F, j);
    constexpr int d = cdir(q);
    constexpr int jm1 = cidx(cadd(q, d, -1));
    static_for<NM>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      const double v = out.v[j * NM + m] + ce[d] * Xq[K::xslot(F - 1, jm1, m + 1) * K::NECX + rmeta.rm[d]];
      out.v[j * NM + m] = v;
      if constexpr (F < K::FMAX && m >= 1) {
        if (rmeta.row < K::NECX) Xq[K::xslot(F, j, m) * K::NECX + rmeta.row] = v;
      }
      if constexpr (m == 0 && F >= K::LCv) {
        if constexpr (ACCUM) acc[nc_upto(F - 1) - K::F0 + j] += v;
        else acc[nc_upto(F - 1) - K::F0 + j] = v;
      }
    });
  });
}

template <int LA, int LB, int LC, int LD>
struct RRK : RR<LA, LB, LC, LD> {
  static constexpr int LCv = LC;
};

// HRR in registers for a pair (LX >= LY): in = level 0 blocks x in [LX, LX+LY] (nc(x) each),
// out = nc(LX) * nc(LY) values, index ix * nc(LY) + iy.  V = displacement vector (A-B / C-D).
template <int LX, int LY, int N0>
__device__ __forceinline__ void rr_hrr_regs(const double (&in0)[N0], const double (&V)[3],
                                            double (&out)[nc(LX) * nc(LY)]) {
  using H = Cls<LX, 0, 0, 0>;  // only for hoff/hsize helpers
  if constexpr (LY == 0) {
    static_for<nc(LX)>([&](auto ic) { out[decltype(ic)::value] = in0[decltype(ic)::value]; });
  } else {
    constexpr int S1 = H::hsize(LX, LY, 1);
    double a[S1];
    // level 1 from level 0
    static_for<LY>([&](auto xc) {  // x = LX .. LX+LY-1
      constexpr int x = LX + decltype(xc)::value;
      static_for<nc(x)>([&](auto ixc) {
        constexpr int ix = decltype(ixc)::value;
        static_for<3>([&](auto iyc) {
          constexpr int iy = decltype(iyc)::value;  // p component: x,y,z <-> d = iy
          constexpr int d = iy;
          constexpr int ixp1 = cidx(cadd(cxyz(x, ix), d, +1));
          constexpr int o = H::hoff(LX, x, 1) + ix * 3 + iy;
          constexpr int hi = H::hoff(LX, x + 1, 0) + ixp1;
          constexpr int lo = H::hoff(LX, x, 0) + ix;
          a[o] = in0[hi] + V[d] * in0[lo];
        });
      });
    });
    if constexpr (LY == 1) {
      static_for<nc(LX) * 3>([&](auto ic) { out[decltype(ic)::value] = a[decltype(ic)::value]; });
    } else {
      constexpr int S2 = H::hsize(LX, LY, 2);
      double b[S2];
      static_for<LY - 1>([&](auto xc) {
        constexpr int x = LX + decltype(xc)::value;
        static_for<nc(x)>([&](auto ixc) {
          constexpr int ix = decltype(ixc)::value;
          static_for<6>([&](auto iyc) {
            constexpr int iy = decltype(iyc)::value;
            constexpr C3 qy = cxyz(2, iy);
            constexpr int d = cdir(qy);
            constexpr int iym1 = cidx(cadd(qy, d, -1));
            constexpr int ixp1 = cidx(cadd(cxyz(x, ix), d, +1));
            constexpr int o = H::hoff(LX, x, 2) + ix * 6 + iy;
            constexpr int hi = H::hoff(LX, x + 1, 1) + ixp1 * 3 + iym1;
            constexpr int lo = H::hoff(LX, x, 1) + ix * 3 + iym1;
            b[o] = a[hi] + V[d] * a[lo];
          });
        });
      });
      if constexpr (LY == 2) {
        static_for<nc(LX) * 6>([&](auto ic) { out[decltype(ic)::value] = b[decltype(ic)::value]; });
      } else {
        static_assert(LY == 3, "HRR in registers is written for LY <= 3");
        static_for<nc(LX)>([&](auto ixc) {
          constexpr int ix = decltype(ixc)::value;
          constexpr int x = LX;
          static_for<10>([&](auto iyc) {
            constexpr int iy = decltype(iyc)::value;
            constexpr C3 qy = cxyz(3, iy);
            constexpr int d = cdir(qy);
            constexpr int iym1 = cidx(cadd(qy, d, -1));
            constexpr int ixp1 = cidx(cadd(cxyz(x, ix), d, +1));
            constexpr int hi = H::hoff(LX, x + 1, 2) + ixp1 * 6 + iym1;
            constexpr int lo = H::hoff(LX, x, 2) + ix * 6 + iym1;
            out[ix * 10 + iy] = b[hi] + V[d] * b[lo];
          });
        });
      }
    }
  }
}

// profiling only: one atomic per warp into one of 64 counters (a single hot address would serialise
// billions of atomics and distort the very timings being profiled); every lane of the warp calls it
__device__ __forceinline__ void count_primitives(unsigned long long* counters, int n) {
  const unsigned tot = __reduce_add_sync(0xffffffffu, (unsigned)n);
  if ((threadIdx.x & 31) == 0 && tot)
    atomicAdd(counters + ((blockIdx.x + (threadIdx.x >> 5)) & (kPrimCounters - 1)), (unsigned long long)tot);
}

__device__ __forceinline__ double sel3(int d, double x, double y, double z) {
  return d == 0 ? x : (d == 1 ? y : z);
}

template <int LA, int LB, int LC, int LD, int MODE>
__global__ void __launch_bounds__(RR<LA, LB, LC, LD>::THREADS, RR<LA, LB, LC, LD>::MINB)
eri_rowreg_kernel(const EriParams p, const RowInfo* __restrict__ rows) {
  using K = RRK<LA, LB, LC, LD>;
  constexpr bool FOCK = (MODE == kModeFock);
  constexpr bool PREREQ = (MODE == kModePrereq);
  constexpr int EMAX = K::EMAX, FMAX = K::FMAX, L = K::L, NEC = K::NEC, NECX = K::NECX;
  constexpr int QSIZE = K::qsize(FOCK);
  constexpr bool WL = K::WL;
  constexpr int GROUP = K::GROUP, QPG = K::QPG, NG = K::NG;
  static_assert(FMAX <= 6, "register pyramid is written for LC+LD <= 6");

  extern __shared__ double smem[];
  __shared__ int s_maxit[3];  // rotating: slot r%3 is reduced in round r, slot (r+1)%3 re-zeroed

  const int tid = threadIdx.x;
  const int g = WL ? tid >> 5 : 0;    // sync group of this lane
  const int gl = WL ? tid & 31 : tid; // lane within the group
  const int qg = gl / NEC;            // quartet slot within the group
  const bool lane_on = qg < QPG;      // leftover lanes only take part in barriers and phase 2
  const int q = g * QPG + qg;         // CTA-wide quartet slot (shared-memory region)
  auto sync = [] {
    if constexpr (WL) __syncwarp(); else __syncthreads();
  };
  RowMeta rmeta;
  // leftover lanes get a row index past every "row < NECX" / "row >= ROW0" guarded write
  // (NEC <= 84 < kMaxRows: still a valid RowInfo index)
  rmeta.row = lane_on ? gl - qg * NEC : NEC;
  {
    const RowInfo ri = rows[rmeta.row];
    rmeta.e = ri.e;
    for (int d = 0; d < 3; ++d) {
      rmeta.rm[d] = lane_on ? ri.rm[d] : 0;   // leftover lanes read (and discard) row 0's slots
      rmeta.q[d] = (double)ri.q[d];
    }
  }
  // private chain of build steps for [row 0|00]: slots EMAX-e .. EMAX-1 are real steps
  int cdir_[EMAX > 0 ? EMAX : 1];
  double ccnt_[EMAX > 0 ? EMAX : 1];
  if constexpr (EMAX > 0) {
    const int qx = (int)rmeta.q[0], qy = (int)rmeta.q[1];
    static_for<EMAX>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      const int s = k - (EMAX - rmeta.e);  // index of the real step, < 0: no-op
      int d = -1, c = 0;
      if (s >= 0) {
        if (s < qx) { d = 0; c = s; }
        else if (s < qx + qy) { d = 1; c = s - qx; }
        else { d = 2; c = s - qx - qy; }
      }
      cdir_[k] = d;
      ccnt_[k] = (double)c;
    });
  }
  double* const Q = smem + (size_t)(lane_on ? q : g * QPG) * QSIZE;  // this lane's quartet region

  const unsigned ntasks = p.ntasks_dev ? *p.ntasks_dev : p.ntasks;
  if constexpr (!WL) {
    if (tid < 3) s_maxit[tid] = 0;
    __syncthreads();
  }
  int round = 0;
  // Fock mode: groups pull their next QPG tasks from a device counter (zeroed by the host before
  // the launch).  The quartets of one launch differ by orders of magnitude in primitive count,
  // so a static round-robin leaves SMs idle while a few groups finish the heavy ones.
  // (Splitting one task's primitive loop over several slots was measured too: slower, the
  // per-task HRR + digestion then runs on a fraction of the lanes.)
  __shared__ unsigned s_base[2];
  auto next_base = [&](int r) -> unsigned {
    if constexpr (!FOCK) {
      return (unsigned)((blockIdx.x * NG + g) * QPG) + (unsigned)r * (gridDim.x * NG * QPG);
    } else if constexpr (WL) {
      unsigned b = 0;
      if (gl == 0) b = atomicAdd(p.work_counter, (unsigned)QPG);
      return __shfl_sync(0xffffffffu, b, 0);
    } else {
      if (tid == 0) s_base[r & 1] = atomicAdd(p.work_counter, (unsigned)QPG);
      __syncthreads();
      return s_base[r & 1];
    }
  };
  for (unsigned base = next_base(0); base < ntasks; base = next_base(round)) {
    const unsigned task = base + qg;
    const bool valid = lane_on && task < ntasks;
    int ib = 0, ik = 0, pb0 = 0, nb = 0, pk0 = 0, nk = 0;
    double ln_prec = p.ln_precision, prec = p.precision, deg = 1.0;
    if (valid) {
      int2 tk = make_int2(0, 0);
      if constexpr (PREREQ) {
      } else if constexpr (FOCK) {   // precision and degeneracy come with the task (screen_kernel)
        const int4 ft = p.ftasks[task];
        tk = make_int2(ft.x, ft.y & 0x3fffffff);
        deg = (double)(1 << ((unsigned)ft.y >> 30));
        ln_prec = __hiloint2double(ft.w, ft.z);
      } else {
        tk = p.prod_nk ? make_int2(p.prod_b0 + (int)(task / p.prod_nk), p.prod_k0 + (int)(task % p.prod_nk))
                       : p.tasks[task];
      }
      ib = p.swap_tasks ? tk.y : tk.x;
      ik = p.swap_tasks ? tk.x : tk.y;
      if constexpr (PREREQ) {   // one run of caller-made records per task; no pair blocks
        ib = ik = (int)task;
        pb0 = p.prereq_off[task];
        nb = p.prereq_off[task + 1] - pb0;
        nk = 1;
      } else {
        pb0 = p.bra.prim_off[ib];
        nb = p.bra.prim_off[ib + 1] - pb0;
        pk0 = p.ket.prim_off[ik];
        nk = p.ket.prim_off[ik + 1] - pk0;
      }
    }
    const int nit = nb * nk;
    if constexpr (!WL) {
      // slot (round+1)%3 was last read in round-2, i.e. before the barrier of round-1
      if (tid == 0) s_maxit[(round + 1) % 3] = 0;
      if (valid && rmeta.row == 0 && nit > 0) atomicMax(&s_maxit[round % 3], nit);
    }

    double CD[3] = {0, 0, 0}, Ac[3] = {0, 0, 0}, Cc[3] = {0, 0, 0};   // C - D, centres A and C
    if (valid) {
      if constexpr (PREREQ) {   // swap_tasks: this kernel's ket is the caller's bra
        const double* gv = p.prereq_geom + 6 * (size_t)task + (p.swap_tasks ? 0 : 3);
        CD[0] = gv[0]; CD[1] = gv[1]; CD[2] = gv[2];
      } else {
        const PairGeom& gk = p.ket.geom[ik];
        CD[0] = gk.AB[0]; CD[1] = gk.AB[1]; CD[2] = gk.AB[2];
        Cc[0] = gk.A[0]; Cc[1] = gk.A[1]; Cc[2] = gk.A[2];
        if constexpr (EMAX > 0) {
          const PairGeom& gb = p.bra.geom[ib];
          Ac[0] = gb.A[0]; Ac[1] = gb.A[1]; Ac[2] = gb.A[2];
        }
      }
    }
    const bool bra_unit = p.bra.unit_b != 0, ket_unit = p.ket.unit_b != 0;
    int bf4[4] = {0, 0, 0, 0};   // first basis functions of the four shells (Fock mode)
    if constexpr (FOCK) {
      if (valid) {
        const int2 ab = *reinterpret_cast<const int2*>(p.bra.geom[ib].bf);
        const int2 cd = *reinterpret_cast<const int2*>(p.ket.geom[ik].bf);
        bf4[0] = ab.x; bf4[1] = ab.y; bf4[2] = cd.x; bf4[3] = cd.y;
      }
    }
    const double npbraket = (double)nb * (double)nk;
    double acc[K::NFT];
    static_for<K::NFT>([&](auto ic) { acc[decltype(ic)::value] = 0.0; });
    int nsurv = 0;
    sync();   // also separates the previous round's phase 2 from this round's writes
    if constexpr (FOCK)   // density blocks of this quartet: in flight during the K loop
      fock_prefetch_density<LA, LB, LC, LD, NEC>(p, valid, rmeta.row, Q + K::OFF_D, bf4);
    int maxit;
    if constexpr (NEC == 1) maxit = nit;   // no exchange inside the loop: private trip count
    else if constexpr (WL) maxit = __reduce_max_sync(0xffffffffu, valid ? nit : 0);
    else maxit = s_maxit[round % 3];
    ++round;

#ifndef LB200_X_KLOOP
#define LB200_X_KLOOP 0
#endif
    // LB200_X_KLOOP = 1: primitive-pair counters of the K loop (it = ipb * nk + ipk) advanced incrementally and
    // the bra record reloaded only when it changes, instead of an integer division by the run-time nk and two
    // record loads per primitive quartet.  Measured on the (H2O)_64 / def2-TZVP build: 2.29 s -> 2.30 s, twice
    // (the contracted kernels wait on memory, not on the ~20 instructions of the division) -- left off.
    int ipb_c = 0, ipk_c = 0;
    PrimPair bp_c{};
    for (int it = 0; it < maxit; ++it) {
      // ---- prerequisites (engine.impl.h:1331-1367,1389-1392,1602-1641) ----------------
      bool on = valid && it < nit;
      PrimPair bp, kp;
      double pfac = 0.0, Targ = 0.0, rho = 0.0, oogpq = 0.0;
      if constexpr (PREREQ) {
      } else if (on) {
        if constexpr (LB200_X_KLOOP) {
          if (ipk_c == 0) bp_c = p.bra.prim[pb0 + ipb_c];   // the bra record changes every nk iterations only
          bp = bp_c;
          kp = p.ket.prim[pk0 + ipk_c];
          if (++ipk_c == nk) { ipk_c = 0; ++ipb_c; }
        } else {
          const int ipb = it / nk;
          const int ipk = it - ipb * nk;
          bp = p.bra.prim[pb0 + ipb];
          kp = p.ket.prim[pk0 + ipk];
        }
        on = bp.ln_scr + kp.ln_scr > ln_prec;  // engine.impl.h:1313-1314
      }
      if constexpr (PREREQ) {
      } else if (on) {
        const double PQx = bp.P[0] - kp.P[0], PQy = bp.P[1] - kp.P[1], PQz = bp.P[2] - kp.P[2];
        const double PQ2 = PQx * PQx + PQy * PQy + PQz * PQz;
        const double gpq = bp.gamma + kp.gamma;
        oogpq = 1.0 / gpq;
        pfac = bp.Kc * kp.Kc * sqrt(gpq) * oogpq;
        if (p.screening & (kScreenOriginal | kScreenConservative)) {  // engine.impl.h:1371-1386
          double est = fabs(pfac);
          if (p.screening == kScreenConservative)
            est *= fmax(1.0, bp.nonsph * kp.nonsph) * npbraket;
          if (est < prec) on = false;
        }
        rho = bp.gamma * kp.gamma * oogpq;
        Targ = PQ2 * rho;
      }
      double PA[3], WP[3], QC[3], WQ[3], oo2z = 0, roz = 0, koo2e[6], roe = 0, ce[3];
      if constexpr (PREREQ) {
        if (on) {
          ++nsurv;
          const PrereqRec& r = p.prereq[pb0 + it];
          const bool sw = p.swap_tasks != 0;   // kernel bra = caller ket
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            PA[d] = sw ? r.QC[d] : r.PA[d];
            QC[d] = sw ? r.PA[d] : r.QC[d];
            WP[d] = sw ? r.WQ[d] : r.WP[d];
            WQ[d] = sw ? r.WP[d] : r.WQ[d];
          }
          oo2z = sw ? r.oo2e : r.oo2z;
          roz = sw ? r.roe : r.roz;
          const double oo2e = sw ? r.oo2z : r.oo2e;
          koo2e[0] = 0.0; koo2e[1] = oo2e; koo2e[2] = 2.0 * oo2e; koo2e[3] = 3.0 * oo2e;
          koo2e[4] = 4.0 * oo2e; koo2e[5] = 5.0 * oo2e;
          roe = sw ? r.roz : r.roe;
          ce[0] = rmeta.q[0] * r.oo2ze; ce[1] = rmeta.q[1] * r.oo2ze; ce[2] = rmeta.q[2] * r.oo2ze;
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) PA[d] = WP[d] = QC[d] = WQ[d] = ce[d] = 0.0;
          koo2e[0] = koo2e[1] = koo2e[2] = koo2e[3] = koo2e[4] = koo2e[5] = 0.0;
        }
      } else if (on) {
        ++nsurv;
        const double gp = oogpq * bp.gamma, gq = oogpq * kp.gamma;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double W = gp * bp.P[d] + gq * kp.P[d];
          WP[d] = W - bp.P[d];
          WQ[d] = W - kp.P[d];
          PA[d] = bra_unit ? 0.0 : bp.P[d] - Ac[d];   // engine.impl.h:1514-1537
          QC[d] = ket_unit ? 0.0 : kp.P[d] - Cc[d];
        }
        oo2z = 0.5 * bp.oog;
        roz = rho * bp.oog;
        const double oo2e = 0.5 * kp.oog;
        koo2e[0] = 0.0; koo2e[1] = oo2e; koo2e[2] = 2.0 * oo2e; koo2e[3] = 3.0 * oo2e;
        koo2e[4] = 4.0 * oo2e; koo2e[5] = 5.0 * oo2e;
        roe = rho * kp.oog;
        const double oo2ze = 0.5 * oogpq;
        ce[0] = rmeta.q[0] * oo2ze; ce[1] = rmeta.q[1] * oo2ze; ce[2] = rmeta.q[2] * oo2ze;
      } else {
        pfac = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) PA[d] = WP[d] = QC[d] = WQ[d] = ce[d] = 0.0;
        koo2e[0] = koo2e[1] = koo2e[2] = koo2e[3] = koo2e[4] = koo2e[5] = 0.0;
      }

      // ---- Boys: lanes m of the quartet, then broadcast through shared memory ----------
      double F[L + 1];
      if constexpr (PREREQ) {   // (ss|ss)^(m) come with the record; every lane reads its own copy
        static_for<L + 1>([&](auto mc) {
          constexpr int m = decltype(mc)::value;
          F[m] = on ? p.prereq[pb0 + it].F[m] : 0.0;
        });
      } else if constexpr (NEC == 1) {
        if constexpr (FOCK && LB200_BOYS_RECUR) {
          if (on) {
            boys_all<L>(p.boys, Targ, F);
            static_for<L + 1>([&](auto mc) { F[decltype(mc)::value] *= pfac; });
          } else {
            static_for<L + 1>([&](auto mc) { F[decltype(mc)::value] = 0.0; });
          }
        } else {
          static_for<L + 1>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            F[m] = on ? boys_value(p.boys, Targ, m) * pfac : 0.0;
          });
        }
      } else {
        if constexpr (FOCK && LB200_BOYS_RECUR && NEC <= 4) {
          // one lane per quartet evaluates every order from one table row (boys_all); with more rows
          // per quartet the row-per-lane evaluation is already spread thin enough
          if (lane_on && rmeta.row == 0) {
            double Fa[L + 1];
            if (on) boys_all<L>(p.boys, Targ, Fa);
            static_for<L + 1>([&](auto mc) {
              constexpr int m = decltype(mc)::value;
              Q[K::OFF_F + m] = on ? Fa[m] * pfac : 0.0;
            });
          }
        } else if (lane_on) {
          for (int m = rmeta.row; m <= L; m += NEC)
            Q[K::OFF_F + m] = on ? boys_value(p.boys, Targ, m) * pfac : 0.0;
        }
        sync();
        static_for<L + 1>([&](auto mc) { F[decltype(mc)::value] = Q[K::OFF_F + decltype(mc)::value]; });
      }

      // ---- [row 0|00]^(m): private chain (vrr_11_twoprep_11.h:154-222) ------------------
      Lvl<FMAX, 0> l0;
      if constexpr (EMAX == 0) {
        static_for<FMAX + 1>([&](auto mc) { l0.v[decltype(mc)::value] = F[decltype(mc)::value]; });
      } else {
        double cur[L + 1], prv[L + 1];
        static_for<L + 1>([&](auto mc) {
          cur[decltype(mc)::value] = F[decltype(mc)::value];
          prv[decltype(mc)::value] = 0.0;
        });
        static_for<EMAX>([&](auto kc) {
          constexpr int k = decltype(kc)::value;
          const int d = cdir_[k];
          const double pa = d < 0 ? 1.0 : sel3(d, PA[0], PA[1], PA[2]);
          const double wp = d < 0 ? 0.0 : sel3(d, WP[0], WP[1], WP[2]);
          const double c = ccnt_[k] * oo2z;
          static_for<L - k>([&](auto mc) {
            constexpr int m = decltype(mc)::value;
            const double nv = pa * cur[m] + wp * cur[m + 1] + c * (prv[m] - roz * prv[m + 1]);
            prv[m] = cur[m];
            cur[m] = nv;
          });
        });
        static_for<FMAX + 1>([&](auto mc) { l0.v[decltype(mc)::value] = cur[decltype(mc)::value]; });
        // level-0 cross values
        if (lane_on && rmeta.row < NECX)
          static_for<FMAX>([&](auto mc) {
            constexpr int m = decltype(mc)::value + 1;
            Q[K::OFF_X + K::xslot(0, 0, m) * NECX + rmeta.row] = l0.v[m];
          });
      }
      if constexpr (LC == 0) acc[0] += l0.v[0];

      // ---- [row 0|f 0]^(m), f = 1..FMAX (vrr_11_twoprep_11.h:305-383) -------------------
      double* Xq = Q + K::OFF_X;
      if constexpr (FMAX >= 1) {
        if constexpr (EMAX > 0) sync();
        Lvl<FMAX, 1> l1;
        rr_build_level<K, 1>(l1, l0, l0, QC, WQ, koo2e, roe, ce, Xq, rmeta, acc);
        if constexpr (FMAX >= 2) {
          if constexpr (EMAX > 0) sync();
          Lvl<FMAX, 2> l2;
          rr_build_level<K, 2>(l2, l1, l0, QC, WQ, koo2e, roe, ce, Xq, rmeta, acc);
          if constexpr (FMAX >= 3) {
            if constexpr (EMAX > 0) sync();
            Lvl<FMAX, 3> l3;
            rr_build_level<K, 3>(l3, l2, l1, QC, WQ, koo2e, roe, ce, Xq, rmeta, acc);
            if constexpr (FMAX >= 4) {
              if constexpr (EMAX > 0) sync();
              Lvl<FMAX, 4> l4;
              rr_build_level<K, 4>(l4, l3, l2, QC, WQ, koo2e, roe, ce, Xq, rmeta, acc);
              if constexpr (FMAX >= 5) {   // (fd|, (ff| unrolled: correct, but the pyramid spills
                if constexpr (EMAX > 0) sync();
                Lvl<FMAX, 5> l5;
                rr_build_level<K, 5>(l5, l4, l3, QC, WQ, koo2e, roe, ce, Xq, rmeta, acc);
                if constexpr (FMAX >= 6) {
                  if constexpr (EMAX > 0) sync();
                  Lvl<FMAX, 6> l6;
                  rr_build_level<K, 6>(l6, l5, l4, QC, WQ, koo2e, roe, ce, Xq, rmeta, acc);
                }
              }
            }
          }
        }
      }
      // the next iteration's first shared write (Boys values / level-0 cross values) must not
      // overtake this iteration's last cross-term reads
      if constexpr (EMAX > 0) sync();
    }

    // ---- ket HRR in registers: (row 0|c d) from (row 0|f 0), hrr.h:324 -------------------
    double H[K::NCD];
    rr_hrr_regs<LC, LD>(acc, CD, H);
    const bool screened_out = (nsurv == 0);
    if constexpr (FOCK) {   // K_eff of the flop model, counted only when profiling
      if (p.prim_counter) count_primitives(p.prim_counter, (valid && rmeta.row == 0) ? nsurv : 0);
    }
    if (screened_out) static_for<K::NCD>([&](auto ic) { H[decltype(ic)::value] = 0.0; });

    {
      // ---- hand the rows over to (quartet, cd) lanes ------------------------------------
      // (the K loop ended with a barrier, or had no shared traffic: the region is free)
      constexpr int TBOFF = K::OFF_B2;
      if constexpr (LB > 0) {   // A - B is only needed by the bra HRR
        if (valid && rmeta.row == 0) {
          if constexpr (PREREQ) {
            const double* gv = p.prereq_geom + 6 * (size_t)task + (p.swap_tasks ? 3 : 0);
            Q[0] = gv[0]; Q[1] = gv[1]; Q[2] = gv[2];
          } else {
            const PairGeom& gb = p.bra.geom[ib];
            Q[0] = gb.AB[0]; Q[1] = gb.AB[1]; Q[2] = gb.AB[2];
          }
        }
      }
      if constexpr (LB > 0) {
        if (valid && rmeta.row >= K::ROW0)
          static_for<K::NCD>([&](auto ic) {
            Q[TBOFF + decltype(ic)::value * K::RTP + (rmeta.row - K::ROW0)] = H[decltype(ic)::value];
          });
      } else {  // LB == 0: the rows are the final (a 0|c d) integrals
        if (valid && rmeta.row >= K::ROW0)
          static_for<K::NCD>([&](auto ic) {
            Q[K::HDR + (rmeta.row - K::ROW0) * K::CS + decltype(ic)::value] = H[decltype(ic)::value];
          });
      }
      sync();
      if constexpr (LB > 0) {
        // ---- bra HRR in registers: (a b|c d) from (e 0|c d), hrr.h:246 -------------------
        for (int item = gl; item < QPG * K::NCD; item += GROUP) {
          const int q2 = item / K::NCD, cd = item - q2 * K::NCD;
          if (base + q2 >= ntasks) continue;
          double* Q2 = smem + (size_t)(g * QPG + q2) * QSIZE;
          const double ABv[3] = {Q2[0], Q2[1], Q2[2]};
          double colin[K::NRT];
          static_for<K::NRT>([&](auto rc) {
            colin[decltype(rc)::value] = Q2[TBOFF + cd * K::RTP + decltype(rc)::value];
          });
          double O[K::NAB];
          rr_hrr_regs<LA, LB>(colin, ABv, O);
          if (!FOCK && !p.transpose_out) {
            // natural orientation: lanes of a warp are adjacent in cd -> coalesced rows
            double* __restrict__ o = p.out + (size_t)(base + q2) * (K::NAB * K::NCD);
            static_for<K::NAB>([&](auto ic) { o[decltype(ic)::value * K::NCD + cd] = O[decltype(ic)::value]; });
          } else {
            double* fin = Q2 + K::HDR;
            static_for<K::NAB>([&](auto ic) { fin[decltype(ic)::value * K::CS + cd] = O[decltype(ic)::value]; });
          }
        }
        if (FOCK || p.transpose_out) sync();
      }
      if constexpr (!FOCK) {
        // ---- coalesced copy-out: the round's quartets are one contiguous run of the output
        // (layout ((a*nb+b)*nc+c)*nd+d per shell set, doc/progman/progman.tex:472-474; [cd][ab]
        // when the caller's bra is this kernel's unrolled side) -----------------------------
        constexpr int BLK = K::NAB * K::NCD;
        const unsigned left = ntasks - base;
        const int nvalid = left < (unsigned)QPG ? (int)left : QPG;
        double* __restrict__ o = p.out + (size_t)base * BLK;
        if (!p.transpose_out) {
          if constexpr (LB == 0) {
            for (int idx = gl; idx < nvalid * BLK; idx += GROUP) {
              const int q2 = idx / BLK, i = idx - q2 * BLK;
              const int ab = i / K::NCD, cd = i - ab * K::NCD;
              o[idx] = smem[(size_t)(g * QPG + q2) * QSIZE + K::HDR + ab * K::CS + cd];
            }
          }
        } else {
          for (int idx = gl; idx < nvalid * BLK; idx += GROUP) {
            const int q2 = idx / BLK, i = idx - q2 * BLK;
            const int cd = i / K::NAB, ab = i - cd * K::NAB;
            o[idx] = smem[(size_t)(g * QPG + q2) * QSIZE + K::HDR + ab * K::CS + cd];
          }
        }
      } else {
        // ---- cart -> pure, then 6-way digestion, by the NEC lanes of each quartet --------
        fock_digest<LA, LB, LC, LD, NEC, WL>(p, valid && !screened_out, rmeta.row, Q + K::HDR, K::CS,
                                         Q + K::OFF_B2, Q + K::OFF_D, ib, ik, deg, bf4);
      }
    }
  }
}

}  // namespace lb200
