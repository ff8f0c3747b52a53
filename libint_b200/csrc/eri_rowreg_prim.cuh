// Uncontracted ("primitive") specialization of the row-register class kernel (eri_rowreg.cuh)
// with a software pipeline across quartets.
//
// libint generates separate code for uncontracted and contracted shell sets
// (LIBINT_CONTRACTED_INTS, src/bin/libint/dg.cc:1128-1188: the contraction loop is emitted only
// for contracted targets); this is the same split for the GPU path: when every shell pair of
// both pair blocks holds at most one primitive pair there is no K loop, no accumulator set and
// no per-iteration on/off scaffolding, and -- because one round is one primitive quartet -- the
// dependent global-memory chain  task -> primitive offsets -> pair records -> Boys table  of
// round r+1 can be issued while round r computes:
//
//   round r   top        offsets(r+1) <- prim_off[task(r+1)],  task(r+2) <- tasks[]   (registers)
//             VRR        prerequisites + F_m from stage[r&1] (shared, broadcast reads)
//             mid-VRR    cp.async pair records, A-B, C-D of round r+1 -> stage[(r+1)&1]
//             ket HRR, transpose to (quartet, cd) lanes
//             Boys lanes wait for the records, form T / pfac / 1/(zeta+eta) / rho of round r+1,
//                        publish them and touch their Boys-table row (L1/L2 prefetch by load)
//             bra HRR, copy-out
//             Boys lanes evaluate F_m(T)*pfac of round r+1 -> stage[(r+1)&1]
//
// so the only exposed memory latency is the first round of a group.  Everything else (row /
// register-pyramid layout, recurrences, HRR, copy-out) is eri_rowreg.cuh's.
#pragma once
#include "eri_rowreg.cuh"

namespace lb200 {

// TR: the caller's bra is this kernel's unrolled side, results leave as [cd][ab] through the
// transposing copy-out.  TR = false with LB > 0 stores straight from the bra-HRR lanes and needs
// no staging buffer for the final integrals: a third less shared memory, one more CTA per SM.
// FOCK: the finished shell set is digested into the Fock matrix (fock_digest.cuh) instead of
// stored; tasks come from the screening kernel's device-side list, the per-quartet engine
// precision and degeneracy (hartree-fock++.cc:1667-1703) are prefetched with the task.
template <int LA, int LB, int LC, int LD, bool TR, bool FOCK = false>
struct RRP : RRK<LA, LB, LC, LD> {
  using B = RRK<LA, LB, LC, LD>;
  // CTA shape of this kernel: 35 rows per quartet ((dd| / (fp| rows) waste 18 % of a 128-lane CTA
  // (3 quartets); 256 lanes would hold 7 (96 %).  Measured: no gain ((dd|dd) 36.6 vs 36.5 ms per
  // 10^7 quartets) -- the kernel is bound by the latency of its FP64 dependency chains at 4 warps
  // per scheduler, not by lane utilisation -- so the narrow CTA stays the default.
#ifndef LB200_PRIM_WIDE35
#define LB200_PRIM_WIDE35 0
#endif
  // LB200_PRIM_T35: CTA width of the 35-row kernels (experiments: 160 = 4 quartets, 87.5 % of the lanes)
#ifndef LB200_PRIM_T35
#define LB200_PRIM_T35 0
#endif
  static constexpr int THREADS = (B::NEC == 35 && LB200_PRIM_T35) ? LB200_PRIM_T35
                               : ((B::NEC == 35 && LB200_PRIM_WIDE35) ? 256 : B::THREADS);
  static constexpr int GROUP = B::WL ? 32 : THREADS;
  static constexpr int QPG = GROUP / B::NEC;
  static constexpr int NG = THREADS / GROUP;
  static constexpr int QPC = NG * QPG;
  // pipeline stage (doubles): bra record 8 | ket record 8 | bra PairGeom 8 (A, A-B, ints) |
  // ket PairGeom 8 (C, C-D, ints) | prep 4 | F_m
  static constexpr int S_BP = 0, S_KP = 8, S_GB = 16, S_GK = 24, S_AB = S_GB + 3, S_CD = S_GK + 3,
                       S_PREP = 32, S_F = 36;
  static constexpr int PSTAGE = (S_F + B::L + 1 + 1) & ~1;
  static constexpr int PIPE = 2 * PSTAGE;
  // phase areas behind the pipeline stages
  static constexpr int OFF_X = PIPE;
  static constexpr int VRR_DOUBLES = OFF_X + (B::EMAX > 0 ? B::XSLOTS * B::NECX : 0);
  static constexpr int OFF_FIN = PIPE;                       // final integrals [NAB][CS]
  static constexpr bool HAS_FIN = TR || LB == 0 || FOCK;
  static constexpr int OFF_B2 = OFF_FIN + (HAS_FIN ? B::NAB * B::CS : 0);  // row -> column transpose buffer
  static constexpr int OFF_D =
      cmax(OFF_B2 + cmax(FOCK ? B::NAB * B::NCD : 0, LB > 0 ? B::NCD * B::RTP : 0), (VRR_DOUBLES + 1) & ~1);
  static constexpr int P2_DOUBLES = OFF_D + (FOCK ? fock_dblock_doubles<LA, LB, LC, LD>() : 0);
  static constexpr int QSIZE = B::pad_stride(cmax(VRR_DOUBLES, P2_DOUBLES));
#ifndef LB200_PRIM_MINB_HI
#define LB200_PRIM_MINB_HI 4
#endif
#ifndef LB200_PRIM_MINB_MID
#define LB200_PRIM_MINB_MID 5
#endif
#ifndef LB200_PRIM_MINB_LO
#define LB200_PRIM_MINB_LO 6
#endif
  static constexpr int MINB128 =
      B::FMAX >= 4 ? LB200_PRIM_MINB_HI : (B::FMAX >= 2 ? LB200_PRIM_MINB_MID : LB200_PRIM_MINB_LO);
  static constexpr int MINB = THREADS == 256 ? cmax(1, MINB128 / 2) : (THREADS > 128 ? cmax(1, MINB128 * 128 / THREADS) : MINB128);
};

#ifndef LB200_X_NOZERO
#define LB200_X_NOZERO 1
#endif
// timing decomposition only (results are WRONG when any of these is set): skip the global stores of the bra-HRR
// lanes / the whole bra HRR / the register pyramid / the Boys evaluation
#ifndef LB200_DIAG_NOSTORE
#define LB200_DIAG_NOSTORE 0
#endif
#ifndef LB200_DIAG_NOBRAHRR
#define LB200_DIAG_NOBRAHRR 0
#endif
#ifndef LB200_DIAG_NOVRR
#define LB200_DIAG_NOVRR 0
#endif
#ifndef LB200_DIAG_NOBOYS
#define LB200_DIAG_NOBOYS 0
#endif

template <int LA, int LB, int LC, int LD, bool TR, bool FOCK>
__global__ void __launch_bounds__(RRP<LA, LB, LC, LD, TR, FOCK>::THREADS, RRP<LA, LB, LC, LD, TR, FOCK>::MINB)
eri_rowreg_prim_kernel(const EriParams p, const RowInfo* __restrict__ rows) {
  using K = RRP<LA, LB, LC, LD, TR, FOCK>;
  static_assert(!(FOCK && TR), "Fock mode has no output orientation");
  constexpr int EMAX = K::EMAX, FMAX = K::FMAX, L = K::L, NEC = K::NEC, NECX = K::NECX;
  constexpr int QSIZE = K::QSIZE, PSTAGE = K::PSTAGE;
  constexpr bool WL = K::WL;
  constexpr int GROUP = K::GROUP, QPG = K::QPG, NG = K::NG;
  static_assert(NEC > 1, "thread-per-quartet classes use the general kernel");
  static_assert(FMAX <= 6, "register pyramid is written for LC+LD <= 6");

  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const int g = WL ? tid >> 5 : 0;
  const int gl = WL ? tid & 31 : tid;
  const int qg = gl / NEC;
  const bool lane_on = qg < QPG;
  const int q = g * QPG + qg;
  auto sync = [] {
    if constexpr (WL) __syncwarp(); else __syncthreads();
  };
  RowMeta rmeta;
  rmeta.row = lane_on ? gl - qg * NEC : NEC;
  {
    const RowInfo ri = rows[rmeta.row];
    rmeta.e = ri.e;
    for (int d = 0; d < 3; ++d) {
      rmeta.rm[d] = lane_on ? ri.rm[d] : 0;   // leftover lanes read (and discard) row 0's slots
      rmeta.q[d] = (double)ri.q[d];
    }
  }
  int cdir_[EMAX > 0 ? EMAX : 1];
  double ccnt_[EMAX > 0 ? EMAX : 1];
  if constexpr (EMAX > 0) {
    const int qx = (int)rmeta.q[0], qy = (int)rmeta.q[1];
    static_for<EMAX>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      const int s = k - (EMAX - rmeta.e);
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
  double* const Q = smem + (size_t)(lane_on ? q : g * QPG) * QSIZE;
  // Fock mode: one lane per quartet evaluates every order from one table row (boys_all)
  constexpr bool RECUR = FOCK && LB200_BOYS_RECUR;
  const bool boys_lane = lane_on && (RECUR ? rmeta.row == 0 : rmeta.row <= L);

  const bool bra_unit = p.bra.unit_b != 0, ket_unit = p.ket.unit_b != 0;
  const unsigned ntasks = p.ntasks_dev ? *p.ntasks_dev : p.ntasks;
  const unsigned stride = gridDim.x * NG * QPG;
  unsigned base = (blockIdx.x * NG + g) * QPG;
  if (base >= ntasks) return;   // whole group idle (groups never share barriers when WL;
                                // for the CTA-wide group the condition is CTA-uniform)

  // ---- pipeline helpers ---------------------------------------------------------------
  // primitive offsets of a task: (first bra primitive, first ket primitive), -1 if the task is
  // out of range or either pair kept no primitive
  // Fock mode: the task record carries ln(engine precision) and the degeneracy of the quartet
  // (types.cuh, written by the screening kernel)
  struct Task { int2 t; double lnp, deg; };
  auto load_task = [&](unsigned b) -> Task {
    const unsigned t = b + qg;
    Task r{make_int2(-1, -1), 0.0, 1.0};
    if (!lane_on || t >= ntasks) return r;
    int2 tk;
    if constexpr (FOCK) {
      const int4 ft = p.ftasks[t];
      tk = make_int2(ft.x, ft.y & 0x3fffffff);
      r.deg = (double)(1 << ((unsigned)ft.y >> 30));
      r.lnp = __hiloint2double(ft.w, ft.z);
    } else {
      tk = p.prod_nk ? make_int2(p.prod_b0 + (int)(t / p.prod_nk), p.prod_k0 + (int)(t % p.prod_nk)) : p.tasks[t];
    }
    r.t = p.swap_tasks ? make_int2(tk.y, tk.x) : tk;
    return r;
  };
  struct Off { int pb, pk, ib, ik; };   // ib < 0: no task; pb < 0: a pair kept no primitive
  auto load_off = [&](int2 tk) -> Off {
    Off o{-1, -1, -1, -1};
    if (tk.x < 0) return o;
    const int pb0 = p.bra.prim_off[tk.x], pb1 = p.bra.prim_off[tk.x + 1];
    const int pk0 = p.ket.prim_off[tk.y], pk1 = p.ket.prim_off[tk.y + 1];
    o.ib = tk.x; o.ik = tk.y;
    if (pb1 > pb0 && pk1 > pk0) { o.pb = pb0; o.pk = pk0; }
    return o;
  };
  // 16 sixteen-byte chunks per quartet: 4 + 4 pieces of the two records, 4 + 4 of the two PairGeoms
  // (geometry for every real task: the HRR of an all-screened quartet must see finite numbers)
  auto issue_records = [&](const Off& o, double* S) {
    if (lane_on && o.ib >= 0) {
      const char* gb = reinterpret_cast<const char*>(p.bra.prim + (o.pb < 0 ? 0 : o.pb));
      const char* gk = reinterpret_cast<const char*>(p.ket.prim + (o.pk < 0 ? 0 : o.pk));
      const char* hb = reinterpret_cast<const char*>(p.bra.geom + o.ib);
      const char* hk = reinterpret_cast<const char*>(p.ket.geom + o.ik);
      for (int c = rmeta.row; c < 16; c += NEC) {
        if (c < 8) {
          if (o.pb < 0) continue;
          if (c < 4) cp_async16(S + K::S_BP + 2 * c, gb + 16 * c);
          else cp_async16(S + K::S_KP + 2 * (c - 4), gk + 16 * (c - 4));
        } else if (c < 12) {
          cp_async16(S + K::S_GB + 2 * (c - 8), hb + 16 * (c - 8));
        } else {
          cp_async16(S + K::S_GK + 2 * (c - 12), hk + 16 * (c - 12));
        }
      }
    }
  };
  // Boys lanes, first half: T, pfac (registers), 1/(zeta+eta), rho, on (published by row 0)
  struct BoysState { double T, pfac; bool on; };
  auto boys_prepare = [&](const Off& o, double* S, double lnp) -> BoysState {
    BoysState b{0.0, 0.0, false};
    if (!boys_lane) return b;
    bool on = o.pb >= 0;
    double oogpq = 0.0, rho = 0.0;
    if (on) {
      const double lnb = S[K::S_BP + 6], lnk = S[K::S_KP + 6];
      const double ln_prec = FOCK ? lnp : p.ln_precision;   // hartree-fock++.cc:1693-1695
      on = lnb + lnk > ln_prec;   // engine.impl.h:1313-1314
    }
    if (on) {
      const double PQx = S[K::S_BP + 0] - S[K::S_KP + 0], PQy = S[K::S_BP + 1] - S[K::S_KP + 1],
                   PQz = S[K::S_BP + 2] - S[K::S_KP + 2];
      const double PQ2 = PQx * PQx + PQy * PQy + PQz * PQz;
      const double gb = S[K::S_BP + 4], gk = S[K::S_KP + 4];
      const double gpq = gb + gk;
      oogpq = 1.0 / gpq;
      b.pfac = S[K::S_BP + 3] * S[K::S_KP + 3] * sqrt(gpq) * oogpq;
      if (p.screening & (kScreenOriginal | kScreenConservative)) {  // engine.impl.h:1371-1386
        double est = fabs(b.pfac);
        if (p.screening == kScreenConservative)
          est *= fmax(1.0, S[K::S_BP + 7] * S[K::S_KP + 7]);  // npbra * npket = 1
        if (est < p.precision) on = false;
      }
      rho = gb * gk * oogpq;
      b.T = PQ2 * rho;
    }
    b.on = on;
    if (rmeta.row == 0) {
      S[K::S_PREP + 0] = oogpq;
      S[K::S_PREP + 1] = rho;
      S[K::S_PREP + 2] = on ? 1.0 : 0.0;
    }
    return b;
  };
  // touch this lane's first Boys-table row so that boys_finish finds it in L1
  auto boys_touch = [&](const BoysState& b) -> double {
    if (!boys_lane || !b.on || b.T > kBoysTmax) return 0.0;
    int iv = (int)(b.T * 7.0);
    iv = iv > kBoysNInt - 1 ? kBoysNInt - 1 : iv;
    const double* d = p.boys + ((size_t)iv * (kBoysTableMmax + 1) + (RECUR ? L : rmeta.row)) * 8;
    return __ldg(d) + __ldg(d + 4);
  };
  auto boys_finish = [&](const BoysState& b, double* S) {
    if constexpr (RECUR) {
      if (boys_lane) {
        double Fa[L + 1];
        if (b.on) boys_all<L>(p.boys, b.T, Fa);
        static_for<L + 1>([&](auto mc) {
          constexpr int m = decltype(mc)::value;
          S[K::S_F + m] = b.on ? Fa[m] * b.pfac : 0.0;
        });
      }
    } else {
      if (boys_lane)
        for (int m = rmeta.row; m <= L; m += NEC)
          S[K::S_F + m] = b.on ? boys_value(p.boys, b.T, m) * b.pfac : 0.0;
    }
  };

  // ---- prologue: round 0 of this group, unpipelined ---------------------------------------
  Task tk_next;
  Off ocur;
  double deg_cur = 1.0;
  {
    if constexpr (LB200_X_NOZERO) {   // stale stage contents must be finite numbers
      if (lane_on)
        for (int k = rmeta.row; k < K::PIPE; k += NEC) Q[k] = 0.0;
      sync();
    }
    const Task tk0 = load_task(base);
    tk_next = load_task(base + stride);
    ocur = load_off(tk0.t);
    deg_cur = tk0.deg;
    issue_records(ocur, Q);
    cp_async_wait_all();
    sync();
    const BoysState b0 = boys_prepare(ocur, Q, tk0.lnp);
    boys_finish(b0, Q);
  }

  for (int round = 0; base < ntasks; base += stride, ++round) {
    double* const S = Q + (round & 1) * PSTAGE;          // this round's stage
    double* const SN = Q + ((round + 1) & 1) * PSTAGE;   // next round's stage
    const bool more = base + stride < ntasks;            // group-uniform
    // ---- top: next round's offsets, the task after that (consumed later in this round) ----
    const Off onext = more ? load_off(tk_next.t) : Off{-1, -1, -1, -1};
    const Task tk_next2 = (more && base + 2 * stride < ntasks) ? load_task(base + 2 * stride)
                                                               : Task{make_int2(-1, -1), 0.0, 1.0};
    sync();   // stage S complete (prologue / previous round); previous phase 2 finished

    // ---- prerequisites from the stage (engine.impl.h:1331-1367,1389-1392,1602-1641) --------
    const bool on = S[K::S_PREP + 2] != 0.0;
    int bf4[4] = {0, 0, 0, 0};
    if constexpr (FOCK) {   // density blocks of this quartet: in flight during the VRR
      const bool act = lane_on && base + qg < ntasks && on;
      if (act) {
        const int2 ab = *reinterpret_cast<const int2*>(S + K::S_GB + 6);
        const int2 cd = *reinterpret_cast<const int2*>(S + K::S_GK + 6);
        bf4[0] = ab.x; bf4[1] = ab.y; bf4[2] = cd.x; bf4[3] = cd.y;
      }
      fock_prefetch_density<LA, LB, LC, LD, NEC>(p, act, rmeta.row, Q + K::OFF_D, bf4);
    }
    double PA[3], WP[3], QC[3], WQ[3], oo2z, roz, koo2e[6], roe, ce[3];
    {
      const double oogpq = S[K::S_PREP + 0], rho = S[K::S_PREP + 1];
      const double gb = S[K::S_BP + 4], gk = S[K::S_KP + 4];
      const double oogb = S[K::S_BP + 5], oogk = S[K::S_KP + 5];
      const double gp = oogpq * gb, gq = oogpq * gk;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double Pb = S[K::S_BP + d], Pk = S[K::S_KP + d];
        const double W = gp * Pb + gq * Pk;
        WP[d] = W - Pb;
        WQ[d] = W - Pk;
        PA[d] = bra_unit ? 0.0 : Pb - S[K::S_GB + d];   // engine.impl.h:1514-1537
        QC[d] = ket_unit ? 0.0 : Pk - S[K::S_GK + d];
      }
      oo2z = 0.5 * oogb;
      roz = rho * oogb;
      const double oo2e = 0.5 * oogk;
      koo2e[0] = 0.0; koo2e[1] = oo2e; koo2e[2] = 2.0 * oo2e; koo2e[3] = 3.0 * oo2e;
      koo2e[4] = 4.0 * oo2e; koo2e[5] = 5.0 * oo2e;
      roe = rho * oogk;
      const double oo2ze = 0.5 * oogpq;
      ce[0] = rmeta.q[0] * oo2ze; ce[1] = rmeta.q[1] * oo2ze; ce[2] = rmeta.q[2] * oo2ze;
    }
    double CD[3] = {S[K::S_CD], S[K::S_CD + 1], S[K::S_CD + 2]};
    double F[L + 1];
    static_for<L + 1>([&](auto mc) { F[decltype(mc)::value] = S[K::S_F + decltype(mc)::value]; });

    // ---- [row 0|00]^(m): private chain (vrr_11_twoprep_11.h:154-222) ----------------------
    double acc[K::NFT];
    Lvl<FMAX, 0> l0;
    {
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
      if (lane_on && rmeta.row < NECX)
        static_for<FMAX>([&](auto mc) {
          constexpr int m = decltype(mc)::value + 1;
          Q[K::OFF_X + K::xslot(0, 0, m) * NECX + rmeta.row] = l0.v[m];
        });
    }
    if constexpr (LC == 0) acc[0] = l0.v[0];

    // next round's records: the stage SN was last read before this round's top barrier
    issue_records(onext, SN);

    // ---- [row 0|f 0]^(m), f = 1..FMAX (vrr_11_twoprep_11.h:305-383) -----------------------
    double* Xq = Q + K::OFF_X;
#ifndef LB200_X_SPLIT
#define LB200_X_SPLIT 0
#endif
    // one level: barrier, then the level (default) -- or the barrier inside the level, after its register-only
    // part (LB200_X_SPLIT = 1, rr_build_level_split)
    auto level = [&](auto fc, auto& out, const auto& p1, const auto& p2) {
      constexpr int F = decltype(fc)::value;
      if constexpr (LB200_X_SPLIT) {
        rr_build_level_split<K, F, false>(out, p1, p2, QC, WQ, koo2e, roe, ce, Xq, rmeta, acc, sync);
      } else {
        sync();
        rr_build_level<K, F, false>(out, p1, p2, QC, WQ, koo2e, roe, ce, Xq, rmeta, acc);
      }
    };
    if constexpr (FMAX >= 1 && !LB200_DIAG_NOVRR) {
      Lvl<FMAX, 1> l1;
      level(std::integral_constant<int, 1>{}, l1, l0, l0);
      if constexpr (FMAX >= 2) {
        Lvl<FMAX, 2> l2;
        level(std::integral_constant<int, 2>{}, l2, l1, l0);
        if constexpr (FMAX >= 3) {
          Lvl<FMAX, 3> l3;
          level(std::integral_constant<int, 3>{}, l3, l2, l1);
          if constexpr (FMAX >= 4) {
            Lvl<FMAX, 4> l4;
            level(std::integral_constant<int, 4>{}, l4, l3, l2);
            if constexpr (FMAX >= 5) {
              Lvl<FMAX, 5> l5;
              level(std::integral_constant<int, 5>{}, l5, l4, l3);
              if constexpr (FMAX >= 6) {
                Lvl<FMAX, 6> l6;
                level(std::integral_constant<int, 6>{}, l6, l5, l4);
              }
            }
          }
        }
      }
    }

    // ---- ket HRR in registers: (row 0|c d) from (row 0|f 0), hrr.h:324 ---------------------
    double H[K::NCD];
    rr_hrr_regs<LC, LD>(acc, CD, H);
    // LB200_X_NOZERO (default): a screened-out quartet has F_m = 0 (boys_finish) and finite prerequisites (the stages are
    // zeroed once in the prologue), so every recurrence value is already an exact zero
    if constexpr (!LB200_X_NOZERO) {
      if (!on) static_for<K::NCD>([&](auto ic) { H[decltype(ic)::value] = 0.0; });
    }

    const unsigned task = base + qg;
    const bool valid = lane_on && task < ntasks;
    if constexpr (FOCK) {
      if (p.prim_counter) count_primitives(p.prim_counter, (valid && rmeta.row == 0 && on) ? 1 : 0);
    }
    sync();   // last cross-term reads done: the phase-2 areas alias the cross-term area
    constexpr int TBOFF = K::OFF_B2;
#ifndef LB200_X_FINPACK
#define LB200_X_FINPACK 1
#endif
    // row stride of the final integrals of an (a 0|c d) class: padded (odd, conflict-free writes by the row lanes)
    // or, in store mode for the (d s| rows, dense -- the copy-out is then a linear copy of the quartet's block
    // without the two index divisions per element.  Measured per 2^20 quartets (profiles/r03_variants.txt):
    // (ds|dd) 0.881 -> 0.801 ms, (ds|dp) 0.529 -> 0.515, (ds|ds) 0.393 -> 0.380; with (p s| rows (three row lanes
    // per quartet, eight quartets per warp) the dense rows collide in the banks: (ps|dd) 0.391 -> 0.418 -- padded.
#ifndef LB200_X_FINPACK_MINAB_TMA
#define LB200_X_FINPACK_MINAB_TMA 3
#endif
#ifndef LB200_X_TMASTORE
#define LB200_X_TMASTORE 1
#endif
    // Dense rows also for the (p s| rows when (and only when) the block can leave by TMA (LB200_X_FINPACK_MINAB_TMA = 3;
    // 6 = off).  Measured per 2^20 quartets on one B200 (profiles/r04_variants_tmastore.txt), lane copy-out -> TMA:
    // (ds|dd) 0.805 -> 0.721 ms, (ds|dp) 0.517 -> 0.468, (ds|ds) 0.381 -> 0.353; padded rows + lane copy-out ->
    // dense rows + TMA: (ps|dd) 0.394 -> 0.342, (ps|dp) 0.248 -> 0.232, (ps|ds) 0.162 -> 0.160.  One issuing lane per
    // group instead of one per quartet (LB200_X_TMASTORE = 2): 0.731 / 0.480 / 0.371 for the (ds| classes -- slower.
    constexpr bool TMA_SHAPE = LB200_X_TMASTORE && (K::NAB * K::NCD) % 2 == 0 && QSIZE % 2 == 0 && K::OFF_FIN % 2 == 0;
    constexpr bool FINPACK = LB200_X_FINPACK && LB == 0 && !TR && !FOCK &&
                             (K::NAB >= 6 || (TMA_SHAPE && K::NAB >= LB200_X_FINPACK_MINAB_TMA));
    constexpr int FCS = FINPACK ? K::NCD : K::CS;
    // LB200_X_TMASTORE: the dense block of a FINPACK class leaves through one TMA bulk copy per quartet
    // (cp.async.bulk shared -> global, issued by the quartet's row-0 lane) instead of the lanes' LDS + STG loop;
    // it drains while the Boys lanes evaluate F_m of the next round.  Needs 16-byte aligned source / destination /
    // size: the block is an even number of doubles, the stage offsets are even, the output pointer is checked.
    constexpr bool TMASTORE = FINPACK && TMA_SHAPE;
    const bool tma_out = TMASTORE && (reinterpret_cast<size_t>(p.out) & 15) == 0;
    if constexpr (LB > 0) {
      if (valid && rmeta.row >= K::ROW0)
        static_for<K::NCD>([&](auto ic) {
          Q[TBOFF + decltype(ic)::value * K::RTP + (rmeta.row - K::ROW0)] = H[decltype(ic)::value];
        });
    } else {
      if (valid && rmeta.row >= K::ROW0)
        static_for<K::NCD>([&](auto ic) {
          Q[K::OFF_FIN + (rmeta.row - K::ROW0) * FCS + decltype(ic)::value] = H[decltype(ic)::value];
        });
    }
    if constexpr (TMASTORE) {
      if (tma_out) bulk_store_fence();   // this lane's part of the final block -> visible to the async proxy
    }
    cp_async_wait_all();   // own copies of the next round's records have landed ...
    sync();                // ... and everybody else's; transposed rows visible

    // ---- next round, Boys lanes: T, pfac, 1/(zeta+eta), rho; touch the table row ----------
    const BoysState bn = boys_prepare(onext, SN, tk_next.lnp);
    const double touched = boys_touch(bn);

    if constexpr (LB > 0 && !LB200_DIAG_NOBRAHRR) {
      // ---- bra HRR in registers: (a b|c d) from (e 0|c d), hrr.h:246 -----------------------
      for (int item = gl; item < QPG * K::NCD; item += GROUP) {
        const int q2 = item / K::NCD, cd = item - q2 * K::NCD;
        if (base + q2 >= ntasks) continue;
        double* Q2 = smem + (size_t)(g * QPG + q2) * QSIZE;
        const double* S2 = Q2 + (round & 1) * PSTAGE;
        const double ABv[3] = {S2[K::S_AB], S2[K::S_AB + 1], S2[K::S_AB + 2]};
        double colin[K::NRT];
        static_for<K::NRT>([&](auto rc) {
          colin[decltype(rc)::value] = Q2[TBOFF + cd * K::RTP + decltype(rc)::value];
        });
        double O[K::NAB];
        rr_hrr_regs<LA, LB>(colin, ABv, O);
        if constexpr (!TR && !FOCK) {
#ifndef LB200_X_STOREPTR
#define LB200_X_STOREPTR 0
#endif
          if constexpr (LB200_DIAG_NOSTORE) {
            double t = 0.0;
            static_for<K::NAB>([&](auto ic) { t += O[decltype(ic)::value]; });
            if (t == 123.456) p.out[0] = t;
          } else if constexpr (LB200_X_STOREPTR) {   // one address, 36 immediate offsets
            double* __restrict__ o = p.out + (size_t)(base + q2) * (K::NAB * K::NCD) + cd;
            static_for<K::NAB>([&](auto ic) { o[decltype(ic)::value * K::NCD] = O[decltype(ic)::value]; });
          } else {
            double* __restrict__ o = p.out + (size_t)(base + q2) * (K::NAB * K::NCD);
            static_for<K::NAB>([&](auto ic) { o[decltype(ic)::value * K::NCD + cd] = O[decltype(ic)::value]; });
          }
        } else {
          double* fin = Q2 + K::OFF_FIN;
          static_for<K::NAB>([&](auto ic) { fin[decltype(ic)::value * K::CS + cd] = O[decltype(ic)::value]; });
        }
      }
      if constexpr (TR || FOCK) sync();
    }
    if constexpr (FOCK) {
      // ---- cart -> pure, then 6-way digestion, by the NEC lanes of each quartet ------------
      fock_digest<LA, LB, LC, LD, NEC, WL>(p, valid && on, rmeta.row, Q + K::OFF_FIN, K::CS,
                                           Q + K::OFF_B2, Q + K::OFF_D, ocur.ib, ocur.ik, deg_cur, bf4);
    } else {
      // ---- coalesced copy-out (see eri_rowreg.cuh) ----------------------------------------
      constexpr int BLK = K::NAB * K::NCD;
      const unsigned left = ntasks - base;
      const int nvalid = left < (unsigned)QPG ? (int)left : QPG;
      double* __restrict__ o = p.out + (size_t)base * BLK;
      if constexpr (!TR) {
        if constexpr (LB == 0) {
          if constexpr (TMASTORE) {
            if (tma_out) {
              if constexpr (LB200_X_TMASTORE == 2) {   // one lane of the group issues every quartet's copy
                if (gl == 0)
                  for (int q2 = 0; q2 < nvalid; ++q2)
                    bulk_store(o + (size_t)q2 * BLK, smem + (size_t)(g * QPG + q2) * QSIZE + K::OFF_FIN,
                               (unsigned)(BLK * sizeof(double)));
              } else {
                if (valid && rmeta.row == 0)
                  bulk_store(p.out + (size_t)task * BLK, Q + K::OFF_FIN, (unsigned)(BLK * sizeof(double)));
              }
            }
          }
          if constexpr (FINPACK) {
            if (!tma_out)
            for (int q2 = 0; q2 < nvalid; ++q2) {
              const double* __restrict__ src = smem + (size_t)(g * QPG + q2) * QSIZE + K::OFF_FIN;
              double* __restrict__ dst = o + (size_t)q2 * BLK;
              for (int i = gl; i < BLK; i += GROUP) dst[i] = src[i];
            }
          } else {
            for (int idx = gl; idx < nvalid * BLK; idx += GROUP) {
              const int q2 = idx / BLK, i = idx - q2 * BLK;
              const int ab = i / K::NCD, cd = i - ab * K::NCD;
              o[idx] = smem[(size_t)(g * QPG + q2) * QSIZE + K::OFF_FIN + ab * K::CS + cd];
            }
          }
        }
      } else {
        for (int idx = gl; idx < nvalid * BLK; idx += GROUP) {
          const int q2 = idx / BLK, i = idx - q2 * BLK;
          const int cd = i / K::NAB, ab = i - cd * K::NAB;
          o[idx] = smem[(size_t)(g * QPG + q2) * QSIZE + K::OFF_FIN + ab * K::CS + cd];
        }
      }
    }
    // ---- next round, Boys lanes: F_m(T) * pfac ---------------------------------------------
    {
      BoysState b2 = bn;
      // keep the touch alive without changing any value: (touched != touched) is false for
      // every finite table entry
      if (touched != touched) b2.pfac = touched;
      if constexpr (!LB200_DIAG_NOBOYS) boys_finish(b2, SN);
    }
    if constexpr (TMASTORE) {
      // the bulk copy has read its block: the next round's cross terms may overwrite it after the top barrier
      if (tma_out && (LB200_X_TMASTORE == 2 ? gl == 0 : (valid && rmeta.row == 0))) bulk_store_wait_read();
    }
    ocur = onext;
    deg_cur = tk_next.deg;
    tk_next = tk_next2;
  }
}

}  // namespace lb200
