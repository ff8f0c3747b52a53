// cart -> pure transform and 6-way J/K digestion of one shell quartet held in shared memory,
// executed by the T lanes that own the quartet; every thread of the sync group (warp if
// WARP_SYNC, else CTA) must call it, lanes of idle quartets with active == false.
//
// Reference: tests/hartree-fock/hartree-fock++.cc:1703-1743 (g_12 += D_34 v, g_34 += D_12 v,
// g_13 -= D_24 v/4, g_24 -= D_13 v/4, g_14 -= D_23 v/4, g_23 -= D_14 v/4, v = (12|34) deg)
// and the solid-harmonic transform Engine::compute2 applies before returning
// (engine.impl.h:1965-1985, solidharmonics.h:281-463; standard convention pure iff l >= 2).
#pragma once
#include "eri_kernel.cuh"

namespace lb200 {

#ifdef LB200_EXPERIMENT_NO_RED   // diagnostics only: contractions without the atomics
#define LB200_FOCK_RED(addr, val) do { if ((val) == -12345.678) *(addr) = (val); } while (0)
#else
#define LB200_FOCK_RED(addr, val) atomicAdd(addr, val)
#endif

// doubles of shared memory fock_digest needs for the staged density blocks (Cartesian upper bound)
// LB200_DIGEST_STAGE: 0 = the contractions read D from global memory (L1/L2), 1 = the quartet's lanes
// copy the six blocks into shared memory right before the contractions (measured: slower, the copy's
// latency is exposed once more), 2 = the copy is issued with cp.async at the top of the round, one
// integral evaluation ahead of its use, so the contractions find the blocks in shared memory.
#ifndef LB200_DIGEST_STAGE
#define LB200_DIGEST_STAGE 0
#endif
// no staging for thread-per-quartet kernels with an (fd| / (ff| unrolled side: 128 quartets per CTA
// already fill the shared memory there
template <int LA, int LB, int LC, int LD>
constexpr bool fock_stage_enabled() {
  return LB200_DIGEST_STAGE != 0 && !(LA + LB == 0 && LC + LD >= 5);
}
template <int LA, int LB, int LC, int LD>
constexpr int fock_dblock_doubles() {
  if (!fock_stage_enabled<LA, LB, LC, LD>()) return 0;
  return nc(LA) * nc(LB) + nc(LC) * nc(LD) + nc(LA) * nc(LC) + nc(LB) * nc(LD) + nc(LA) * nc(LD) +
         nc(LB) * nc(LC);
}

// Run-time purity path.  The compile-time digestion below is specialised for the standard
// convention (pure iff l >= 2, what BasisSet gives cc-pVXZ / def2 bases); a block whose shells
// deviate -- Cartesian d of 6-31G* (basis.h.in:368-386), BasisSet::set_pure(false), a pure p
// shell -- takes this generic routine: densify, one sparse pass per pure index
// (solidharmonics.h:281-463 in the order of engine.impl.h:1965-1985), then the same six
// contractions with run-time block sizes.  One non-inlined copy serves every class kernel.
struct GenericDigest {
  const double* D;
  double* F;
  int nbf;
  const int *rowptr, *col, *base;
  const double* val;
  int l[4], pure[4], bf[4];
};

static __device__ __noinline__ void fock_digest_generic(const GenericDigest g, bool active, int lane, int T,
                                                 bool warp_sync, double* __restrict__ fin, int CS,
                                                 double* __restrict__ buf2, double deg) {
  constexpr int RP = 2 * kMaxShellL + 2;
  int n[4] = {nc(g.l[0]), nc(g.l[1]), nc(g.l[2]), nc(g.l[3])};
  {
    const int NCD = n[2] * n[3], NAB = n[0] * n[1];
    if (active)
      for (int i = lane; i < NAB * NCD; i += T) buf2[i] = fin[(i / NCD) * CS + i % NCD];
    if (warp_sync) __syncwarp(); else __syncthreads();
  }
  double* cur = buf2;
  double* oth = fin;
  for (int ax = 0; ax < 4; ++ax) {
    if (!(g.pure[ax] && g.l[ax] > 0)) continue;
    const int L = g.l[ax], np = 2 * L + 1, nin = n[ax];
    int outer = 1, inner = 1;
    for (int x = 0; x < ax; ++x) outer *= n[x];
    for (int x = ax + 1; x < 4; ++x) inner *= n[x];
    const int b0 = g.base[L];
    if (active)
      for (int i = lane; i < outer * np * inner; i += T) {
        const int o = i / (np * inner), r = i - o * np * inner;
        const int m = r / inner, in = r - m * inner;
        double acc = 0.0;
        for (int k = g.rowptr[L * RP + m]; k < g.rowptr[L * RP + m + 1]; ++k)
          acc += g.val[b0 + k] * cur[(o * nin + g.col[b0 + k]) * inner + in];
        oth[i] = acc;
      }
    if (warp_sync) __syncwarp(); else __syncthreads();
    double* t = cur; cur = oth; oth = t;
    n[ax] = np;
  }
  if (!active) return;
  const int na = n[0], nb = n[1], nc_ = n[2], nd = n[3], nbf = g.nbf;
  const int bfa = g.bf[0], bfb = g.bf[1], bfc = g.bf[2], bfd = g.bf[3];
  auto I = [&](int a, int b, int c, int d) -> double { return cur[((a * nb + b) * nc_ + c) * nd + d]; };
  const double* __restrict__ D = g.D;
  double* __restrict__ F = g.F;
  const double kdeg = -0.25 * deg;
  for (int i = lane; i < na * nb; i += T) {  // F(a,b) += D(c,d) v
    const int a = i / nb, b = i - a * nb;
    double s = 0.0;
    for (int c = 0; c < nc_; ++c)
      for (int d = 0; d < nd; ++d) s += I(a, b, c, d) * D[(bfc + c) * nbf + bfd + d];
    atomicAdd(&F[(bfa + a) * nbf + bfb + b], s * deg);
  }
  for (int i = lane; i < nc_ * nd; i += T) {  // F(c,d) += D(a,b) v
    const int c = i / nd, d = i - c * nd;
    double s = 0.0;
    for (int a = 0; a < na; ++a)
      for (int b = 0; b < nb; ++b) s += I(a, b, c, d) * D[(bfa + a) * nbf + bfb + b];
    atomicAdd(&F[(bfc + c) * nbf + bfd + d], s * deg);
  }
  for (int i = lane; i < na * nc_; i += T) {  // F(a,c) -= 1/4 D(b,d) v
    const int a = i / nc_, c = i - a * nc_;
    double s = 0.0;
    for (int b = 0; b < nb; ++b)
      for (int d = 0; d < nd; ++d) s += I(a, b, c, d) * D[(bfb + b) * nbf + bfd + d];
    atomicAdd(&F[(bfa + a) * nbf + bfc + c], s * kdeg);
  }
  for (int i = lane; i < nb * nd; i += T) {  // F(b,d) -= 1/4 D(a,c) v
    const int b = i / nd, d = i - b * nd;
    double s = 0.0;
    for (int a = 0; a < na; ++a)
      for (int c = 0; c < nc_; ++c) s += I(a, b, c, d) * D[(bfa + a) * nbf + bfc + c];
    atomicAdd(&F[(bfb + b) * nbf + bfd + d], s * kdeg);
  }
  for (int i = lane; i < na * nd; i += T) {  // F(a,d) -= 1/4 D(b,c) v
    const int a = i / nd, d = i - a * nd;
    double s = 0.0;
    for (int b = 0; b < nb; ++b)
      for (int c = 0; c < nc_; ++c) s += I(a, b, c, d) * D[(bfb + b) * nbf + bfc + c];
    atomicAdd(&F[(bfa + a) * nbf + bfd + d], s * kdeg);
  }
  for (int i = lane; i < nb * nc_; i += T) {  // F(b,c) -= 1/4 D(a,d) v
    const int b = i / nc_, c = i - b * nc_;
    double s = 0.0;
    for (int a = 0; a < na; ++a)
      for (int d = 0; d < nd; ++d) s += I(a, b, c, d) * D[(bfa + a) * nbf + bfd + d];
    atomicAdd(&F[(bfb + b) * nbf + bfc + c], s * kdeg);
  }
}

template <int LA, int LB, int LC, int LD, int T, bool WARP_SYNC>
__device__ __forceinline__ void fock_digest(const EriParams& p, bool active, int lane,
                                            double* __restrict__ fin /* [NAB][CS] */, int,
                                            double* __restrict__ buf2 /* NAB*NCD dense */,
                                            double* __restrict__ dsm /* fock_dblock_doubles */, int ib,
                                            int ik, double deg, const int (&bf4)[4]) {
  constexpr int NA = nc(LA), NB = nc(LB), NC = nc(LC), ND = nc(LD), NCD = NC * ND, CS = NCD | 1;
  constexpr int PA_ = LA >= 2, PB_ = LB >= 2, PC_ = LC >= 2, PD_ = LD >= 2;
  constexpr int na = PA_ ? npure(LA) : NA, nb = PB_ ? npure(LB) : NB;
  constexpr int nc_ = PC_ ? npure(LC) : NC, nd = PD_ ? npure(LD) : ND;
  constexpr int NPASS = PA_ + PB_ + PC_ + PD_;
  {
    // grid-uniform: all pairs of a block share one purity pattern (PairBlock)
    auto std_ok = [](int l, int pure) { return l == 0 || (pure != 0) == (l >= 2); };
    if (!(std_ok(LA, p.bra.pure_a) && std_ok(LB, p.bra.pure_b) && std_ok(LC, p.ket.pure_a) &&
          std_ok(LD, p.ket.pure_b))) {
      GenericDigest g;
      g.D = p.D; g.F = p.F; g.nbf = p.nbf;
      g.rowptr = p.sph_rowptr; g.col = p.sph_col; g.base = p.sph_base; g.val = p.sph_val;
      g.l[0] = LA; g.l[1] = LB; g.l[2] = LC; g.l[3] = LD;
      g.pure[0] = p.bra.pure_a; g.pure[1] = p.bra.pure_b; g.pure[2] = p.ket.pure_a; g.pure[3] = p.ket.pure_b;
      g.bf[0] = bf4[0]; g.bf[1] = bf4[1]; g.bf[2] = bf4[2]; g.bf[3] = bf4[3];
      if constexpr (LB200_DIGEST_STAGE == 2) cp_async_wait_all();   // (nothing was issued; keeps the group count clean)
      fock_digest_generic(g, active, lane, T, WARP_SYNC, fin, CS, buf2, deg);
      return;
    }
  }
  auto fin_at = [&](int ab, int cd) -> double { return fin[ab * CS + cd]; };
  auto sync = [] {
    if constexpr (WARP_SYNC) __syncwarp(); else __syncthreads();
  };
  // pass n reads: n == 0 the strided HRR layout, n odd buf2, n even (>0) the dense fin region;
  // writes: n even buf2, n odd the fin region (dense)
  if constexpr (PA_) {
    if (active)
      pure_pass<T, LA, 1, NB * NCD>(
          lane,
          [&](int, int k, int in) {
            const int b = in / NCD, cd = in - b * NCD;
            return fin_at(k * NB + b, cd);
          },
          buf2);
    sync();
  }
  if constexpr (PB_) {
    constexpr int n = PA_;
    double* out = (n % 2 == 0) ? buf2 : fin;
    if (active) {
      if constexpr (n == 0) {
        pure_pass<T, LB, na, NCD>(lane, [&](int o, int k, int in) { return fin_at(o * NB + k, in); }, out);
      } else {
        const double* in_ = buf2;
        pure_pass<T, LB, na, NCD>(lane, [&](int o, int k, int in) { return in_[(o * NB + k) * NCD + in]; }, out);
      }
    }
    sync();
  }
  if constexpr (PC_) {
    constexpr int n = PA_ + PB_;
    double* out = (n % 2 == 0) ? buf2 : fin;
    if (active) {
      if constexpr (n == 0) {
        pure_pass<T, LC, na * nb, ND>(lane, [&](int o, int k, int in) { return fin_at(o, k * ND + in); }, out);
      } else {
        const double* in_ = (n % 2 == 1) ? buf2 : fin;
        pure_pass<T, LC, na * nb, ND>(lane, [&](int o, int k, int in) { return in_[(o * NC + k) * ND + in]; }, out);
      }
    }
    sync();
  }
  if constexpr (PD_) {
    constexpr int n = PA_ + PB_ + PC_;
    double* out = (n % 2 == 0) ? buf2 : fin;
    if (active) {
      if constexpr (n == 0) {
        pure_pass<T, LD, na * nb * nc_, 1>(
            lane,
            [&](int o, int k, int) {
              const int ab = o / NC, c = o - ab * NC;
              return fin_at(ab, c * ND + k);
            },
            out);
      } else {
        const double* in_ = (n % 2 == 1) ? buf2 : fin;
        pure_pass<T, LD, na * nb * nc_, 1>(lane, [&](int o, int k, int) { return in_[o * ND + k]; }, out);
      }
    }
    sync();
  }
  const double* cur = (NPASS % 2 == 1) ? buf2 : fin;
  auto I = [&](int a, int b, int c, int d) -> double {
    if constexpr (NPASS > 0)
      return cur[((a * nb + b) * nc_ + c) * nd + d];
    else
      return fin_at(a * nb + b, c * nd + d);
  };
  const int bfa = bf4[0], bfb = bf4[1], bfc = bf4[2], bfd = bf4[3];   // loaded with the task
  const int n = p.nbf;
  const double* __restrict__ D = p.D;
  double* __restrict__ F = p.F;
  // The six density blocks of the quartet are fetched ONCE, by the quartet's lanes together, into
  // shared memory: the contractions below read each D element |F block| times, and a scattered
  // 8-byte global load costs the L1 a wavefront per lane where a shared-memory read of the same
  // value is a broadcast (Fock-mode ncu: l1tex 90-96 % busy, FP64 pipe 7-15 %).
  constexpr bool STAGE = fock_stage_enabled<LA, LB, LC, LD>();
  constexpr bool ASYNC = STAGE && LB200_DIGEST_STAGE == 2;
  constexpr int O_AB = 0, O_CD = O_AB + na * nb, O_AC = O_CD + nc_ * nd, O_BD = O_AC + na * nc_,
                O_AD = O_BD + nb * nd, O_BC = O_AD + na * nd, NDB = O_BC + nb * nc_;
  if constexpr (ASYNC) {   // issued by fock_prefetch_density at the top of the round
    cp_async_wait_all();
    sync();
  } else if constexpr (STAGE) {
  if (active)
    for (int i = lane; i < NDB; i += T) {
      int r, c, r0, c0;
      if (i < O_CD) { r = i / nb; c = i - r * nb; r0 = bfa; c0 = bfb; }
      else if (i < O_AC) { const int k = i - O_CD; r = k / nd; c = k - r * nd; r0 = bfc; c0 = bfd; }
      else if (i < O_BD) { const int k = i - O_AC; r = k / nc_; c = k - r * nc_; r0 = bfa; c0 = bfc; }
      else if (i < O_AD) { const int k = i - O_BD; r = k / nd; c = k - r * nd; r0 = bfb; c0 = bfd; }
      else if (i < O_BC) { const int k = i - O_AD; r = k / nd; c = k - r * nd; r0 = bfa; c0 = bfd; }
      else { const int k = i - O_BC; r = k / nc_; c = k - r * nc_; r0 = bfb; c0 = bfc; }
      dsm[i] = __ldg(&D[(size_t)(r0 + r) * n + c0 + c]);
    }
  sync();
  }
  if (!active) return;
#ifdef LB200_EXPERIMENT_NO_DIGEST   // diagnostics only: integrals without the six contractions
  if (deg != -12345.0) return;
#endif
  // element (r, c) of a density block: from the staged copy, or straight from global memory
  auto Dab = [&](int a, int b) { return STAGE ? dsm[O_AB + a * nb + b] : __ldg(&D[(size_t)(bfa + a) * n + bfb + b]); };
  auto Dcd = [&](int c, int d) { return STAGE ? dsm[O_CD + c * nd + d] : __ldg(&D[(size_t)(bfc + c) * n + bfd + d]); };
  auto Dac = [&](int a, int c) { return STAGE ? dsm[O_AC + a * nc_ + c] : __ldg(&D[(size_t)(bfa + a) * n + bfc + c]); };
  auto Dbd = [&](int b, int d) { return STAGE ? dsm[O_BD + b * nd + d] : __ldg(&D[(size_t)(bfb + b) * n + bfd + d]); };
  auto Dad = [&](int a, int d) { return STAGE ? dsm[O_AD + a * nd + d] : __ldg(&D[(size_t)(bfa + a) * n + bfd + d]); };
  auto Dbc = [&](int b, int c) { return STAGE ? dsm[O_BC + b * nc_ + c] : __ldg(&D[(size_t)(bfb + b) * n + bfc + c]); };
  // Two phases: first every contraction of this lane (loads of D and of the integrals, FMAs -- all
  // independent, so their latencies overlap), then every atomic.  With one loop per block and its
  // atomics in between, the loads of a block queued up behind the atomics of the previous one and
  // each of the six blocks paid a full memory latency (Fock-mode ncu: the first FMA of every block
  // waits on the long scoreboard).  Falls back to block-by-block when the partial sums of one lane
  // would not fit in registers.
  constexpr int KAB = (na * nb + T - 1) / T, KCD = (nc_ * nd + T - 1) / T, KAC = (na * nc_ + T - 1) / T,
                KBD = (nb * nd + T - 1) / T, KAD = (na * nd + T - 1) / T, KBC = (nb * nc_ + T - 1) / T;
#ifndef LB200_DIGEST_TWO_PHASE_MAX
#define LB200_DIGEST_TWO_PHASE_MAX 40
#endif
  constexpr bool TWO_PHASE = KAB + KCD + KAC + KBD + KAD + KBC <= LB200_DIGEST_TWO_PHASE_MAX;
  const double kdeg = -0.25 * deg;
  auto f_ab = [&](int i) {   // F(a,b) += D(c,d) v
    const int a = i / nb, b = i - a * nb;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < nc_; ++c)
#pragma unroll
      for (int d = 0; d < nd; ++d) s += I(a, b, c, d) * Dcd(c, d);
    return s * deg;
  };
  auto f_cd = [&](int i) {   // F(c,d) += D(a,b) v
    const int c = i / nd, d = i - c * nd;
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < na; ++a)
#pragma unroll
      for (int b = 0; b < nb; ++b) s += I(a, b, c, d) * Dab(a, b);
    return s * deg;
  };
  auto f_ac = [&](int i) {   // F(a,c) -= 1/4 D(b,d) v
    const int a = i / nc_, c = i - a * nc_;
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < nb; ++b)
#pragma unroll
      for (int d = 0; d < nd; ++d) s += I(a, b, c, d) * Dbd(b, d);
    return s * kdeg;
  };
  auto f_bd = [&](int i) {   // F(b,d) -= 1/4 D(a,c) v
    const int b = i / nd, d = i - b * nd;
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < na; ++a)
#pragma unroll
      for (int c = 0; c < nc_; ++c) s += I(a, b, c, d) * Dac(a, c);
    return s * kdeg;
  };
  auto f_ad = [&](int i) {   // F(a,d) -= 1/4 D(b,c) v
    const int a = i / nd, d = i - a * nd;
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < nb; ++b)
#pragma unroll
      for (int c = 0; c < nc_; ++c) s += I(a, b, c, d) * Dbc(b, c);
    return s * kdeg;
  };
  auto f_bc = [&](int i) {   // F(b,c) -= 1/4 D(a,d) v
    const int b = i / nc_, c = i - b * nc_;
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < na; ++a)
#pragma unroll
      for (int d = 0; d < nd; ++d) s += I(a, b, c, d) * Dad(a, d);
    return s * kdeg;
  };
  auto p_ab = [&](int i) { const int a = i / nb, b = i - a * nb; return &F[(size_t)(bfa + a) * n + bfb + b]; };
  auto p_cd = [&](int i) { const int c = i / nd, d = i - c * nd; return &F[(size_t)(bfc + c) * n + bfd + d]; };
  auto p_ac = [&](int i) { const int a = i / nc_, c = i - a * nc_; return &F[(size_t)(bfa + a) * n + bfc + c]; };
  auto p_bd = [&](int i) { const int b = i / nd, d = i - b * nd; return &F[(size_t)(bfb + b) * n + bfd + d]; };
  auto p_ad = [&](int i) { const int a = i / nd, d = i - a * nd; return &F[(size_t)(bfa + a) * n + bfd + d]; };
  auto p_bc = [&](int i) { const int b = i / nc_, c = i - b * nc_; return &F[(size_t)(bfb + b) * n + bfc + c]; };
  if constexpr (TWO_PHASE) {
    double sab[KAB], scd[KCD], sac[KAC], sbd[KBD], sad[KAD], sbc[KBC];
    static_for<KAB>([&](auto k) { const int i = lane + decltype(k)::value * T; sab[decltype(k)::value] = i < na * nb ? f_ab(i) : 0.0; });
    static_for<KCD>([&](auto k) { const int i = lane + decltype(k)::value * T; scd[decltype(k)::value] = i < nc_ * nd ? f_cd(i) : 0.0; });
    static_for<KAC>([&](auto k) { const int i = lane + decltype(k)::value * T; sac[decltype(k)::value] = i < na * nc_ ? f_ac(i) : 0.0; });
    static_for<KBD>([&](auto k) { const int i = lane + decltype(k)::value * T; sbd[decltype(k)::value] = i < nb * nd ? f_bd(i) : 0.0; });
    static_for<KAD>([&](auto k) { const int i = lane + decltype(k)::value * T; sad[decltype(k)::value] = i < na * nd ? f_ad(i) : 0.0; });
    static_for<KBC>([&](auto k) { const int i = lane + decltype(k)::value * T; sbc[decltype(k)::value] = i < nb * nc_ ? f_bc(i) : 0.0; });
    static_for<KAB>([&](auto k) { const int i = lane + decltype(k)::value * T; if (i < na * nb) LB200_FOCK_RED(p_ab(i), sab[decltype(k)::value]); });
    static_for<KCD>([&](auto k) { const int i = lane + decltype(k)::value * T; if (i < nc_ * nd) LB200_FOCK_RED(p_cd(i), scd[decltype(k)::value]); });
    static_for<KAC>([&](auto k) { const int i = lane + decltype(k)::value * T; if (i < na * nc_) LB200_FOCK_RED(p_ac(i), sac[decltype(k)::value]); });
    static_for<KBD>([&](auto k) { const int i = lane + decltype(k)::value * T; if (i < nb * nd) LB200_FOCK_RED(p_bd(i), sbd[decltype(k)::value]); });
    static_for<KAD>([&](auto k) { const int i = lane + decltype(k)::value * T; if (i < na * nd) LB200_FOCK_RED(p_ad(i), sad[decltype(k)::value]); });
    static_for<KBC>([&](auto k) { const int i = lane + decltype(k)::value * T; if (i < nb * nc_) LB200_FOCK_RED(p_bc(i), sbc[decltype(k)::value]); });
  } else {
    for (int i = lane; i < na * nb; i += T) LB200_FOCK_RED(p_ab(i), f_ab(i));
    for (int i = lane; i < nc_ * nd; i += T) LB200_FOCK_RED(p_cd(i), f_cd(i));
    for (int i = lane; i < na * nc_; i += T) LB200_FOCK_RED(p_ac(i), f_ac(i));
    for (int i = lane; i < nb * nd; i += T) LB200_FOCK_RED(p_bd(i), f_bd(i));
    for (int i = lane; i < na * nd; i += T) LB200_FOCK_RED(p_ad(i), f_ad(i));
    for (int i = lane; i < nb * nc_; i += T) LB200_FOCK_RED(p_bc(i), f_bc(i));
  }
}

// Issues the cp.async copies of the quartet's six density blocks into `dsm` (layout as in fock_digest:
// ab | cd | ac | bd | ad | bc, pure sizes of the standard convention).  Called by every lane of the
// quartet at the top of a round; completion is awaited inside fock_digest.  Blocks with a
// non-standard purity pattern are digested by fock_digest_generic straight from global memory.
template <int LA, int LB, int LC, int LD, int T>
__device__ __forceinline__ void fock_prefetch_density(const EriParams& p, bool active, int lane,
                                                      double* __restrict__ dsm, const int (&bf4)[4]) {
#if LB200_DIGEST_STAGE == 2
  if constexpr (!fock_stage_enabled<LA, LB, LC, LD>()) return;
  constexpr int NA = nc(LA), NB = nc(LB), NC = nc(LC), ND = nc(LD);
  constexpr int na = LA >= 2 ? npure(LA) : NA, nb = LB >= 2 ? npure(LB) : NB;
  constexpr int nc_ = LC >= 2 ? npure(LC) : NC, nd = LD >= 2 ? npure(LD) : ND;
  constexpr int O_CD = na * nb, O_AC = O_CD + nc_ * nd, O_BD = O_AC + na * nc_, O_AD = O_BD + nb * nd,
                O_BC = O_AD + na * nd, NDB = O_BC + nb * nc_;
  auto std_ok = [](int l, int pure) { return l == 0 || (pure != 0) == (l >= 2); };
  if (!(std_ok(LA, p.bra.pure_a) && std_ok(LB, p.bra.pure_b) && std_ok(LC, p.ket.pure_a) &&
        std_ok(LD, p.ket.pure_b)))
    return;
  if (!active) return;
  const double* __restrict__ D = p.D;
  const size_t n = (size_t)p.nbf;
  for (int i = lane; i < NDB; i += T) {
    int r, c, r0, c0;
    if (i < O_CD) { r = i / nb; c = i - r * nb; r0 = bf4[0]; c0 = bf4[1]; }
    else if (i < O_AC) { const int k = i - O_CD; r = k / nd; c = k - r * nd; r0 = bf4[2]; c0 = bf4[3]; }
    else if (i < O_BD) { const int k = i - O_AC; r = k / nc_; c = k - r * nc_; r0 = bf4[0]; c0 = bf4[2]; }
    else if (i < O_AD) { const int k = i - O_BD; r = k / nd; c = k - r * nd; r0 = bf4[1]; c0 = bf4[3]; }
    else if (i < O_BC) { const int k = i - O_AD; r = k / nd; c = k - r * nd; r0 = bf4[0]; c0 = bf4[3]; }
    else { const int k = i - O_BC; r = k / nc_; c = k - r * nc_; r0 = bf4[1]; c0 = bf4[2]; }
    cp_async8(dsm + i, D + (size_t)(r0 + r) * n + c0 + c);
  }
#endif
}

}  // namespace lb200
