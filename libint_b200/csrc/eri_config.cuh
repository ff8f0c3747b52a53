// Launch configuration of the class kernels: team size T, warps per CTA, shared memory.
#pragma once
#include "eri_kernel.cuh"

namespace lb200 {

constexpr int kSmemLimit = 227 * 1024;  // usable dynamic shared memory per CTA / SM on sm_100

template <int LA, int LB, int LC, int LD, int MODE>
struct Cfg {
  using K = Cls<LA, LB, LC, LD>;
  static constexpr bool FOCK = MODE == kModeFock;
  static constexpr int TEAM_DOUBLES = K::buf_size(1, FOCK) + K::buf_size(0, FOCK);
  static constexpr int TEAM_BYTES = TEAM_DOUBLES * 8;
  static constexpr bool FITS = TEAM_BYTES <= kSmemLimit;

  // lane efficiency of a team size: rows are dealt to lanes round-robin, 32/T teams share a warp
  static constexpr double eff(int t) {
    const int lanes_used = t <= 32 ? (32 / t) * t : t;
    const int iters = (K::NEC + t - 1) / t;
    const double e = (double)K::NEC / (iters * t) * lanes_used / (t <= 32 ? 32 : t);
    // resident warps per SM given the shared-memory footprint
    const int tpw = t <= 32 ? 32 / t : 1;
    const int teams = kSmemLimit / TEAM_BYTES;
    if (teams < 1) return 0.0;
    double warps = t <= 32 ? (double)teams / tpw : (double)(t / 32) * cmin(teams, 1);
    if (t > 32 && teams >= 2) warps = (double)(t / 32) * cmin(teams, 32);
    const double occ = warps >= 8.0 ? 1.0 : warps / 8.0;
    return e * occ;
  }
  static constexpr int pick() {
    constexpr int cand[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20, 24, 32, 64, 128};
    int best = 32;
    double be = -1.0;
    for (int t : cand) {
      const double e = eff(t);
      if (e > be * 1.0001) { be = e; best = t; }
    }
    return best;
  }
  static constexpr int T = pick();
  static constexpr int TPW = T <= 32 ? 32 / T : 1;
  // warps per CTA (T <= 32) : as many as fit, at most 4
  static constexpr int WARPS =
      T <= 32 ? cmax(1, cmin(4, kSmemLimit / cmax(1, TPW * TEAM_BYTES))) : T / 32;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int TEAMS_PER_CTA = T <= 32 ? WARPS * TPW : 1;
  static constexpr int SMEM_BYTES = TEAMS_PER_CTA * TEAM_BYTES;
  static constexpr int CTAS_PER_SM = cmax(1, cmin(16, kSmemLimit / cmax(1, SMEM_BYTES + 1024)));
};

}  // namespace lb200
