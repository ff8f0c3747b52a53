// Test driver for include/libint_b200.hpp (the C++ mirror of libint2::Engine / the direct-SCF
// Fock builder).  Reads a shell table written by tests/test_cxx_api.py, then
//   eri  i j k l ...   prints Engine::compute(shells[i], shells[j], shells[k], shells[l]) per quartet
//   eri3 i k l ...     BraKet::xs_xx
//   fock Dfile prec    prints compute_2body_fock(D)
//   forces Dfile prec natoms a0 a1 ...   prints compute_2body_forces(D, shell2atom)
// Exit code 3 = the library reported that no GPU is usable (there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "libint_b200.hpp"

using libint_b200::Shell;

static std::vector<Shell> read_shells(const char* path) {
  std::ifstream is(path);
  size_t n;
  is >> n;
  std::vector<Shell> out;
  for (size_t s = 0; s < n; ++s) {
    int l, pure, np;
    std::array<double, 3> O;
    is >> l >> pure >> np >> O[0] >> O[1] >> O[2];
    std::vector<double> a(np), c(np);
    for (auto& x : a) is >> x;
    for (auto& x : c) is >> x;
    out.emplace_back(a, l, pure != 0, c, O, /*embed_normalization=*/false);
  }
  if (!is) throw std::runtime_error("bad shell file");
  return out;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    const auto shells = read_shells(argv[1]);
    size_t max_nprim = 1;
    int max_l = 0;
    for (const auto& s : shells) {
      max_nprim = std::max(max_nprim, s.nprim());
      max_l = std::max(max_l, s.l);
    }
    std::printf("%s", "");
    if (!std::strcmp(argv[2], "eri") || !std::strcmp(argv[2], "eri3")) {
      const bool three = !std::strcmp(argv[2], "eri3");
      libint_b200::Engine engine(libint_b200::Operator::coulomb, max_nprim, max_l, 0, 0.0,
                                 three ? libint_b200::BraKet::xs_xx : libint_b200::BraKet::xx_xx);
      const int per = three ? 3 : 4;
      for (int a = 3; a + per <= argc; a += per) {
        int id[4];
        for (int k = 0; k < per; ++k) id[k] = std::atoi(argv[a + k]);
        const auto& res = three ? engine.compute(shells[id[0]], shells[id[1]], shells[id[2]])
                                : engine.compute(shells[id[0]], shells[id[1]], shells[id[2]], shells[id[3]]);
        size_t n = 1;
        for (int k = 0; k < per; ++k) n *= shells[id[k]].size();
        if (res[0] == nullptr) {
          std::printf("null\n");
          continue;
        }
        for (size_t i = 0; i < n; ++i) std::printf("%.17g ", res[0][i]);
        std::printf("\n");
      }
    } else if (!std::strcmp(argv[2], "fock")) {
      libint_b200::FockBuilder fb(shells);
      const int n = fb.nbf();
      std::vector<double> D((size_t)n * n);
      std::ifstream is(argv[3]);
      for (auto& x : D) is >> x;
      const auto G = fb.compute_2body_fock(D, std::atof(argv[4]));
      for (double g : G) std::printf("%.17g ", g);
      std::printf("\n");
    } else if (!std::strcmp(argv[2], "forces")) {
      libint_b200::FockBuilder fb(shells);
      const int n = fb.nbf();
      std::vector<double> D((size_t)n * n);
      std::ifstream is(argv[3]);
      for (auto& x : D) is >> x;
      const int natoms = std::atoi(argv[5]);
      std::vector<int> s2a;
      for (int a = 6; a < argc; ++a) s2a.push_back(std::atoi(argv[a]));
      const auto F2 = fb.compute_2body_forces(D, s2a, natoms, std::atof(argv[4]), /*use_schwarz=*/false);
      for (double g : F2) std::printf("%.17g ", g);
      std::printf("\n");
    } else if (!std::strcmp(argv[2], "lmax")) {
      try {
        libint_b200::Engine engine(libint_b200::Operator::coulomb, 1, LB200_MAX_AM + 1);
      } catch (const libint_b200::lmax_exceeded& e) {
        std::printf("lmax_exceeded %d %d\n", e.lmax_limit(), e.lmax_requested());
        return 0;
      }
      return 1;
    }
  } catch (const libint_b200::error& e) {
    std::fprintf(stderr, "libint_b200::error: %s\n", e.what());
    return std::strstr(e.what(), "(-2)") ? 3 : 4;
  }
  return 0;
}
