// Test driver for include/libint_b200_basis.hpp (host only, no GPU): builds libint_b200::BasisSet(name, atoms) for the
// water molecule of the reference's tests and prints what libint2::BasisSet exposes (basis.h.in:257-333), for
// tests/test_cxx_api.py to compare with the Python mirror (itself checked against the reference's own reader).
//   basis_api_driver geometry.xyz name [name ...]      (LIBINT_B200_DATA_PATH points at libint_b200/data)
#include <cstdio>
#include <fstream>

#include "libint_b200_basis.hpp"

int main(int argc, char** argv) {
  using namespace libint_b200;
  try {
    std::ifstream xyz(argv[1]);
    const std::vector<Atom> atoms = read_dotxyz(xyz);
    std::printf("natoms %zu Z", atoms.size());
    for (const Atom& a : atoms) std::printf(" %d", a.atomic_number);
    std::printf(" x0 %.17g\n", atoms[1].x);
    for (int k = 2; k < argc; ++k) {
      BasisSet bs(argv[k], atoms);
      std::printf("%s nshell %zu nbf %ld max_nprim %zu max_l %ld\n", argv[k], bs.size(), bs.nbf(), bs.max_nprim(), bs.max_l());
      std::printf("shell2bf");
      for (size_t x : bs.shell2bf()) std::printf(" %zu", x);
      std::printf("\nshell2atom");
      for (long x : bs.shell2atom(atoms)) std::printf(" %ld", x);
      std::printf("\natom2shell");
      for (const auto& v : bs.atom2shell(atoms)) std::printf(" %zu", v.size());
      std::printf("\npure");
      for (const Shell& s : bs) std::printf(" %d", s.pure ? 1 : 0);
      double csum = 0;
      for (const Shell& s : bs)
        for (double c : s.coeff) csum += c;
      std::printf("\ncoeffsum %.15g\n", csum);
      bs.set_pure(false);
      std::printf("cartesian nbf %ld\n", bs.nbf());
      bs.set_pure(true);
      std::printf("solid nbf %ld\n", bs.nbf());
    }
    // error contracts: unknown basis file -> std::ios_base::failure; missing element with throw_if_no_match -> logic_error
    try { BasisSet("no-such-basis", atoms); std::printf("unknown: no throw\n"); }
    catch (const std::ios_base::failure&) { std::printf("unknown: ios_base::failure\n"); }
    std::vector<Atom> neon = {{10, 0., 0., 0.}};
    std::printf("quiet omit: %zu shells\n", BasisSet("sto-3g", std::vector<Atom>{{79, 0., 0., 0.}}).size());
    try { BasisSet("sto-3g", std::vector<Atom>{{79, 0., 0., 0.}}, true); std::printf("missing: no throw\n"); }
    catch (const std::logic_error&) { std::printf("missing: logic_error\n"); }
    (void)neon;
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 4;
  }
}
