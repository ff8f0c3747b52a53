// TEST INFRASTRUCTURE (not part of the product): compiles the __host__ __device__ core of the GPU one-body force
// kernel (libint_b200/csrc/onebody_deriv.cuh) with g++, so that the arithmetic of a kernel that can only run on the
// GPU box is checked in the CPU suite against the numpy derivative integrals (libint_b200/onebody.py), which in turn
// reproduce the reference's golden forces (tests/test_scf.py).  Built by tests/test_deriv_host.py into a temporary
// shared object; nothing under libint_b200/ links or loads it.
#include "onebody_deriv.cuh"

namespace {
struct HostAcc {
  double* F;
  int n3;
  void add(int which, int idx, double v) { F[which * n3 + idx] += v; }
};
}  // namespace

// Dc, Wc: densities in the Cartesian functions of every shell, [nbfc][nbfc]; F: [2][3 natom] (F1, F_Pulay), zeroed here
extern "C" void ob1_host_forces(int nshell, const int* l, const int* nprim, const int* off, const double* O,
                                const double* alpha, const double* coeff, const int* shell2cbf, const int* shell2atom,
                                int natom, const double* charges, const double* Dc, const double* Wc, int nbfc, double* F) {
  using namespace lb200::ob1;
  HostAcc acc{F, 3 * natom};
  for (int i = 0; i < 6 * natom; ++i) F[i] = 0.0;
  for (int a = 0; a < nshell; ++a)
    for (int b = 0; b <= a; ++b) {
      const int na = (l[a] + 1) * (l[a] + 2) / 2, nb = (l[b] + 1) * (l[b] + 2) / 2;
      double wD[kNC * kNC], wW[kNC * kNC];
      for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) {
          const long ab = (long)(shell2cbf[a] + i) * nbfc + shell2cbf[b] + j, ba = (long)(shell2cbf[b] + j) * nbfc + shell2cbf[a] + i;
          wD[i * nb + j] = a == b ? Dc[ab] : Dc[ab] + Dc[ba];
          wW[i * nb + j] = a == b ? Wc[ab] : Wc[ab] + Wc[ba];
        }
      pair_forces(l[a], l[b], O + 3 * a, O + 3 * b, nprim[a], alpha + off[a], coeff + off[a], nprim[b], alpha + off[b],
                  coeff + off[b], wD, wW, shell2atom[a], shell2atom[b], natom, charges, acc);
    }
}
