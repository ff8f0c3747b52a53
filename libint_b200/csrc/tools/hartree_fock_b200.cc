// hartree-fock-b200: the reference's direct-SCF test driver (tests/hartree-fock/hartree-fock++.cc, main() :233-716)
// as a C++ host program on the B200 library -- plain C++17 above include/libint_b200.hpp (the header-only mirror
// of libint2::Shell / Engine / the Fock builder on the C ABI), no torch, no Python:
//
//   hartree-fock-b200 geometry.xyz basis.json [more-basis.json ...] [--codata2010] [--dump-basis]
//   hartree-fock-b200 geometry.xyz basis-name                         (e.g. aug-cc-pVDZ, the reference's default)
//
// geometry: XYZ file in Angstrom (libint2::read_dotxyz, atom.h:83-160); basis: packed files under
// libint_b200/data/basis (the reference's lib/basis/*.g94 re-packed by tools/pack_basis.py); several files are
// the components of one basis in the reference's sense (aug-cc-pVDZ = cc-pvdz.json augmentation-cc-pvdz.json,
// basis.h.in:388-400).
// Steps and conventions follow the reference: nuclear repulsion (:245-255), S/T/V (:267-275; here lb200_onebody
// on the GPU), conditioned orthogonalizer (:281-290,:1957-2006), core-Hamiltonian start, D = C_occ C_occ^T,
// E = sum D o (H + F) + E_nuc (:472), error ||FDS - SDF|| / n^2 (:476-477), libint2::DIIS (start 2, depth 5),
// the Fock precision schedule (:463-466), convergence 1e-12 (:412); the two-electron part of every Fock matrix is
// lb200_fock_build; the force block (:596-716) is lb200_onebody_forces + lb200_fock_grad.  Prints the lines the reference's
// validation scripts parse (hartree-fock++-validate.py:60-70,128-132).  Dense linear algebra of the (small) test
// systems is a cyclic Jacobi eigensolver on the host.  Exit code 3 = no usable GPU (there is no CPU fallback).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "libint_b200_basis.hpp"

namespace {

using Matrix = std::vector<double>;   // row-major n x n (or n x m where noted)
using libint_b200::Atom;       // libint2::Atom, read_dotxyz, BasisSet: include/libint_b200_basis.hpp
using libint_b200::BasisSet;

// ---- small dense linear algebra -------------------------------------------------------------------------------
Matrix matmul(const Matrix& A, const Matrix& B, int n, int k, int m, bool tA = false, bool tB = false) {
  Matrix C((size_t)n * m, 0.0);   // C[n x m] = op(A)[n x k] op(B)[k x m]
  for (int i = 0; i < n; ++i)
    for (int p = 0; p < k; ++p) {
      const double a = tA ? A[(size_t)p * n + i] : A[(size_t)i * k + p];
      if (a == 0.0) continue;
      for (int j = 0; j < m; ++j) C[(size_t)i * m + j] += a * (tB ? B[(size_t)j * k + p] : B[(size_t)p * m + j]);
    }
  return C;
}

// cyclic Jacobi: A (n x n, symmetric) = V diag(w) V^T, eigenvalues ascending
void jacobi_eigh(Matrix A, int n, std::vector<double>& w, Matrix& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (std::abs(apq) < 1e-300) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {   // columns p, q
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {   // rows p, q
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> ord(n);
  for (int i = 0; i < n; ++i) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
  w.resize(n);
  Matrix Vs((size_t)n * n);
  for (int j = 0; j < n; ++j) {
    w[j] = A[(size_t)ord[j] * n + ord[j]];
    for (int i = 0; i < n; ++i) Vs[(size_t)i * n + j] = V[(size_t)i * n + ord[j]];
  }
  V.swap(Vs);
}

// solve B c = rhs (small, dense) by Gaussian elimination with partial pivoting; false if singular
bool solve(std::vector<double> B, std::vector<double> rhs, int n, std::vector<double>& c) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int i = k + 1; i < n; ++i)
      if (std::abs(B[(size_t)i * n + k]) > std::abs(B[(size_t)piv * n + k])) piv = i;
    if (std::abs(B[(size_t)piv * n + k]) < 1e-300) return false;
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(B[(size_t)k * n + j], B[(size_t)piv * n + j]);
      std::swap(rhs[k], rhs[piv]);
    }
    for (int i = k + 1; i < n; ++i) {
      const double f = B[(size_t)i * n + k] / B[(size_t)k * n + k];
      for (int j = k; j < n; ++j) B[(size_t)i * n + j] -= f * B[(size_t)k * n + j];
      rhs[i] -= f * rhs[k];
    }
  }
  c.assign(n, 0.0);
  for (int i = n - 1; i >= 0; --i) {
    double s = rhs[i];
    for (int j = i + 1; j < n; ++j) s -= B[(size_t)i * n + j] * c[j];
    c[i] = s / B[(size_t)i * n + i];
  }
  return true;
}

// libint2::DIIS (include/libint2/diis.h), start 2, depth 5
struct DIIS {
  std::vector<Matrix> x, e;
  int iter = 0;
  Matrix extrapolate(const Matrix& F, const Matrix& err) {
    ++iter;
    x.push_back(F);
    e.push_back(err);
    if (x.size() > 5) { x.erase(x.begin()); e.erase(e.begin()); }
    const int n = (int)x.size();
    if (iter < 2 || n < 2) return F;
    std::vector<double> B((size_t)(n + 1) * (n + 1), 0.0), rhs(n + 1, 0.0), c;
    double scale = 1e-300;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (size_t k = 0; k < err.size(); ++k) s += e[i][k] * e[j][k];
        B[(size_t)i * (n + 1) + j] = s;
        scale = std::max(scale, std::abs(s));
      }
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) B[(size_t)i * (n + 1) + j] /= scale;
      B[(size_t)i * (n + 1) + n] = B[(size_t)n * (n + 1) + i] = -1.0;
    }
    rhs[n] = -1.0;
    if (!solve(B, rhs, n + 1, c)) return F;
    Matrix out(F.size(), 0.0);
    for (int i = 0; i < n; ++i)
      for (size_t k = 0; k < F.size(); ++k) out[k] += c[i] * x[i][k];
    return out;
  }
};

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s geometry.xyz (basis-name | basis.json [more-basis.json ...]) [--codata2010]\n", argv[0]);
    return 2;
  }
  try {
    bool codata2010 = false, dump_basis = false;
    std::vector<std::string> basis_files;
    for (int a = 2; a < argc; ++a) {
      if (!std::strcmp(argv[a], "--codata2010")) codata2010 = true;
      else if (!std::strcmp(argv[a], "--dump-basis")) dump_basis = true;   // host-only: print the shells and stop
      else basis_files.push_back(argv[a]);
    }
    if (basis_files.empty()) throw std::runtime_error("no basis file given");
    const double b2a = codata2010 ? 0.52917721092 /* atom.h:63, hartree-fock.cc:306 */ : 0.529177210903 /* atom.h:53 */;
    std::ifstream xyz(argv[1]);
    if (!xyz) throw std::runtime_error(std::string("cannot open geometry file ") + argv[1]);
    const std::vector<Atom> atoms = libint_b200::read_dotxyz(xyz, b2a);
    // BasisSet(name, atoms) from explicit component files (basis.h.in:134-180); a missing element is an error here
    // or, like `hartree-fock++ geometry.xyz basis-name` (:236-244), a single basis NAME looked up under
    // $LIBINT_B200_DATA_PATH/basis (default: the data directory of the tree this driver was built in)
    const bool by_name = basis_files.size() == 1 && basis_files[0].find(".json") == std::string::npos;
    const BasisSet obs = by_name ? BasisSet(basis_files[0], atoms, /*throw_if_no_match=*/true)
                                 : BasisSet::from_files(basis_files, atoms, /*throw_if_no_match=*/true);
    std::vector<int> shell2atom;
    for (long a : obs.shell2atom(atoms)) shell2atom.push_back((int)a);
    if (dump_basis) {   // l pure nprim Ox Oy Oz, then exponents, then normalization-embedded coefficients
      std::printf("%zu\n", obs.size());
      for (size_t i = 0; i < obs.size(); ++i) {
        const auto& sh = obs[i];
        std::printf("%d %d %zu %.17g %.17g %.17g %d\n", sh.l, sh.pure ? 1 : 0, sh.nprim(), sh.O[0], sh.O[1], sh.O[2], shell2atom[i]);
        for (double x : sh.alpha) std::printf("%.17g ", x);
        std::printf("\n");
        for (double x : sh.coeff) std::printf("%.17g ", x);
        std::printf("\n");
      }
      return 0;
    }
    std::printf("Atomic Cartesian coordinates (a.u.):\n");
    int nelec = 0;
    for (const Atom& a : atoms) {
      std::printf("%d %.10f %.10f %.10f\n", a.atomic_number, a.x, a.y, a.z);
      nelec += a.atomic_number;
    }
    if (nelec % 2) throw std::runtime_error("RHF needs an even number of electrons");
    const int ndocc = nelec / 2;
    double enuc = 0.0;   // :245-255
    for (size_t i = 0; i < atoms.size(); ++i)
      for (size_t j = 0; j < i; ++j)
        enuc += atoms[i].atomic_number * atoms[j].atomic_number / std::sqrt(std::pow(atoms[i].x - atoms[j].x, 2) + std::pow(atoms[i].y - atoms[j].y, 2) +
                                                    std::pow(atoms[i].z - atoms[j].z, 2));

    libint_b200::FockBuilder fb(obs);
    const int n = fb.nbf();
    std::printf("orbital basis set rank = %d\n", n);
    std::printf("Nuclear repulsion energy = %.12f\n", enuc);
    std::vector<std::array<double, 4>> charges;
    for (const Atom& a : atoms) charges.push_back({{(double)a.atomic_number, a.x, a.y, a.z}});
    const auto STV = fb.compute_1body_ints(charges);
    const Matrix& S = STV[0];
    Matrix H((size_t)n * n);
    for (size_t k = 0; k < H.size(); ++k) H[k] = STV[1][k] + STV[2][k];

    // conditioning orthogonalizer (:281-290,:1957-2006): X = U w^-1/2 on the eigenvalues >= w_max / 1e8
    std::vector<double> w;
    Matrix U;
    jacobi_eigh(S, n, w, U);
    int first = 0;
    while (first < n && w[first] < w[n - 1] / 1e8) ++first;
    const int r = n - first;
    const double cond = w[n - 1] / w[first];
    Matrix X((size_t)n * r);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < r; ++j) X[(size_t)i * r + j] = U[(size_t)i * n + first + j] / std::sqrt(w[first + j]);

    Matrix C;
    std::vector<double> evals;
    auto density = [&](const Matrix& F) {
      const Matrix XtF = matmul(X, F, r, n, n, true);
      const Matrix Fp = matmul(XtF, X, r, n, r);
      Matrix Cp;
      jacobi_eigh(Fp, r, evals, Cp);
      C = matmul(X, Cp, n, r, r);
      Matrix D((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double s = 0.0;
          for (int o = 0; o < ndocc; ++o) s += C[(size_t)i * r + o] * C[(size_t)j * r + o];
          D[(size_t)i * n + j] = s;
        }
      return D;
    };

    Matrix D = density(H), F = H;
    DIIS diis;
    double ehf = 0.0, rms = 1.0, ediff_rel = 0.0;
    const double eps = std::numeric_limits<double>::epsilon();
    int it = 0;
    std::printf("\n\nIter         E(HF)                 D(E)/E         RMS([F,D])/nn\n");
    while (true) {
      ++it;
      const double ehf_last = ehf;
      const double precision = std::min(std::min(1e-3 / cond, 1e-7), std::max(rms / 1e4, eps));   // :463-466
      const Matrix G = fb.compute_2body_fock(D, precision);
      for (size_t k = 0; k < F.size(); ++k) F[k] = H[k] + G[k];
      ehf = 0.0;
      for (size_t k = 0; k < F.size(); ++k) ehf += D[k] * (H[k] + F[k]);
      ediff_rel = std::abs((ehf - ehf_last) / ehf);
      const Matrix FD = matmul(F, D, n, n, n), FDS = matmul(FD, S, n, n, n);
      Matrix comm((size_t)n * n);
      double nrm = 0.0;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          comm[(size_t)i * n + j] = FDS[(size_t)i * n + j] - FDS[(size_t)j * n + i];   // SDF = (FDS)^T
          nrm += comm[(size_t)i * n + j] * comm[(size_t)i * n + j];
        }
      rms = std::sqrt(nrm) / ((double)n * n);
      D = density(diis.extrapolate(F, comm));
      std::printf(" %02d %20.12f %20.12e %20.12e\n", it, ehf + enuc, ediff_rel, rms);
      if (!((ediff_rel > 1e-12 || rms > 1e-12) && it < 100)) break;
    }
    const bool converged = ediff_rel <= 1e-12 && rms <= 1e-12;
    std::printf("%s\n", converged ? "converged" : "NOT converged");
    std::printf("** Hartree-Fock energy = %20.12f\n", ehf + enuc);

    // the force block of the reference driver (:596-716), in its output format (parsed by
    // hartree-fock++-validate.py:128-132): one-body and Pulay parts from lb200_onebody_forces (:601-627), the two-body
    // part from lb200_fock_grad (compute_2body_fock_deriv<1> traced with D, :642-656), nuclear repulsion (:668-701)
    const size_t n3 = 3 * atoms.size();
    auto print_forces = [&](const char* key, const std::vector<double>& f) {
      std::printf("** %s forces = ", key);
      for (double v : f) std::printf("%.15g ", v);
      std::printf("\n");
    };
    Matrix W((size_t)n * n, 0.0);   // orbital-energy-weighted density C_occ eps_occ C_occ^T (:617-619)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int o = 0; o < ndocc; ++o) s += C[(size_t)i * r + o] * evals[o] * C[(size_t)j * r + o];
        W[(size_t)i * n + j] = s;
      }
    const auto F1P = fb.compute_1body_forces(charges, shell2atom, D, W);
    print_forces("1-body", F1P[0]);
    print_forces("Pulay", F1P[1]);
    std::vector<double> F2;
    try {
      F2 = fb.compute_2body_forces(D, shell2atom, (int)atoms.size(), eps, /*use_schwarz=*/false);
      print_forces("2-body", F2);
    } catch (const libint_b200::lmax_exceeded& e) {
      std::printf("2-body forces skipped: %s\n", e.what());
    }
    std::vector<double> FN(n3, 0.0);
    for (size_t a1 = 1; a1 < atoms.size(); ++a1)
      for (size_t a2 = 0; a2 < a1; ++a2) {
        const double d[3] = {atoms[a1].x - atoms[a2].x, atoms[a1].y - atoms[a2].y, atoms[a1].z - atoms[a2].z};
        const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const double f = atoms[a1].atomic_number * atoms[a2].atomic_number / (std::sqrt(r2) * r2);
        for (int k = 0; k < 3; ++k) { FN[3 * a1 + k] -= d[k] * f; FN[3 * a2 + k] += d[k] * f; }
      }
    print_forces("nuclear repulsion", FN);
    if (!F2.empty()) {
      std::vector<double> Ftot(n3);
      for (size_t k = 0; k < n3; ++k) Ftot[k] = F1P[0][k] + F1P[1][k] + F2[k] + FN[k];
      print_forces("Hartree-Fock", Ftot);
    }
    return converged ? 0 : 1;
  } catch (const libint_b200::error& e) {
    std::fprintf(stderr, "libint_b200::error: %s (there is no CPU fallback)\n", e.what());
    return std::strstr(e.what(), "(-2)") ? 3 : 4;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 4;
  }
}
