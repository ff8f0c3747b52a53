// libint_b200.hpp -- header-only C++ mirror of the libint2 API surface of the Coulomb-ERI path,
// implemented on the C ABI of libint_b200.h (link with -llibint_b200).
//
// What it mirrors (evaleev/libint, file:line):
//   libint2::Shell                      include/libint2/shell.h:720-1012   (one contraction per shell,
//                                       as Engine::compute2 requires, engine.impl.h:1167-1171)
//   libint2::Operator / BraKet          include/libint2/engine.h:84-246, :256-266
//   libint2::Engine                     include/libint2/engine.h:503-526 (ctor), :787-791
//                                       (compute2), :731 (results), :809-826 (set_precision),
//                                       :893-916 (lmax_exceeded)
//   compute_2body_fock (driver)         tests/hartree-fock/hartree-fock++.cc:1574-1772
//
// Engine::compute2 here is the *correctness* path of INTEGRATION.md section 1: one shell set per
// call, one host->device and one device->host transfer; results()[0] points at a host buffer owned
// by the engine, valid until the next compute call, or is nullptr when every primitive pair was
// screened out (engine.impl.h:1781-1784).  Throughput code calls lb200_eri_batch /
// lb200_fock_build on task lists instead (FockBuilder below; libint_b200/df3c.py).
#ifndef LIBINT_B200_HPP
#define LIBINT_B200_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "libint_b200.h"

namespace libint_b200 {

class error : public std::runtime_error {
 public:
  using std::runtime_error::runtime_error;
};

/// Engine::lmax_exceeded (engine.h:893-916)
class lmax_exceeded : public std::logic_error {
 public:
  lmax_exceeded(const std::string& what, int lmax_limit, int lmax_requested)
      : std::logic_error("libint_b200: " + what + ": angular momentum " +
                         std::to_string(lmax_requested) + " exceeds the limit " +
                         std::to_string(lmax_limit)),
        lmax_limit_(lmax_limit), lmax_requested_(lmax_requested) {}
  int lmax_limit() const { return lmax_limit_; }
  int lmax_requested() const { return lmax_requested_; }

 private:
  int lmax_limit_, lmax_requested_;
};

/// libint2::Shell with a single contraction (shell.h:720).  `coeff` holds normalization-embedded
/// coefficients exactly as Shell::contr[0].coeff does after Shell::renorm().
struct Shell {
  std::vector<double> alpha;   ///< exponents
  std::vector<double> coeff;   ///< contraction coefficients (renorm()ed unless constructed raw)
  int l = 0;                   ///< angular momentum
  bool pure = false;           ///< solid harmonics if true, Cartesian otherwise
  std::array<double, 3> O{{0., 0., 0.}};

  Shell() = default;
  /// coefficients as in a basis-set file; embeds the normalization like the reference's ctor
  Shell(std::vector<double> exponents, int am, bool solid, std::vector<double> coefficients,
        std::array<double, 3> origin, bool embed_normalization = true,
        bool enforce_unit_normalization = true)
      : alpha(std::move(exponents)), coeff(std::move(coefficients)), l(am), pure(solid), O(origin) {
    if (alpha.size() != coeff.size()) throw std::invalid_argument("Shell: alpha/coeff size mismatch");
    if (embed_normalization && !alpha.empty()) {
      if (lb200_shell_renorm(l, (int)alpha.size(), alpha.data(), coeff.data(),
                             enforce_unit_normalization ? 1 : 0, nullptr) != LB200_OK)
        throw error("lb200_shell_renorm failed");
    }
  }
  /// Shell::unit(), shell.h:906-909
  static Shell unit() {
    Shell s;
    s.alpha = {0.0};
    s.coeff = {1.0};
    return s;
  }
  size_t nprim() const { return alpha.size(); }
  size_t cartesian_size() const { return (size_t)(l + 1) * (l + 2) / 2; }
  size_t size() const { return pure ? (size_t)(2 * l + 1) : cartesian_size(); }
};

enum class Operator { coulomb };
enum class BraKet { xx_xx, xs_xx, xs_xs };
enum class ScreeningMethod : int {   // values of shell.h:1041-1059
  Original = LB200_SCREEN_ORIGINAL,
  Conservative = LB200_SCREEN_CONSERVATIVE
};

namespace detail {
inline void check(lb200_context* ctx, int rc, const char* what) {
  if (rc == LB200_OK) return;
  const char* msg = ctx ? lb200_last_error(ctx) : nullptr;
  throw error(std::string(what) + " failed (" + std::to_string(rc) + ")" +
              (msg && *msg ? std::string(": ") + msg : std::string()));
}
struct flat_basis {
  std::vector<int> l, pure, nprim;
  std::vector<double> O, alpha, coeff;
  void add(const Shell& s) {
    l.push_back(s.l);
    pure.push_back(s.pure ? 1 : 0);
    nprim.push_back((int)s.nprim());
    O.insert(O.end(), s.O.begin(), s.O.end());
    alpha.insert(alpha.end(), s.alpha.begin(), s.alpha.end());
    coeff.insert(coeff.end(), s.coeff.begin(), s.coeff.end());
  }
  lb200_basis* upload(lb200_context* ctx) const {
    lb200_basis* bs = nullptr;
    check(ctx, lb200_basis_create(ctx, (int)l.size(), l.data(), pure.data(), nprim.data(), O.data(),
                                  alpha.data(), coeff.data(), &bs), "lb200_basis_create");
    return bs;
  }
};
}  // namespace detail

/// libint2::Engine for Operator::coulomb, derivative order 0
class Engine {
 public:
  using target_ptr_vec = std::vector<const double*>;

  Engine(Operator oper, size_t max_nprim, int max_l, int deriv_order = 0,
         double precision = std::numeric_limits<double>::epsilon(), BraKet braket = BraKet::xx_xx,
         ScreeningMethod screening = ScreeningMethod::Original, int device = 0)
      : max_nprim_(max_nprim), max_l_(max_l), precision_(precision), braket_(braket),
        screening_(screening), targets_(1, nullptr) {
    (void)oper;
    if (deriv_order != 0) throw std::invalid_argument("libint_b200::Engine: deriv_order must be 0");
    if (max_l > LB200_MAX_AM) throw lmax_exceeded("Engine", LB200_MAX_AM, max_l);
    detail::check(nullptr, lb200_context_create(device, &ctx_), "lb200_context_create");
  }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  ~Engine() { lb200_context_destroy(ctx_); }

  Engine& set_precision(double prec) { precision_ = prec; return *this; }
  double precision() const { return precision_; }
  Engine& set(BraKet bk) { braket_ = bk; return *this; }
  Engine& set(ScreeningMethod s) { screening_ = s; return *this; }
  const target_ptr_vec& results() const { return targets_; }
  lb200_context* context() const { return ctx_; }

  /// compute2<Operator::coulomb, BraKet::xx_xx, 0>(bra1, bra2, ket1, ket2)
  const target_ptr_vec& compute(const Shell& s1, const Shell& s2, const Shell& s3, const Shell& s4) {
    return compute2(s1, s2, s3, s4);
  }
  /// BraKet::xs_xx: the second bra shell is Shell::unit() (engine.impl.h:139-175)
  const target_ptr_vec& compute(const Shell& s1, const Shell& s3, const Shell& s4) {
    return compute2(s1, Shell::unit(), s3, s4);
  }
  const target_ptr_vec& compute(const Shell& s1, const Shell& s3) {
    return compute2(s1, Shell::unit(), s3, Shell::unit());
  }

  const target_ptr_vec& compute2(const Shell& bra1, const Shell& bra2, const Shell& ket1,
                                 const Shell& ket2) {
    const Shell* sh[4] = {&bra1, &bra2, &ket1, &ket2};
    for (const Shell* s : sh) {
      if (s->l > LB200_MAX_AM) throw lmax_exceeded("Engine::compute2", LB200_MAX_AM, s->l);
      if (s->nprim() > max_nprim_ && !(s->nprim() == 1 && s->alpha[0] == 0.0))
        throw std::invalid_argument("libint_b200::Engine: shell exceeds max_nprim");
    }
    detail::flat_basis fb;
    for (const Shell* s : sh) fb.add(*s);
    lb200_basis* bs = fb.upload(ctx_);
    // the library wants l(first) >= l(second) inside a pair; the swap is undone on the result,
    // as engine.impl.h:1988-2067 does for the CPU kernels
    int b0 = 0, b1 = 1, k0 = 2, k1 = 3;
    const bool swap_bra = fb.l[0] < fb.l[1], swap_ket = fb.l[2] < fb.l[3];
    if (swap_bra) std::swap(b0, b1);
    if (swap_ket) std::swap(k0, k1);
    const double ln_prec = precision_ > 0. ? std::log(precision_) : std::numeric_limits<double>::lowest();
    lb200_pairs *bra = nullptr, *ket = nullptr;
    int rc = lb200_pairs_create(ctx_, bs, bs, 1, &b0, &b1, (int)screening_, ln_prec, nullptr, &bra);
    if (rc == LB200_OK)
      rc = lb200_pairs_create(ctx_, bs, bs, 1, &k0, &k1, (int)screening_, ln_prec, nullptr, &ket);
    if (rc != LB200_OK) {
      lb200_pairs_destroy(bra);
      lb200_basis_destroy(bs);
      detail::check(ctx_, rc, "lb200_pairs_create");
    }
    long long ib[6], ik[6];
    lb200_pairs_info(bra, ib);
    lb200_pairs_info(ket, ik);
    const long long n = lb200_eri_block_size(bra, ket, 1);
    raw_.assign((size_t)n, 0.0);
    const int task[2] = {0, 0};
    rc = lb200_eri_batch(ctx_, bra, ket, 1, task, 0, (int)screening_, precision_, 1, raw_.data(), 0);
    lb200_pairs_destroy(bra);
    lb200_pairs_destroy(ket);
    lb200_basis_destroy(bs);
    if (rc == LB200_ERR_LMAX)
      throw lmax_exceeded("Engine::compute2", LB200_MAX_AM, *std::max_element(fb.l.begin(), fb.l.end()));
    detail::check(ctx_, rc, "lb200_eri_batch");
    if (ib[3] == 0 || ik[3] == 0) {   // every primitive pair screened out (engine.impl.h:1781-1784)
      targets_[0] = nullptr;
      return targets_;
    }
    const size_t n1 = bra1.size(), n2 = bra2.size(), n3 = ket1.size(), n4 = ket2.size();
    if (!swap_bra && !swap_ket) {
      targets_[0] = raw_.data();
      return targets_;
    }
    // raw_ is laid out (b_first, b_second, k_first, k_second) in the library's pair order
    result_.resize(raw_.size());
    const size_t m1 = swap_bra ? n2 : n1, m2 = swap_bra ? n1 : n2, m3 = swap_ket ? n4 : n3,
                 m4 = swap_ket ? n3 : n4;
    (void)m1;
    for (size_t a = 0; a < n1; ++a)
      for (size_t b = 0; b < n2; ++b)
        for (size_t c = 0; c < n3; ++c)
          for (size_t d = 0; d < n4; ++d) {
            const size_t i1 = swap_bra ? b : a, i2 = swap_bra ? a : b;
            const size_t i3 = swap_ket ? d : c, i4 = swap_ket ? c : d;
            result_[((a * n2 + b) * n3 + c) * n4 + d] = raw_[((i1 * m2 + i2) * m3 + i3) * m4 + i4];
          }
    targets_[0] = result_.data();
    return targets_;
  }

 private:
  lb200_context* ctx_ = nullptr;
  size_t max_nprim_;
  int max_l_;
  double precision_;
  BraKet braket_;
  ScreeningMethod screening_;
  target_ptr_vec targets_;
  std::vector<double> raw_, result_;
};

/// The direct-SCF two-electron builder: compute_shellpairs + compute_schwarz_ints once
/// (hartree-fock++.cc:1305-1436, :1230-1298), then compute_2body_fock per iteration (:1574-1772).
class FockBuilder {
 public:
  FockBuilder(const std::vector<Shell>& obs, double pair_threshold = 1e-12, int device = 0,
              int rank = 0, int nranks = 1)
      : rank_(rank), nranks_(nranks) {
    detail::check(nullptr, lb200_context_create(device, &ctx_), "lb200_context_create");
    detail::flat_basis fb;
    for (const Shell& s : obs) fb.add(s);
    bs_ = fb.upload(ctx_);
    nbf_ = lb200_basis_nbf(bs_);
    long long np = 0;
    detail::check(ctx_, lb200_significant_pairs(bs_, pair_threshold, nullptr, nullptr, 0, &np),
                  "lb200_significant_pairs");
    std::vector<int> s1((size_t)np), s2((size_t)np);
    detail::check(ctx_, lb200_significant_pairs(bs_, pair_threshold, s1.data(), s2.data(), np, &np),
                  "lb200_significant_pairs");
    detail::check(ctx_, lb200_fock_create(ctx_, bs_, np, s1.data(), s2.data(), &fock_), "lb200_fock_create");
  }
  FockBuilder(const FockBuilder&) = delete;
  FockBuilder& operator=(const FockBuilder&) = delete;
  ~FockBuilder() {
    lb200_fock_destroy(fock_);
    lb200_basis_destroy(bs_);
    lb200_context_destroy(ctx_);
  }
  int nbf() const { return nbf_; }
  /// G = J - K/2 contribution of this rank for density D (row-major nbf x nbf, host memory);
  /// with nranks > 1 the caller sums the partial G's (MPI_Allreduce / ncclAllReduce).
  std::vector<double> compute_2body_fock(const std::vector<double>& D, double precision) const {
    if ((long long)D.size() != (long long)nbf_ * nbf_) throw std::invalid_argument("FockBuilder: D size");
    std::vector<double> G(D.size());
    detail::check(ctx_, lb200_fock_build(fock_, D.data(), 0, precision, 1, rank_, nranks_, G.data(), 0, nullptr),
                  "lb200_fock_build");
    return G;
  }
  /// Two-body forces F2[3 * atom + xyz] = sum_ij G1[3 atom + xyz]_ij D_ij with G1 = compute_2body_fock_deriv<1>
  /// (hartree-fock++.cc:1775-2055, traced with D at :648-656); this rank's share.  shell2atom =
  /// BasisSet::shell2atom(atoms).  Throws lmax_exceeded when a raised derivative class has no kernel.
  std::vector<double> compute_2body_forces(const std::vector<double>& D, const std::vector<int>& shell2atom,
                                           int natoms, double precision, bool use_schwarz = true) const {
    if ((long long)D.size() != (long long)nbf_ * nbf_) throw std::invalid_argument("FockBuilder: D size");
    if ((int)shell2atom.size() != lb200_basis_nshell(bs_)) throw std::invalid_argument("FockBuilder: shell2atom size");
    std::vector<double> F2((size_t)3 * natoms);
    const int rc = lb200_fock_grad(fock_, D.data(), 0, precision, use_schwarz ? 1 : 0, rank_, nranks_, natoms,
                                   shell2atom.data(), F2.data(), nullptr);
    if (rc == LB200_ERR_LMAX) throw lmax_exceeded("FockBuilder::compute_2body_forces", LB200_MAX_AM - 1, LB200_MAX_AM);
    detail::check(ctx_, rc, "lb200_fock_grad");
    return F2;
  }
  /// S, T, V of the basis with point charges {Z, x, y, z} per atom: the three compute_1body_ints calls of
  /// hartree-fock++.cc:267-275 (lb200_onebody on the GPU); row-major nbf x nbf each
  std::array<std::vector<double>, 3> compute_1body_ints(const std::vector<std::array<double, 4>>& charges) const {
    std::array<std::vector<double>, 3> M;
    for (auto& m : M) m.assign((size_t)nbf_ * nbf_, 0.0);
    std::vector<double> ch;
    for (const auto& c : charges) ch.insert(ch.end(), c.begin(), c.end());
    detail::check(ctx_, lb200_onebody(ctx_, bs_, (int)charges.size(), ch.data(), M[0].data(), M[1].data(),
                                      M[2].data(), 0), "lb200_onebody");
    return M;
  }
  /// One-body and Pulay contributions to the forces, {F1, F_Pulay}, 3 * natoms each (hartree-fock++.cc:601-627):
  /// F1 = 2 sum (T1 + V1) o D, F_Pulay = -2 sum S1 o W with the first-derivative integrals of
  /// compute_1body_ints_deriv (:1154-1228) contracted on the GPU (lb200_onebody_forces); W = C_occ eps_occ C_occ^T.
  std::array<std::vector<double>, 2> compute_1body_forces(const std::vector<std::array<double, 4>>& charges,
                                                          const std::vector<int>& shell2atom,
                                                          const std::vector<double>& D,
                                                          const std::vector<double>& W) const {
    if ((long long)D.size() != (long long)nbf_ * nbf_ || W.size() != D.size())
      throw std::invalid_argument("FockBuilder: D / W size");
    if ((int)shell2atom.size() != lb200_basis_nshell(bs_)) throw std::invalid_argument("FockBuilder: shell2atom size");
    std::array<std::vector<double>, 2> F;
    for (auto& f : F) f.assign(3 * charges.size(), 0.0);
    std::vector<double> ch;
    for (const auto& c : charges) ch.insert(ch.end(), c.begin(), c.end());
    detail::check(ctx_, lb200_onebody_forces(ctx_, bs_, (int)charges.size(), ch.data(), shell2atom.data(), D.data(),
                                             W.data(), 0, F[0].data(), F[1].data()), "lb200_onebody_forces");
    return F;
  }
  /// the Schwarz matrix of compute_schwarz_ints (nshell x nshell)
  std::vector<double> schwarz() const {
    const int ns = lb200_basis_nshell(bs_);
    std::vector<double> K((size_t)ns * ns);
    detail::check(ctx_, lb200_fock_schwarz(fock_, K.data()), "lb200_fock_schwarz");
    return K;
  }

 private:
  lb200_context* ctx_ = nullptr;
  lb200_basis* bs_ = nullptr;
  lb200_fock* fock_ = nullptr;
  int nbf_ = 0, rank_, nranks_;
};

}  // namespace libint_b200

#endif  // LIBINT_B200_HPP
