// oracle/oracle_capi.cc -- TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// ctypes-friendly C entry points around the reference's own, unmodified,
// header-only C++ API (libint2::Engine, Shell, ShellPair, BasisSet,
// FmEval_Chebyshev7, and the closed-form eri() of src/bin/test_eri/eri.h),
// compiled from /root/reference/include where the headers lie, on top of the
// restated kernels of oracle_kernels.cc.  The Fock driver below restates the
// direct-SCF consumer of the path (tests/hartree-fock/hartree-fock++.cc); each
// function cites the lines it follows.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.

#include <libint2.hpp>
#include <libint2/boys.h>

#define LIBINT2_REF_REALTYPE double
#include <eri.h>  // /root/reference/src/bin/test_eri/eri.h

#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <vector>

using libint2::BraKet;
using libint2::Engine;
using libint2::Operator;
using libint2::ScreeningMethod;
using libint2::Shell;
using libint2::ShellPair;

namespace {

ScreeningMethod to_screening(int s) {
  switch (s) {
    case 0x0001: return ScreeningMethod::Original;
    case 0x0010: return ScreeningMethod::Conservative;
    case 0x0100: return ScreeningMethod::Schwarz;
    case 0x1000: return ScreeningMethod::SchwarzInf;
    default: return ScreeningMethod::Invalid;
  }
}

// flat shell description -> libint2::Shell
// coeff_is_raw != 0: coefficients as in the basis file, Shell ctor embeds the
//                    normalization (shell.h:958-999);
// coeff_is_raw == 0: coefficients already carry the normalization.
std::vector<Shell> make_shells(int nshell, const int* l, const int* pure, const int* nprim,
                               const double* O, const double* alpha, const double* coeff,
                               int coeff_is_raw) {
  std::vector<Shell> shells;
  shells.reserve(nshell);
  size_t off = 0;
  for (int s = 0; s < nshell; ++s) {
    libint2::svector<double> a(alpha + off, alpha + off + nprim[s]);
    libint2::svector<double> c(coeff + off, coeff + off + nprim[s]);
    libint2::svector<Shell::Contraction> contr;
    contr.push_back(Shell::Contraction{l[s], pure[s] != 0, c});
    shells.emplace_back(a, contr, std::array<double, 3>{{O[3 * s], O[3 * s + 1], O[3 * s + 2]}},
                        coeff_is_raw != 0);
    off += nprim[s];
  }
  return shells;
}

template <typename F>
void parallel_do(int nthreads, F&& f) {  // hartree-fock++.cc:187-206
  std::vector<std::thread> threads;
  for (int t = 1; t < nthreads; ++t) threads.emplace_back(f, t);
  f(0);
  for (auto& t : threads) t.join();
}

}  // namespace

extern "C" {

int lbo_init() {
  if (!libint2::initialized()) libint2::initialize();
  // libint2::initialized() lives in an inline (process-wide unique) singleton: when a second library
  // built from this file on another kernel library is loaded into the same process, initialize() is
  // skipped for it -- fill this library's own function tables explicitly (idempotent).
  static bool tables_filled = false;
  if (!tables_filled) {
    libint2_static_init();
    tables_filled = true;
  }
  return 0;
}

int lbo_set_unit_normalization(int flag) {  // shell.h:895-902
  Shell::do_enforce_unit_normalization(flag != 0);
  return 0;
}

// FmEval_Chebyshev7::eval, boys.h:345-454 (compiled without -mavx => Horner branch)
void lbo_boys_cheb7(double T, int mmax, int table_mmax, double* out) {
  auto ev = libint2::FmEval_Chebyshev7<double>::instance(table_mmax);
  ev->eval(out, T, mmax);
}

// FmEval_Reference2, boys.h:214-247
void lbo_boys_reference(double T, int mmax, double* out) {
  libint2::FmEval_Reference2<double>::eval(out, T, mmax);
}

// closed-form primitive ERI, src/bin/test_eri/eri.h:121-380
double lbo_eri_closed(const int* lmn, const double* alpha, const double* centers,
                      int norm_flag) {
  return eri(lmn[0], lmn[1], lmn[2], alpha[0], centers + 0, lmn[3], lmn[4], lmn[5], alpha[1],
             centers + 3, lmn[6], lmn[7], lmn[8], alpha[2], centers + 6, lmn[9], lmn[10],
             lmn[11], alpha[3], centers + 9, norm_flag);
}

// Shell ctor normalization (shell.h:958-999): returns embedded coefficients and
// max_ln_coeff
int lbo_shell_renorm(int l, int nprim, const double* alpha, const double* coeff,
                     double* out_coeff, double* out_max_ln_coeff) {
  int pure = 0;
  double O[3] = {0, 0, 0};
  auto sh = make_shells(1, &l, &pure, &nprim, O, alpha, coeff, 1);
  for (int p = 0; p < nprim; ++p) {
    out_coeff[p] = sh[0].contr[0].coeff[p];
    out_max_ln_coeff[p] = sh[0].max_ln_coeff[p];
  }
  return 0;
}

// solid-harmonic coefficient, solidharmonics.h:114-174
double lbo_solidharmonic_coeff(int l, int m, int lx, int ly, int lz) {
  return libint2::solidharmonics::SolidHarmonicsCoefficients<double>::coeff(l, m, lx, ly, lz);
}

// One shell set through the reference Engine.
//   braket: 0 = xx_xx (4 shells), 1 = xs_xx (3 shells: bra1, ket1, ket2), 2 = xs_xs (2 shells)
//   returns number of doubles written, 0 if the whole set was screened out
//   (results()[0] == nullptr, engine.impl.h:1781-1784), <0 on error
long lbo_compute2(int braket, const int* l, const int* pure, const int* nprim, const double* O,
                  const double* alpha, const double* coeff, int coeff_is_raw, int screening,
                  double precision, int uniform_cart_norm, double* out, long out_cap) {
  lbo_init();
  const int ns = braket == 0 ? 4 : (braket == 1 ? 3 : 2);
  auto sh = make_shells(ns, l, pure, nprim, O, alpha, coeff, coeff_is_raw);
  int max_nprim = 0, max_l = 0;
  for (auto& s : sh) {
    max_nprim = std::max<int>(max_nprim, s.nprim());
    max_l = std::max<int>(max_l, s.contr[0].l);
  }
  const BraKet bk = braket == 0 ? BraKet::xx_xx : (braket == 1 ? BraKet::xs_xx : BraKet::xs_xs);
  try {
    Engine engine(Operator::coulomb, max_nprim, max_l, 0, precision,
                  libint2::operator_traits<Operator::coulomb>::default_params(), bk,
                  to_screening(screening));
    if (uniform_cart_norm) engine.set(libint2::CartesianShellNormalization::uniform);
    const auto& buf = engine.results();
    size_t n = 1;
    for (auto& s : sh) n *= s.size();
    if (braket == 0)
      engine.compute(sh[0], sh[1], sh[2], sh[3]);
    else if (braket == 1)
      engine.compute(sh[0], sh[1], sh[2]);
    else
      engine.compute(sh[0], sh[1]);
    if (buf[0] == nullptr) return 0;
    if ((long)n > out_cap) return -2;
    std::memcpy(out, buf[0], n * sizeof(double));
    return (long)n;
  } catch (std::exception& e) {
    std::fprintf(stderr, "lbo_compute2: %s\n", e.what());
    return -1;
  }
}

// ShellPair::init with Original / Conservative screening (shell.h:1138-1256).
// Writes per primitive pair: P[3], K, one_over_gamma, nonsph_screen_fac, ln_scr, p1, p2
// (9 doubles each); returns the number of surviving primitive pairs; AB -> out_AB[3].
int lbo_shellpair(const int* l, const int* pure, const int* nprim, const double* O,
                  const double* alpha, const double* coeff, int coeff_is_raw, double ln_prec,
                  int screening, double* out, int out_cap_pairs, double* out_AB) {
  auto sh = make_shells(2, l, pure, nprim, O, alpha, coeff, coeff_is_raw);
  ShellPair sp(sh[0], sh[1], ln_prec, to_screening(screening));
  int n = (int)sp.primpairs.size();
  if (n > out_cap_pairs) return -2;
  for (int i = 0; i < n; ++i) {
    const auto& pp = sp.primpairs[i];
    double* o = out + 9 * i;
    o[0] = pp.P[0]; o[1] = pp.P[1]; o[2] = pp.P[2];
    o[3] = pp.K; o[4] = pp.one_over_gamma; o[5] = pp.nonsph_screen_fac; o[6] = pp.ln_scr;
    o[7] = pp.p1; o[8] = pp.p2;
  }
  for (int k = 0; k < 3; ++k) out_AB[k] = sp.AB[k];
  return n;
}

// -----------------------------------------------------------------------------
// Direct Fock build, restating tests/hartree-fock/hartree-fock++.cc
// -----------------------------------------------------------------------------
struct LboFock {
  std::vector<Shell> obs;
  std::vector<size_t> shell2bf;
  size_t nbf = 0;
  int max_nprim = 0, max_l = 0;
  int nthreads = 1;
  // significant shell pairs, splist[s1] sorted ascending (hartree-fock++.cc:1305-1381)
  std::vector<std::vector<size_t>> splist;
  std::vector<std::vector<std::shared_ptr<ShellPair>>> spdata;
  std::vector<double> K;  // Schwarz matrix nsh x nsh (hartree-fock++.cc:1230-1298)
  ScreeningMethod screening = ScreeningMethod::SchwarzInf;  // hartree-fock++.cc:67
};

// pair list given as npair (s1,s2) tuples with s1 >= s2 -- the overlap-based
// significance test (hartree-fock++.cc:1353-1361) needs 1-body kernels the
// oracle does not have, so the caller supplies the list
void* lbo_fock_create(int nshell, const int* l, const int* pure, const int* nprim, const double* O,
                      const double* alpha, const double* coeff, int coeff_is_raw, int npair,
                      const int* pair_s1, const int* pair_s2, int nthreads) {
  lbo_init();
  auto* f = new LboFock;
  f->obs = make_shells(nshell, l, pure, nprim, O, alpha, coeff, coeff_is_raw);
  f->nthreads = std::max(1, nthreads);
  f->shell2bf.resize(nshell);
  size_t n = 0;
  for (int s = 0; s < nshell; ++s) {
    f->shell2bf[s] = n;
    n += f->obs[s].size();
    f->max_nprim = std::max<int>(f->max_nprim, f->obs[s].nprim());
    f->max_l = std::max<int>(f->max_l, f->obs[s].contr[0].l);
  }
  f->nbf = n;
  f->splist.resize(nshell);
  for (int i = 0; i < npair; ++i) f->splist[pair_s1[i]].push_back(pair_s2[i]);
  for (auto& v : f->splist) std::sort(v.begin(), v.end());

  const int nthr = f->nthreads;
  // --- Schwarz matrix: K(s1,s2) = sqrt(||(s1 s2|s1 s2)||_inf), engine precision 0
  //     (hartree-fock++.cc:1244-1289)
  f->K.assign((size_t)nshell * nshell, 0.0);
  {
    std::vector<Engine> engines(nthr);
    engines[0] = Engine(Operator::coulomb, f->max_nprim, f->max_l, 0, 0.);
    for (int i = 1; i < nthr; ++i) engines[i] = engines[0];
    parallel_do(nthr, [&](int tid) {
      const auto& buf = engines[tid].results();
      for (long s1 = 0, s12 = 0; s1 != nshell; ++s1) {
        const auto n1 = f->obs[s1].size();
        for (long s2 = 0; s2 <= s1; ++s2, ++s12) {
          if (s12 % nthr != tid) continue;
          const auto n2 = f->obs[s2].size();
          const auto n12 = n1 * n2;
          engines[tid].compute2<Operator::coulomb, BraKet::xx_xx, 0>(f->obs[s1], f->obs[s2],
                                                                     f->obs[s1], f->obs[s2]);
          // lpNorm<Infinity> of the n12 x n12 matrix = max |element|
          double nrm = 0;
          for (size_t i = 0; i < n12 * n12; ++i) nrm = std::max(nrm, std::abs(buf[0][i]));
          f->K[s1 * nshell + s2] = f->K[s2 * nshell + s1] = std::sqrt(nrm);
        }
      }
    });
  }
  // --- shell-pair data with SchwarzInf primitive factors (hartree-fock++.cc:1383-1431)
  {
    const double max_engine_precision = std::numeric_limits<double>::epsilon() / 1e10;  // :64
    const auto ln_max_engine_precision = std::log(max_engine_precision);
    std::vector<Engine> engines(nthr);
    engines[0] = Engine(Operator::coulomb, f->max_nprim, f->max_l, 0);
    engines[0].set_precision(0.);
    for (int i = 1; i < nthr; ++i) engines[i] = engines[0];
    f->spdata.resize(nshell);
    parallel_do(nthr, [&](int tid) {
      auto schwarz_factor_evaluator = [&](const Shell& s1, size_t p1, const Shell& s2,
                                          size_t p2) -> double {
        auto& engine = engines[tid];
        auto& buf = engine.results();
        auto ps1 = s1.extract_primitive(p1, false);
        auto ps2 = s2.extract_primitive(p2, false);
        const auto n12 = ps1.size() * ps2.size();
        engine.compute(ps1, ps2, ps1, ps2);
        if (buf[0]) {
          double nrm = 0;
          for (size_t i = 0; i < n12 * n12; ++i) nrm = std::max(nrm, std::abs(buf[0][i]));
          return std::sqrt(nrm);
        } else
          return 0.;
      };
      for (long s1 = 0; s1 != nshell; ++s1) {
        if (s1 % nthr != tid) continue;
        for (const auto& s2 : f->splist[s1])
          f->spdata[s1].emplace_back(std::make_shared<ShellPair>(
              f->obs[s1], f->obs[s2], ln_max_engine_precision, f->screening,
              schwarz_factor_evaluator));
      }
    });
  }
  return f;
}

void lbo_fock_destroy(void* h) { delete static_cast<LboFock*>(h); }
long lbo_fock_nbf(void* h) { return (long)static_cast<LboFock*>(h)->nbf; }
void lbo_fock_schwarz(void* h, double* K) {
  auto* f = static_cast<LboFock*>(h);
  std::memcpy(K, f->K.data(), f->K.size() * sizeof(double));
}
// primitive pair data of pair (s1,s2): 9 doubles per primitive pair as in lbo_shellpair
int lbo_fock_pairdata(void* h, int s1, int s2, double* out, int cap) {
  auto* f = static_cast<LboFock*>(h);
  const auto& lst = f->splist[s1];
  auto it = std::find(lst.begin(), lst.end(), (size_t)s2);
  if (it == lst.end()) return -1;
  const auto& sp = *f->spdata[s1][it - lst.begin()];
  int n = (int)sp.primpairs.size();
  if (n > cap) return -2;
  for (int i = 0; i < n; ++i) {
    const auto& pp = sp.primpairs[i];
    double* o = out + 9 * i;
    o[0] = pp.P[0]; o[1] = pp.P[1]; o[2] = pp.P[2];
    o[3] = pp.K; o[4] = pp.one_over_gamma; o[5] = pp.nonsph_screen_fac; o[6] = pp.ln_scr;
    o[7] = pp.p1; o[8] = pp.p2;
  }
  return n;
}

// compute_2body_fock (hartree-fock++.cc:1574-1772): G = sym(sum deg * 6-way digestion).
// task_stride/task_offset restrict the build to quartets with
// (s1234 % task_stride) == task_offset *before* the thread round-robin, which is
// how bench.py samples a bounded subset of a large build (pass 1, 0 for all).
// stats[0] = #integrals computed, stats[1] = #shell quartets computed,
// stats[2] = wall seconds of the threaded region.
int lbo_fock_build(void* h, const double* D, double precision, int use_schwarz, long task_stride,
                   long task_offset, double* G_out, double* stats) {
  auto* f = static_cast<LboFock*>(h);
  const auto& obs = f->obs;
  const long nshells = (long)obs.size();
  const size_t n = f->nbf;
  const int nthr = f->nthreads;
  const auto& shell2bf = f->shell2bf;
  std::vector<std::vector<double>> G(nthr, std::vector<double>(n * n, 0.0));

  // compute_shellblock_norm (hartree-fock++.cc:939-957): inf-norm of shell blocks
  std::vector<double> Dn((size_t)nshells * nshells, 0.0);
  double Dmax = 0;
  for (long s1 = 0; s1 < nshells; ++s1)
    for (long s2 = 0; s2 < nshells; ++s2) {
      double v = 0;
      for (size_t i = 0; i < obs[s1].size(); ++i)
        for (size_t j = 0; j < obs[s2].size(); ++j)
          v = std::max(v, std::abs(D[(shell2bf[s1] + i) * n + shell2bf[s2] + j]));
      Dn[s1 * nshells + s2] = v;
      Dmax = std::max(Dmax, v);
    }
  const bool do_schwarz = use_schwarz != 0;
  const double fock_precision = precision;
  const double needed_engine_precision = fock_precision / Dmax;  // :1588

  std::vector<Engine> engines(nthr);
  engines[0] = Engine(Operator::coulomb, f->max_nprim, f->max_l, 0);
  engines[0].set(f->screening);
  engines[0].set_precision(needed_engine_precision);
  for (int i = 1; i < nthr; ++i) engines[i] = engines[0];
  std::atomic<size_t> num_ints{0}, num_quartets{0};
  const auto& K = f->K;
  auto Dnorm = [&](long a, long b) { return Dn[a * nshells + b]; };

  const auto t0 = std::chrono::high_resolution_clock::now();
  parallel_do(nthr, [&](int tid) {
    auto& engine = engines[tid];
    auto& g = G[tid];
    const auto& buf = engine.results();
    size_t my_ints = 0, my_q = 0;
    long s1234 = 0, sampled = 0;
    for (long s1 = 0; s1 != nshells; ++s1) {
      const auto bf1_first = shell2bf[s1];
      const auto n1 = obs[s1].size();
      auto sp12_iter = f->spdata[s1].begin();
      for (const auto& s2 : f->splist[s1]) {
        const auto bf2_first = shell2bf[s2];
        const auto n2 = obs[s2].size();
        const auto* sp12 = sp12_iter->get();
        ++sp12_iter;
        const auto Dnorm12 = do_schwarz ? Dnorm(s1, s2) : 0.;
        for (long s3 = 0; s3 <= s1; ++s3) {
          const auto bf3_first = shell2bf[s3];
          const auto n3 = obs[s3].size();
          const auto Dnorm123 =
              do_schwarz ? std::max(Dnorm(s1, s3), std::max(Dnorm(s2, s3), Dnorm12)) : 0.;
          auto sp34_iter = f->spdata[s3].begin();
          const auto s4_max = (s1 == s3) ? (long)s2 : s3;
          for (const auto& s4 : f->splist[s3]) {
            if ((long)s4 > s4_max) break;
            const auto* sp34 = sp34_iter->get();
            ++sp34_iter;
            const long id = s1234++;
            if (task_stride > 1 && (id % task_stride) != task_offset) continue;
            if ((sampled++) % nthr != tid) continue;
            const auto Dnorm1234 =
                do_schwarz ? std::max(Dnorm(s1, s4),
                                      std::max(Dnorm(s2, s4), std::max(Dnorm(s3, s4), Dnorm123)))
                           : 0.;
            if (do_schwarz &&
                Dnorm1234 * K[s1 * nshells + s2] * K[s3 * nshells + s4] < fock_precision)
              continue;
            const auto bf4_first = shell2bf[s4];
            const auto n4 = obs[s4].size();
            const auto s12_deg = (s1 == (long)s2) ? 1 : 2;
            const auto s34_deg = (s3 == (long)s4) ? 1 : 2;
            const auto s12_34_deg = (s1 == s3) ? (s2 == s4 ? 1 : 2) : 2;
            const double deg = s12_deg * s34_deg * s12_34_deg;
            engine.set_precision(Dnorm1234 != 0. ? fock_precision / Dnorm1234
                                                 : needed_engine_precision);
            engine.compute2<Operator::coulomb, BraKet::xx_xx, 0>(obs[s1], obs[s2], obs[s3],
                                                                 obs[s4], sp12, sp34);
            const auto* buf_1234 = buf[0];
            if (buf_1234 == nullptr) continue;
            my_ints += n1 * n2 * n3 * n4;
            ++my_q;
            for (size_t f1 = 0, f1234 = 0; f1 != n1; ++f1) {
              const auto bf1 = f1 + bf1_first;
              for (size_t f2 = 0; f2 != n2; ++f2) {
                const auto bf2 = f2 + bf2_first;
                for (size_t f3 = 0; f3 != n3; ++f3) {
                  const auto bf3 = f3 + bf3_first;
                  for (size_t f4 = 0; f4 != n4; ++f4, ++f1234) {
                    const auto bf4 = f4 + bf4_first;
                    const auto v = buf_1234[f1234] * deg;
                    g[bf1 * n + bf2] += D[bf3 * n + bf4] * v;
                    g[bf3 * n + bf4] += D[bf1 * n + bf2] * v;
                    g[bf1 * n + bf3] -= 0.25 * D[bf2 * n + bf4] * v;
                    g[bf2 * n + bf4] -= 0.25 * D[bf1 * n + bf3] * v;
                    g[bf1 * n + bf4] -= 0.25 * D[bf2 * n + bf3] * v;
                    g[bf2 * n + bf3] -= 0.25 * D[bf1 * n + bf4] * v;
                  }
                }
              }
            }
          }
        }
      }
    }
    num_ints += my_ints;
    num_quartets += my_q;
  });
  const auto t1 = std::chrono::high_resolution_clock::now();
  for (int i = 1; i < nthr; ++i)
    for (size_t k = 0; k < n * n; ++k) G[0][k] += G[i][k];
  for (size_t i = 0; i < n; ++i)
    for (size_t j = 0; j < n; ++j) G_out[i * n + j] = 0.5 * (G[0][i * n + j] + G[0][j * n + i]);
  if (stats) {
    stats[0] = (double)num_ints.load();
    stats[1] = (double)num_quartets.load();
    stats[2] = std::chrono::duration<double>(t1 - t0).count();
  }
  return 0;
}

// Time the reference Engine on a list of shell quartets of one class, nthreads
// independent engines (the reference's parallel model,
// doc/wiki/using-modern-CPlusPlus-API.md:387-389). quartets index into the shell
// table; returns wall seconds, checksum in *sum.
// use_pairs != 0: the ShellPair of every distinct (bra1,bra2) / (ket1,ket2) is precomputed
// outside the timed region and handed to compute2, as the reference's own direct-SCF driver
// does (hartree-fock++.cc:1383-1431,1697); 0: compute2 rebuilds both pairs per quartet
// (engine.impl.h:1259-1276).
double lbo_time_quartets(int nshell, const int* l, const int* pure, const int* nprim,
                         const double* O, const double* alpha, const double* coeff,
                         int coeff_is_raw, long nq, const int* q4, int nthreads, int use_pairs,
                         double* sum) {
  lbo_init();
  auto sh = make_shells(nshell, l, pure, nprim, O, alpha, coeff, coeff_is_raw);
  int max_nprim = 0, max_l = 0;
  for (auto& s : sh) {
    max_nprim = std::max<int>(max_nprim, s.nprim());
    max_l = std::max<int>(max_l, s.contr[0].l);
  }
  const int nthr = std::max(1, nthreads);
  std::vector<Engine> engines(nthr);
  engines[0] = Engine(Operator::coulomb, max_nprim, max_l, 0);
  for (int i = 1; i < nthr; ++i) engines[i] = engines[0];
  std::vector<double> sums(nthr, 0.0);
  std::unordered_map<long long, std::shared_ptr<ShellPair>> pairs;
  std::vector<const ShellPair*> spb, spk;
  if (use_pairs) {
    const double ln_prec = std::log(engines[0].precision());
    auto get = [&](int a, int b) {
      const long long key = (long long)a * nshell + b;
      auto it = pairs.find(key);
      if (it == pairs.end())
        it = pairs.emplace(key, std::make_shared<ShellPair>(sh[a], sh[b], ln_prec,
                                                            ScreeningMethod::Original)).first;
      return it->second.get();
    };
    spb.resize(nq);
    spk.resize(nq);
    for (long q = 0; q < nq; ++q) {
      spb[q] = get(q4[4 * q], q4[4 * q + 1]);
      spk[q] = get(q4[4 * q + 2], q4[4 * q + 3]);
    }
  }
  const auto t0 = std::chrono::high_resolution_clock::now();
  parallel_do(nthr, [&](int tid) {
    auto& e = engines[tid];
    const auto& buf = e.results();
    double s = 0;
    for (long q = tid; q < nq; q += nthr) {
      e.compute2<Operator::coulomb, BraKet::xx_xx, 0>(sh[q4[4 * q]], sh[q4[4 * q + 1]],
                                                      sh[q4[4 * q + 2]], sh[q4[4 * q + 3]],
                                                      use_pairs ? spb[q] : nullptr,
                                                      use_pairs ? spk[q] : nullptr);
      if (buf[0]) s += buf[0][0];
    }
    sums[tid] = s;
  });
  const auto t1 = std::chrono::high_resolution_clock::now();
  double s = 0;
  for (auto v : sums) s += v;
  if (sum) *sum = s;
  return std::chrono::duration<double>(t1 - t0).count();
}

// Time the reference Engine (xs_xx: bra1, unit shell, ket1, ket2) on a list of shell triplets, the
// DF set-up loop of hartree-fock++.cc:2215-2262 (one Engine per thread, round-robin).  t3 = nt x
// {DF shell in table 1, orbital shell, orbital shell in table 2}.
double lbo_time_triplets(int ns1, const int* l1, const int* pure1, const int* nprim1, const double* O1,
                         const double* alpha1, const double* coeff1, int ns2, const int* l2, const int* pure2,
                         const int* nprim2, const double* O2, const double* alpha2, const double* coeff2,
                         long nt, const int* t3, int nthreads, double* sum) {
  lbo_init();
  auto dfs = make_shells(ns1, l1, pure1, nprim1, O1, alpha1, coeff1, 0);
  auto obs = make_shells(ns2, l2, pure2, nprim2, O2, alpha2, coeff2, 0);
  int max_nprim = 0, max_l = 0;
  for (auto* v : {&dfs, &obs})
    for (auto& s : *v) {
      max_nprim = std::max<int>(max_nprim, s.nprim());
      max_l = std::max<int>(max_l, s.contr[0].l);
    }
  const int nthr = std::max(1, nthreads);
  std::vector<Engine> engines(nthr);
  engines[0] = Engine(Operator::coulomb, max_nprim, max_l, 0);
  engines[0].set(BraKet::xs_xx);
  for (int i = 1; i < nthr; ++i) engines[i] = engines[0];
  std::vector<double> sums(nthr, 0.0);
  const auto& unit = Shell::unit();
  const auto t0 = std::chrono::high_resolution_clock::now();
  parallel_do(nthr, [&](int tid) {
    auto& e = engines[tid];
    const auto& buf = e.results();
    double s = 0;
    for (long q = tid; q < nt; q += nthr) {
      e.compute2<Operator::coulomb, BraKet::xs_xx, 0>(dfs[t3[3 * q]], unit, obs[t3[3 * q + 1]], obs[t3[3 * q + 2]]);
      if (buf[0]) s += buf[0][0];
    }
    sums[tid] = s;
  });
  const auto t1 = std::chrono::high_resolution_clock::now();
  double s = 0;
  for (auto v : sums) s += v;
  if (sum) *sum = s;
  return std::chrono::duration<double>(t1 - t0).count();
}

// A list of shell quartets of one class through the reference Engine (xx_xx), results
// (row-major n1*n2*n3*n4, pure where flagged) written to out[q * blk ...]; a set the Engine
// screens out entirely (results()[0] == nullptr) gives zeros.  Returns blk, < 0 on error.
long lbo_compute_batch(int nshell, const int* l, const int* pure, const int* nprim,
                       const double* O, const double* alpha, const double* coeff,
                       int coeff_is_raw, long nq, const int* q4, int nthreads, double precision,
                       double* out) {
  lbo_init();
  if (nq <= 0) return 0;
  auto sh = make_shells(nshell, l, pure, nprim, O, alpha, coeff, coeff_is_raw);
  int max_nprim = 0, max_l = 0;
  for (auto& s : sh) {
    max_nprim = std::max<int>(max_nprim, s.nprim());
    max_l = std::max<int>(max_l, s.contr[0].l);
  }
  const long blk = (long)(sh[q4[0]].size() * sh[q4[1]].size() * sh[q4[2]].size() * sh[q4[3]].size());
  const int nthr = std::max(1, nthreads);
  try {
    std::vector<Engine> engines(nthr);
    engines[0] = Engine(Operator::coulomb, max_nprim, max_l, 0, precision);
    for (int i = 1; i < nthr; ++i) engines[i] = engines[0];
    parallel_do(nthr, [&](int tid) {
      auto& e = engines[tid];
      const auto& buf = e.results();
      for (long q = tid; q < nq; q += nthr) {
        e.compute2<Operator::coulomb, BraKet::xx_xx, 0>(sh[q4[4 * q]], sh[q4[4 * q + 1]],
                                                        sh[q4[4 * q + 2]], sh[q4[4 * q + 3]]);
        if (buf[0])
          std::memcpy(out + q * blk, buf[0], blk * sizeof(double));
        else
          std::memset(out + q * blk, 0, blk * sizeof(double));
      }
    });
  } catch (std::exception& e) {
    std::fprintf(stderr, "lbo_compute_batch: %s\n", e.what());
    return -1;
  }
  return blk;
}

// BasisSet(name, atoms) through the reference's own G94 reader (basis.h.in:473-617);
// LIBINT_DATA_PATH must point at a directory holding basis/<name>.g94.
// Two-call protocol: first with caps 0 to get counts (returns nshell, *nprim_total),
// then with buffers.
int lbo_basis_load(const char* name, int natom, const int* Z, const double* xyz_bohr,
                   int cap_shell, int cap_prim, int* l, int* pure, int* nprim, double* O,
                   double* alpha, double* coeff, int* nprim_total) {
  lbo_init();
  try {
    std::vector<libint2::Atom> atoms(natom);
    for (int i = 0; i < natom; ++i) {
      atoms[i].atomic_number = Z[i];
      atoms[i].x = xyz_bohr[3 * i];
      atoms[i].y = xyz_bohr[3 * i + 1];
      atoms[i].z = xyz_bohr[3 * i + 2];
    }
    libint2::BasisSet bs(name, atoms);
    int ns = (int)bs.size(), np = 0;
    for (auto& s : bs) np += (int)s.nprim();
    *nprim_total = np;
    if (cap_shell < ns || cap_prim < np) return ns;
    size_t off = 0;
    for (int i = 0; i < ns; ++i) {
      l[i] = bs[i].contr[0].l;
      pure[i] = bs[i].contr[0].pure;
      nprim[i] = (int)bs[i].nprim();
      for (int k = 0; k < 3; ++k) O[3 * i + k] = bs[i].O[k];
      for (size_t p = 0; p < bs[i].nprim(); ++p) {
        alpha[off + p] = bs[i].alpha[p];
        coeff[off + p] = bs[i].contr[0].coeff[p];  // normalization embedded
      }
      off += bs[i].nprim();
    }
    return ns;
  } catch (std::exception& e) {
    std::fprintf(stderr, "lbo_basis_load: %s\n", e.what());
    return -1;
  }
}


// Engine(Operator::coulomb, max_nprim, max_l, deriv_order = 1).compute2<coulomb, xx_xx, 1>: the twelve shell
// sets results()[0..11] (3 * centre + xyz, caller's shell order, pure where flagged) of one quartet.  Only in a
// build whose generated headers provide derivative order 1 -- i.e. librefengine_b200.so, the reference Engine
// on the GPU library's Libint_t boundary; the oracle's own shim headers stop at order 0 (returns -3 there).
long lbo_compute2_deriv1(const int* l, const int* pure, const int* nprim, const double* O, const double* alpha,
                         const double* coeff, int coeff_is_raw, double precision, double* out, long out_cap) {
#if LIBINT2_MAX_DERIV_ORDER >= 1
  lbo_init();
  auto sh = make_shells(4, l, pure, nprim, O, alpha, coeff, coeff_is_raw);
  int max_nprim = 0, max_l = 0;
  for (auto& s : sh) {
    max_nprim = std::max<int>(max_nprim, s.nprim());
    max_l = std::max<int>(max_l, s.contr[0].l);
  }
  try {
    Engine engine(Operator::coulomb, max_nprim, max_l, 1, precision);
    const auto& buf = engine.results();
    size_t n = 1;
    for (auto& s : sh) n *= s.size();
    engine.compute2<Operator::coulomb, BraKet::xx_xx, 1>(sh[0], sh[1], sh[2], sh[3]);
    if (buf[0] == nullptr) return 0;
    if ((long)(12 * n) > out_cap) return -2;
    for (int d = 0; d < 12; ++d) std::memcpy(out + d * n, buf[d], n * sizeof(double));
    return (long)n;
  } catch (std::exception& e) {
    std::fprintf(stderr, "lbo_compute2_deriv1: %s\n", e.what());
    return -1;
  }
#else
  (void)l; (void)pure; (void)nprim; (void)O; (void)alpha; (void)coeff; (void)coeff_is_raw; (void)precision;
  (void)out; (void)out_cap;
  return -3;
#endif
}

// ---- first geometric derivatives -------------------------------------------------------------------
// The reference validates its generated derivative kernels against the closed-form eri() with a
// derivative index (tests/eri/test.cc:381-445; eri.h:383-460: one 2*alpha*(a+1) - a*(a-1) step per
// derivative).  The generated eri1 kernels cannot be built here, so that closed form IS the derivative
// oracle: contracted CARTESIAN shell sets, out[d][n1*n2*n3*n4] for d = 3*centre + xyz, shells in the
// order given (coefficients carry the normalization, coeff_is_raw as in make_shells).
namespace {

void deriv1_closed_set(const Shell& s0, const Shell& s1, const Shell& s2, const Shell& s3, double* out) {
  const Shell* sh[4] = {&s0, &s1, &s2, &s3};
  int l[4];
  size_t nc[4];
  for (int c = 0; c < 4; ++c) {
    l[c] = sh[c]->contr[0].l;
    nc[c] = (size_t)(l[c] + 1) * (l[c] + 2) / 2;
  }
  const size_t blk = nc[0] * nc[1] * nc[2] * nc[3];
  std::fill(out, out + 12 * blk, 0.0);
  // Cartesian components in the STANDARD order (cgshell_ordering.h)
  std::vector<std::array<unsigned, 3>> q[4];
  for (int c = 0; c < 4; ++c)
    for (int x = l[c]; x >= 0; --x)
      for (int y = l[c] - x; y >= 0; --y)
        q[c].push_back({{(unsigned)x, (unsigned)y, (unsigned)(l[c] - x - y)}});
  for (size_t p0 = 0; p0 < sh[0]->nprim(); ++p0)
    for (size_t p1 = 0; p1 < sh[1]->nprim(); ++p1)
      for (size_t p2 = 0; p2 < sh[2]->nprim(); ++p2)
        for (size_t p3 = 0; p3 < sh[3]->nprim(); ++p3) {
          const double c0123 = sh[0]->contr[0].coeff[p0] * sh[1]->contr[0].coeff[p1] *
                               sh[2]->contr[0].coeff[p2] * sh[3]->contr[0].coeff[p3];
          size_t e = 0;
          for (auto& a : q[0])
            for (auto& b : q[1])
              for (auto& c : q[2])
                for (auto& d : q[3]) {
                  for (unsigned di = 0; di < 12; ++di) {
                    unsigned idx[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                    idx[di] = 1;
                    out[di * blk + e] +=
                        c0123 * eri(idx, a[0], a[1], a[2], sh[0]->alpha[p0], sh[0]->O.data(), b[0], b[1], b[2],
                                    sh[1]->alpha[p1], sh[1]->O.data(), c[0], c[1], c[2], sh[2]->alpha[p2],
                                    sh[2]->O.data(), d[0], d[1], d[2], sh[3]->alpha[p3], sh[3]->O.data(), 0);
                  }
                  ++e;
                }
        }
}

}  // namespace

long lbo_deriv1_closed(const int* l, const int* nprim, const double* O, const double* alpha,
                       const double* coeff, int coeff_is_raw, double* out, long out_cap) {
  lbo_init();
  const int pure[4] = {0, 0, 0, 0};
  auto sh = make_shells(4, l, pure, nprim, O, alpha, coeff, coeff_is_raw);
  long blk = 1;
  for (auto& s : sh) blk *= (long)s.cartesian_size();
  if (12 * blk > out_cap) return -2;
  deriv1_closed_set(sh[0], sh[1], sh[2], sh[3], out);
  return blk;
}

// Two-body forces exactly as the reference forms them: G1 = compute_2body_fock_deriv<1>
// (hartree-fock++.cc:1775-2055: unique quartets s1 >= s2, s3 <= s1, s4 <= (s1 == s3 ? s2 : s3), degeneracy
// weights, the six-fold digestion of every derivative shell set into G[3*atom + xyz], symmetrisation) and
// F2(atom, xyz) = sum G1[i] o D (:648-656) -- with the closed-form derivative sets above in place of the
// Engine, every pair significant, no screening.  CARTESIAN shells only (the caller back-transforms the
// density of pure shells).  grad = 3 * natoms doubles.
int lbo_fock_grad_closed(int nshell, const int* l, const int* nprim, const double* O, const double* alpha,
                         const double* coeff, int coeff_is_raw, const double* D, int natoms,
                         const int* shell2atom, int nthreads, double* grad) {
  lbo_init();
  std::vector<int> pure(nshell, 0);
  auto obs = make_shells(nshell, l, pure.data(), nprim, O, alpha, coeff, coeff_is_raw);
  std::vector<size_t> shell2bf(nshell);
  size_t n = 0;
  for (int s = 0; s < nshell; ++s) { shell2bf[s] = n; n += obs[s].size(); }
  const int nthr = std::max(1, nthreads);
  const size_t nderiv = 3 * (size_t)natoms;
  std::vector<std::vector<double>> G(nthr, std::vector<double>(nderiv * n * n, 0.0));
  parallel_do(nthr, [&](int tid) {
    auto& g_all = G[tid];
    std::vector<double> buf;
    long s1234 = 0;
    for (long s1 = 0; s1 != nshell; ++s1) {
      const auto bf1_first = shell2bf[s1];
      const auto n1 = obs[s1].size();
      for (long s2 = 0; s2 <= s1; ++s2) {
        const auto bf2_first = shell2bf[s2];
        const auto n2 = obs[s2].size();
        for (long s3 = 0; s3 <= s1; ++s3) {
          const auto bf3_first = shell2bf[s3];
          const auto n3 = obs[s3].size();
          const long s4_max = (s1 == s3) ? s2 : s3;
          for (long s4 = 0; s4 <= s4_max; ++s4) {
            if ((s1234++) % nthr != tid) continue;
            const auto bf4_first = shell2bf[s4];
            const auto n4 = obs[s4].size();
            const size_t n1234 = n1 * n2 * n3 * n4;
            const double s12_deg = (s1 == s2) ? 1.0 : 2.0;
            const double s34_deg = (s3 == s4) ? 1.0 : 2.0;
            const double s12_34_deg = (s1 == s3) ? (s2 == s4 ? 1.0 : 2.0) : 2.0;
            const double deg = s12_deg * s34_deg * s12_34_deg;
            buf.resize(12 * n1234);
            deriv1_closed_set(obs[s1], obs[s2], obs[s3], obs[s4], buf.data());
            const int atoms4[4] = {shell2atom[s1], shell2atom[s2], shell2atom[s3], shell2atom[s4]};
            for (int d = 0; d != 12; ++d) {
              const size_t coord = (size_t)atoms4[d / 3] * 3 + d % 3;
              double* g = g_all.data() + coord * n * n;
              const double* shset = buf.data() + (size_t)d * n1234;
              for (size_t f1 = 0, f1234 = 0; f1 != n1; ++f1) {
                const auto bf1 = f1 + bf1_first;
                for (size_t f2 = 0; f2 != n2; ++f2) {
                  const auto bf2 = f2 + bf2_first;
                  for (size_t f3 = 0; f3 != n3; ++f3) {
                    const auto bf3 = f3 + bf3_first;
                    for (size_t f4 = 0; f4 != n4; ++f4, ++f1234) {
                      const auto bf4 = f4 + bf4_first;
                      const double wv = shset[f1234] * deg;
                      g[bf1 * n + bf2] += D[bf3 * n + bf4] * wv;
                      g[bf3 * n + bf4] += D[bf1 * n + bf2] * wv;
                      g[bf1 * n + bf3] -= 0.25 * D[bf2 * n + bf4] * wv;
                      g[bf2 * n + bf4] -= 0.25 * D[bf1 * n + bf3] * wv;
                      g[bf1 * n + bf4] -= 0.25 * D[bf2 * n + bf3] * wv;
                      g[bf2 * n + bf3] -= 0.25 * D[bf1 * n + bf4] * wv;
                    }
                  }
                }
              }
            }
          }
        }
      }
    }
  });
  for (size_t c = 0; c < nderiv; ++c) {
    double f = 0.0;
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j) {
        double gij = 0.0, gji = 0.0;
        for (int t = 0; t < nthr; ++t) {
          gij += G[t][c * n * n + i * n + j];
          gji += G[t][c * n * n + j * n + i];
        }
        f += 0.5 * (gij + gji) * D[i * n + j];   // GG = (G + G^T)/2 (:2044), then G1[i].cwiseProduct(D).sum()
      }
    grad[c] = f;
  }
  return 0;
}

}  // extern "C"
