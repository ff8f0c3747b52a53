// Three-centre Coulomb integrals (P|mu nu) for density fitting, delivered as dense slabs:
//   Z[P - P0][mu][nu],  P over the functions of DF shells [P0, P0 + nP),  mu, nu over the orbital basis.
// Reference: the DF set-up of the direct-SCF driver, one
//   Engine(coulomb, xs_xx).compute2(dfbs[s1], Shell::unit(), obs[s2], obs[s3])
// per shell triplet, copied into Zxy[ndf][n][n] (tests/hartree-fock/hartree-fock++.cc:2215-2262), and
// the two-centre metric (P|Q) of compute_2body_2index_ints (:1517-1571).  Here a slab is one call: the
// triplets are grouped by class (L s|lc ld) and contraction, every group is an implicit Cartesian
// product (DF shells of the slab) x (significant orbital pairs) evaluated by the store-mode class
// kernels (no task list), solid-harmonic transformed and scattered, with its transpose, into the slab.
// Optional Schwarz-type screening |(P|mu nu)| <= sqrt|(P|P)| sqrt|(mu nu|mu nu)| drops whole ket
// suffixes per group (the reference computes every triplet).  Slabs are independent: N ranks take
// disjoint DF-shell ranges with no collective.
#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <map>
#include <numeric>

#include "internal.h"

using namespace lb200;

namespace lb200 {
int diag_schwarz(lb200_context* ctx, const lb200_pairs* P, std::vector<double>& out);
}

struct Df3cGroup {
  int la = 0, lb = 0, contracted = 0;
  lb200_pairs* pairs = nullptr;
  std::vector<int> shell;       // bra groups: DF shell of every pair, ascending
  std::vector<double> bound;    // Schwarz-type factor per pair (ket groups: sorted descending)
};

struct lb200_df3c {
  lb200_context* ctx = nullptr;
  lb200_basis obs, dfbs;
  lb200_basis* unit = nullptr;
  std::vector<Df3cGroup> bras, kets;
  long long npair = 0;
};

namespace {

// chunk of [ntask][n0][n1][n2] (bra = (P, unit): n0 functions of P; ket pair functions n1 x n2) -> slab
__global__ void scatter_slab_kernel(const double* __restrict__ in, long long ntask, int n0, int n1, int n2,
                                    const PairGeom* __restrict__ gbra, const PairGeom* __restrict__ gket,
                                    int b0, int k0, unsigned nk, int rowbase, int nbf, double* __restrict__ Z) {
  const long long blk = (long long)n0 * n1 * n2, total = ntask * blk;
  const size_t n = (size_t)nbf;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long t = g / blk;
    int r = (int)(g - t * blk);
    const int i2 = r % n2; r /= n2;
    const int i1 = r % n1; const int i0 = r / n1;
    const int ib = b0 + (int)(t / nk), ik = k0 + (int)(t % nk);
    const size_t P = (size_t)(gbra[ib].bf[0] - rowbase + i0);
    const size_t c = (size_t)(gket[ik].bf[0] + i1), d = (size_t)(gket[ik].bf[1] + i2);
    const double v = in[g];
    Z[(P * n + c) * n + d] = v;
    Z[(P * n + d) * n + c] = v;
  }
}

// chunk of [ntask][n0][n2] two-centre blocks -> V[ndf][ndf]
__global__ void scatter_metric_kernel(const double* __restrict__ in, long long ntask, int n0, int n2,
                                      const PairGeom* __restrict__ gbra, const PairGeom* __restrict__ gket,
                                      int b0, unsigned nk, int ndf, double* __restrict__ V) {
  const long long blk = (long long)n0 * n2, total = ntask * blk;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long t = g / blk;
    const int r = (int)(g - t * blk);
    const int i0 = r / n2, i2 = r - i0 * n2;
    const int ib = b0 + (int)(t / nk), ik = (int)(t % nk);
    V[(size_t)(gbra[ib].bf[0] + i0) * ndf + gket[ik].bf[0] + i2] = in[g];
  }
}

int grow(lb200_context* ctx, int slot, size_t bytes, void** out) {
  if (ctx->scratch_bytes[slot] < bytes) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(ctx->d_scratch[slot]);
    ctx->d_scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    int rc = check_cuda(ctx, cudaMalloc(&ctx->d_scratch[slot], bytes), "cudaMalloc(scratch)");
    if (rc) return rc;
    ctx->scratch_bytes[slot] = bytes;
  }
  *out = ctx->d_scratch[slot];
  return LB200_OK;
}

constexpr size_t kChunkBytes = (size_t)1 << 28;   // Cartesian integrals per launch

}  // namespace

extern "C" {

int lb200_df3c_create(lb200_context* ctx, const lb200_basis* obs, const lb200_basis* dfbs, long long npair,
                      const int* s1, const int* s2, lb200_df3c** out) {
  if (!ctx || !obs || !dfbs || !out || npair < 0) return LB200_ERR_INVALID;
  cudaSetDevice(ctx->device);
  auto* f = new lb200_df3c;
  f->ctx = ctx; f->obs = *obs; f->dfbs = *dfbs; f->npair = npair;
  int rc = lb200_basis_create_unit(ctx, &f->unit);
  // bra groups: DF shells by (l, pure, contracted), ascending shell index inside a group
  std::map<std::array<int, 3>, std::vector<int>> bg;
  for (int s = 0; s < dfbs->nshell && !rc; ++s) bg[std::array<int, 3>{dfbs->l[s], dfbs->pure[s], dfbs->nprim[s] > 1 ? 1 : 0}].push_back(s);
  for (auto& kv : bg) {
    if (rc) break;
    Df3cGroup g;
    g.la = kv.first[0]; g.contracted = kv.first[2]; g.shell = kv.second;
    std::vector<int> zero(g.shell.size(), 0);
    rc = build_pairs(ctx, dfbs, f->unit, (int)g.shell.size(), g.shell.data(), zero.data(), kScreenOriginal,
                     std::numeric_limits<double>::lowest(), nullptr, nullptr, &g.pairs);
    if (!rc) rc = diag_schwarz(ctx, g.pairs, g.bound);   // sqrt(max |(P|P)|)
    f->bras.push_back(std::move(g));
  }
  // ket groups: significant orbital pairs by class (first shell = higher AM), purity and contraction,
  // sorted by sqrt(max |(mu nu|mu nu)|) descending so that screening keeps a prefix
  std::map<std::array<int, 5>, std::pair<std::vector<int>, std::vector<int>>> kg;
  for (long long i = 0; i < npair && !rc; ++i) {
    int a = s1[i], b = s2[i];
    if (a < 0 || b < 0 || a >= obs->nshell || b >= obs->nshell) { rc = LB200_ERR_INVALID; break; }
    if (obs->l[a] < obs->l[b]) std::swap(a, b);
    auto& e = kg[std::array<int, 5>{obs->l[a], obs->l[b], obs->pure[a], obs->pure[b], obs->nprim[a] * obs->nprim[b] > 1 ? 1 : 0}];
    e.first.push_back(a);
    e.second.push_back(b);
  }
  for (auto& kv : kg) {
    if (rc) break;
    auto& a = kv.second.first;
    auto& b = kv.second.second;
    const int n = (int)a.size();
    lb200_pairs* tmp = nullptr;
    rc = build_pairs(ctx, obs, obs, n, a.data(), b.data(), kScreenOriginal, std::numeric_limits<double>::lowest(),
                     nullptr, nullptr, &tmp);
    std::vector<double> bound;
    if (!rc) rc = diag_schwarz(ctx, tmp, bound);
    lb200_pairs_destroy(tmp);
    if (rc) break;
    std::vector<int> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return bound[x] > bound[y]; });
    Df3cGroup g;
    g.la = kv.first[0]; g.lb = kv.first[1]; g.contracted = kv.first[4];
    std::vector<int> as(n), bs(n);
    g.bound.resize(n);
    for (int i = 0; i < n; ++i) { as[i] = a[ord[i]]; bs[i] = b[ord[i]]; g.bound[i] = bound[ord[i]]; }
    // ShellPair::init with the Engine's default precision (engine.h:503-526: epsilon)
    rc = build_pairs(ctx, obs, obs, n, as.data(), bs.data(), kScreenOriginal,
                     std::log(std::numeric_limits<double>::epsilon()), nullptr, nullptr, &g.pairs);
    f->kets.push_back(std::move(g));
  }
  if (rc) { lb200_df3c_destroy(f); return rc; }
  *out = f;
  return LB200_OK;
}

int lb200_df3c_destroy(lb200_df3c* f) {
  if (!f) return LB200_OK;
  for (auto& g : f->bras) lb200_pairs_destroy(g.pairs);
  for (auto& g : f->kets) lb200_pairs_destroy(g.pairs);
  lb200_basis_destroy(f->unit);
  delete f;
  return LB200_OK;
}

int lb200_df3c_slab(lb200_df3c* f, int P0, int nP, double threshold, double precision, double* Z_dev,
                    double* stats) {
  if (!f || !Z_dev || P0 < 0 || nP < 0 || P0 + nP > f->dfbs.nshell) return LB200_ERR_INVALID;
  lb200_context* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int n = f->obs.nbf;
  if (nP == 0) return LB200_OK;
  const int rowbase = f->dfbs.shell2bf[P0];
  const int rowend = (P0 + nP < f->dfbs.nshell) ? f->dfbs.shell2bf[P0 + nP] : f->dfbs.nbf;
  int rc = check_cuda(ctx, cudaMemsetAsync(Z_dev, 0, (size_t)(rowend - rowbase) * n * n * sizeof(double), st),
                      "memset slab");
  double ntrip = 0, nall = 0;
  for (auto& B : f->bras) {
    if (rc) break;
    const int lo = (int)(std::lower_bound(B.shell.begin(), B.shell.end(), P0) - B.shell.begin());
    const int hi = (int)(std::lower_bound(B.shell.begin(), B.shell.end(), P0 + nP) - B.shell.begin());
    if (hi <= lo) continue;
    const double Kmax = *std::max_element(B.bound.begin() + lo, B.bound.begin() + hi);
    for (auto& Kt : f->kets) {
      if (rc) break;
      int nk = Kt.pairs->dev.npair;
      nall += (double)(hi - lo) * nk;
      if (threshold > 0 && Kmax > 0) {   // kets whose bound can still reach the threshold: a prefix
        const double thr = threshold / Kmax;
        nk = (int)(std::lower_bound(Kt.bound.begin(), Kt.bound.end(), thr, [](double b, double t) { return b >= t; }) -
                   Kt.bound.begin());
      }
      if (nk == 0) continue;
      const int l[4] = {B.la, 0, Kt.la, Kt.lb};
      const int pure[4] = {B.pairs->dev.pure_a, 0, Kt.pairs->dev.pure_a, Kt.pairs->dev.pure_b};
      const bool tform = (pure[0] && l[0] > 0) || (pure[2] && l[2] > 0) || (pure[3] && l[3] > 0);
      const long long cblk = (long long)nc(l[0]) * nc(l[2]) * nc(l[3]);
      const int n0 = pure[0] ? npure(l[0]) : nc(l[0]), n1 = pure[2] ? npure(l[2]) : nc(l[2]),
                n2 = pure[3] ? npure(l[3]) : nc(l[3]);
      const long long per_row = (long long)nk * cblk * 8;
      // chunk: whole bra rows when they fit, else one row split over ket ranges
      const int rows = (int)std::max<long long>(1, std::min<long long>(hi - lo, (long long)kChunkBytes / per_row));
      const int kets_per = per_row <= (long long)kChunkBytes ? nk : (int)std::max<long long>(1, (long long)kChunkBytes / (cblk * 8));
      for (int b0 = lo; b0 < hi && !rc; b0 += rows) {
        const int nb = std::min(rows, hi - b0);
        for (int k0 = 0; k0 < nk && !rc; k0 += kets_per) {
          const int nkk = std::min(kets_per, nk - k0);
          const long long ntask = (long long)nb * nkk;
          double *d_cart = nullptr, *d_pure = nullptr;
          rc = grow(ctx, 1, (size_t)ntask * cblk * 8, reinterpret_cast<void**>(&d_cart));
          if (!rc && tform) rc = grow(ctx, 3, (size_t)ntask * n0 * n1 * n2 * 8, reinterpret_cast<void**>(&d_pure));
          if (rc) break;
          ProductTasks pt{b0, nb, k0, nkk};
          rc = run_store(ctx, B.pairs, Kt.pairs, ntask, nullptr, kScreenOriginal, precision, d_cart, &pt);
          if (rc) break;
          const double* src = d_cart;
          if (tform) {
            rc = check_cuda(ctx, launch_pure_transform(ctx, d_cart, d_pure, ntask, l, pure, st), "pure transform");
            ++ctx->launches;
            src = d_pure;
          }
          const long long total = ntask * n0 * n1 * n2;
          const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 32);
          scatter_slab_kernel<<<grid, 256, 0, st>>>(src, ntask, n0, n1, n2, B.pairs->dev.geom, Kt.pairs->dev.geom,
                                                   b0, k0, (unsigned)nkk, rowbase, n, Z_dev);
          ++ctx->launches;
          ntrip += (double)ntask;
        }
      }
    }
  }
  if (!rc) rc = check_cuda(ctx, cudaGetLastError(), "df3c slab");
  if (stats) { stats[0] = ntrip; stats[1] = nall; }
  return rc;
}

int lb200_df3c_metric(lb200_df3c* f, double* V_dev) {
  if (!f || !V_dev) return LB200_ERR_INVALID;
  lb200_context* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int ndf = f->dfbs.nbf;
  int rc = check_cuda(ctx, cudaMemsetAsync(V_dev, 0, (size_t)ndf * ndf * sizeof(double), st), "memset V");
  for (auto& B : f->bras)
    for (auto& Kt : f->bras) {
      if (rc) return rc;
      const int nbp = B.pairs->dev.npair, nkp = Kt.pairs->dev.npair;
      const int l[4] = {B.la, 0, Kt.la, 0};
      const int pure[4] = {B.pairs->dev.pure_a, 0, Kt.pairs->dev.pure_a, 0};
      const bool tform = (pure[0] && l[0] > 0) || (pure[2] && l[2] > 0);
      const long long cblk = (long long)nc(l[0]) * nc(l[2]);
      const int n0 = pure[0] ? npure(l[0]) : nc(l[0]), n2 = pure[2] ? npure(l[2]) : nc(l[2]);
      const int rows = (int)std::max<long long>(1, std::min<long long>(nbp, (long long)kChunkBytes / ((long long)nkp * cblk * 8)));
      for (int b0 = 0; b0 < nbp && !rc; b0 += rows) {
        const int nb = std::min(rows, nbp - b0);
        const long long ntask = (long long)nb * nkp;
        double *d_cart = nullptr, *d_pure = nullptr;
        rc = grow(ctx, 1, (size_t)ntask * cblk * 8, reinterpret_cast<void**>(&d_cart));
        if (!rc && tform) rc = grow(ctx, 3, (size_t)ntask * n0 * n2 * 8, reinterpret_cast<void**>(&d_pure));
        if (rc) break;
        ProductTasks pt{b0, nb, 0, nkp};
        rc = run_store(ctx, B.pairs, Kt.pairs, ntask, nullptr, kScreenOriginal, 0.0, d_cart, &pt);
        if (rc) break;
        const double* src = d_cart;
        if (tform) {
          rc = check_cuda(ctx, launch_pure_transform(ctx, d_cart, d_pure, ntask, l, pure, st), "pure transform");
          ++ctx->launches;
          src = d_pure;
        }
        const long long total = ntask * n0 * n2;
        const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 32);
        scatter_metric_kernel<<<grid, 256, 0, st>>>(src, ntask, n0, n2, B.pairs->dev.geom, Kt.pairs->dev.geom, b0,
                                                   (unsigned)nkp, ndf, V_dev);
        ++ctx->launches;
      }
    }
  if (!rc) rc = check_cuda(ctx, cudaGetLastError(), "df3c metric");
  return rc;
}

int lb200_df3c_info(const lb200_df3c* f, long long* info) {
  if (!f || !info) return LB200_ERR_INVALID;
  long long nt = 0;
  for (auto& B : f->bras) for (auto& K : f->kets) nt += (long long)B.pairs->dev.npair * K.pairs->dev.npair;
  info[0] = f->obs.nbf; info[1] = f->dfbs.nbf; info[2] = f->dfbs.nshell; info[3] = f->npair; info[4] = nt;
  info[5] = (long long)f->bras.size() * (long long)f->kets.size();
  return LB200_OK;
}

}  // extern "C"
