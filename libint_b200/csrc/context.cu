// Context, basis and shell-pair objects of the C ABI, and the batched ERI entry point.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <limits>

#include "internal.h"

extern "C" const unsigned char lb200_boys_table_begin[];
extern "C" const unsigned char lb200_boys_table_end[];

namespace lb200 {

int set_error(const lb200_context* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
int check_cuda(const lb200_context* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return LB200_OK;
  return set_error(ctx, LB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

namespace {

// (2k-1)!! as used by Shell::renorm (df_Kminus1 of 2l, shell.h:966-975)
double df_kminus1(int k) {
  double r = 1.0;
  for (int n = k - 1; n > 1; n -= 2) r *= n;
  return r;
}

void fill_rows(std::vector<RowInfo>& rows) {
  rows.resize(kMaxRows);
  for (int e = 0; e <= 2 * kMaxShellL; ++e)
    for (int i = 0; i < nc(e); ++i) {
      RowInfo ri{};
      const C3 q = cxyz(e, i);
      ri.e = e;
      ri.q[0] = q.x; ri.q[1] = q.y; ri.q[2] = q.z;
      for (int d = 0; d < 3; ++d)
        ri.rm[d] = cget(q, d) > 0 ? nc_upto(e - 2) + cidx(cadd(q, d, -1)) : 0;
      ri.dir = cdir(q);
      rows[nc_upto(e - 1) + i] = ri;
    }
}

template <int L>
void fill_sph_one(std::vector<int>& rowptr, std::vector<int>& col, std::vector<double>& val,
                  std::vector<int>& base) {
  constexpr int RP = 2 * kMaxShellL + 2;
  base[L] = (int)col.size();
  for (int m = 0; m <= 2 * L + 1; ++m) rowptr[L * RP + m] = Sph<L>::rowptr[m];
  for (int k = 0; k < Sph<L>::nnz; ++k) {
    col.push_back(Sph<L>::e[k].c);
    val.push_back(Sph<L>::e[k].v);
  }
}

}  // namespace
}  // namespace lb200

using namespace lb200;

extern "C" {

int lb200_version(void) { return 100; }

int lb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* lb200_last_error(const lb200_context* ctx) { return ctx ? ctx->err.c_str() : ""; }

int lb200_context_create(int device, lb200_context** out) {
  if (!out) return LB200_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return LB200_ERR_CUDA;
  auto* ctx = new lb200_context;
  ctx->device = device;
  int rc;
  if ((rc = check_cuda(ctx, cudaSetDevice(device), "cudaSetDevice"))) { delete ctx; return rc; }
  cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
  if ((rc = check_cuda(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking),
                       "cudaStreamCreate"))) { delete ctx; return rc; }
  ctx->own_stream = true;
  // Boys table
  const size_t nb = lb200_boys_table_end - lb200_boys_table_begin;
  const size_t expect = (size_t)kBoysNInt * (kBoysTableMmax + 1) * 8 * sizeof(double);
  if (nb != expect) { delete ctx; return LB200_ERR_INVALID; }
  cudaMalloc(&ctx->d_boys, nb);
  cudaMemcpy(ctx->d_boys, lb200_boys_table_begin, nb, cudaMemcpyHostToDevice);
  // row table
  std::vector<RowInfo> rows;
  fill_rows(rows);
  cudaMalloc(&ctx->d_rows, rows.size() * sizeof(RowInfo));
  cudaMemcpy(ctx->d_rows, rows.data(), rows.size() * sizeof(RowInfo), cudaMemcpyHostToDevice);
  // solid-harmonic CSR tables
  constexpr int RP = 2 * kMaxShellL + 2;
  std::vector<int> rowptr((kMaxShellL + 1) * RP, 0), col, base(kMaxShellL + 1, 0);
  std::vector<double> val;
  fill_sph_one<0>(rowptr, col, val, base);
  fill_sph_one<1>(rowptr, col, val, base);
  fill_sph_one<2>(rowptr, col, val, base);
  fill_sph_one<3>(rowptr, col, val, base);
  fill_sph_one<4>(rowptr, col, val, base);
  cudaMalloc(&ctx->d_sph_rowptr, rowptr.size() * sizeof(int));
  cudaMalloc(&ctx->d_sph_col, col.size() * sizeof(int));
  cudaMalloc(&ctx->d_sph_val, val.size() * sizeof(double));
  cudaMalloc(&ctx->d_sph_base, base.size() * sizeof(int));
  cudaMemcpy(ctx->d_sph_rowptr, rowptr.data(), rowptr.size() * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(ctx->d_sph_col, col.data(), col.size() * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(ctx->d_sph_val, val.data(), val.size() * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemcpy(ctx->d_sph_base, base.data(), base.size() * sizeof(int), cudaMemcpyHostToDevice);
  if ((rc = check_cuda(ctx, cudaDeviceSynchronize(), "context init"))) { delete ctx; return rc; }
  *out = ctx;
  return LB200_OK;
}

int lb200_context_destroy(lb200_context* ctx) {
  if (!ctx) return LB200_OK;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->d_boys);
  cudaFree(ctx->d_rows);
  cudaFree(ctx->d_sph_rowptr);
  cudaFree(ctx->d_sph_col);
  cudaFree(ctx->d_sph_val);
  cudaFree(ctx->d_sph_base);
  for (int i = 0; i < lb200_context::kScratchSlots; ++i) cudaFree(ctx->d_scratch[i]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (int b = 0; b < 2; ++b) {
    if (ctx->ev_done[b]) cudaEventDestroy(ctx->ev_done[b]);
    if (ctx->ev_free[b]) cudaEventDestroy(ctx->ev_free[b]);
  }
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return LB200_OK;
}

int lb200_context_set_stream(lb200_context* ctx, void* cuda_stream) {
  if (!ctx) return LB200_ERR_INVALID;
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  return LB200_OK;
}

int lb200_context_synchronize(lb200_context* ctx) {
  if (!ctx) return LB200_ERR_INVALID;
  cudaSetDevice(ctx->device);
  return check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "synchronize");
}

long long lb200_context_launch_count(const lb200_context* ctx) { return ctx ? ctx->launches : 0; }

// Shell::renorm, include/libint2/shell.h:958-999
int lb200_shell_renorm(int l, int nprim, const double* alpha, double* coeff,
                       int enforce_unit_normalization, double* max_ln_coeff) {
  if (l < 0 || nprim < 1 || !alpha || !coeff) return LB200_ERR_INVALID;
  const double sqrt_Pi_cubed = 5.56832799683170784528481798212;
  for (int p = 0; p < nprim; ++p) {
    if (alpha[p] < 0) return LB200_ERR_INVALID;
    if (alpha[p] != 0) {
      const double two_alpha = 2 * alpha[p];
      const double two_alpha_to_am32 = std::pow(two_alpha, l + 1) * std::sqrt(two_alpha);
      const double nf =
          std::sqrt(std::pow(2, l) * two_alpha_to_am32 / (sqrt_Pi_cubed * df_kminus1(2 * l)));
      coeff[p] *= nf;
    }
  }
  if (enforce_unit_normalization) {
    double norm = 0;
    for (int p = 0; p < nprim; ++p)
      for (int q = 0; q <= p; ++q) {
        const double gamma = alpha[p] + alpha[q];
        norm += (p == q ? 1 : 2) * df_kminus1(2 * l) * sqrt_Pi_cubed * coeff[p] * coeff[q] /
                (std::pow(2, l) * std::pow(gamma, l + 1) * std::sqrt(gamma));
      }
    const double nf = 1 / std::sqrt(norm);
    for (int p = 0; p < nprim; ++p) coeff[p] *= nf;
  }
  if (max_ln_coeff)
    for (int p = 0; p < nprim; ++p) max_ln_coeff[p] = std::log(std::abs(coeff[p]));
  return LB200_OK;
}

int lb200_basis_create(lb200_context* ctx, int nshell, const int* l, const int* pure,
                       const int* nprim, const double* origin, const double* alpha,
                       const double* coeff, lb200_basis** out) {
  if (!ctx || !out || nshell < 0) return LB200_ERR_INVALID;
  auto* b = new lb200_basis;
  b->ctx = ctx;
  b->nshell = nshell;
  b->l.assign(l, l + nshell);
  b->pure.assign(pure, pure + nshell);
  b->nprim.assign(nprim, nprim + nshell);
  b->O.assign(origin, origin + 3 * nshell);
  b->off.resize(nshell + 1);
  b->shell2bf.resize(nshell);
  int o = 0, nbf = 0;
  for (int s = 0; s < nshell; ++s) {
    if (l[s] < 0 || l[s] > LB200_MAX_AM) {
      delete b;
      return set_error(ctx, LB200_ERR_LMAX, "shell angular momentum exceeds LB200_MAX_AM");
    }
    if (nprim[s] < 1) { delete b; return set_error(ctx, LB200_ERR_INVALID, "empty shell"); }
    b->off[s] = o;
    o += nprim[s];
    b->shell2bf[s] = nbf;
    b->pure[s] = pure[s] ? 1 : 0;
    nbf += b->size(s);
  }
  b->off[nshell] = o;
  b->nbf = nbf;
  b->alpha.assign(alpha, alpha + o);
  b->coeff.assign(coeff, coeff + o);
  b->max_ln_coeff.resize(o);
  for (int i = 0; i < o; ++i) b->max_ln_coeff[i] = std::log(std::abs(b->coeff[i]));  // shell.h:1001-1011
  *out = b;
  return LB200_OK;
}

int lb200_basis_create_unit(lb200_context* ctx, lb200_basis** out) {
  const int l = 0, pure = 0, nprim = 1;
  const double O[3] = {0, 0, 0}, alpha = 0.0, coeff = 1.0;
  return lb200_basis_create(ctx, 1, &l, &pure, &nprim, O, &alpha, &coeff, out);
}

int lb200_basis_destroy(lb200_basis* bs) {
  delete bs;
  return LB200_OK;
}
int lb200_basis_nbf(const lb200_basis* bs) { return bs ? bs->nbf : LB200_ERR_INVALID; }
int lb200_basis_nshell(const lb200_basis* bs) { return bs ? bs->nshell : LB200_ERR_INVALID; }
int lb200_basis_shell2bf(const lb200_basis* bs, int* out) {
  if (!bs || !out) return LB200_ERR_INVALID;
  std::memcpy(out, bs->shell2bf.data(), sizeof(int) * bs->nshell);
  return LB200_OK;
}

}  // extern "C"

namespace lb200 {

// one device allocation: prim (+ K, p1p2 of device-built blocks) | geom | schwarz | prim_off | shell | gidx
int upload_pairs(lb200_context* ctx, lb200_pairs* P, const std::vector<PairGeom>& geom,
                 const double* pair_schwarz, bool with_prims) {
  PairBlock& d = P->dev;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t n = (size_t)d.npair;
  const size_t np = (size_t)P->nprim_total;
  const size_t o_prim = 0, o_kraw = al(o_prim + np * sizeof(PrimPair));
  const size_t o_pp = with_prims ? o_kraw : al(o_kraw + np * 8);
  const size_t o_geom = with_prims ? o_kraw : al(o_pp + np * sizeof(int2));
  const size_t o_sw = al(o_geom + n * sizeof(PairGeom)), o_po = al(o_sw + n * 8), o_sh = al(o_po + (n + 1) * 4);
  const size_t o_gi = al(o_sh + 2 * n * 4), total = al(o_gi + n * 4) + 256;
  cudaSetDevice(ctx->device);
  int rc = check_cuda(ctx, cudaMalloc(&P->d_block, total), "cudaMalloc(pairs)");
  if (rc) return rc;
  char* base = static_cast<char*>(P->d_block);
  std::vector<int> gidx(n);
  for (size_t i = 0; i < n; ++i) {
    const long long hi = std::max(P->shell[2 * i], P->shell[2 * i + 1]);
    const long long lo = std::min(P->shell[2 * i], P->shell[2 * i + 1]);
    gidx[i] = (int)(hi * (hi + 1) / 2 + lo);   // < 2^31 for < 65536 shells (checked by lb200_fock_create)
  }
  std::vector<double> sw(n, 0.0);
  if (pair_schwarz) sw.assign(pair_schwarz, pair_schwarz + n);
  if (with_prims)
    cudaMemcpy(base + o_prim, P->prim.data(), np * sizeof(PrimPair), cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_geom, geom.data(), n * sizeof(PairGeom), cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_sw, sw.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_po, P->prim_off.data(), (n + 1) * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_sh, P->shell.data(), 2 * n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(base + o_gi, gidx.data(), n * 4, cudaMemcpyHostToDevice);
  rc = check_cuda(ctx, cudaGetLastError(), "upload pairs");
  if (rc) { cudaFree(P->d_block); P->d_block = nullptr; return rc; }
  d.prim = reinterpret_cast<const PrimPair*>(base + o_prim);
  d.geom = reinterpret_cast<const PairGeom*>(base + o_geom);
  d.schwarz = reinterpret_cast<const double*>(base + o_sw);
  d.prim_off = reinterpret_cast<const int*>(base + o_po);
  d.shell = reinterpret_cast<const int*>(base + o_sh);
  d.gidx = reinterpret_cast<const int*>(base + o_gi);
  if (!with_prims) {
    P->d_Kraw = reinterpret_cast<const double*>(base + o_kraw);
    P->d_p1p2 = reinterpret_cast<const int2*>(base + o_pp);
  }
  return LB200_OK;
}

int ctx_scratch(lb200_context* ctx, int slot, size_t bytes, void** out) {
  if (ctx->scratch_bytes[slot] < bytes) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(ctx->d_scratch[slot]);
    ctx->d_scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    int r = check_cuda(ctx, cudaMalloc(&ctx->d_scratch[slot], bytes), "cudaMalloc(scratch)");
    if (r) return r;
    ctx->scratch_bytes[slot] = bytes;
  }
  *out = ctx->d_scratch[slot];
  return LB200_OK;
}

int pairs_host_mirror(const lb200_pairs* Pc) {
  if (Pc->host_valid) return LB200_OK;
  auto* P = const_cast<lb200_pairs*>(Pc);
  const size_t np = (size_t)P->nprim_total;
  cudaSetDevice(P->ctx->device);
  P->prim.resize(np);
  P->Kraw.resize(np);
  P->p1p2.resize(2 * np);
  cudaMemcpy(P->prim.data(), P->dev.prim, np * sizeof(PrimPair), cudaMemcpyDeviceToHost);
  cudaMemcpy(P->Kraw.data(), P->d_Kraw, np * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(P->p1p2.data(), P->d_p1p2, np * sizeof(int2), cudaMemcpyDeviceToHost);
  const int rc = check_cuda(P->ctx, cudaGetLastError(), "download pair records");
  if (!rc) P->host_valid = true;
  return rc;
}

// ShellPair::init, include/libint2/shell.h:1138-1256 (Original/Conservative) and
// :1259-1328 (Schwarz variants), for a whole block of pairs; PA/gamma/c_a*c_b are the
// per-quartet prerequisites of engine.impl.h:1331-1367,1514-1537 hoisted to the pair.
// Blocks of at least LB200_PAIRS_DEVICE_MIN pairs (default 2048) run the primitive-pair loop on the GPU
// (pairs_device.cu: count, prefix sum, fill); smaller ones on the host.
int build_pairs(lb200_context* ctx, const lb200_basis* bs1, const lb200_basis* bs2, int npair,
                const int* s1, const int* s2, int screening, double ln_prec,
                const double* prim_schwarz, const double* pair_schwarz, lb200_pairs** out) {
  if (!ctx || !bs1 || !bs2 || !out || npair < 0) return LB200_ERR_INVALID;
  if (screening != kScreenOriginal && screening != kScreenConservative &&
      screening != kScreenSchwarz && screening != kScreenSchwarzInf)
    return set_error(ctx, LB200_ERR_INVALID, "invalid screening method");
  const bool schwarz = screening == kScreenSchwarz || screening == kScreenSchwarzInf;
  std::vector<double> computed;
  // the library's own evaluator is the SchwarzInf one (sqrt of the largest |(ab|ab)| of the primitive
  // pair, hartree-fock++.cc:1403-1409); ScreeningMethod::Schwarz wants the Frobenius norm, which only a
  // caller-supplied table can provide
  if (screening == kScreenSchwarz && !prim_schwarz && npair > 0)
    return set_error(ctx, LB200_ERR_INVALID,
                     "LB200_SCREEN_SCHWARZ needs prim_schwarz (the built-in evaluator is SchwarzInf)");
  if (schwarz && !prim_schwarz && npair > 0) {
    int rc = compute_prim_schwarz(ctx, bs1, bs2, npair, s1, s2, computed);
    if (rc) return rc;
    prim_schwarz = computed.data();
  }
  auto* P = new lb200_pairs;
  P->ctx = ctx;
  P->alpha1 = bs1->alpha; P->off1 = bs1->off;
  P->alpha2 = bs2->alpha; P->off2 = bs2->off;
  PairBlock& d = P->dev;
  d.npair = npair;
  if (npair > 0) {
    d.la = bs1->l[s1[0]]; d.lb = bs2->l[s2[0]];
    d.pure_a = bs1->pure[s1[0]]; d.pure_b = bs2->pure[s2[0]];
  } else {
    d.la = d.lb = d.pure_a = d.pure_b = 0;
    d.unit_b = 0;
  }
  P->prim_off.assign(npair + 1, 0);
  P->shell.resize(2 * (size_t)npair);
  P->bf.resize(2 * (size_t)npair);
  P->AB.resize(3 * (size_t)npair);
  P->A.resize(3 * (size_t)npair);
  // ---- per pair: validation, geometry ---------------------------------------------------------
  for (int i = 0; i < npair; ++i) {
    const int a = s1[i], b = s2[i];
    if (a < 0 || a >= bs1->nshell || b < 0 || b >= bs2->nshell) {
      delete P;
      return set_error(ctx, LB200_ERR_INVALID, "shell index out of range");
    }
    if (bs1->l[a] != d.la || bs2->l[b] != d.lb || bs1->pure[a] != d.pure_a ||
        bs2->pure[b] != d.pure_b || d.la < d.lb) {
      delete P;
      return set_error(ctx, LB200_ERR_INVALID,
                       "pairs of one block must share one class with l(s1) >= l(s2)");
    }
    P->shell[2 * i] = a; P->shell[2 * i + 1] = b;
    P->bf[2 * i] = bs1->shell2bf[a]; P->bf[2 * i + 1] = bs2->shell2bf[b];
    for (int k = 0; k < 3; ++k) {
      P->A[3 * i + k] = bs1->O[3 * a + k];
      P->AB[3 * i + k] = bs1->O[3 * a + k] - bs2->O[3 * b + k];
    }
    const bool unit_b = bs2->is_unit(b);
    if (i == 0) d.unit_b = unit_b ? 1 : 0;
    if ((d.unit_b != 0) != unit_b) {
      delete P;
      return set_error(ctx, LB200_ERR_INVALID, "unit and ordinary second shells cannot share a block");
    }
  }
  const size_t n = (size_t)npair;
  std::vector<PairGeom> geom(n);
  for (size_t i = 0; i < n; ++i) {
    PairGeom& g = geom[i];
    for (int k = 0; k < 3; ++k) {
      g.A[k] = bs1->O[3 * (size_t)P->shell[2 * i] + k];
      g.AB[k] = P->AB[3 * i + k];
    }
    g.bf[0] = P->bf[2 * i]; g.bf[1] = P->bf[2 * i + 1];
    g.shell[0] = P->shell[2 * i]; g.shell[1] = P->shell[2 * i + 1];
  }
  // ---- primitive pairs --------------------------------------------------------------------------
  // read per call: tests build one block both ways
  const char* env_min = std::getenv("LB200_PAIRS_DEVICE_MIN");
  const int device_min = env_min ? std::atoi(env_min) : 2048;
  if (npair >= device_min && npair > 0) {
    DevicePrimBuilder bld;
    int rc = bld.init(ctx, bs1, bs2, npair, s1, s2, screening, ln_prec, prim_schwarz);
    std::vector<int> counts;
    if (!rc) rc = bld.count(counts);
    if (rc) { delete P; return rc; }
    long long tot = 0;
    for (int i = 0; i < npair; ++i) {
      tot += counts[i];
      if (tot > 0x7fffffffll) { delete P; return set_error(ctx, LB200_ERR_INVALID, "too many primitive pairs in one block"); }
      P->prim_off[i + 1] = (int)tot;
      d.max_nprim = std::max(d.max_nprim, counts[i]);
    }
    P->nprim_total = tot;
    P->host_valid = false;
    rc = upload_pairs(ctx, P, geom, pair_schwarz, false);
    if (!rc) rc = bld.fill(d.prim_off, const_cast<PrimPair*>(d.prim), const_cast<double*>(P->d_Kraw),
                           const_cast<int2*>(P->d_p1p2));
    if (rc) { cudaFree(P->d_block); delete P; return rc; }
    *out = P;
    return LB200_OK;
  }
  size_t fac_off = 0;
  for (int i = 0; i < npair; ++i) {
    const int a = s1[i], b = s2[i];
    const double* A = &bs1->O[3 * a];
    const double* B = &bs2->O[3 * b];
    double AB2 = 0.;
    for (int k = 0; k < 3; ++k) AB2 += P->AB[3 * i + k] * P->AB[3 * i + k];
    const int np1 = bs1->nprim[a], np2 = bs2->nprim[b];
    const int l1 = bs1->l[a], l2 = bs2->l[b];
    for (int p1 = 0; p1 < np1; ++p1)
      for (int p2 = 0; p2 < np2; ++p2) {
        const double a1 = bs1->alpha[bs1->off[a] + p1], a2 = bs2->alpha[bs2->off[b] + p2];
        const double mlc1 = bs1->max_ln_coeff[bs1->off[a] + p1];
        const double mlc2 = bs2->max_ln_coeff[bs2->off[b] + p2];
        const double gamma = a1 + a2;
        const double oogamma = 1 / gamma;
        const double rho = a1 * a2 * oogamma;
        const double minus_rho_times_AB2 = -rho * AB2;
        double ln_screen_fac;
        if (schwarz) {
          ln_screen_fac =
              std::log((double)(np1 * np2) * prim_schwarz[fac_off + (size_t)p1 * np2 + p2]) + mlc1 + mlc2;
          if (ln_screen_fac < ln_prec) continue;
        } else {
          ln_screen_fac = minus_rho_times_AB2 + mlc1 + mlc2;
          if (screening == kScreenOriginal && ln_screen_fac < ln_prec) continue;
        }
        double Pc[3];
        if (AB2 == 0.) {
          Pc[0] = A[0]; Pc[1] = A[1]; Pc[2] = A[2];
        } else {
          Pc[0] = (a1 * A[0] + a2 * B[0]) * oogamma;
          Pc[1] = (a1 * A[1] + a2 * B[1]) * oogamma;
          Pc[2] = (a1 * A[2] + a2 * B[2]) * oogamma;
        }
        double nonsph = 0;
        if (screening == kScreenConservative) {  // shell.h:1196-1232
          const double mpa = std::pow(std::max(std::max(std::abs(Pc[0] - A[0]), std::abs(Pc[1] - A[1])),
                                               std::abs(Pc[2] - A[2])), l1);
          const double mpb = std::pow(std::max(std::max(std::abs(Pc[0] - B[0]), std::abs(Pc[1] - B[1])),
                                               std::abs(Pc[2] - B[2])), l2);
          double f1 = 1, f2 = 1;
          for (int k = 2; k <= l1; ++k) f1 *= k;
          for (int k = 2; k <= l2; ++k) f2 *= k;
          const double fl = f1 * f2 * std::pow(oogamma, l1 + l2);
          nonsph = std::max(mpa * mpb, fl);
          const double ln_nonsph = std::log(std::max(nonsph, 1.0));
          const double ln_sph_extra = 1.777485947591722872387900 + std::log(oogamma);
          const double ln_nprim = std::log((double)(np1 * np2));
          ln_screen_fac += ln_sph_extra + ln_nonsph + ln_nprim;
          if (ln_screen_fac < ln_prec) continue;
        }
        PrimPair pp{};
        pp.P[0] = Pc[0]; pp.P[1] = Pc[1]; pp.P[2] = Pc[2];
        const double K = 5.9149671727956128778 * std::exp(minus_rho_times_AB2) * oogamma;
        pp.Kc = K * (bs1->coeff[bs1->off[a] + p1] * bs2->coeff[bs2->off[b] + p2]);
        pp.gamma = gamma;
        pp.oog = oogamma;
        pp.ln_scr = ln_screen_fac;
        pp.nonsph = nonsph;
        P->prim.push_back(pp);
        P->Kraw.push_back(K);  // unscaled K kept for lb200_pairs_get
        P->p1p2.push_back(p1);
        P->p1p2.push_back(p2);
      }
    fac_off += (size_t)np1 * np2;
    P->prim_off[i + 1] = (int)P->prim.size();
    d.max_nprim = std::max(d.max_nprim, P->prim_off[i + 1] - P->prim_off[i]);
  }
  P->nprim_total = (long long)P->prim.size();
  const int rc = upload_pairs(ctx, P, geom, pair_schwarz, true);
  if (rc) { delete P; return rc; }
  *out = P;
  return LB200_OK;
}

namespace {

struct BatchPlan {
  bool swap;        // caller's bra becomes the kernel's ket
  int la, lb, lc, ld;
};

int plan_batch(const lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket,
               BatchPlan& pl) {
  const int kb = order_key(bra->dev.la, bra->dev.lb), kk = order_key(ket->dev.la, ket->dev.lb);
  pl.swap = kb < kk;
  const PairBlock& B = pl.swap ? ket->dev : bra->dev;
  const PairBlock& Kt = pl.swap ? bra->dev : ket->dev;
  pl.la = B.la; pl.lb = B.lb; pl.lc = Kt.la; pl.ld = Kt.lb;
  if (!class_supported(pl.la, pl.lb, pl.lc, pl.ld))
    return set_error(ctx, LB200_ERR_LMAX, "no kernel built for this angular-momentum class");
  return LB200_OK;
}

}  // namespace

// run the store-mode kernel for `ntasks` device-resident tasks into device buffer `d_out`
// (Cartesian, caller's bra-ket order)
int run_store(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket,
              long long ntasks, const int2* d_tasks, int screening, double precision,
              double* d_out, const ProductTasks* prod) {
  BatchPlan pl;
  int rc = plan_batch(ctx, bra, ket, pl);
  if (rc) return rc;
  if (ntasks < 0 || ntasks > 0xffffffffll)
    return set_error(ctx, LB200_ERR_INVALID, "more than 2^32 - 1 tasks in one launch");
  EriParams p{};
  p.bra = pl.swap ? ket->dev : bra->dev;
  p.ket = pl.swap ? bra->dev : ket->dev;
  p.tasks = d_tasks;
  if (prod) {   // implicit (bra range) x (ket range) list
    p.tasks = nullptr;
    p.prod_nk = (unsigned)prod->nk; p.prod_b0 = prod->b0; p.prod_k0 = prod->k0;
  }
  p.ntasks_dev = nullptr;
  p.ntasks = (unsigned)ntasks;
  p.swap_tasks = pl.swap ? 1 : 0;
  // LB200_NO_PRIM_KERNEL=1 forces the general (contraction-loop) kernel, for A/B timing
  static const bool no_prim = getenv("LB200_NO_PRIM_KERNEL") != nullptr;
  p.uncontracted = (!no_prim && p.bra.max_nprim <= 1 && p.ket.max_nprim <= 1) ? 1 : 0;
  p.boys = ctx->d_boys;
  p.screening = screening;
  if (precision > 0.) {  // Engine::set_precision, engine.h:809-826
    p.precision = precision;
    p.ln_precision = std::log(precision);
  } else {
    p.precision = 0.;
    p.ln_precision = std::numeric_limits<double>::lowest();
  }
  p.out = d_out;
  p.out_stride = (long long)nc(pl.la) * nc(pl.lb) * nc(pl.lc) * nc(pl.ld);
  p.transpose_out = pl.swap ? 1 : 0;
  cudaError_t e = launch_eri(pl.la, pl.lb, pl.lc, pl.ld, p, ctx->d_rows, kModeStoreCart,
                             ctx->num_sms, ctx->stream);
  ++ctx->launches;
  return check_cuda(ctx, e, "launch eri store kernel");
}

}  // namespace lb200

extern "C" {

int lb200_pairs_create(lb200_context* ctx, const lb200_basis* bs1, const lb200_basis* bs2,
                       int npair, const int* s1, const int* s2, int screening, double ln_prec,
                       const double* prim_schwarz, lb200_pairs** out) {
  return build_pairs(ctx, bs1, bs2, npair, s1, s2, screening, ln_prec, prim_schwarz, nullptr, out);
}

int lb200_pairs_destroy(lb200_pairs* p) {
  if (!p) return LB200_OK;
  cudaSetDevice(p->ctx->device);
  free_deriv_blocks(p);
  cudaFree(p->d_block);
  delete p;
  return LB200_OK;
}

int lb200_pairs_info(const lb200_pairs* p, long long* info) {
  if (!p || !info) return LB200_ERR_INVALID;
  info[0] = p->dev.la; info[1] = p->dev.lb; info[2] = p->dev.npair;
  info[3] = p->nprim_total; info[4] = p->dev.pure_a; info[5] = p->dev.pure_b;
  return LB200_OK;
}

int lb200_pairs_get(const lb200_pairs* p, int i, double* out, int cap) {
  if (!p || i < 0 || i >= p->dev.npair) return LB200_ERR_INVALID;
  if (pairs_host_mirror(p)) return LB200_ERR_CUDA;
  const int b = p->prim_off[i], e = p->prim_off[i + 1];
  if (e - b > cap) return LB200_ERR_INVALID;
  for (int k = b; k < e; ++k) {
    const PrimPair& pp = p->prim[k];
    double* o = out + 9 * (k - b);
    o[0] = pp.P[0]; o[1] = pp.P[1]; o[2] = pp.P[2];
    o[3] = p->Kraw[k]; o[4] = pp.oog; o[5] = pp.nonsph; o[6] = pp.ln_scr;
    o[7] = p->p1p2[2 * k]; o[8] = p->p1p2[2 * k + 1];
  }
  return e - b;
}

int lb200_eri_class_supported(int la, int lb, int lc, int ld) {
  if (la < lb || lc < ld || lb < 0 || ld < 0) return 0;
  if (order_key(la, lb) < order_key(lc, ld)) return class_supported(lc, ld, la, lb) ? 1 : 0;
  return class_supported(la, lb, lc, ld) ? 1 : 0;
}

long long lb200_eri_block_size(const lb200_pairs* bra, const lb200_pairs* ket, int pure_out) {
  if (!bra || !ket) return LB200_ERR_INVALID;
  auto sz = [&](int l, int pure) { return (pure_out && pure) ? npure(l) : nc(l); };
  return (long long)sz(bra->dev.la, bra->dev.pure_a) * sz(bra->dev.lb, bra->dev.pure_b) *
         sz(ket->dev.la, ket->dev.pure_a) * sz(ket->dev.lb, ket->dev.pure_b);
}

int lb200_eri_product(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket, int b0, int nb,
                      int k0, int nk, int screening, double precision, int pure_out, double* out) {
  if (!ctx || !bra || !ket || !out || b0 < 0 || nb < 0 || k0 < 0 || nk < 0 || b0 + nb > bra->dev.npair ||
      k0 + nk > ket->dev.npair)
    return LB200_ERR_INVALID;
  const long long ntasks = (long long)nb * nk;
  if (ntasks == 0) return LB200_OK;
  cudaSetDevice(ctx->device);
  const int l[4] = {bra->dev.la, bra->dev.lb, ket->dev.la, ket->dev.lb};
  const int pure[4] = {bra->dev.pure_a, bra->dev.pure_b, ket->dev.pure_a, ket->dev.pure_b};
  const bool need_tform = pure_out && ((pure[0] && l[0] > 0) || (pure[1] && l[1] > 0) ||
                                       (pure[2] && l[2] > 0) || (pure[3] && l[3] > 0));
  ProductTasks pt{b0, nb, k0, nk};
  if (!need_tform) return run_store(ctx, bra, ket, ntasks, nullptr, screening, precision, out, &pt);
  // Cartesian integrals into scratch, transformed into the caller's buffer
  const long long ncart_blk = (long long)nc(l[0]) * nc(l[1]) * nc(l[2]) * nc(l[3]);
  const size_t need = (size_t)ntasks * ncart_blk * 8;
  if (ctx->scratch_bytes[1] < need) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_scratch[1]);
    ctx->d_scratch[1] = nullptr; ctx->scratch_bytes[1] = 0;
    int rc = check_cuda(ctx, cudaMalloc(&ctx->d_scratch[1], need), "cudaMalloc(scratch)");
    if (rc) return rc;
    ctx->scratch_bytes[1] = need;
  }
  double* d_cart = static_cast<double*>(ctx->d_scratch[1]);
  int rc = run_store(ctx, bra, ket, ntasks, nullptr, screening, precision, d_cart, &pt);
  if (rc) return rc;
  rc = check_cuda(ctx, launch_pure_transform(ctx, d_cart, out, ntasks, l, pure, ctx->stream), "pure transform");
  ++ctx->launches;
  return rc;
}

int lb200_eri_batch(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket,
                    long long ntasks, const int* tasks, int tasks_on_device, int screening,
                    double precision, int pure_out, double* out, int out_on_device) {
  if (!ctx || !bra || !ket || ntasks < 0 || (ntasks > 0 && (!tasks || !out)))
    return LB200_ERR_INVALID;
  if (ntasks == 0) return LB200_OK;
  cudaSetDevice(ctx->device);
  const int l[4] = {bra->dev.la, bra->dev.lb, ket->dev.la, ket->dev.lb};
  const int pure[4] = {bra->dev.pure_a, bra->dev.pure_b, ket->dev.pure_a, ket->dev.pure_b};
  const bool need_tform = pure_out && ((pure[0] && l[0] > 0) || (pure[1] && l[1] > 0) ||
                                       (pure[2] && l[2] > 0) || (pure[3] && l[3] > 0));
  const long long ncart_blk = (long long)nc(l[0]) * nc(l[1]) * nc(l[2]) * nc(l[3]);
  const long long nout_blk = lb200_eri_block_size(bra, ket, need_tform ? 1 : 0);
  // chunking: bounded device scratch when the result goes to the host or needs a transform
  const long long chunk_bytes = 1ll << 28;  // 256 MiB of Cartesian integrals per chunk
  long long chunk = std::max(1ll, chunk_bytes / (ncart_blk * 8));
  if (chunk > ntasks) chunk = ntasks;
  const bool direct = out_on_device && !need_tform;
  int rc = LB200_OK;
  // device scratch lives in the context and only grows: no cudaMalloc / cudaFree (= device
  // synchronisation) per call
  auto scratch = [&](int slot, size_t bytes, void** out_ptr) -> int {
    if (ctx->scratch_bytes[slot] < bytes) {
      cudaStreamSynchronize(ctx->stream);
      if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
      cudaFree(ctx->d_scratch[slot]);
      ctx->d_scratch[slot] = nullptr;
      ctx->scratch_bytes[slot] = 0;
      int r = check_cuda(ctx, cudaMalloc(&ctx->d_scratch[slot], bytes), "cudaMalloc(scratch)");
      if (r) return r;
      ctx->scratch_bytes[slot] = bytes;
    }
    *out_ptr = ctx->d_scratch[slot];
    return LB200_OK;
  };
  int2* d_tasks = nullptr;
  double* d_cart[2] = {nullptr, nullptr};
  double* d_pure[2] = {nullptr, nullptr};
  if (tasks_on_device) {
    d_tasks = reinterpret_cast<int2*>(const_cast<int*>(tasks));
  } else {
    if ((rc = scratch(0, ntasks * sizeof(int2), reinterpret_cast<void**>(&d_tasks)))) return rc;
    cudaMemcpyAsync(d_tasks, tasks, ntasks * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream);
  }
  if (direct) {
    rc = run_store(ctx, bra, ket, ntasks, d_tasks, screening, precision, out);
  } else {
    const int nbuf = (!out_on_device && ntasks > chunk) ? 2 : 1;
    for (int b = 0; b < nbuf && !rc; ++b) {
      rc = scratch(1 + b, chunk * ncart_blk * 8, reinterpret_cast<void**>(&d_cart[b]));
      if (!rc && need_tform && !out_on_device)
        rc = scratch(3 + b, chunk * nout_blk * 8, reinterpret_cast<void**>(&d_pure[b]));
    }
    cudaStream_t copy_stream = nullptr;
    if (!rc && !out_on_device) {
      if (!ctx->copy_stream) {
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
        for (int b = 0; b < 2; ++b) {
          cudaEventCreateWithFlags(&ctx->ev_done[b], cudaEventDisableTiming);
          cudaEventCreateWithFlags(&ctx->ev_free[b], cudaEventDisableTiming);
        }
      }
      copy_stream = ctx->copy_stream;
    }
    int b = 0;
    bool used[2] = {false, false};
    for (long long t0 = 0; t0 < ntasks && !rc; t0 += chunk, b = (b + 1) % nbuf) {
      const long long nt = std::min(chunk, ntasks - t0);
      // D2H of this buffer's previous chunk done (events of earlier calls are complete: every
      // host-output call ends with a copy-stream synchronisation)
      if (copy_stream && used[b]) cudaStreamWaitEvent(ctx->stream, ctx->ev_free[b], 0);
      rc = run_store(ctx, bra, ket, nt, d_tasks + t0, screening, precision, d_cart[b]);
      if (rc) break;
      const double* src = d_cart[b];
      if (need_tform) {
        double* dst = out_on_device ? out + t0 * nout_blk : d_pure[b];
        rc = check_cuda(ctx, launch_pure_transform(ctx, d_cart[b], dst, nt, l, pure, ctx->stream),
                        "pure transform");
        ++ctx->launches;
        src = dst;
      }
      if (!out_on_device) {
        cudaEventRecord(ctx->ev_done[b], ctx->stream);
        cudaStreamWaitEvent(copy_stream, ctx->ev_done[b], 0);
        cudaMemcpyAsync(out + t0 * nout_blk, src, nt * nout_blk * 8, cudaMemcpyDeviceToHost,
                        copy_stream);
        cudaEventRecord(ctx->ev_free[b], copy_stream);
        used[b] = true;
      }
    }
    if (copy_stream) cudaStreamSynchronize(copy_stream);
  }
  if (!rc && (!out_on_device || !tasks_on_device))
    rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "eri_batch");
  else if (!rc)
    rc = check_cuda(ctx, cudaGetLastError(), "eri_batch");
  return rc;
}

}  // extern "C"

#include "internal_host.h"
lb200_basis_view lb200_view(const lb200_basis* bs) {
  return lb200_basis_view{bs->nshell,    bs->l.data(), bs->pure.data(),  bs->nprim.data(),
                          bs->off.data(), bs->O.data(), bs->alpha.data(), bs->coeff.data()};
}
