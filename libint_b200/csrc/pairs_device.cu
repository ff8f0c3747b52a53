// ShellPair::init (include/libint2/shell.h:1138-1328) for a whole block of shell pairs on the GPU: one
// thread per pair runs the primitive-pair loop -- screening value, P, K*c_a*c_b, gamma, 1/gamma -- first
// to count the survivors, then (after a prefix sum) to write the 64-byte records.  Same formulas, same
// constants and the same operation order as the host loop of build_pairs (context.cu), which stays for
// small blocks; exp/log come from the CUDA math library instead of libm (<= 1 ulp apart).
#include <algorithm>
#include <vector>

#include "internal.h"

using namespace lb200;

namespace lb200 {

struct DevBasis {
  const int *l, *nprim, *off;
  const double *O, *alpha, *coeff, *mlc;
};

struct PairPrimParams {
  int npair;
  const int *s1, *s2;
  DevBasis b1, b2;
  int screening;
  double ln_prec;
  const double* prim_schwarz;     // [sum np1*np2] or null
  const long long* fac_off;       // [npair] offsets into prim_schwarz
  int* count;                     // [npair] (count pass)
  const int* prim_off;            // [npair+1] (fill pass)
  PrimPair* prim;
  double* Kraw;
  int2* p1p2;
};

template <bool FILL>
__global__ void pair_prims_kernel(const PairPrimParams p) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.npair; i += gridDim.x * blockDim.x) {
    const int a = p.s1[i], b = p.s2[i];
    const double A[3] = {p.b1.O[3 * a], p.b1.O[3 * a + 1], p.b1.O[3 * a + 2]};
    const double B[3] = {p.b2.O[3 * b], p.b2.O[3 * b + 1], p.b2.O[3 * b + 2]};
    double AB2 = 0.;
    for (int k = 0; k < 3; ++k) AB2 += (A[k] - B[k]) * (A[k] - B[k]);
    const int np1 = p.b1.nprim[a], np2 = p.b2.nprim[b];
    const int l1 = p.b1.l[a], l2 = p.b2.l[b];
    const int o1 = p.b1.off[a], o2 = p.b2.off[b];
    const bool schwarz = p.screening == kScreenSchwarz || p.screening == kScreenSchwarzInf;
    int n = 0;
    const int base = FILL ? p.prim_off[i] : 0;
    for (int p1 = 0; p1 < np1; ++p1)
      for (int p2 = 0; p2 < np2; ++p2) {
        const double a1 = p.b1.alpha[o1 + p1], a2 = p.b2.alpha[o2 + p2];
        const double mlc1 = p.b1.mlc[o1 + p1], mlc2 = p.b2.mlc[o2 + p2];
        const double gamma = a1 + a2;
        const double oogamma = 1 / gamma;
        const double rho = a1 * a2 * oogamma;
        const double minus_rho_times_AB2 = -rho * AB2;
        double ln_screen_fac;
        if (schwarz) {
          ln_screen_fac = log((double)(np1 * np2) * p.prim_schwarz[p.fac_off[i] + (long long)p1 * np2 + p2]) + mlc1 + mlc2;
          if (ln_screen_fac < p.ln_prec) continue;
        } else {
          ln_screen_fac = minus_rho_times_AB2 + mlc1 + mlc2;
          if (p.screening == kScreenOriginal && ln_screen_fac < p.ln_prec) continue;
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
        if (p.screening == kScreenConservative) {  // shell.h:1196-1232
          const double mpa = pow(fmax(fmax(fabs(Pc[0] - A[0]), fabs(Pc[1] - A[1])), fabs(Pc[2] - A[2])), (double)l1);
          const double mpb = pow(fmax(fmax(fabs(Pc[0] - B[0]), fabs(Pc[1] - B[1])), fabs(Pc[2] - B[2])), (double)l2);
          double f1 = 1, f2 = 1;
          for (int k = 2; k <= l1; ++k) f1 *= k;
          for (int k = 2; k <= l2; ++k) f2 *= k;
          const double fl = f1 * f2 * pow(oogamma, (double)(l1 + l2));
          nonsph = fmax(mpa * mpb, fl);
          const double ln_nonsph = log(fmax(nonsph, 1.0));
          const double ln_sph_extra = 1.777485947591722872387900 + log(oogamma);
          const double ln_nprim = log((double)(np1 * np2));
          ln_screen_fac += ln_sph_extra + ln_nonsph + ln_nprim;
          if (ln_screen_fac < p.ln_prec) continue;
        }
        if constexpr (FILL) {
          const double K = 5.9149671727956128778 * exp(minus_rho_times_AB2) * oogamma;
          PrimPair pp;
          pp.P[0] = Pc[0]; pp.P[1] = Pc[1]; pp.P[2] = Pc[2];
          pp.Kc = K * (p.b1.coeff[o1 + p1] * p.b2.coeff[o2 + p2]);
          pp.gamma = gamma;
          pp.oog = oogamma;
          pp.ln_scr = ln_screen_fac;
          pp.nonsph = nonsph;
          p.prim[base + n] = pp;
          p.Kraw[base + n] = K;
          p.p1p2[base + n] = make_int2(p1, p2);
        }
        ++n;
      }
    if constexpr (!FILL) p.count[i] = n;
  }
}

}  // namespace lb200

namespace lb200 {

DevicePrimBuilder::~DevicePrimBuilder() {
  if (d_tmp) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
  }
  delete static_cast<PairPrimParams*>(params);
}

int DevicePrimBuilder::init(lb200_context* c, const lb200_basis* bs1, const lb200_basis* bs2, int np,
                            const int* s1, const int* s2, int screening_, double ln_prec_,
                            const double* prim_schwarz) {
  ctx = c; npair = np; screening = screening_; ln_prec = ln_prec_;
  cudaSetDevice(ctx->device);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  // primitive-factor offsets of the Schwarz variants: [pair][p1][p2]
  std::vector<long long> fac_off;
  size_t nfac = 0;
  if (prim_schwarz) {
    fac_off.resize(np);
    for (int i = 0; i < np; ++i) {
      fac_off[i] = (long long)nfac;
      nfac += (size_t)bs1->nprim[s1[i]] * bs2->nprim[s2[i]];
    }
  }
  struct Arr { const void* src; size_t bytes; size_t off; };
  const lb200_basis* bs[2] = {bs1, bs2};
  Arr arr[18];
  int na = 0;
  size_t total = 0;
  auto add = [&](const void* src, size_t bytes) {
    arr[na] = Arr{src, bytes, total};
    total = al(total + std::max<size_t>(bytes, 8));
    return na++;
  };
  int ib[2][7];
  for (int k = 0; k < 2; ++k) {
    const lb200_basis* b = bs[k];
    ib[k][0] = add(b->l.data(), b->l.size() * 4);
    ib[k][1] = add(b->nprim.data(), b->nprim.size() * 4);
    ib[k][2] = add(b->off.data(), b->off.size() * 4);
    ib[k][3] = add(b->O.data(), b->O.size() * 8);
    ib[k][4] = add(b->alpha.data(), b->alpha.size() * 8);
    ib[k][5] = add(b->coeff.data(), b->coeff.size() * 8);
    ib[k][6] = add(b->max_ln_coeff.data(), b->max_ln_coeff.size() * 8);
  }
  const int i_s1 = add(s1, (size_t)np * 4), i_s2 = add(s2, (size_t)np * 4);
  const int i_fo = add(fac_off.data(), fac_off.size() * 8);
  const int i_fa = add(prim_schwarz, nfac * 8);
  const size_t o_count = total;
  total = al(total + (size_t)np * 4);
  int rc = check_cuda(ctx, cudaMalloc(&d_tmp, total), "cudaMalloc(pair builder)");
  if (rc) return rc;
  for (int k = 0; k < na; ++k)
    if (arr[k].bytes) cudaMemcpyAsync(d_tmp + arr[k].off, arr[k].src, arr[k].bytes, cudaMemcpyHostToDevice, ctx->stream);
  // the sources are host vectors of the caller / of this frame: finish the copies before returning
  rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "upload pair builder inputs");
  if (rc) return rc;
  auto P = new PairPrimParams;
  params = P;
  auto ptr = [&](int k) { return d_tmp + arr[k].off; };
  DevBasis* db[2] = {&P->b1, &P->b2};
  for (int k = 0; k < 2; ++k) {
    db[k]->l = reinterpret_cast<const int*>(ptr(ib[k][0]));
    db[k]->nprim = reinterpret_cast<const int*>(ptr(ib[k][1]));
    db[k]->off = reinterpret_cast<const int*>(ptr(ib[k][2]));
    db[k]->O = reinterpret_cast<const double*>(ptr(ib[k][3]));
    db[k]->alpha = reinterpret_cast<const double*>(ptr(ib[k][4]));
    db[k]->coeff = reinterpret_cast<const double*>(ptr(ib[k][5]));
    db[k]->mlc = reinterpret_cast<const double*>(ptr(ib[k][6]));
  }
  P->npair = np;
  P->s1 = reinterpret_cast<const int*>(ptr(i_s1));
  P->s2 = reinterpret_cast<const int*>(ptr(i_s2));
  P->screening = screening;
  P->ln_prec = ln_prec;
  P->prim_schwarz = prim_schwarz ? reinterpret_cast<const double*>(ptr(i_fa)) : nullptr;
  P->fac_off = prim_schwarz ? reinterpret_cast<const long long*>(ptr(i_fo)) : nullptr;
  P->count = reinterpret_cast<int*>(d_tmp + o_count);
  P->prim_off = nullptr; P->prim = nullptr; P->Kraw = nullptr; P->p1p2 = nullptr;
  return LB200_OK;
}

int DevicePrimBuilder::count(std::vector<int>& counts) {
  auto* P = static_cast<PairPrimParams*>(params);
  const int grid = std::min((npair + 127) / 128, ctx->num_sms * 16);
  pair_prims_kernel<false><<<grid, 128, 0, ctx->stream>>>(*P);
  ++ctx->launches;
  counts.resize(npair);
  cudaMemcpyAsync(counts.data(), P->count, (size_t)npair * 4, cudaMemcpyDeviceToHost, ctx->stream);
  return check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "count primitive pairs");
}

int DevicePrimBuilder::fill(const int* d_prim_off, PrimPair* d_prim, double* d_Kraw, int2* d_p1p2) {
  auto* P = static_cast<PairPrimParams*>(params);
  P->prim_off = d_prim_off; P->prim = d_prim; P->Kraw = d_Kraw; P->p1p2 = d_p1p2;
  const int grid = std::min((npair + 127) / 128, ctx->num_sms * 16);
  pair_prims_kernel<true><<<grid, 128, 0, ctx->stream>>>(*P);
  ++ctx->launches;
  const int rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "fill primitive pairs");
  delete P;
  params = nullptr;
  return rc;
}

}  // namespace lb200
