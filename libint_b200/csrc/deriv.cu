// First geometric derivatives of the Coulomb integrals and the two-body gradient that consumes them:
// Engine::compute2<Operator::coulomb, BraKet::xx_xx, 1> (include/libint2/engine.impl.h:1151-2113 with
// deriv_order_ = 1: twelve shell sets per quartet, derivative index 3 * centre + xyz) and the force
// contraction of compute_2body_fock_deriv<1> (tests/hartree-fock/hartree-fock++.cc:1775-2055) followed
// by F2(atom, xyz) = sum G1[i] o D (:648-656).
//
// GPU design.  libint generates a second family of kernels for derivatives.  Here the derivative of a
// contracted shell set is assembled from ORDINARY shell sets of the production class kernels with one
// angular momentum shifted,
//     d/dA_x (ab|cd) = (a+1_x b|cd)[coefficients * 2 alpha_a] - a_x (a-1_x b|cd),
// the relation the reference's own derivative check is built on (src/bin/test_eri/eri.h:383-460).  The
// exponent factor is folded into K * c_a * c_b of a TWIN of the pair block (same pairs, same surviving
// primitives, same geometry; class (la+1 lb|, (la-1 lb|, (la lb+1|, (la lb-1|), so one task list drives
// six store-mode launches per (bra class, ket class): A+, A-, B+, B-, C+, C-.  The fourth centre follows
// from translational invariance.  A combine kernel then either writes the twelve sets (batched
// compute2<..., 1>) or contracts them with the two-particle density of the Fock digestion
// (2 D_ab D_cd - 1/2 D_ac D_bd - 1/2 D_ad D_bc, the trace of :1832-1852 with D) straight into the
// 3 * natoms gradient -- the 3 * natoms Fock-derivative matrices of the reference are never formed.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "internal.h"

using namespace lb200;

struct lb200_deriv_blocks {
  lb200_pairs* v[4] = {nullptr, nullptr, nullptr, nullptr};   // A+, A-, B+, B-
  int swapped[4] = {0, 0, 0, 0};   // 1: the twin lists the pair as (second shell, first shell)
};

namespace {

// twin of S with one angular momentum shifted: which = 0 (la+1, Kc * 2 alpha_a), 1 (la-1), 2 (lb+1,
// Kc * 2 alpha_b), 3 (lb-1).  *out = null if the lowered shell does not exist.
int derive_pairs(const lb200_pairs* S, int which, lb200_pairs** out, int* swapped) {
  *out = nullptr;
  *swapped = 0;
  const int la = S->dev.la, lb = S->dev.lb;
  const int na = la + (which == 0) - (which == 1), nb = lb + (which == 2) - (which == 3);
  if (na < 0 || nb < 0) return LB200_OK;
  int rc = pairs_host_mirror(S);
  if (rc) return rc;
  lb200_context* ctx = S->ctx;
  const bool sw = na < nb;
  auto* P = new lb200_pairs;
  P->ctx = ctx;
  PairBlock& d = P->dev;
  d.npair = S->dev.npair;
  d.la = sw ? nb : na;
  d.lb = sw ? na : nb;
  d.pure_a = d.pure_b = 0;
  d.unit_b = 0;
  d.max_nprim = S->dev.max_nprim;
  P->prim_off = S->prim_off;
  P->prim = S->prim;
  P->nprim_total = (long long)P->prim.size();
  const size_t n = (size_t)d.npair;
  if (which == 0 || which == 2)
    for (size_t i = 0; i < n; ++i) {
      const int sa = S->shell[2 * i], sb = S->shell[2 * i + 1];
      for (int k = S->prim_off[i]; k < S->prim_off[i + 1]; ++k) {
        const double alpha = which == 0 ? S->alpha1[S->off1[sa] + S->p1p2[2 * (size_t)k]]
                                        : S->alpha2[S->off2[sb] + S->p1p2[2 * (size_t)k + 1]];
        P->prim[k].Kc *= 2.0 * alpha;
      }
    }
  P->shell.resize(2 * n);
  P->bf.assign(2 * n, 0);
  std::vector<PairGeom> geom(n);
  for (size_t i = 0; i < n; ++i) {
    P->shell[2 * i] = S->shell[2 * i + (sw ? 1 : 0)];
    P->shell[2 * i + 1] = S->shell[2 * i + (sw ? 0 : 1)];
    PairGeom& g = geom[i];
    for (int k = 0; k < 3; ++k) {
      const double A = S->A[3 * i + k], AB = S->AB[3 * i + k];
      g.A[k] = sw ? A - AB : A;    // first shell of the twin: B = A - (A - B)
      g.AB[k] = sw ? -AB : AB;
    }
    g.bf[0] = g.bf[1] = 0;
    g.shell[0] = P->shell[2 * i]; g.shell[1] = P->shell[2 * i + 1];
  }
  rc = upload_pairs(ctx, P, geom, nullptr, true);
  if (rc) { delete P; return rc; }
  *out = P;
  *swapped = sw ? 1 : 0;
  return LB200_OK;
}

int get_deriv_blocks(const lb200_pairs* Pc, lb200_deriv_blocks** out) {
  auto* P = const_cast<lb200_pairs*>(Pc);
  if (!P->deriv) {
    if (P->dev.unit_b)
      return set_error(P->ctx, LB200_ERR_INVALID, "derivatives of unit-shell (3-centre) blocks are not built");
    auto* db = new lb200_deriv_blocks;
    for (int w = 0; w < 4; ++w) {
      const int rc = derive_pairs(P, w, &db->v[w], &db->swapped[w]);
      if (rc) {
        for (int k = 0; k < 4; ++k)
          if (db->v[k]) lb200_pairs_destroy(db->v[k]);
        delete db;
        return rc;
      }
    }
    P->deriv = db;
  }
  *out = P->deriv;
  return LB200_OK;
}

struct CombineParams {
  long long ntasks;
  int l[4], n[4];
  DerivBuf buf[6];
};

// component index of q + e_d in the shell l + 1 / of q - e_d in the shell l - 1 (STANDARD ordering, cart.cuh)
__device__ __forceinline__ int comp_shift(int l, C3 q, int d, int s) {
  const C3 r = cadd(q, d, s);
  return cidx(l + s, r.x, r.y);
}

// the nine explicit derivative values of one Cartesian element (a, b, c, d) of task t: dv[3 * centre + xyz]
__device__ __forceinline__ void deriv_element(const CombineParams& p, long long t, const int (&i)[4], double (&dv)[9]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const DerivBuf& bp = p.buf[2 * c];
    const DerivBuf& bm = p.buf[2 * c + 1];
    const C3 q = cxyz(p.l[c], i[c]);
    long long rest_p = t * bp.blk, rest_m = t * bm.blk;
#pragma unroll
    for (int x = 0; x < 4; ++x)
      if (x != c) {
        rest_p += (long long)i[x] * bp.s[x];
        rest_m += (long long)i[x] * bm.s[x];
      }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double v = bp.p[rest_p + (long long)comp_shift(p.l[c], q, d, +1) * bp.s[c]];
      const int qd = cget(q, d);
      if (qd > 0) v -= (double)qd * bm.p[rest_m + (long long)comp_shift(p.l[c], q, d, -1) * bm.s[c]];
      dv[3 * c + d] = v;
    }
  }
}

__global__ void deriv_store_kernel(const CombineParams p, double* __restrict__ out) {
  const long long nblk = (long long)p.n[0] * p.n[1] * p.n[2] * p.n[3];
  const long long total = p.ntasks * nblk;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long t = g / nblk;
    int r = (int)(g - t * nblk);
    int i[4];
    i[3] = r % p.n[3]; r /= p.n[3];
    i[2] = r % p.n[2]; r /= p.n[2];
    i[1] = r % p.n[1]; r /= p.n[1];
    i[0] = r;
    double dv[9];
    deriv_element(p, t, i, dv);
    double* o = out + t * 12 * nblk + (g - t * nblk);
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k * nblk] = dv[k];
    // translational invariance: d/dD = -(d/dA + d/dB + d/dC)
#pragma unroll
    for (int d = 0; d < 3; ++d) o[(9 + d) * nblk] = -(dv[d] + dv[3 + d] + dv[6 + d]);
  }
}

// one warp per task
__global__ void deriv_grad_kernel(const CombineParams p, const DerivGradParams gp) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nblk = p.n[0] * p.n[1] * p.n[2] * p.n[3];
  for (long long t = w0; t < p.ntasks; t += nw) {
    const int2 tk = gp.tasks[t];
    const double deg = (double)(1 << ((unsigned)gp.ftasks[t].y >> 30));
    int sh[4];
    sh[0] = gp.bra_shell[2 * tk.x]; sh[1] = gp.bra_shell[2 * tk.x + 1];
    sh[2] = gp.ket_shell[2 * tk.y]; sh[3] = gp.ket_shell[2 * tk.y + 1];
    int cb[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) cb[x] = gp.shell2cbf[sh[x]];
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    const double* Dc = gp.Dc;
    const long long nb = gp.nbfc;
    for (int e = lane; e < nblk; e += 32) {
      int r = e;
      int i[4];
      i[3] = r % p.n[3]; r /= p.n[3];
      i[2] = r % p.n[2]; r /= p.n[2];
      i[1] = r % p.n[1]; r /= p.n[1];
      i[0] = r;
      const long long a = cb[0] + i[0], b = cb[1] + i[1], c = cb[2] + i[2], d = cb[3] + i[3];
      const double gam = 2.0 * Dc[a * nb + b] * Dc[c * nb + d] - 0.5 * Dc[a * nb + c] * Dc[b * nb + d] -
                         0.5 * Dc[a * nb + d] * Dc[b * nb + c];
      double dv[9];
      deriv_element(p, t, i, dv);
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[k] += gam * dv[k];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      double v = acc[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[k] = v * deg;
    }
    if (lane == 0) {
      int at[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) at[x] = gp.shell2atom[sh[x]];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        atomicAdd(gp.grad + 3 * at[0] + d, acc[d]);
        atomicAdd(gp.grad + 3 * at[1] + d, acc[3 + d]);
        atomicAdd(gp.grad + 3 * at[2] + d, acc[6 + d]);
        atomicAdd(gp.grad + 3 * at[3] + d, -(acc[d] + acc[3 + d] + acc[6 + d]));
      }
    }
  }
}

struct CartDensParams {
  const double* D;
  double* Dc;
  int nbf, nbfc;
  const int *l, *pure, *shell2bf, *shell2cbf, *cbf2shell;
  const int *rowptr, *col, *base;
  const double* val;
};

// C[p][c] of shell (l, pure): sparse row p of the cart -> pure table, or the identity
__device__ __forceinline__ double sph_coef(const CartDensParams& p, int l, int pure, int pi, int ci) {
  if (!pure) return pi == ci ? 1.0 : 0.0;
  constexpr int RP = 2 * kMaxShellL + 2;
  for (int k = p.rowptr[l * RP + pi]; k < p.rowptr[l * RP + pi + 1]; ++k)
    if (p.col[p.base[l] + k] == ci) return p.val[p.base[l] + k];
  return 0.0;
}

__global__ void cartesianize_density_kernel(const CartDensParams p) {
  const long long total = (long long)p.nbfc * p.nbfc;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(g / p.nbfc), j = (int)(g % p.nbfc);
    const int s1 = p.cbf2shell[i], s2 = p.cbf2shell[j];
    const int c1 = i - p.shell2cbf[s1], c2 = j - p.shell2cbf[s2];
    const int l1 = p.l[s1], l2 = p.l[s2], pu1 = p.pure[s1], pu2 = p.pure[s2];
    const int n1 = pu1 ? 2 * l1 + 1 : nc(l1), n2 = pu2 ? 2 * l2 + 1 : nc(l2);
    const int b1 = p.shell2bf[s1], b2 = p.shell2bf[s2];
    double v = 0.0;
    for (int p1 = 0; p1 < n1; ++p1) {
      const double x1 = sph_coef(p, l1, pu1, p1, c1);
      if (x1 == 0.0) continue;
      for (int p2 = 0; p2 < n2; ++p2) {
        const double x2 = sph_coef(p, l2, pu2, p2, c2);
        if (x2 != 0.0) v += x1 * x2 * p.D[(long long)(b1 + p1) * p.nbf + b2 + p2];
      }
    }
    p.Dc[g] = v;
  }
}

__global__ void unpack_tasks_kernel(const int4* __restrict__ ft, const unsigned* __restrict__ count,
                                    int2* __restrict__ tasks, long long cap) {
  const long long n = min((long long)*count, cap);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int4 f = ft[t];
    tasks[t] = make_int2(f.x, f.y & 0x3fffffff);
  }
}

CombineParams make_combine(const DerivSets& ds, long long ntasks) {
  CombineParams p;
  p.ntasks = ntasks;
  for (int x = 0; x < 4; ++x) { p.l[x] = ds.l[x]; p.n[x] = ds.n[x]; }
  for (int k = 0; k < 6; ++k) p.buf[k] = ds.buf[k];
  return p;
}

}  // namespace

namespace lb200 {

void free_deriv_blocks(lb200_pairs* P) {
  if (!P->deriv) return;
  for (int k = 0; k < 4; ++k)
    if (P->deriv->v[k]) lb200_pairs_destroy(P->deriv->v[k]);
  delete P->deriv;
  P->deriv = nullptr;
}

int deriv_plan(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket, DerivSets& ds) {
  return deriv_plan_class(ctx, bra->dev.la, bra->dev.lb, ket->dev.la, ket->dev.lb, ds);
}

int deriv_plan_class(lb200_context* ctx, int la, int lb, int lc, int ld, DerivSets& ds) {
  ds.l[0] = la; ds.l[1] = lb; ds.l[2] = lc; ds.l[3] = ld;
  for (int x = 0; x < 4; ++x) ds.n[x] = nc(ds.l[x]);
  ds.doubles_per_task = 0;
  for (int k = 0; k < 6; ++k) {
    const int c = k / 2, sgn = (k & 1) ? -1 : +1;
    DerivBuf& b = ds.buf[k];
    b.p = nullptr;
    b.blk = 0;
    for (int x = 0; x < 4; ++x) b.s[x] = 0;
    if (ds.l[c] + sgn < 0) continue;
    int m[4];
    for (int x = 0; x < 4; ++x) m[x] = nc(ds.l[x] + (x == c ? sgn : 0));
    // class of the shifted side, ordered as its twin block lists it (higher angular momentum first)
    int pa = ds.l[c < 2 ? 0 : 2] + (c == 0 || c == 2 ? sgn : 0);
    int pb = ds.l[c < 2 ? 1 : 3] + (c == 1 ? sgn : 0);
    const bool sw = pa < pb;
    if (sw) std::swap(pa, pb);
    const int oa = c < 2 ? ds.l[2] : ds.l[0], ob = c < 2 ? ds.l[3] : ds.l[1];   // the unshifted side
    if (!lb200_eri_class_supported(pa, pb, oa, ob))
      return set_error(ctx, LB200_ERR_LMAX, "no kernel built for a shifted class of this derivative");
    // layout written by run_store: [bra.first][bra.second][ket.first][ket.second] of the blocks handed to it
    if (c < 2) {
      const int f = sw ? 1 : 0, s = sw ? 0 : 1;   // which original index comes first inside the bra twin
      b.s[3] = 1; b.s[2] = m[3];
      b.s[s] = m[2] * m[3];
      b.s[f] = m[s] * m[2] * m[3];
    } else {
      b.s[3] = 1; b.s[2] = m[3];   // only the first ket shell is shifted; A+ / A- twins with la == lb swap
      if (sw) { b.s[2] = 1; b.s[3] = m[2]; }
      b.s[1] = m[2] * m[3];
      b.s[0] = m[1] * m[2] * m[3];
    }
    b.blk = (long long)m[0] * m[1] * m[2] * m[3];
    ds.doubles_per_task += b.blk;
  }
  return LB200_OK;
}

int deriv_eval(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket, long long ntasks,
               const int2* d_tasks, int screening, double precision, double* scratch, DerivSets& ds) {
  lb200_deriv_blocks *db = nullptr, *dk = nullptr;
  int rc = get_deriv_blocks(bra, &db);
  if (!rc) rc = get_deriv_blocks(ket, &dk);
  if (rc) return rc;
  double* cur = scratch;
  for (int k = 0; k < 6 && !rc; ++k) {
    DerivBuf& b = ds.buf[k];
    if (b.blk == 0) { b.p = nullptr; continue; }
    const lb200_pairs* B = k < 4 ? db->v[k] : bra;
    const lb200_pairs* K = k < 4 ? ket : dk->v[k - 4];
    rc = run_store(ctx, B, K, ntasks, d_tasks, screening, precision, cur);
    b.p = cur;
    cur += b.blk * ntasks;
  }
  return rc;
}

cudaError_t launch_deriv_store(const DerivSets& ds, long long ntasks, double* out, cudaStream_t st) {
  if (ntasks == 0) return cudaSuccess;
  const CombineParams p = make_combine(ds, ntasks);
  const long long total = ntasks * ds.n[0] * ds.n[1] * ds.n[2] * ds.n[3];
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 32);
  deriv_store_kernel<<<grid, 256, 0, st>>>(p, out);
  return cudaGetLastError();
}

cudaError_t launch_deriv_grad(const DerivSets& ds, long long ntasks, const DerivGradParams& gp, cudaStream_t st) {
  if (ntasks == 0) return cudaSuccess;
  const CombineParams p = make_combine(ds, ntasks);
  const int grid = (int)std::min<long long>((ntasks + 3) / 4, 148 * 16);
  deriv_grad_kernel<<<grid, 128, 0, st>>>(p, gp);
  return cudaGetLastError();
}

cudaError_t launch_cartesianize_density(const lb200_context* ctx, const double* D, int nbf, double* Dc, int nbfc,
                                        int /*nshell*/, const int* d_l, const int* d_pure, const int* d_shell2bf,
                                        const int* d_shell2cbf, const int* d_cbf2shell, cudaStream_t st) {
  CartDensParams p;
  p.D = D; p.Dc = Dc; p.nbf = nbf; p.nbfc = nbfc;
  p.l = d_l; p.pure = d_pure; p.shell2bf = d_shell2bf; p.shell2cbf = d_shell2cbf; p.cbf2shell = d_cbf2shell;
  p.rowptr = ctx->d_sph_rowptr; p.col = ctx->d_sph_col; p.base = ctx->d_sph_base; p.val = ctx->d_sph_val;
  const long long total = (long long)nbfc * nbfc;
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 32);
  cartesianize_density_kernel<<<grid, 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_unpack_tasks(const int4* ftasks, const unsigned* count, int2* tasks, cudaStream_t st,
                                long long cap) {
  unpack_tasks_kernel<<<148 * 4, 256, 0, st>>>(ftasks, count, tasks, cap);
  return cudaGetLastError();
}

}  // namespace lb200

extern "C" {

int lb200_eri_deriv1_plan(int la, int lb, int lc, int ld, long long* plan) {
  if (!plan || la < lb || lc < ld || lb < 0 || ld < 0) return LB200_ERR_INVALID;
  DerivSets ds;
  const int rc = deriv_plan_class(nullptr, la, lb, lc, ld, ds);
  if (rc) return rc;
  for (int k = 0; k < 6; ++k) {
    plan[5 * k] = ds.buf[k].blk;
    for (int x = 0; x < 4; ++x) plan[5 * k + 1 + x] = ds.buf[k].s[x];
  }
  return LB200_OK;
}

int lb200_eri_deriv1_batch(lb200_context* ctx, const lb200_pairs* bra, const lb200_pairs* ket, long long ntasks,
                           const int* tasks, int tasks_on_device, int screening, double precision, int pure_out,
                           double* out, int out_on_device) {
  if (!ctx || !bra || !ket || ntasks < 0 || (ntasks > 0 && (!tasks || !out))) return LB200_ERR_INVALID;
  if (ntasks == 0) return LB200_OK;
  cudaSetDevice(ctx->device);
  DerivSets ds;
  int rc = deriv_plan(ctx, bra, ket, ds);
  if (rc) return rc;
  const int pure[4] = {bra->dev.pure_a, bra->dev.pure_b, ket->dev.pure_a, ket->dev.pure_b};
  const bool tform = pure_out && ((pure[0] && ds.l[0] > 0) || (pure[1] && ds.l[1] > 0) ||
                                  (pure[2] && ds.l[2] > 0) || (pure[3] && ds.l[3] > 0));
  const long long ncart = (long long)ds.n[0] * ds.n[1] * ds.n[2] * ds.n[3];
  const long long nout = lb200_eri_block_size(bra, ket, tform ? 1 : 0);
  const long long per_task = ds.doubles_per_task + 12 * ncart + (tform ? 12 * nout : 0);
  const long long chunk = std::max(1ll, std::min(ntasks, (1ll << 29) / (per_task * 8)));
  int2* d_tasks = nullptr;
  if (tasks_on_device) {
    d_tasks = reinterpret_cast<int2*>(const_cast<int*>(tasks));
  } else {
    if ((rc = ctx_scratch(ctx, 0, ntasks * sizeof(int2), reinterpret_cast<void**>(&d_tasks)))) return rc;
    cudaMemcpyAsync(d_tasks, tasks, ntasks * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream);
  }
  double *d_sets = nullptr, *d_cart = nullptr, *d_pure = nullptr;
  if ((rc = ctx_scratch(ctx, 5, chunk * ds.doubles_per_task * 8, reinterpret_cast<void**>(&d_sets)))) return rc;
  const bool direct = out_on_device && !tform;
  if (!direct && (rc = ctx_scratch(ctx, 6, chunk * 12 * ncart * 8, reinterpret_cast<void**>(&d_cart)))) return rc;
  if (tform && !out_on_device &&
      (rc = ctx_scratch(ctx, 7, chunk * 12 * nout * 8, reinterpret_cast<void**>(&d_pure))))
    return rc;
  for (long long t0 = 0; t0 < ntasks && !rc; t0 += chunk) {
    const long long nt = std::min(chunk, ntasks - t0);
    rc = deriv_eval(ctx, bra, ket, nt, d_tasks + t0, screening, precision, d_sets, ds);
    if (rc) break;
    double* cart = direct ? out + t0 * 12 * ncart : d_cart;
    rc = check_cuda(ctx, launch_deriv_store(ds, nt, cart, ctx->stream), "derivative combine");
    ++ctx->launches;
    if (rc) break;
    const double* src = cart;
    if (tform) {
      double* dst = out_on_device ? out + t0 * 12 * nout : d_pure;
      rc = check_cuda(ctx, launch_pure_transform(ctx, cart, dst, nt * 12, ds.l, pure, ctx->stream), "pure transform");
      ++ctx->launches;
      src = dst;
    }
    if (!rc && !out_on_device) {
      cudaMemcpyAsync(out + t0 * 12 * nout, src, nt * 12 * nout * 8, cudaMemcpyDeviceToHost, ctx->stream);
      rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "eri_deriv1_batch");
    }
  }
  if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "eri_deriv1_batch");
  return rc;
}

}  // extern "C"
