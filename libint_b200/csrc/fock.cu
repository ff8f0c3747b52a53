// Direct Fock build on the GPU: the consumer of the ERI path,
// compute_2body_fock of tests/hartree-fock/hartree-fock++.cc:1574-1772, plus its set-up
// (Schwarz matrix :1230-1298, SchwarzInf shell-pair data :1383-1431, shell-block norms of
// D :939-957).  Quartet enumeration is re-designed for the GPU: significant shell pairs are
// grouped by class and sorted by Schwarz bound; for every (bra class, ket class) a
// screening kernel compacts the surviving (bra pair, ket pair) tasks of a row chunk into a
// task list that the fused ERI+digestion kernel of that class consumes without a host
// round trip.  The set of quartets computed and their per-quartet precision are exactly
// the reference's (same predicate, same unique-quartet rule).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>

#include "internal.h"


using namespace lb200;

struct FockClass {
  int la, lb, pa, pb;
  int bucket = 0;   // contraction-degree bucket of the block's pairs
  double* d_dn = nullptr;   // device, [npair]: Dnorm(shell1, shell2) of every pair, refreshed per build
  lb200_pairs* pairs = nullptr;
  std::vector<double> schwarz;  // sorted descending, same order as pairs
};

struct lb200_fock {
  lb200_context* ctx = nullptr;
  lb200_basis obs;  // private copy
  std::vector<double> K;  // Schwarz matrix
  std::vector<FockClass> classes;
  // device work space (lazily sized)
  double* d_D = nullptr;
  double* d_F = nullptr;
  double* d_Dnorm = nullptr;
  double* d_scalar = nullptr;
  int* d_shell2bf = nullptr;
  int* d_shellsize = nullptr;
  // Class pairs are issued round-robin on a few side streams, each with its own task buffer and
  // counters: the tail of one class kernel (a few long quartets on a few SMs) overlaps the screening
  // and the start of the next pair instead of idling the GPU ~1200 times per build.
  static constexpr int kMaxStreams = 8;
  int nstreams = 0;
  cudaStream_t streams[kMaxStreams] = {};
  cudaEvent_t ev_ready = nullptr, ev_done[kMaxStreams] = {};
  int4* d_tasks[kMaxStreams] = {};   // task records (types.cuh: EriParams::ftasks)
  unsigned* d_count[kMaxStreams] = {};   // [0] task count, [1] dynamic-scheduling cursor of the class kernel
  long long task_cap = 0;
  // per (bra class, ket class) timings of the last build, filled when profiling is on
  bool profile = false;
  unsigned long long* d_primcount = nullptr;
  struct ProfRow { int c[6]; double ms, nq, nprim; };
  std::vector<ProfRow> prof;
};

namespace {

// sqrt(max |(ab|ab)|) over the functions of each pair of a block (pure where flagged):
// the quantity both Schwarz set-ups of the reference take from Engine results
// (hartree-fock++.cc:1283-1286 and :1403-1409)
int diag_schwarz_impl(lb200_context* ctx, const lb200_pairs* P, std::vector<double>& out) {
  const long long n = P->dev.npair;
  out.assign(n, 0.0);
  if (n == 0) return LB200_OK;
  const int l[4] = {P->dev.la, P->dev.lb, P->dev.la, P->dev.lb};
  const int pure[4] = {P->dev.pure_a, P->dev.pure_b, P->dev.pure_a, P->dev.pure_b};
  const bool tform = (pure[0] && l[0] > 0) || (pure[1] && l[1] > 0);
  const long long ncart = (long long)nc(l[0]) * nc(l[1]) * nc(l[0]) * nc(l[1]);
  long long npure = 1;
  for (int x = 0; x < 4; ++x) npure *= pure[x] ? lb200::npure(l[x]) : nc(l[x]);
  const long long chunk = std::max(1ll, std::min(n, (1ll << 27) / (ncart * 8)));
  int2* d_tasks = nullptr;
  double *d_cart = nullptr, *d_pure = nullptr, *d_max = nullptr;
  int rc = check_cuda(ctx, cudaMalloc(&d_tasks, chunk * sizeof(int2)), "cudaMalloc");
  if (!rc) rc = check_cuda(ctx, cudaMalloc(&d_cart, chunk * ncart * 8), "cudaMalloc");
  if (!rc && tform) rc = check_cuda(ctx, cudaMalloc(&d_pure, chunk * npure * 8), "cudaMalloc");
  if (!rc) rc = check_cuda(ctx, cudaMalloc(&d_max, chunk * 8), "cudaMalloc");
  std::vector<int2> tasks(chunk);
  for (long long t0 = 0; t0 < n && !rc; t0 += chunk) {
    const long long nt = std::min(chunk, n - t0);
    for (long long i = 0; i < nt; ++i) tasks[i] = make_int2((int)(t0 + i), (int)(t0 + i));
    cudaMemcpyAsync(d_tasks, tasks.data(), nt * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream);
    rc = run_store(ctx, P, P, nt, d_tasks, kScreenOriginal, 0., d_cart);
    if (rc) break;
    const double* src = d_cart;
    long long nblk = ncart;
    if (tform) {
      rc = check_cuda(ctx, launch_pure_transform(ctx, d_cart, d_pure, nt, l, pure, ctx->stream),
                      "pure transform");
      ++ctx->launches;
      src = d_pure;
      nblk = npure;
    }
    if (!rc) rc = check_cuda(ctx, launch_block_absmax(src, d_max, nt, nblk, ctx->stream), "absmax");
    ++ctx->launches;
    if (!rc)
      rc = check_cuda(ctx, cudaMemcpyAsync(out.data() + t0, d_max, nt * 8, cudaMemcpyDeviceToHost,
                                           ctx->stream), "copy");
    if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "diag_schwarz");
  }
  cudaFree(d_tasks); cudaFree(d_cart); cudaFree(d_pure); cudaFree(d_max);
  for (auto& v : out) v = std::sqrt(v);
  return rc;
}

// basis whose shells are the individual primitives of `bs`, coefficient 1
// (Shell::extract_primitive(p, false), shell.h:927-936)
void primitive_basis(const lb200_basis& bs, lb200_basis& out) {
  out.ctx = bs.ctx;
  out.nshell = (int)bs.alpha.size();
  out.l.clear(); out.pure.clear(); out.nprim.clear(); out.off.clear(); out.shell2bf.clear();
  out.O.clear(); out.alpha = bs.alpha; out.coeff.assign(bs.alpha.size(), 1.0);
  out.max_ln_coeff.assign(bs.alpha.size(), 0.0);
  int nbf = 0;
  for (int s = 0; s < bs.nshell; ++s)
    for (int p = 0; p < bs.nprim[s]; ++p) {
      out.l.push_back(bs.l[s]);
      out.pure.push_back(bs.pure[s]);
      out.nprim.push_back(1);
      out.off.push_back((int)out.off.size());
      out.shell2bf.push_back(nbf);
      nbf += bs.size(s);
      for (int k = 0; k < 3; ++k) out.O.push_back(bs.O[3 * s + k]);
    }
  out.off.push_back((int)out.alpha.size());
  out.nbf = nbf;
}

__global__ void shellblock_norm_kernel(const double* __restrict__ D, int nbf, int nshell,
                                       const int* __restrict__ shell2bf,
                                       const int* __restrict__ shellsize,
                                       double* __restrict__ Dnorm) {
  // compute_shellblock_norm, hartree-fock++.cc:939-957: inf-norm (max |element|) per block
  const long long n2 = (long long)nshell * nshell;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n2;
       g += (long long)gridDim.x * blockDim.x) {
    const int s1 = (int)(g / nshell), s2 = (int)(g % nshell);
    const int b1 = shell2bf[s1], b2 = shell2bf[s2], n1 = shellsize[s1], n2_ = shellsize[s2];
    double m = 0.0;
    for (int i = 0; i < n1; ++i)
      for (int j = 0; j < n2_; ++j) m = fmax(m, fabs(D[(long long)(b1 + i) * nbf + b2 + j]));
    Dnorm[g] = m;
  }
}

__global__ void absmax_kernel(const double* __restrict__ x, long long n, double* out) {
  __shared__ double sm[32];
  double m = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    m = fmax(m, fabs(x[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) {
      // values are non-negative: integer compare of the bit patterns orders them
      atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
    }
  }
}

// which rank computes the quartet (bra pair gi | ket pair gj): a mixing hash of the canonical
// index of the BRA pair (the row of the task matrix: the pair of the class with the larger order
// key, or the larger canonical index inside one class).  Whole rows are owned by one rank, so a
// rank enumerates and screens only its own rows (the host zeroes the candidate count of the
// others) -- with ~10^5 rows per build the shares are even to about a percent.  Replaces the
// reference's serial round-robin s1234 % nthreads (hartree-fock++.cc:1665).
__host__ __device__ inline int task_owner(int gi, int /*gj*/, int nranks) {
  unsigned h = (unsigned)gi * 2654435761u;
  h ^= h >> 15;
  h *= 2246822519u;
  h ^= h >> 13;
  return (int)(h % (unsigned)nranks);
}

struct ScreenParams {
  PairBlock bra, ket;
  int same_class;
  int row0, nrow;            // bra rows of this chunk
  int nket;                  // pairs in the ket block
  double thr_num;            // fock_precision / Dmax * (1 - 1e-12): candidate-prefix threshold numerator
  const double* bra_dn;      // [npair] Dnorm of the pair's own block (D12 / D34)
  const double* ket_dn;
  int stage_rows;            // 1: the two Dnorm rows of the bra shells fit in shared memory
  const double* Dnorm;
  int nshell;
  double fock_precision;
  double ln_needed_engine_precision;   // used when max |D| of the quartet is 0 (hf++:1693-1695)
  int use_schwarz;
  int rank, nranks;
  int4* tasks;
  unsigned* count;
  unsigned cap;
};

// per-pair Dnorm(shell1, shell2): the D12 / D34 operand of the screen, one coalesced array per
// block instead of a random gather per candidate
__global__ void pair_dnorm_kernel(const int* __restrict__ shell, int npair, const double* __restrict__ Dn,
                                  int ns, double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += gridDim.x * blockDim.x)
    out[i] = Dn[(size_t)shell[2 * i] * ns + shell[2 * i + 1]];
}

// unique-quartet rule and Schwarz x density screen, hartree-fock++.cc:1628-1677.
// One CTA per bra row: the rows Dnorm(s1, :) and Dnorm(s2, :) -- four of the six |D| operands of
// every candidate of that row -- are staged in shared memory when the row has enough candidates
// to pay for it; D12 and D34 come from the per-pair arrays.  What is left per candidate is 28
// coalesced bytes of ket-pair data and four shared-memory gathers.
__global__ void screen_kernel(const ScreenParams p) {
  extern __shared__ double s_rows[];   // [2][nshell] when p.stage_rows
  const int lane = threadIdx.x & 31;
  const int ns = p.nshell;
  __shared__ unsigned s_jm;
  for (int r = blockIdx.x; r < p.nrow; r += gridDim.x) {
    const int i = p.row0 + r;
    const double Ki = p.bra.schwarz[i];
    const int gi = p.bra.gidx[i];
    // candidate prefix of this row: the kets are sorted by Schwarz bound (descending), and
    // K_i * K_j * Dmax >= precision is necessary for survival -- a binary search instead of a
    // host-made table (the host evaluates the same expression for its capacity planning);
    // rows of other ranks have no candidates
    if (threadIdx.x == 0) {
      unsigned jm0 = (unsigned)p.nket;
      if (p.use_schwarz) {
        const double thr = p.thr_num / Ki;
        int lo = 0, hi = p.nket;   // first j with schwarz[j] < thr
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (p.ket.schwarz[mid] < thr) hi = mid; else lo = mid + 1;
        }
        jm0 = (unsigned)lo;
      }
      if (p.nranks > 1 && task_owner(gi, 0, p.nranks) != p.rank) jm0 = 0;
      s_jm = jm0;
    }
    __syncthreads();
    const unsigned jm = s_jm;
    __syncthreads();   // s_jm is rewritten by the next row
    if (jm == 0) continue;
    const int s1 = p.bra.shell[2 * i], s2 = p.bra.shell[2 * i + 1];
    const double D12 = p.bra_dn[i];
    const double* r1 = p.Dnorm + (size_t)s1 * ns;
    const double* r2 = p.Dnorm + (size_t)s2 * ns;
    const bool staged = p.stage_rows && jm >= 4u * blockDim.x;
    if (staged) {
      __syncthreads();   // previous row's readers are done
      for (int k = threadIdx.x; k < ns; k += blockDim.x) {
        s_rows[k] = r1[k];
        s_rows[ns + k] = r2[k];
      }
      __syncthreads();
      r1 = s_rows;
      r2 = s_rows + ns;
    }
    for (unsigned j0 = 0; j0 < jm; j0 += blockDim.x) {
      const unsigned j = j0 + threadIdx.x;
      bool keep = false;
      double dn = 0.0;   // Dnorm1234 = 0 without Schwarz screening (hartree-fock++.cc:1666-1672)
      int code = 0;      // log2 of the permutational degeneracy (hartree-fock++.cc:1683-1687)
      if (j < jm) {
        const int gj = p.ket.gidx[j];
        keep = !p.same_class || gi >= gj;
        if (keep && p.nranks > 1) keep = task_owner(gi, gj, p.nranks) == p.rank;
        if (keep) {
          const int2 s34 = reinterpret_cast<const int2*>(p.ket.shell)[j];
          code = (s1 != s2) + (s34.x != s34.y) + (gi != gj);
          if (p.use_schwarz) {
            dn = fmax(D12, r1[s34.x]);
            dn = fmax(dn, r2[s34.x]);
            dn = fmax(dn, r1[s34.y]);
            dn = fmax(dn, r2[s34.y]);
            dn = fmax(dn, p.ket_dn[j]);
            const double Kj = p.ket.schwarz[j];
            // reference multiplies Dnorm * K(s1,s2) * K(s3,s4) with (s1,s2) the larger pair
            const double est = gi >= gj ? dn * Ki * Kj : dn * Kj * Ki;
            keep = !(est < p.fock_precision);
          }
        }
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, keep);
      if (ballot) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(p.count, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
          const unsigned pos = base + __popc(ballot & ((1u << lane) - 1u));
          // engine precision of this quartet (hartree-fock++.cc:1693-1695), as its logarithm:
          // the class kernels compare it with the primitive-pair screening sums
          const double lnp = dn != 0.0 ? log(p.fock_precision / dn) : p.ln_needed_engine_precision;
          if (pos < p.cap)
            p.tasks[pos] = make_int4(i, (int)(j | ((unsigned)code << 30)), __double2loint(lnp),
                                     __double2hiint(lnp));
        }
      }
    }
  }
}

__global__ void symmetrize_kernel(const double* __restrict__ F, double* __restrict__ G, int n) {
  // hartree-fock++.cc:1766: GG = 0.5 * (G + G^T)
  const long long n2 = (long long)n * n;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n2;
       g += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(g / n), j = (int)(g % n);
    G[g] = 0.5 * (F[g] + F[(long long)j * n + i]);
  }
}

struct DevArrays {   // device arrays of one call, freed on every exit path
  std::vector<void*> p;
  ~DevArrays() { for (void* x : p) cudaFree(x); }
  template <class T> int get(lb200_context* c, T** out, size_t count) {
    void* q = nullptr;
    const int r = check_cuda(c, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)), "cudaMalloc(grad)");
    if (!r) { p.push_back(q); *out = static_cast<T*>(q); }
    return r;
  }
};

}  // namespace

namespace lb200 {

int diag_schwarz(lb200_context* ctx, const lb200_pairs* P, std::vector<double>& out) {
  return ::diag_schwarz_impl(ctx, P, out);
}

int compute_prim_schwarz(lb200_context* ctx, const lb200_basis* bs1, const lb200_basis* bs2,
                         int npair, const int* s1, const int* s2, std::vector<double>& out) {
  lb200_basis pb1, pb2;
  primitive_basis(*bs1, pb1);
  primitive_basis(*bs2, pb2);
  std::vector<int> q1, q2;
  for (int i = 0; i < npair; ++i)
    for (int p1 = 0; p1 < bs1->nprim[s1[i]]; ++p1)
      for (int p2 = 0; p2 < bs2->nprim[s2[i]]; ++p2) {
        q1.push_back(bs1->off[s1[i]] + p1);
        q2.push_back(bs2->off[s2[i]] + p2);
      }
  lb200_pairs* P = nullptr;
  int rc = build_pairs(ctx, &pb1, &pb2, (int)q1.size(), q1.data(), q2.data(), kScreenOriginal,
                       std::numeric_limits<double>::lowest(), nullptr, nullptr, &P);
  if (rc) return rc;
  rc = diag_schwarz_impl(ctx, P, out);
  lb200_pairs_destroy(P);
  return rc;
}

}  // namespace lb200

extern "C" {

int lb200_fock_create(lb200_context* ctx, const lb200_basis* obs, long long npair, const int* s1,
                      const int* s2, lb200_fock** out) {
  if (!ctx || !obs || !out || npair < 0) return LB200_ERR_INVALID;
  // canonical pair indices s1(s1+1)/2+s2 are kept in 32 bits (PairBlock::gidx, task ownership)
  if (obs->nshell >= 65536)
    return set_error(ctx, LB200_ERR_INVALID, "Fock builder supports fewer than 65536 shells");
  cudaSetDevice(ctx->device);
  auto* f = new lb200_fock;
  f->ctx = ctx;
  f->obs = *obs;
  const int ns = obs->nshell;
  f->K.assign((size_t)ns * ns, 0.0);
  // group by class (first shell = higher AM) and by contraction degree: the quartets of one
  // launch advance through their primitive loops in lockstep (per warp / per CTA), so pairs
  // with very different numbers of primitive pairs must not share a block.  Bucket upper
  // bounds on nprim(a)*nprim(b); LB200_FOCK_BUCKETS="1,6" style override for experiments.
  std::vector<int> bounds = {1, 6};
  if (const char* e = std::getenv("LB200_FOCK_BUCKETS")) {
    bounds.clear();
    for (const char* q = e; *q;) {
      char* end = nullptr;
      const long v = std::strtol(q, &end, 10);
      if (end == q) break;
      if (v > 0) bounds.push_back((int)v);
      q = (*end == ',') ? end + 1 : end;
    }
  }
  auto bucket_of = [&](int npp) {
    int k = 0;
    while (k < (int)bounds.size() && npp > bounds[k]) ++k;
    return k;
  };
  std::map<std::array<int, 5>, std::pair<std::vector<int>, std::vector<int>>> groups;
  for (long long i = 0; i < npair; ++i) {
    int a = s1[i], b = s2[i];
    if (a < 0 || b < 0 || a >= ns || b >= ns) { delete f; return LB200_ERR_INVALID; }
    if (obs->l[a] < obs->l[b]) std::swap(a, b);
    auto& g = groups[{obs->l[a], obs->l[b], obs->pure[a], obs->pure[b],
                      bucket_of(obs->nprim[a] * obs->nprim[b])}];
    g.first.push_back(a);
    g.second.push_back(b);
  }
  const double max_engine_precision = std::numeric_limits<double>::epsilon() / 1e10;  // hf++:64
  const double ln_max_engine_precision = std::log(max_engine_precision);
  int rc = LB200_OK;
  for (auto& kv : groups) {
    auto& a = kv.second.first;
    auto& b = kv.second.second;
    const int n = (int)a.size();
    if (a.size() >= (size_t)1 << 30) { rc = LB200_ERR_INVALID; break; }   // task records keep 30 bits per pair index
    // shell-level Schwarz bound, no primitive screening (hartree-fock++.cc:1244-1247)
    lb200_pairs* all = nullptr;
    rc = build_pairs(ctx, obs, obs, n, a.data(), b.data(), kScreenOriginal,
                     std::numeric_limits<double>::lowest(), nullptr, nullptr, &all);
    std::vector<double> ksh;
    if (!rc) rc = diag_schwarz_impl(ctx, all, ksh);
    lb200_pairs_destroy(all);
    if (rc) break;
    // sort the class by Schwarz bound, descending (stable for reproducibility)
    std::vector<int> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return ksh[x] > ksh[y]; });
    FockClass fc;
    fc.la = kv.first[0]; fc.lb = kv.first[1]; fc.pa = kv.first[2]; fc.pb = kv.first[3];
    fc.bucket = kv.first[4];
    std::vector<int> as(n), bs_(n);
    fc.schwarz.resize(n);
    for (int i = 0; i < n; ++i) {
      as[i] = a[ord[i]]; bs_[i] = b[ord[i]]; fc.schwarz[i] = ksh[ord[i]];
      f->K[(size_t)as[i] * ns + bs_[i]] = f->K[(size_t)bs_[i] * ns + as[i]] = fc.schwarz[i];
    }
    // SchwarzInf shell-pair data (hartree-fock++.cc:1383-1431)
    rc = build_pairs(ctx, obs, obs, n, as.data(), bs_.data(), kScreenSchwarzInf,
                     ln_max_engine_precision, nullptr, fc.schwarz.data(), &fc.pairs);
    if (rc) break;
    if (cudaMalloc(&fc.d_dn, std::max(1, n) * sizeof(double)) != cudaSuccess) { rc = LB200_ERR_NOMEM; break; }
    f->classes.push_back(std::move(fc));
  }
  if (rc) { lb200_fock_destroy(f); return rc; }
  // kernel orientation wants bra key >= ket key: keep classes sorted by key
  std::stable_sort(f->classes.begin(), f->classes.end(), [](const FockClass& x, const FockClass& y) {
    return order_key(x.la, x.lb) < order_key(y.la, y.lb);
  });
  std::vector<int> ssz(ns);
  for (int s = 0; s < ns; ++s) ssz[s] = obs->size(s);
  // per-device attribute: set for this context's device whenever a builder is created
  cudaFuncSetAttribute(screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaMalloc(&f->d_shell2bf, ns * sizeof(int));
  cudaMalloc(&f->d_shellsize, ns * sizeof(int));
  cudaMalloc(&f->d_Dnorm, (size_t)ns * ns * 8);
  cudaMalloc(&f->d_scalar, 8);
  {
    int nsr = 4;
    if (const char* e = std::getenv("LB200_FOCK_STREAMS")) nsr = std::atoi(e);
    f->nstreams = std::max(1, std::min(nsr, (int)lb200_fock::kMaxStreams));
    cudaEventCreateWithFlags(&f->ev_ready, cudaEventDisableTiming);
    for (int k = 0; k < f->nstreams; ++k) {
      cudaMalloc(&f->d_count[k], 8);
      if (k > 0) {   // slot 0 is the context's own stream
        cudaStreamCreateWithFlags(&f->streams[k], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&f->ev_done[k], cudaEventDisableTiming);
      }
    }
  }
  cudaMemcpy(f->d_shell2bf, obs->shell2bf.data(), ns * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(f->d_shellsize, ssz.data(), ns * sizeof(int), cudaMemcpyHostToDevice);
  rc = check_cuda(ctx, cudaGetLastError(), "fock_create");
  if (rc) { lb200_fock_destroy(f); return rc; }
  *out = f;
  return LB200_OK;
}

int lb200_fock_destroy(lb200_fock* f) {
  if (!f) return LB200_OK;
  cudaSetDevice(f->ctx->device);
  for (auto& c : f->classes) { lb200_pairs_destroy(c.pairs); cudaFree(c.d_dn); }
  cudaFree(f->d_D); cudaFree(f->d_F); cudaFree(f->d_Dnorm); cudaFree(f->d_scalar);
  cudaFree(f->d_shell2bf); cudaFree(f->d_shellsize);
  for (int k = 0; k < lb200_fock::kMaxStreams; ++k) {
    cudaFree(f->d_tasks[k]); cudaFree(f->d_count[k]);
    if (k > 0 && f->streams[k]) cudaStreamDestroy(f->streams[k]);   // slot 0 is the context's stream
    if (f->ev_done[k]) cudaEventDestroy(f->ev_done[k]);
  }
  if (f->ev_ready) cudaEventDestroy(f->ev_ready);
  cudaFree(f->d_primcount);
  delete f;
  return LB200_OK;
}

int lb200_fock_set_profile(lb200_fock* f, int on) {
  if (!f) return LB200_ERR_INVALID;
  f->profile = on != 0;
  return LB200_OK;
}

long long lb200_fock_get_profile(const lb200_fock* f, double* rows, long long cap) {
  if (!f) return LB200_ERR_INVALID;
  const long long n = (long long)f->prof.size();
  if (!rows) return n;
  for (long long i = 0; i < n && i < cap; ++i) {
    const auto& e = f->prof[i];
    double* o = rows + 9 * i;
    for (int k = 0; k < 6; ++k) o[k] = e.c[k];
    o[6] = e.ms; o[7] = e.nq; o[8] = e.nprim;
  }
  return n;
}

int lb200_fock_task_owner(int bra_pair_index, int ket_pair_index, int nranks) {
  if (nranks < 1 || bra_pair_index < 0 || ket_pair_index < 0) return LB200_ERR_INVALID;
  return nranks == 1 ? 0 : task_owner(bra_pair_index, ket_pair_index, nranks);
}

int lb200_fock_schwarz(const lb200_fock* f, double* K) {
  if (!f || !K) return LB200_ERR_INVALID;
  std::memcpy(K, f->K.data(), f->K.size() * 8);
  return LB200_OK;
}

int lb200_fock_build(lb200_fock* f, const double* D, int D_on_device, double precision,
                     int use_schwarz, int rank, int nranks, double* G, int G_on_device,
                     double* stats) {
  if (!f || !D || !G || nranks < 1 || rank < 0 || rank >= nranks) return LB200_ERR_INVALID;
  lb200_context* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int n = f->obs.nbf, ns = f->obs.nshell;
  const size_t n2 = (size_t)n * n;
  int rc = LB200_OK;
  if (!f->d_F) rc = check_cuda(ctx, cudaMalloc(&f->d_F, n2 * 8), "cudaMalloc(F)");
  if (!rc && !f->d_D) rc = check_cuda(ctx, cudaMalloc(&f->d_D, n2 * 8), "cudaMalloc(D)");
  if (rc) return rc;
  const long long launches0 = ctx->launches;
  struct Events {   // destroyed on every exit path
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
  } evs;
  cudaEvent_t &ev0 = evs.e[0], &ev1 = evs.e[1], &pe0 = evs.e[2], &pe1 = evs.e[3];
  cudaEventCreate(&ev0);
  cudaEventCreate(&ev1);
  cudaEventRecord(ev0, st);
  cudaMemcpyAsync(f->d_D, D, n2 * 8, D_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(f->d_F, 0, n2 * 8, st);
  cudaMemsetAsync(f->d_scalar, 0, 8, st);
  shellblock_norm_kernel<<<std::min(1024, (ns * ns + 255) / 256), 256, 0, st>>>(
      f->d_D, n, ns, f->d_shell2bf, f->d_shellsize, f->d_Dnorm);
  absmax_kernel<<<std::min(1024, (ns * ns + 255) / 256), 256, 0, st>>>(f->d_Dnorm, (long long)ns * ns,
                                                                     f->d_scalar);
  ctx->launches += 2;
  for (auto& c : f->classes) {
    const int np = c.pairs->dev.npair;
    if (np == 0) continue;
    pair_dnorm_kernel<<<std::min(1024, (np + 255) / 256), 256, 0, st>>>(c.pairs->dev.shell, np, f->d_Dnorm,
                                                                      ns, c.d_dn);
    ++ctx->launches;
  }
  double Dmax = 0;
  cudaMemcpyAsync(&Dmax, f->d_scalar, 8, cudaMemcpyDeviceToHost, st);
  if ((rc = check_cuda(ctx, cudaStreamSynchronize(st), "fock: D norms"))) return rc;
  const double fock_precision = precision;
  const double needed_engine_precision = fock_precision / Dmax;  // hartree-fock++.cc:1588
  // LB200_FOCK_PROFILE=1: per class-pair device time (one sync per launch; diagnostics only)
  const bool env_profile = stats && std::getenv("LB200_FOCK_PROFILE");
  const bool profile = env_profile || f->profile;
  // streams in use: one (the context's) when per-launch accounting is wanted, else all of them
  const int NS = (profile || stats) ? 1 : f->nstreams;
  // task buffers
  const long long cap = 1ll << 24;
  if (f->task_cap < cap) {
    for (int k = 0; k < f->nstreams; ++k) {
      cudaFree(f->d_tasks[k]);
      f->d_tasks[k] = nullptr;
      if ((rc = check_cuda(ctx, cudaMalloc(&f->d_tasks[k], cap * sizeof(int4)), "cudaMalloc(tasks)")))
        return rc;
    }
    f->task_cap = cap;
  }
  f->streams[0] = st;
  if (NS > 1) {   // side streams start after the density norms are on the device
    cudaEventRecord(f->ev_ready, st);
    for (int k = 1; k < NS; ++k) cudaStreamWaitEvent(f->streams[k], f->ev_ready, 0);
  }
  int next_stream = 0;
  const double thr_num = fock_precision / Dmax * (1.0 - 1e-12);
  double nquartets = 0, ncand = 0;
  std::vector<unsigned> jmax;
  using Prof = lb200_fock::ProfRow;
  std::vector<Prof>& prof = f->prof;
  prof.clear();
  if (profile && !f->d_primcount) cudaMalloc(&f->d_primcount, 8 * kPrimCounters);
  if (profile) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); }
  const size_t ncls = f->classes.size();
  for (size_t X = 0; X < ncls && !rc; ++X)
    for (size_t Y = 0; Y <= X && !rc; ++Y) {
      const FockClass& B = f->classes[X];   // key(B) >= key(Y): kernel orientation
      const FockClass& Kt = f->classes[Y];
      const int nb = B.pairs->dev.npair, nk = Kt.pairs->dev.npair;
      if (nb == 0 || nk == 0) continue;
      if (!class_supported(B.la, B.lb, Kt.la, Kt.lb))
        return set_error(ctx, LB200_ERR_LMAX, "no kernel built for a class of this basis");
      // candidate prefix per bra row from the sorted Schwarz bounds:
      // K_i * K_j * Dmax >= precision is necessary for survival
      jmax.assign(nb, (unsigned)nk);
      if (use_schwarz) {
        // both bound lists are sorted descending, so the prefix length only shrinks from one bra
        // row to the next: one sweep over (rows + kets) instead of a binary search per row (the
        // host loop must stay ahead of eight GPUs' worth of class kernels)
        int lo = nk;   // first j with schwarz[j] < thr
        for (int i = 0; i < nb; ++i) {
          const double thr = thr_num / B.schwarz[i];   // same expression as screen_kernel
          while (lo > 0 && Kt.schwarz[lo - 1] < thr) --lo;
          jmax[i] = (unsigned)lo;
        }
      }
      if (nranks > 1) {   // rows of other ranks: nothing to enumerate
        const std::vector<int>& sh = B.pairs->shell;
        for (int i = 0; i < nb; ++i) {
          const long long hi = std::max(sh[2 * i], sh[2 * i + 1]), lo = std::min(sh[2 * i], sh[2 * i + 1]);
          if (task_owner((int)(hi * (hi + 1) / 2 + lo), 0, nranks) != rank) jmax[i] = 0;
        }
      }
      int row = 0;
      while (row < nb && !rc) {
        long long sum = 0;
        int r1 = row;
        while (r1 < nb && (r1 == row || sum + jmax[r1] <= cap)) sum += jmax[r1++];
        if (sum > cap) return set_error(ctx, LB200_ERR_NOMEM, "task buffer too small for one row");
        if (sum > 0) {
          ncand += (double)sum;
          const int sidx = next_stream;
          next_stream = (next_stream + 1) % NS;
          cudaStream_t st = f->streams[sidx];   // shadows the context's stream inside this launch pair
          int4* const d_tasks = f->d_tasks[sidx];
          unsigned* const d_count = f->d_count[sidx];
          ScreenParams sp{};
          sp.bra = B.pairs->dev; sp.ket = Kt.pairs->dev;
          sp.same_class = (X == Y);
          sp.row0 = row; sp.nrow = r1 - row;
          sp.nket = nk; sp.thr_num = thr_num;
          sp.Dnorm = f->d_Dnorm; sp.nshell = ns;
          sp.bra_dn = B.d_dn; sp.ket_dn = Kt.d_dn;
          sp.fock_precision = fock_precision; sp.use_schwarz = use_schwarz;
          sp.ln_needed_engine_precision = std::log(needed_engine_precision);
          sp.rank = rank; sp.nranks = nranks;
          sp.tasks = d_tasks; sp.count = d_count; sp.cap = (unsigned)cap;
          cudaMemsetAsync(d_count, 0, 8, st);
          const int threads = 128;
          const size_t row_bytes = 2 * (size_t)ns * sizeof(double);
          sp.stage_rows = row_bytes <= (size_t)96 * 1024;
          const size_t smem = sp.stage_rows ? row_bytes : 0;
          const int grid = std::min(ctx->num_sms * 16, sp.nrow);
          screen_kernel<<<grid, threads, smem, st>>>(sp);
          ++ctx->launches;
          EriParams p{};
          p.bra = B.pairs->dev; p.ket = Kt.pairs->dev;
          p.tasks = nullptr; p.ftasks = d_tasks; p.ntasks_dev = d_count; p.ntasks = 0; p.swap_tasks = 0;
          p.work_counter = d_count + 1;
          // uncontracted x uncontracted bucket: the pipelined kernel (LB200_NO_PRIM_KERNEL=1: A/B)
          static const bool no_prim = std::getenv("LB200_NO_PRIM_KERNEL") != nullptr;
          p.uncontracted = (!no_prim && p.bra.max_nprim <= 1 && p.ket.max_nprim <= 1) ? 1 : 0;
          p.boys = ctx->d_boys;
          p.screening = kScreenSchwarzInf;
          p.D = f->d_D; p.F = f->d_F; p.nbf = n; p.Dnorm = f->d_Dnorm; p.nshell = ns;
          p.sph_rowptr = ctx->d_sph_rowptr; p.sph_col = ctx->d_sph_col;
          p.sph_val = ctx->d_sph_val; p.sph_base = ctx->d_sph_base;
          p.fock_precision = fock_precision;
          p.needed_engine_precision = needed_engine_precision;
          p.ln_needed_engine_precision = std::log(needed_engine_precision);
          p.prim_counter = nullptr;
          if (profile) {
            cudaMemsetAsync(f->d_primcount, 0, 8 * kPrimCounters, st);
            p.prim_counter = f->d_primcount;
            cudaEventRecord(pe0, st);
          }
          // LB200_FOCK_SCREEN_ONLY=1 (diagnostics): enumerate and screen, skip the class kernels
          static const bool screen_only = std::getenv("LB200_FOCK_SCREEN_ONLY") != nullptr;
          if (!screen_only)
          rc = check_cuda(ctx, launch_eri(B.la, B.lb, Kt.la, Kt.lb, p, ctx->d_rows, kModeFock,
                                          ctx->num_sms, st), "launch fock kernel");
          ++ctx->launches;
          if (profile) {
            cudaEventRecord(pe1, st);
            cudaEventSynchronize(pe1);
            float ms = 0;
            cudaEventElapsedTime(&ms, pe0, pe1);
            unsigned c = 0;
            unsigned long long np = 0, npc[kPrimCounters];
            cudaMemcpy(&c, d_count, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(npc, f->d_primcount, 8 * kPrimCounters, cudaMemcpyDeviceToHost);
            for (int k = 0; k < kPrimCounters; ++k) np += npc[k];
            bool found = false;
            for (auto& e : prof)
              if (e.c[0] == B.la && e.c[1] == B.lb && e.c[2] == Kt.la && e.c[3] == Kt.lb &&
                  e.c[4] == B.bucket && e.c[5] == Kt.bucket) {
                e.ms += ms; e.nq += c; e.nprim += (double)np; found = true;
              }
            if (!found)
              prof.push_back(Prof{{B.la, B.lb, Kt.la, Kt.lb, B.bucket, Kt.bucket}, ms, (double)c, (double)np});
          }
          if (stats) {  // optional accounting costs a sync per chunk
            unsigned c = 0;
            cudaMemcpyAsync(&c, d_count, 4, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            nquartets += c;
          }
        }
        row = r1;
      }
    }
  if (NS > 1) {   // the context's stream continues after every side stream has drained
    for (int k = 1; k < NS; ++k) {
      cudaEventRecord(f->ev_done[k], f->streams[k]);
      cudaStreamWaitEvent(st, f->ev_done[k], 0);
    }
  }
  if (rc) return rc;
  if (profile) {
    std::sort(prof.begin(), prof.end(), [](const Prof& a, const Prof& b) { return a.ms > b.ms; });
  }
  if (env_profile) {
    double tot = 0;
    for (auto& e : prof) tot += e.ms;
    std::fprintf(stderr, "lb200 fock profile: %zu class pairs, %.2f ms in class kernels\n", prof.size(), tot);
    for (auto& e : prof)
      std::fprintf(stderr, "  (%d%d|%d%d) buckets %d,%d %10.3f ms %5.1f%% %12.0f quartets %8.2f ns/quartet\n",
                   e.c[0], e.c[1], e.c[2], e.c[3], e.c[4], e.c[5], e.ms, 100 * e.ms / tot, e.nq,
                   e.nq > 0 ? 1e6 * e.ms / e.nq : 0.0);
  }
  double* d_G = nullptr;
  if (G_on_device) {
    symmetrize_kernel<<<std::min(4096ll, (long long)(n2 + 255) / 256), 256, 0, st>>>(f->d_F, G, n);
  } else {
    d_G = f->d_D;  // D no longer needed
    symmetrize_kernel<<<std::min(4096ll, (long long)(n2 + 255) / 256), 256, 0, st>>>(f->d_F, d_G, n);
    cudaMemcpyAsync(G, d_G, n2 * 8, cudaMemcpyDeviceToHost, st);
  }
  ++ctx->launches;
  cudaEventRecord(ev1, st);
  rc = check_cuda(ctx, cudaStreamSynchronize(st), "fock build");
  float ms = 0;
  cudaEventElapsedTime(&ms, ev0, ev1);
  if (stats) {
    stats[0] = nquartets;
    stats[1] = (double)(ctx->launches - launches0);
    stats[2] = ms;
    stats[3] = ncand;
  }
  return rc;
}


// Two-body forces: F2(atom, xyz) = sum_ij G1[3 atom + xyz]_ij D_ij with G1 = compute_2body_fock_deriv<1>
// (hartree-fock++.cc:1775-2055, consumed at :648-656), evaluated without forming the 3 * natoms matrices:
// the same quartets as the Fock build (same pair blocks, same screening kernel, same rank ownership), their
// derivative shell sets from deriv.cu, contracted with the two-particle density on the fly.  Every quartet
// runs at the tightest engine precision of the build (fock_precision / max|D|); the reference loosens it
// per quartet (:1975-1977), which only drops primitives below that bound.
int lb200_fock_grad(lb200_fock* f, const double* D, int D_on_device, double precision, int use_schwarz,
                    int rank, int nranks, int natoms, const int* shell2atom, double* grad, double* stats) {
  if (!f || !D || !grad || !shell2atom || natoms < 1 || nranks < 1 || rank < 0 || rank >= nranks)
    return LB200_ERR_INVALID;
  lb200_context* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const lb200_basis& obs = f->obs;
  const int n = obs.nbf, ns = obs.nshell;
  const size_t n2 = (size_t)n * n;
  for (int s = 0; s < ns; ++s)
    if (shell2atom[s] < 0 || shell2atom[s] >= natoms)
      return set_error(ctx, LB200_ERR_INVALID, "shell2atom entry out of range");
  int rc = LB200_OK;
  if (!f->d_D) rc = check_cuda(ctx, cudaMalloc(&f->d_D, n2 * 8), "cudaMalloc(D)");
  if (rc) return rc;
  // Cartesian function offsets
  std::vector<int> s2c(ns), c2s, hl(obs.l), hp(obs.pure);
  int nbfc = 0;
  for (int s = 0; s < ns; ++s) {
    s2c[s] = nbfc;
    for (int k = 0; k < nc(obs.l[s]); ++k) c2s.push_back(s);
    nbfc += nc(obs.l[s]);
  }
  DevArrays dev;   // freed on every exit path
  int *d_l = nullptr, *d_pure = nullptr, *d_s2c = nullptr, *d_c2s = nullptr, *d_s2a = nullptr;
  double *d_Dc = nullptr, *d_grad = nullptr;
  int2* d_tasks2 = nullptr;
  const long long cap = 1ll << 22;   // tasks per screening chunk (the derivative sets are the memory hog)
  if ((rc = dev.get(ctx, &d_l, ns)) || (rc = dev.get(ctx, &d_pure, ns)) || (rc = dev.get(ctx, &d_s2c, ns)) ||
      (rc = dev.get(ctx, &d_c2s, nbfc)) || (rc = dev.get(ctx, &d_s2a, ns)) ||
      (rc = dev.get(ctx, &d_Dc, (size_t)nbfc * nbfc)) || (rc = dev.get(ctx, &d_grad, 3 * (size_t)natoms)) ||
      (rc = dev.get(ctx, &d_tasks2, (size_t)cap)))
    return rc;
  if (f->task_cap < cap) {
    for (int k = 0; k < f->nstreams; ++k) {
      cudaFree(f->d_tasks[k]);
      f->d_tasks[k] = nullptr;
      if ((rc = check_cuda(ctx, cudaMalloc(&f->d_tasks[k], cap * sizeof(int4)), "cudaMalloc(tasks)"))) return rc;
    }
    f->task_cap = cap;
  }
  struct Events {
    cudaEvent_t e[2] = {nullptr, nullptr};
    ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
  } evs;
  cudaEventCreate(&evs.e[0]);
  cudaEventCreate(&evs.e[1]);
  cudaEventRecord(evs.e[0], st);
  cudaMemcpyAsync(d_l, hl.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_pure, hp.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_s2c, s2c.data(), ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_c2s, c2s.data(), (size_t)nbfc * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_s2a, shell2atom, ns * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(f->d_D, D, n2 * 8, D_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(d_grad, 0, 3 * (size_t)natoms * 8, st);
  cudaMemsetAsync(f->d_scalar, 0, 8, st);
  shellblock_norm_kernel<<<std::min(1024, (ns * ns + 255) / 256), 256, 0, st>>>(
      f->d_D, n, ns, f->d_shell2bf, f->d_shellsize, f->d_Dnorm);
  absmax_kernel<<<std::min(1024, (ns * ns + 255) / 256), 256, 0, st>>>(f->d_Dnorm, (long long)ns * ns, f->d_scalar);
  ctx->launches += 2;
  for (auto& c : f->classes) {
    const int np = c.pairs->dev.npair;
    if (np == 0) continue;
    pair_dnorm_kernel<<<std::min(1024, (np + 255) / 256), 256, 0, st>>>(c.pairs->dev.shell, np, f->d_Dnorm, ns, c.d_dn);
    ++ctx->launches;
  }
  rc = check_cuda(ctx, launch_cartesianize_density(ctx, f->d_D, n, d_Dc, nbfc, ns, d_l, d_pure, f->d_shell2bf, d_s2c,
                                                   d_c2s, st), "cartesianize density");
  ++ctx->launches;
  double Dmax = 0;
  cudaMemcpyAsync(&Dmax, f->d_scalar, 8, cudaMemcpyDeviceToHost, st);
  if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(st), "gradient: D norms");
  if (rc) return rc;
  const double fock_precision = precision;
  const double needed_engine_precision = fock_precision / Dmax;   // hartree-fock++.cc:1797
  const double thr_num = fock_precision / Dmax * (1.0 - 1e-12);
  const long long launches0 = ctx->launches;
  double nquartets = 0;
  std::vector<unsigned> jmax;
  int4* const d_ftasks = f->d_tasks[0];
  unsigned* const d_count = f->d_count[0];
  const size_t ncls = f->classes.size();
  for (size_t X = 0; X < ncls && !rc; ++X)
    for (size_t Y = 0; Y <= X && !rc; ++Y) {
      const FockClass& B = f->classes[X];
      const FockClass& Kt = f->classes[Y];
      const int nb = B.pairs->dev.npair, nk = Kt.pairs->dev.npair;
      if (nb == 0 || nk == 0) continue;
      DerivSets ds;
      if ((rc = deriv_plan(ctx, B.pairs, Kt.pairs, ds))) break;
      jmax.assign(nb, (unsigned)nk);
      if (use_schwarz) {
        int lo = nk;
        for (int i = 0; i < nb; ++i) {
          const double thr = thr_num / B.schwarz[i];
          while (lo > 0 && Kt.schwarz[lo - 1] < thr) --lo;
          jmax[i] = (unsigned)lo;
        }
      }
      // tasks of one derivative chunk: bounded by 1 GiB of shifted shell sets
      const long long sub = std::max(1ll, std::min(cap, (1ll << 30) / (ds.doubles_per_task * 8)));
      int row = 0;
      while (row < nb && !rc) {
        long long sum = 0;
        int r1 = row;
        while (r1 < nb && (r1 == row || sum + jmax[r1] <= cap)) sum += jmax[r1++];
        if (sum > cap) return set_error(ctx, LB200_ERR_NOMEM, "task buffer too small for one row");
        if (sum > 0) {
          ScreenParams sp{};
          sp.bra = B.pairs->dev; sp.ket = Kt.pairs->dev;
          sp.same_class = (X == Y);
          sp.row0 = row; sp.nrow = r1 - row;
          sp.nket = nk; sp.thr_num = thr_num;
          sp.Dnorm = f->d_Dnorm; sp.nshell = ns;
          sp.bra_dn = B.d_dn; sp.ket_dn = Kt.d_dn;
          sp.fock_precision = fock_precision; sp.use_schwarz = use_schwarz;
          sp.ln_needed_engine_precision = std::log(needed_engine_precision);
          sp.rank = rank; sp.nranks = nranks;
          sp.tasks = d_ftasks; sp.count = d_count; sp.cap = (unsigned)cap;
          cudaMemsetAsync(d_count, 0, 8, st);
          const size_t row_bytes = 2 * (size_t)ns * sizeof(double);
          sp.stage_rows = row_bytes <= (size_t)96 * 1024;
          screen_kernel<<<std::min(ctx->num_sms * 16, sp.nrow), 128, sp.stage_rows ? row_bytes : 0, st>>>(sp);
          ++ctx->launches;
          rc = check_cuda(ctx, launch_unpack_tasks(d_ftasks, d_count, d_tasks2, st, cap), "unpack tasks");
          ++ctx->launches;
          unsigned cnt = 0;
          cudaMemcpyAsync(&cnt, d_count, 4, cudaMemcpyDeviceToHost, st);
          if (!rc) rc = check_cuda(ctx, cudaStreamSynchronize(st), "gradient: screening");
          if (rc) break;
          nquartets += cnt;
          for (long long t0 = 0; t0 < (long long)cnt && !rc; t0 += sub) {
            const long long nt = std::min(sub, (long long)cnt - t0);
            double* d_sets = nullptr;
            if ((rc = ctx_scratch(ctx, 5, (size_t)nt * ds.doubles_per_task * 8, reinterpret_cast<void**>(&d_sets))))
              break;
            rc = deriv_eval(ctx, B.pairs, Kt.pairs, nt, d_tasks2 + t0, kScreenSchwarzInf, needed_engine_precision,
                            d_sets, ds);
            if (rc) break;
            DerivGradParams gp;
            gp.tasks = d_tasks2 + t0; gp.ftasks = d_ftasks + t0;
            gp.bra_shell = B.pairs->dev.shell; gp.ket_shell = Kt.pairs->dev.shell;
            gp.shell2cbf = d_s2c; gp.shell2atom = d_s2a;
            gp.Dc = d_Dc; gp.nbfc = nbfc; gp.grad = d_grad;
            rc = check_cuda(ctx, launch_deriv_grad(ds, nt, gp, st), "gradient contraction");
            ++ctx->launches;
          }
        }
        row = r1;
      }
    }
  if (rc) return rc;
  cudaMemcpyAsync(grad, d_grad, 3 * (size_t)natoms * 8, cudaMemcpyDeviceToHost, st);
  cudaEventRecord(evs.e[1], st);
  rc = check_cuda(ctx, cudaStreamSynchronize(st), "fock gradient");
  float ms = 0;
  cudaEventElapsedTime(&ms, evs.e[0], evs.e[1]);
  if (stats) {
    stats[0] = nquartets;
    stats[1] = (double)(ctx->launches - launches0);
    stats[2] = ms;
  }
  return rc;
}

}  // extern "C"
