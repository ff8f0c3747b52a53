// Generic (any purity pattern) Cartesian -> real-solid-harmonic transform of a batch of
// shell sets, and a per-block |max| reduction.  Reference semantics: the four index
// transforms of include/libint2/solidharmonics.h:281-463 applied by Engine::compute2
// (engine.impl.h:1965-1985); here each output element gathers its sparse 4-index product
// directly.  Used by the store path with pure_out and by the Schwarz set-up; the Fock
// kernels fuse their own compile-time transform.
#include "internal.h"

namespace lb200 {

namespace {

struct TformParams {
  const double* in;
  double* out;
  long long ntasks;
  int l[4], pure[4], nin[4], nout[4];
  const int* rowptr;  // [(kMaxShellL+1)][2*kMaxShellL+2]
  const int* col;
  const double* val;
  const int* base;
};

__global__ void pure_transform_kernel(const TformParams p) {
  const long long nout_blk = (long long)p.nout[0] * p.nout[1] * p.nout[2] * p.nout[3];
  const long long nin_blk = (long long)p.nin[0] * p.nin[1] * p.nin[2] * p.nin[3];
  const long long total = p.ntasks * nout_blk;
  constexpr int RP = 2 * kMaxShellL + 2;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long t = g / nout_blk;
    int r = (int)(g - t * nout_blk);
    int idx[4];
    idx[3] = r % p.nout[3]; r /= p.nout[3];
    idx[2] = r % p.nout[2]; r /= p.nout[2];
    idx[1] = r % p.nout[1]; r /= p.nout[1];
    idx[0] = r;
    // sparse rows of each index: pure -> CSR row of (l, m); cart -> identity
    int k0[4], k1[4];
    for (int x = 0; x < 4; ++x) {
      if (p.pure[x]) {
        k0[x] = p.rowptr[p.l[x] * RP + idx[x]];
        k1[x] = p.rowptr[p.l[x] * RP + idx[x] + 1];
      } else {
        k0[x] = 0;
        k1[x] = 1;
      }
    }
    const double* in = p.in + t * nin_blk;
    double acc = 0.0;
    for (int a = k0[0]; a < k1[0]; ++a) {
      const int ca = p.pure[0] ? p.col[p.base[p.l[0]] + a] : idx[0];
      const double va = p.pure[0] ? p.val[p.base[p.l[0]] + a] : 1.0;
      for (int b = k0[1]; b < k1[1]; ++b) {
        const int cb = p.pure[1] ? p.col[p.base[p.l[1]] + b] : idx[1];
        const double vb = p.pure[1] ? p.val[p.base[p.l[1]] + b] : 1.0;
        for (int c = k0[2]; c < k1[2]; ++c) {
          const int cc = p.pure[2] ? p.col[p.base[p.l[2]] + c] : idx[2];
          const double vc = p.pure[2] ? p.val[p.base[p.l[2]] + c] : 1.0;
          for (int d = k0[3]; d < k1[3]; ++d) {
            const int cd = p.pure[3] ? p.col[p.base[p.l[3]] + d] : idx[3];
            const double vd = p.pure[3] ? p.val[p.base[p.l[3]] + d] : 1.0;
            acc += va * vb * vc * vd *
                   in[((ca * p.nin[1] + cb) * p.nin[2] + cc) * p.nin[3] + cd];
          }
        }
      }
    }
    p.out[g] = acc;
  }
}

__global__ void block_absmax_kernel(const double* __restrict__ in, double* __restrict__ out,
                                    long long ntasks, long long n) {
  // one warp per block of n doubles
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= ntasks) return;
  const double* b = in + w * n;
  double m = 0.0;
  for (long long i = lane; i < n; i += 32) m = fmax(m, fabs(b[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[w] = m;
}

}  // namespace

cudaError_t launch_pure_transform(const lb200_context* ctx, const double* in, double* out,
                                  long long ntasks, const int l[4], const int pure[4],
                                  cudaStream_t stream) {
  TformParams p;
  p.in = in;
  p.out = out;
  p.ntasks = ntasks;
  long long nout = 1;
  for (int x = 0; x < 4; ++x) {
    p.l[x] = l[x];
    p.pure[x] = pure[x];
    p.nin[x] = nc(l[x]);
    p.nout[x] = pure[x] ? npure(l[x]) : nc(l[x]);
    nout *= p.nout[x];
  }
  p.rowptr = ctx->d_sph_rowptr;
  p.col = ctx->d_sph_col;
  p.val = ctx->d_sph_val;
  p.base = ctx->d_sph_base;
  const long long total = ntasks * nout;
  if (total == 0) return cudaSuccess;
  long long grid = (total + 255) / 256;
  const long long cap = (long long)ctx->num_sms * 16;
  if (grid > cap) grid = cap;
  pure_transform_kernel<<<(unsigned)grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_block_absmax(const double* in, double* out, long long ntasks, long long n,
                                cudaStream_t stream) {
  if (ntasks == 0) return cudaSuccess;
  const long long threads = ntasks * 32;
  block_absmax_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(in, out, ntasks, n);
  return cudaGetLastError();
}

}  // namespace lb200

// ---------------------------------------------------------------------------------------
// FP64 FMA throughput probe: 8 independent FMA chains per thread, all in registers.
// ---------------------------------------------------------------------------------------
namespace lb200 {
namespace {
__global__ void __launch_bounds__(256) fma_probe_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}
}  // namespace
}  // namespace lb200

extern "C" int lb200_fp64_peak_probe(lb200_context* ctx, int iters, double* tflops, double* ms_out) {
  if (!ctx || iters < 1 || !tflops) return LB200_ERR_INVALID;
  cudaSetDevice(ctx->device);
  double* d = nullptr;
  int rc = lb200::check_cuda(ctx, cudaMalloc(&d, 8), "cudaMalloc");
  if (rc) return rc;
  const int grid = ctx->num_sms * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  lb200::fma_probe_kernel<<<grid, block, 0, ctx->stream>>>(d, iters / 8 + 1, 0.999999, 1e-9);  // warm-up
  cudaEventRecord(e0, ctx->stream);
  lb200::fma_probe_kernel<<<grid, block, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
  cudaEventRecord(e1, ctx->stream);
  ctx->launches += 2;
  rc = lb200::check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "fp64 probe");
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double flops = 2.0 * 8 * 16 * (double)iters * grid * block;
  *tflops = ms > 0 ? flops / (ms * 1e-3) / 1e12 : 0.0;
  if (ms_out) *ms_out = ms;
  return rc;
}
