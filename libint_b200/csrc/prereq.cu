// lb200_eri_prereq_batch: contracted Cartesian shell sets from caller-made per-primitive
// prerequisites -- the batched, GPU form of the reference's innermost plugin call
//   libint2_build_eri[la][lb][lc][ld](const Libint_t* inteval /* [contrdepth] */)
// (src/bin/libint/iface.cc:114-185; consumed at include/libint2/engine.impl.h:1898-1899).
// The class kernels are the production ones (eri_rowreg.cuh) in kModePrereq: VRR, contraction
// and HRR run on the device, only the prerequisite set-up stays with the caller, exactly as it
// does behind the reference's Libint_t boundary.  iface.cc builds the Libint_t ABI on top.
#include <cstring>

#include "internal.h"

using namespace lb200;

namespace {

int grow(lb200_context* ctx, int slot, size_t bytes, void** out) {
  if (ctx->scratch_bytes[slot] < bytes) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(ctx->d_scratch[slot]);
    ctx->d_scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    const size_t want = bytes < 4096 ? 4096 : bytes;
    int rc = check_cuda(ctx, cudaMalloc(&ctx->d_scratch[slot], want), "cudaMalloc(scratch)");
    if (rc) return rc;
    ctx->scratch_bytes[slot] = want;
  }
  *out = ctx->d_scratch[slot];
  return LB200_OK;
}

}  // namespace

extern "C" int lb200_eri_prereq_batch(lb200_context* ctx, int la, int lb, int lc, int ld,
                                      long long ntasks, const int* prim_off, const double* recs,
                                      const double* geom, double* out) {
  if (!ctx || ntasks < 0 || (ntasks > 0 && (!prim_off || !recs || !geom || !out)))
    return LB200_ERR_INVALID;
  if (la < lb || lc < ld || lb < 0 || ld < 0)
    return set_error(ctx, LB200_ERR_INVALID, "class must have la >= lb and lc >= ld");
  if (ntasks == 0) return LB200_OK;
  if (ntasks > 0x7fffffffll) return set_error(ctx, LB200_ERR_INVALID, "too many tasks");
  static_assert(sizeof(PrereqRec) == LB200_PREREQ_DOUBLES * sizeof(double), "PrereqRec layout");
  cudaSetDevice(ctx->device);
  const bool swap = order_key(la, lb) < order_key(lc, ld);
  const int ka = swap ? lc : la, kb = swap ? ld : lb, kc = swap ? la : lc, kd = swap ? lb : ld;
  if (!class_supported(ka, kb, kc, kd))
    return set_error(ctx, LB200_ERR_LMAX, "no kernel built for this angular-momentum class");
  if (la + lb + lc + ld > kBoysTableMmax)
    return set_error(ctx, LB200_ERR_LMAX, "total angular momentum exceeds the prerequisite record");
  const long long nrec = prim_off[ntasks];
  const long long blk = (long long)nc(la) * nc(lb) * nc(lc) * nc(ld);
  // one device block: offsets | geometry | records ; results in a second one
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_off = 0, o_geom = al((ntasks + 1) * sizeof(int));
  const size_t o_rec = al(o_geom + 6 * ntasks * sizeof(double));
  const size_t total = o_rec + nrec * sizeof(PrereqRec);
  char* d_in = nullptr;
  double* d_out = nullptr;
  int rc = grow(ctx, 1, total, reinterpret_cast<void**>(&d_in));
  if (!rc) rc = grow(ctx, 2, blk * ntasks * sizeof(double), reinterpret_cast<void**>(&d_out));
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  cudaMemcpyAsync(d_in + o_off, prim_off, (ntasks + 1) * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_in + o_geom, geom, 6 * ntasks * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_in + o_rec, recs, nrec * sizeof(PrereqRec), cudaMemcpyHostToDevice, st);
  EriParams p{};
  p.prereq_off = reinterpret_cast<const int*>(d_in + o_off);
  p.prereq_geom = reinterpret_cast<const double*>(d_in + o_geom);
  p.prereq = reinterpret_cast<const PrereqRec*>(d_in + o_rec);
  p.ntasks = (unsigned)ntasks;
  p.swap_tasks = swap ? 1 : 0;
  p.uncontracted = 0;
  p.boys = ctx->d_boys;
  p.out = d_out;
  p.out_stride = blk;
  p.transpose_out = swap ? 1 : 0;
  cudaError_t e = launch_eri(ka, kb, kc, kd, p, ctx->d_rows, kModePrereq, ctx->num_sms, st);
  ++ctx->launches;
  if ((rc = check_cuda(ctx, e, "launch eri prereq kernel"))) return rc;
  cudaMemcpyAsync(out, d_out, blk * ntasks * sizeof(double), cudaMemcpyDeviceToHost, st);
  return check_cuda(ctx, cudaStreamSynchronize(st), "eri_prereq_batch");
}
