// Multi-GPU entries of the C ABI (gpl_dist_*): one process per GPU, the NCCL communicator lives
// INSIDE the library (ncclBroadcast / ncclAllReduce on its own stream), so an R host -- or any FFI --
// drives the sharded full GP exactly like the single-GPU entries.  The schedule (who does what, on
// which stream, in which order) is dist_sched.hpp; this file is its CUDA + NCCL back end plus the
// Newton / EP / gradient loops on top:
//   gpl_dist_fit_full     fit_laplace / fit_ep .method_full (R/gp_fit.R:91-99,150-161) with
//                         laplace.method_full (R/laplace.R:13-38,274-297) / ep.method_full
//                         (R/ep.R:16-49,286-330) over the sharded factor
//   gpl_dist_energy_grad  gp_energy (R/gp.R:138-147) + the analytic gradient, lik_gaussian
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a torch process that is torch's own copy,
// otherwise the system library), so single-GPU users of the library do not need NCCL at all.
#include "dist_sched.hpp"
#include "fit.cuh"

#include <dlfcn.h>
#include <limits.h>
#include <math.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <vector>

namespace gpl {

namespace {

// ---- NCCL, bound lazily ---------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;   // optional
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  char load_error[512] = {0};
};

NcclApi* nccl_api_load(NcclApi& api);

// loaded once per process (std::call_once: several host threads may create communicators)
NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] { nccl_api_load(api); });
  if (!api.handle) set_error("%s", api.load_error);       // the message is thread-local: repeat it per caller
  return api.handle ? &api : nullptr;
}

NcclApi* nccl_api_load(NcclApi& api) {
  const char* names[] = {getenv("GPL_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    snprintf(api.load_error, sizeof(api.load_error), "cannot load NCCL (libnccl.so.2): %s", dlerror());
    return nullptr;
  }
  bool ok = true;
  auto sym = [&](const char* s) {
    void* p = dlsym(api.handle, s);
    if (!p) ok = false;
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  api.CommInitRankConfig = (decltype(api.CommInitRankConfig))dlsym(api.handle, "ncclCommInitRankConfig");
  if (!ok) {
    snprintf(api.load_error, sizeof(api.load_error), "libnccl.so.2 lacks an expected symbol");
    dlclose(api.handle);
    api.handle = nullptr;
    return nullptr;
  }
  return &api;
}

#define GPL_NCCL(d, call)                                                                      \
  do {                                                                                         \
    ncclResult_t r__ = (call);                                                                 \
    if (r__ != ncclSuccess) {                                                                  \
      set_error("%s:%d NCCL error: %s", __FILE__, __LINE__, (d)->api->GetErrorString(r__));    \
      return GPL_ERR_NCCL;                                                                     \
    }                                                                                          \
  } while (0)

}  // namespace

// one per (process, GPU): communicator, its stream and the large buffers kept between calls
struct Dist {
  Ctx* ctx = nullptr;
  NcclApi* api = nullptr;
  ncclComm_t comm = nullptr;       // capped at a few CTAs: broadcasts that overlap the trailing updates
  ncclComm_t comm_wide = nullptr;  // NCCL's default CTA count: the latency-bound tail, where the SMs idle anyway
  int rank = 0, world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaStream_t aux_stream = nullptr;    // second high-priority stream of the panel chain
  cudaEvent_t ev[dist::EV_COUNT] = {};
  // cached device buffers (cudaMalloc, not the stream-ordered pool: they are tens of GB)
  double* A = nullptr; size_t capA = 0;
  double* X = nullptr; size_t capX = 0;
  double* panel[dist::MAX_WINDOW] = {}; size_t capP[dist::MAX_WINDOW] = {};
  double* tmp = nullptr; size_t capT = 0;
  double* dinv = nullptr; size_t capD = 0;
  int* d_info = nullptr;        // [0] failing pivot (min), [1] scratch
  double* d_acc = nullptr;      // device scalars: [0] logdet
  double nccl_bytes = 0.0;      // payload bytes this rank has put through collectives
};

namespace {

int ensure(double** p, size_t* cap, size_t doubles) {
  if (*cap >= doubles && *p) return GPL_OK;
  if (*p) GPL_CUDA(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  cudaError_t e = cudaMalloc(p, sizeof(double) * (doubles ? doubles : 2));
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    set_error("multi-GPU workspace: device allocation of %.2f GB failed", 8.0 * doubles / 1e9);
    return GPL_ERR_NOMEM;
  }
  GPL_CUDA(e);
  *cap = doubles;
  return GPL_OK;
}

__global__ void scale_sym_block_kernel(int64_t m, int w, double* __restrict__ A, int64_t lda,
                                       const double* __restrict__ sr, const double* __restrict__ sc,
                                       double jitter) {
  // A (m x w; its first w rows are the diagonal block) <- diag(sr) (A + jitter I) diag(sc) + I
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double si = sr[i];
  for (int j = blockIdx.y; j < w; j += gridDim.y) {
    double v = A[(int64_t)j * lda + i];
    if (i == j) v += jitter;
    v *= si * sc[j];
    if (i == j) v += 1.0;
    A[(int64_t)j * lda + i] = v;
  }
}

__global__ void diag_logsum_acc_kernel(const double* __restrict__ A, int64_t lda, int w,
                                       double* __restrict__ acc) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < w; i += blockDim.x) s += log(A[(int64_t)i * lda + i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) acc[0] += 2.0 * s;
  }
}

__global__ void diag_dot_s2_kernel(const double* __restrict__ A, int64_t lda, int w,
                                   const double* __restrict__ s, double* __restrict__ acc) {
  // acc += sum_j s_j^2 A_jj over a diagonal block
  __shared__ double red[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < w; i += blockDim.x) v += s[i] * s[i] * A[(int64_t)i * lda + i];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) acc[0] += v;
  }
}

__global__ void acc_vec_kernel(int n, const double* __restrict__ src, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// CUDA + NCCL back end of dist_sched.hpp
struct GpuOps {
  Dist* d = nullptr;
  Ctx* ctx = nullptr;
  dist::Layout* lay = nullptr;
  // what fill_block builds: B = I + S (K + jitter I) S from the prepared features
  const CfPlan* plan = nullptr;
  const double* Z = nullptr;
  const double* s = nullptr;      // device n-vector V^-1/2
  double jitter = 0.0;

  cudaStream_t st(dist::Stream s_) const {
    return s_ == dist::S_MAIN ? main_stream
           : (s_ == dist::S_PANEL ? ctx->panel_stream : (s_ == dist::S_AUX ? d->aux_stream : d->comm_stream));
  }
  cudaStream_t main_stream = nullptr;
  // every library routine launches on ctx->stream: point it at the requested stream for one call
  struct Scope {
    Ctx* c;
    cudaStream_t saved;
    Scope(Ctx* c_, cudaStream_t s_) : c(c_), saved(c_->stream) { c->stream = s_; }
    ~Scope() { c->stream = saved; }
  };
  double* dinv_of(int64_t k) const {
    const int64_t per = (lay->nb + NB - 1) / NB * NB * NB;
    return d->dinv + (k / lay->world) * per;
  }

  int fill_block(dist::Stream S, int64_t b, double* dst, int64_t ld) {
    Scope sc(ctx, st(S));
    const int64_t c0 = lay->col0(b), w = lay->width(b), m = lay->n - c0;
    GPL_TRY(cf_pairwise(ctx, *plan, Z + c0, m, lay->n, Z + c0, w, lay->n, false, 0.0, dst, ld));
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)(w < 64 ? w : 64));
    scale_sym_block_kernel<<<grid, 256, 0, ctx->stream>>>(m, (int)w, dst, ld, s + c0, s + c0, jitter);
    GPL_LAUNCH_CHECK(ctx);
    return GPL_OK;
  }
  int potrf_diag(dist::Stream S, int64_t k, int64_t w, double* A, int64_t lda, int pivot0) {
    Scope sc(ctx, st(S));
    return dpotrf_nosync(ctx, dinv_of(k), w, A, lda, d->d_info, pivot0);
  }
  int trsm_below(dist::Stream S, int64_t k, int64_t c0, int64_t wc, int64_t m, int64_t w, const double* L,
                 int64_t ldl, double* B, int64_t ldb) {
    (void)wc;
    Scope sc(ctx, st(S));
    return dtrsm_right_with(ctx, dinv_of(k) + (c0 / NB) * NB * NB, true, m, w, L, ldl, B, ldb);
  }
  int inv_diag(dist::Stream S, int64_t k, int64_t w, const double* L, int64_t ldl, double* out, int64_t ldo) {
    Scope sc(ctx, st(S));
    return dtrtri_block(ctx, dinv_of(k), w, L, ldl, out, ldo);
  }
  int gemm(dist::Stream S, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
           int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower,
           const dist::Stair* stair) {
    Scope sc(ctx, st(S));
    GemmStair gs{0, 0, 0, 0};
    if (stair) gs = GemmStair{stair->nb, stair->first, stair->step, stair->bskip};
    return dgemm(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower, 0, 0, stair ? &gs : nullptr);
  }
  int copy2d(dist::Stream S, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows,
             int64_t cols) {
    if (rows <= 0 || cols <= 0) return GPL_OK;
    GPL_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * lds, sizeof(double) * rows,
                               cols, cudaMemcpyDeviceToDevice, st(S)));
    return GPL_OK;
  }
  int zero(dist::Stream S, double* p, int64_t count) {
    if (count > 0) GPL_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * count, st(S)));
    return GPL_OK;
  }
  int trsv(dist::Stream S, int64_t k, bool trans, int64_t w, const double* L, int64_t ldl, double* b,
           double* x) {
    Scope sc(ctx, st(S));
    return trsv_lower(ctx, dinv_of(k), trans, w, L, ldl, b, x);
  }
  int gemv(dist::Stream S, bool trans, int64_t m, int64_t n, const double* A, int64_t lda, const double* x,
           double* y) {
    Scope sc(ctx, st(S));
    return trans ? gemv_t(ctx, m, n, A, lda, x, y) : gemv_n(ctx, m, n, A, lda, x, y);
  }
  int axpy(dist::Stream S, int64_t n, double a, const double* x, double* y) {
    Scope sc(ctx, st(S));
    return vec_op(ctx, VOP_AXPBY, n, x, y, y, a, 1.0);
  }
  int sub(dist::Stream S, int64_t n, const double* a, const double* b, double* out) {
    Scope sc(ctx, st(S));
    return vec_op(ctx, VOP_SUB, n, a, b, out);
  }
  int colsumsq(dist::Stream S, int64_t m, int64_t n, const double* A, int64_t lda, double* out) {
    Scope sc(ctx, st(S));
    return gpl::colsumsq(ctx, m, n, A, lda, out);
  }
  int diag_logsum(dist::Stream S, int64_t w, const double* A, int64_t lda, double* acc) {
    diag_logsum_acc_kernel<<<1, 1024, 0, st(S)>>>(A, lda, (int)w, acc);
    GPL_LAUNCH_CHECK(ctx);
    return GPL_OK;
  }
  // GPL_DIST_TRACE=<prefix>: every schedule event is also time-stamped (a timing event on the same
  // stream) and dumped per rank as <prefix>.<rank>.csv after the factorisation -- the device-side timeline
  struct TraceRec { int ev; int stream; cudaEvent_t t; };
  std::vector<TraceRec> trace;
  cudaEvent_t trace_t0 = nullptr;
  bool tracing = false;
  int record(dist::Event e, dist::Stream S) {
    GPL_CUDA(cudaEventRecord(d->ev[e], st(S)));
    if (tracing) {
      cudaEvent_t t;
      GPL_CUDA(cudaEventCreate(&t));
      GPL_CUDA(cudaEventRecord(t, st(S)));
      trace.push_back({(int)e, (int)S, t});
    }
    return GPL_OK;
  }
  int trace_begin() {
    const char* p = getenv("GPL_DIST_TRACE");
    tracing = p && *p;
    if (!tracing) return GPL_OK;
    trace.clear();
    if (!trace_t0) GPL_CUDA(cudaEventCreate(&trace_t0));
    GPL_CUDA(cudaEventRecord(trace_t0, main_stream));
    return GPL_OK;
  }
  int trace_dump() {
    if (!tracing) return GPL_OK;
    GPL_CUDA(cudaDeviceSynchronize());
    char path[512];
    snprintf(path, sizeof(path), "%s.%d.csv", getenv("GPL_DIST_TRACE"), d->rank);
    FILE* f = fopen(path, "w");
    if (f) fprintf(f, "event,stream,ms\n");
    for (const TraceRec& r : trace) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, trace_t0, r.t);
      if (f) fprintf(f, "%d,%d,%.4f\n", r.ev, r.stream, ms);
      cudaEventDestroy(r.t);
    }
    if (f) fclose(f);
    trace.clear();
    tracing = false;
    return GPL_OK;
  }
  int wait(dist::Stream S, dist::Event e) {
    GPL_CUDA(cudaStreamWaitEvent(st(S), d->ev[e], 0));
    return GPL_OK;
  }
  int sync(dist::Stream S) {
    GPL_CUDA(cudaStreamSynchronize(st(S)));
    return GPL_OK;
  }
  int bcast(dist::Stream S, double* buf, int64_t count, int root) {
    if (d->world == 1 || count <= 0) return GPL_OK;
    // diagnostics only (wrong results): GPL_DIST_NOCOMM=1 skips the panel broadcasts so that a run times
    // the compute schedule alone
    static const int nocomm = getenv("GPL_DIST_NOCOMM") ? atoi(getenv("GPL_DIST_NOCOMM")) : 0;
    if (nocomm && S == dist::S_COMM) return GPL_OK;
    // panels shorter than ~4 nb P rows belong to the latency-bound tail of a sweep (the trailing updates no
    // longer cover the panel chain): there the uncapped communicator is used
    const bool tail = d->comm_wide && S == dist::S_COMM && count < 4 * lay->nb * lay->nb * (int64_t)lay->world;
    GPL_NCCL(d, d->api->Broadcast(buf, buf, (size_t)count, ncclDouble, root, tail ? d->comm_wide : d->comm, st(S)));
    d->nccl_bytes += 8.0 * (double)count;
    return GPL_OK;
  }
  int allreduce_sum(dist::Stream S, double* buf, int64_t count) {
    if (d->world == 1 || count <= 0) return GPL_OK;
    GPL_NCCL(d, d->api->AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, d->comm, st(S)));
    d->nccl_bytes += 8.0 * (double)count;
    return GPL_OK;
  }
  int check_info(int* first_bad) {
    if (d->world > 1)
      GPL_NCCL(d, d->api->AllReduce(d->d_info, d->d_info, 1, ncclInt32, ncclMin, d->comm, main_stream));
    int h = 0;
    GPL_CUDA(cudaMemcpyAsync(&h, d->d_info, sizeof(int), cudaMemcpyDeviceToHost, main_stream));
    GPL_CUDA(cudaStreamSynchronize(main_stream));
    *first_bad = (h == INT_MAX) ? 0 : h;
    return GPL_OK;
  }
};

using Chol = dist::ShardedChol<GpuOps>;

// everything one distributed fit needs on the device
struct DistState {
  Dist* d = nullptr;
  Ctx* ctx = nullptr;
  dist::Layout lay;
  GpuOps ops;
  Chol ch;
  HostPlan hp;
  DevBuf x, y, trials, offset, Z, Zaux;
  DevBuf v[20];
  LikDev lik;
  int64_t n = 0;
  int dcols = 0;
};

enum { D_F = 0, D_Z, D_V, D_S, D_ALPHA, D_FNEW, D_T1, D_T2, D_RHS, D_SOL, D_W0, D_W1, D_W2, D_W3, D_FVAR,
       D_P, D_CM, D_CV, D_LOGZ, D_DIAG };

int dist_setup(Dist* d, DistState& st, const gpl_cf_model* cf, const gpl_lik* lik, const double* x, int64_t n,
               int dcols, const double* y, const double* trials, const double* offset, double jitter,
               int64_t nb, bool want_inverse, bool want_aux) {
  Ctx* ctx = d->ctx;
  if (n <= 0 || dcols <= 0 || !x || !y || !cf || !lik) {
    set_error("gpl_dist_*: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  if (nb <= 0) nb = 1024;
  if (nb % NB) {
    set_error("gpl_dist_*: the block width nb must be a multiple of %d", NB);
    return GPL_ERR_BAD_ARG;
  }
  if ((lik->kind == GPL_LIK_BINOMIAL || lik->kind == GPL_LIK_BETABINOM) && !trials) {
    set_error("trials must be provided for the binomial / beta-binomial likelihood");
    return GPL_ERR_BAD_ARG;
  }
  GPL_CUDA(cudaSetDevice(ctx->device));
  st.d = d;
  st.ctx = ctx;
  st.n = n;
  st.dcols = dcols;
  st.lay.init(n, nb, d->world, d->rank);
  const dist::Layout& L = st.lay;
  auto up = [&](DevBuf& b, const double* h, size_t cnt) -> int {
    GPL_TRY(b.alloc(ctx, sizeof(double) * cnt));
    GPL_CUDA(cudaMemcpyAsync(b.p, h, sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx->stream));
    return GPL_OK;
  };
  GPL_TRY(up(st.x, x, (size_t)n * dcols));
  GPL_TRY(up(st.y, y, (size_t)n));
  st.lik = LikDev{lik->kind, lik->link, lik->param, nullptr, nullptr};
  if (trials) { GPL_TRY(up(st.trials, trials, (size_t)n)); st.lik.trials = st.trials.d(); }
  if (offset) { GPL_TRY(up(st.offset, offset, (size_t)n)); st.lik.offset = st.offset.d(); }
  GPL_TRY(make_plan(cf, dcols, 0, &st.hp, want_aux));
  const int F = st.hp.plan.nfeat > 0 ? st.hp.plan.nfeat : 1, FA = st.hp.plan.naux;
  GPL_TRY(st.Z.alloc(ctx, sizeof(double) * n * F));
  if (FA > 0) GPL_TRY(st.Zaux.alloc(ctx, sizeof(double) * n * FA));
  GPL_TRY(prep_features(ctx, st.hp, st.x.d(), n, n, st.Z.d(), n, FA > 0 ? st.Zaux.d() : nullptr, n));
  for (int i = 0; i < 20; ++i) GPL_TRY(st.v[i].alloc(ctx, sizeof(double) * n));
  // the large buffers, kept by the communicator object between calls
  const size_t slab = (size_t)(L.nloc > 0 ? L.nloc : 1) * L.ld;
  GPL_TRY(ensure(&d->A, &d->capA, slab));
  if (want_inverse) GPL_TRY(ensure(&d->X, &d->capX, slab));
  // panel window: look-ahead depth D = W - 1 (capped by the world size)
  static const int forced_w = getenv("GPL_DIST_WINDOW") ? atoi(getenv("GPL_DIST_WINDOW")) : 0;
  int W = d->world >= 8 ? 8 : (d->world >= 4 ? 5 : (d->world >= 2 ? 3 : 2));
  if (forced_w >= 2 && forced_w <= dist::MAX_WINDOW) W = forced_w;
  for (int i = 0; i < W; ++i) GPL_TRY(ensure(&d->panel[i], &d->capP[i], (size_t)nb * L.ld));
  if (want_inverse) {
    const int64_t tc = L.nloc > nb ? L.nloc : nb;
    GPL_TRY(ensure(&d->tmp, &d->capT, (size_t)nb * tc));
  }
  const int64_t nown = (L.nblk + d->world - 1) / d->world;
  GPL_TRY(ensure(&d->dinv, &d->capD, (size_t)nown * ((nb + NB - 1) / NB) * NB * NB));
  st.ops.d = d;
  st.ops.ctx = ctx;
  st.ops.lay = &st.lay;
  st.ops.plan = &st.hp.plan;
  st.ops.Z = st.Z.d();
  st.ops.s = st.v[D_S].d();
  st.ops.jitter = jitter;
  st.ops.main_stream = ctx->stream;
  st.ch.lay = st.lay;
  st.ch.ops = &st.ops;
  st.ch.A = d->A;
  st.ch.X = want_inverse ? d->X : nullptr;
  st.ch.window = W;
  st.ch.chunked_tail = !(getenv("GPL_DIST_TAIL") && atoi(getenv("GPL_DIST_TAIL")) == 0);
  for (int i = 0; i < W; ++i) st.ch.panel[i] = d->panel[i];
  st.ch.tmp = d->tmp;
  st.ch.vec[0] = st.v[D_W0].d();
  st.ch.vec[1] = st.v[D_W1].d();
  st.ch.vec[2] = st.v[D_W2].d();
  st.ch.vec[3] = st.v[D_W3].d();
  st.ch.logdet_acc = d->d_acc;
  return GPL_OK;
}

// B(s) = I + S (K + jitter I) S from v[D_S]; factor; logdet(B) on every rank
int dist_factor(DistState& st, double* logdet) {
  Dist* d = st.d;
  Ctx* ctx = st.ctx;
  const int big = INT_MAX;
  GPL_CUDA(cudaMemcpyAsync(d->d_info, &big, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  int info = 0;
  GPL_TRY(st.ops.trace_begin());
  GPL_TRY(st.ch.factor(&info));
  GPL_TRY(st.ops.trace_dump());
  if (info != 0) {
    set_error("the leading minor of order %d is not positive definite", info);
    return GPL_ERR_NOT_PD;
  }
  GPL_TRY(st.ops.allreduce_sum(dist::S_MAIN, d->d_acc, 1));
  GPL_CUDA(cudaMemcpyAsync(logdet, d->d_acc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  return GPL_OK;
}

// alpha = S B^-1 S z  (R/laplace.R:29), fnew = z - V alpha (== K alpha, since (K + V) alpha = z)
int dist_alpha(DistState& st) {
  Ctx* ctx = st.ctx;
  const int64_t n = st.n;
  GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[D_Z].d(), st.v[D_S].d(), st.v[D_RHS].d()));
  GPL_TRY(st.ch.solve(st.v[D_RHS].d(), st.v[D_SOL].d()));
  GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[D_SOL].d(), st.v[D_S].d(), st.v[D_ALPHA].d()));
  GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[D_V].d(), st.v[D_ALPHA].d(), st.v[D_T1].d()));
  GPL_TRY(vec_op(ctx, VOP_SUB, n, st.v[D_Z].d(), st.v[D_T1].d(), st.v[D_FNEW].d()));
  return GPL_OK;
}

}  // namespace

// ---- interface for the row-sharded FITC fit (fitc.cu) ---------------------------------------------
Ctx* dist_ctx(Dist* d) { return d->ctx; }
int dist_rank(const Dist* d) { return d->rank; }
int dist_world(const Dist* d) { return d->world; }

int dist_allreduce_f64(Dist* d, Ctx* ctx, double* buf, int64_t count, int op) {
  if (d->world == 1 || count <= 0) return GPL_OK;
  const ncclRedOp_t o = op == 0 ? ncclSum : (op == 1 ? ncclMax : ncclMin);
  // m x m sums are bandwidth-sized (134 MB at m = 4096): the uncapped communicator when there is one
  ncclComm_t c = (count >= (1 << 20) && d->comm_wide) ? d->comm_wide : d->comm;
  GPL_NCCL(d, d->api->AllReduce(buf, buf, (size_t)count, ncclDouble, o, c, ctx->stream));
  d->nccl_bytes += 8.0 * (double)count;
  return GPL_OK;
}

int dist_allreduce_i32_max(Dist* d, Ctx* ctx, int* buf, int64_t count) {
  if (d->world == 1 || count <= 0) return GPL_OK;
  GPL_NCCL(d, d->api->AllReduce(buf, buf, (size_t)count, ncclInt32, ncclMax, d->comm, ctx->stream));
  return GPL_OK;
}

// ---- Laplace / EP over the sharded factor --------------------------------------------------------
int dist_fit_full_impl(Dist* d, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                       const double* x, int64_t n, int dcols, const double* y, const double* trials,
                       const double* offset, double jitter, int64_t nb, gpl_fit_out* out, double* h_fmean,
                       double* h_z, double* h_V, double* h_alpha, double* h_fvar, double* h_cm, double* h_cv,
                       gpl_dist_timings* tm) {
  if (!out || !approx) {
    set_error("gpl_dist_fit_full: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  const auto t_start = std::chrono::steady_clock::now();
  const bool ep = approx->kind == GPL_APPROX_EP;
  DistState st;
  d->nccl_bytes = 0.0;
  GPL_TRY(dist_setup(d, st, cf, lik, x, n, dcols, y, trials, offset, jitter, nb, ep, false));
  Ctx* ctx = st.ctx;
  memset(out, 0, sizeof(*out));
  out->quad_ok = 1;
  double sc[8], logdet = 0.0, log_evidence = 0.0;
  auto dl = [&](double* h, const double* dv) -> int {
    if (h) GPL_CUDA(cudaMemcpyAsync(h, dv, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    return GPL_OK;
  };
  double t_factor = 0.0, t_solve = 0.0, t_trtri = 0.0;
  auto now = [&]() {
    cudaStreamSynchronize(ctx->stream);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  };

  if (!ep) {
    int maxiter = approx->maxiter;
    if (lik->kind == GPL_LIK_GAUSSIAN) maxiter = 1;                         // R/laplace.R:281-283
    if (maxiter < 1) maxiter = 1;
    GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, st.v[D_F].d(), 0.0));   // R/laplace.R:280
    int it = 0;
    for (it = 1; it <= maxiter; ++it) {
      // one laplace_iter.method_full (R/laplace.R:13-38) on the sharded matrix
      GPL_TRY(lik_pseudodata_la(ctx, st.lik, n, st.v[D_F].d(), st.y.d(), st.v[D_Z].d(), st.v[D_V].d()));
      GPL_TRY(reduce(ctx, RED_MIN, n, st.v[D_V].d(), nullptr, 0));
      GPL_TRY(fetch_scalars(ctx, 1, sc));
      if (!(sc[0] > 0.0)) {
        set_error("pseudo-variance <= 0: the non-log-concave branch (R/laplace.R:40-117) is single-GPU only");
        return GPL_ERR_NOT_LOGCONCAVE;
      }
      GPL_TRY(vec_op(ctx, VOP_RSQRT_INV, n, st.v[D_V].d(), nullptr, st.v[D_S].d()));
      double t0 = now();
      GPL_TRY(dist_factor(st, &logdet));                                    // C_logdet - V_logdet = logdet(B)
      double t1 = now();
      GPL_TRY(dist_alpha(st));
      double t2 = now();
      t_factor += t1 - t0;
      t_solve += t2 - t1;
      GPL_TRY(lik_loglik_sum(ctx, st.lik, n, st.v[D_FNEW].d(), st.y.d(), 0));
      GPL_TRY(reduce(ctx, RED_DOT, n, st.v[D_FNEW].d(), st.v[D_ALPHA].d(), 1));
      GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, st.v[D_F].d(), st.v[D_FNEW].d(), 2));
      GPL_TRY(fetch_scalars(ctx, 3, sc));
      log_evidence = sc[0] - 0.5 * sc[1] - 0.5 * logdet;                    // R/laplace.R:33-38
      std::swap(st.v[D_F].p, st.v[D_FNEW].p);
      if (sc[2] < approx->tol) break;
    }
    if (it > maxiter) it = maxiter;
    out->log_evidence = log_evidence;
    out->iters = it;
    out->accepted = it;
    out->converged = !(maxiter > 1 && it == maxiter);
    GPL_TRY(dl(h_fmean, st.v[D_F].d()));
    GPL_TRY(dl(h_z, st.v[D_Z].d()));
    GPL_TRY(dl(h_V, st.v[D_V].d()));
    GPL_TRY(dl(h_alpha, st.v[D_ALPHA].d()));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  } else {
    // parallel EP (R/ep.R:16-49, 286-330); marginal variances from diag(B^-1) (the X sweep):
    // fvar = V - V diag(B^-1), the same identity as on one GPU instead of the n x n solve of R/ep.R:42,141
    if (!approx->quad_x || !approx->quad_w || approx->quad_order < 1) {
      set_error("EP needs Gauss-Hermite nodes and weights");
      return GPL_ERR_BAD_ARG;
    }
    DevBuf qx, qw, flags, fmean_old, fvar_old, z_old, P_old;
    GPL_TRY(qx.alloc(ctx, sizeof(double) * approx->quad_order));
    GPL_TRY(qw.alloc(ctx, sizeof(double) * approx->quad_order));
    GPL_CUDA(cudaMemcpyAsync(qx.p, approx->quad_x, sizeof(double) * approx->quad_order, cudaMemcpyHostToDevice,
                             ctx->stream));
    GPL_CUDA(cudaMemcpyAsync(qw.p, approx->quad_w, sizeof(double) * approx->quad_order, cudaMemcpyHostToDevice,
                             ctx->stream));
    GPL_TRY(flags.alloc(ctx, sizeof(int) * 4));
    for (DevBuf* b : {&fmean_old, &fvar_old, &z_old, &P_old}) GPL_TRY(b->alloc(ctx, sizeof(double) * n));
    GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, fmean_old.d(), 0.0));
    GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, z_old.d(), 0.0));
    GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, P_old.d(), 0.0));
    GPL_TRY(cf_diag(ctx, st.hp.plan, st.Z.d(), n, st.Z.d(), n, n, jitter, fvar_old.d()));   // diag(K), R/ep.R:297
    double diff_old = INFINITY, damp = approx->damping, diff_new = 0.0;
    const int maxiter = approx->maxiter < 1 ? 1 : approx->maxiter;
    int it = 0, quad_ok = 1, accepted = 0;
    for (it = 1; it <= maxiter; ++it) {
      EpSiteArgs ea;
      ea.fmean = fmean_old.d(); ea.fvar = fvar_old.d(); ea.z_old = z_old.d(); ea.P_old = P_old.d();
      ea.y = st.y.d();
      ea.z_new = st.v[D_Z].d(); ea.V_new = st.v[D_V].d(); ea.P_new = st.v[D_P].d();
      ea.cav_mean = st.v[D_CM].d(); ea.cav_var = st.v[D_CV].d(); ea.logZ = st.v[D_LOGZ].d();
      ea.damping = damp;
      ea.quad_order = approx->quad_order;
      ea.quad_x = qx.d(); ea.quad_w = qw.d();
      GPL_TRY(lik_ep_sites(ctx, st.lik, n, ea, static_cast<int*>(flags.p)));
      int hflags = 0;
      GPL_CUDA(cudaMemcpyAsync(&hflags, flags.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      GPL_CUDA(cudaStreamSynchronize(ctx->stream));
      quad_ok = (hflags & 1) ? 0 : 1;
      if (hflags & 2) {                                                       // R/ep.R:30-32
        out->log_evidence = -INFINITY;
        out->iters = it;
        out->quad_ok = quad_ok;
        set_error("EP: NaN pseudo-data");
        return GPL_ERR_NONFINITE;
      }
      if (hflags & 4) {
        set_error("pseudo-variance <= 0: the non-log-concave branch (R/ep.R:50-138) is single-GPU only");
        return GPL_ERR_NOT_LOGCONCAVE;
      }
      GPL_TRY(vec_op(ctx, VOP_RSQRT_INV, n, st.v[D_V].d(), nullptr, st.v[D_S].d()));
      double t0 = now();
      GPL_TRY(dist_factor(st, &logdet));
      double t1 = now();
      GPL_TRY(dist_alpha(st));
      double t2 = now();
      GPL_TRY(st.ch.x_sweep());
      GPL_TRY(st.ch.diag_binv(st.v[D_DIAG].d()));
      double t3 = now();
      t_factor += t1 - t0;
      t_solve += t2 - t1;
      t_trtri += t3 - t2;
      GPL_TRY(vec_op(ctx, VOP_MUL, n, st.v[D_V].d(), st.v[D_DIAG].d(), st.v[D_T2].d()));
      GPL_TRY(vec_op(ctx, VOP_SUB, n, st.v[D_V].d(), st.v[D_T2].d(), st.v[D_FVAR].d()));     // diag(V C^-1 K)
      GPL_TRY(reduce(ctx, RED_SUMLOG, n, st.v[D_V].d(), nullptr, 0));
      GPL_TRY(reduce(ctx, RED_DOT, n, st.v[D_Z].d(), st.v[D_ALPHA].d(), 1));
      GPL_TRY(ep_evidence_terms(ctx, n, st.v[D_LOGZ].d(), st.v[D_CV].d(), st.v[D_CM].d(), st.v[D_Z].d(),
                                st.v[D_V].d(), 2));
      GPL_TRY(reduce(ctx, RED_MAXABSDIFF, n, fmean_old.d(), st.v[D_FNEW].d(), 3));
      GPL_TRY(fetch_scalars(ctx, 4, sc));
      const double C_logdet = sc[0] + logdet;
      log_evidence = -0.5 * C_logdet - 0.5 * sc[1] + sc[2];                    // R/ep.R:45-49
      diff_new = sc[3];
      if (diff_new > diff_old) {                                              // R/ep.R:306-310
        damp *= approx->damping;
        continue;
      }
      damp = damp > approx->damping_min ? damp : approx->damping_min;
      GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[D_FNEW].d(), nullptr, fmean_old.d()));
      GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[D_FVAR].d(), nullptr, fvar_old.d()));
      GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[D_Z].d(), nullptr, z_old.d()));
      GPL_TRY(vec_op(ctx, VOP_COPY, n, st.v[D_P].d(), nullptr, P_old.d()));
      diff_old = diff_new;
      ++accepted;
      if (diff_new < approx->tol) break;
    }
    if (it > maxiter) it = maxiter;
    out->iters = it;
    out->accepted = accepted;
    out->damping_final = damp;
    out->converged = !(maxiter > 1 && it == maxiter);
    out->quad_ok = quad_ok;
    if (maxiter > 1 && it == maxiter) log_evidence = -INFINITY;               // R/ep.R:322-325
    if (!quad_ok) log_evidence = -INFINITY;
    out->log_evidence = log_evidence;
    GPL_TRY(dl(h_fmean, st.v[D_FNEW].d()));
    GPL_TRY(dl(h_fvar, st.v[D_FVAR].d()));
    GPL_TRY(dl(h_z, st.v[D_Z].d()));
    GPL_TRY(dl(h_V, st.v[D_V].d()));
    GPL_TRY(dl(h_alpha, st.v[D_ALPHA].d()));
    GPL_TRY(dl(h_cm, st.v[D_CM].d()));
    GPL_TRY(dl(h_cv, st.v[D_CV].d()));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (tm) {
    memset(tm, 0, sizeof(*tm));
    tm->factor_s = t_factor;          // includes the K build (about 1 % of a factorisation)
    tm->solve_s = t_solve;
    tm->trtri_s = t_trtri;
    tm->total_s = now();
    tm->nccl_bytes = d->nccl_bytes;
  }
  return GPL_OK;
}

// ---- energy + analytic gradient, lik_gaussian (SURVEY 8e "C^-1 + trace term") --------------------
int dist_energy_grad_impl(Dist* d, const gpl_cf_model* cf, const gpl_lik* lik, const double* x, int64_t n,
                          int dcols, const double* y, double jitter, int64_t nb, int nparam,
                          const int32_t* param_factor, const int32_t* param_slot, double* energy, double* grad,
                          double* h_alpha, gpl_dist_timings* tm) {
  if (!lik || lik->kind != GPL_LIK_GAUSSIAN) {
    set_error("gpl_dist_energy_grad: lik_gaussian + method_full");
    return GPL_ERR_BAD_ARG;
  }
  if (!energy || nparam < 0 || nparam > MAX_GRAD_PARAMS || (nparam > 0 && (!grad || !param_factor || !param_slot))) {
    set_error("gpl_dist_energy_grad: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  const auto t_start = std::chrono::steady_clock::now();
  const bool want_grad = nparam > 0;
  bool want_aux = false;
  for (int j = 0; j < nparam; ++j)
    if (param_factor[j] >= 0 && param_slot[j] == 3) want_aux = true;
  DistState st;
  d->nccl_bytes = 0.0;
  GPL_TRY(dist_setup(d, st, cf, lik, x, n, dcols, y, nullptr, nullptr, jitter, nb, want_grad, want_aux));
  Ctx* ctx = st.ctx;
  const dist::Layout& L = st.lay;
  GradParam gp[MAX_GRAD_PARAMS];
  int pos[MAX_GRAD_PARAMS], ncf = 0;
  for (int j = 0; j < nparam; ++j) {
    if (param_factor[j] < 0) continue;
    if (param_factor[j] >= st.hp.plan.nfactors) {
      set_error("gpl_dist_energy_grad: parameter %d refers to factor %d of %d", j, param_factor[j],
                st.hp.plan.nfactors);
      return GPL_ERR_BAD_ARG;
    }
    gp[ncf].factor = param_factor[j];
    gp[ncf].slot = param_slot[j];
    pos[ncf++] = j;
  }
  double sc[8], logdet = 0.0;
  auto now = [&]() {
    cudaStreamSynchronize(ctx->stream);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  };
  GPL_TRY(vec_op(ctx, VOP_FILL, n, nullptr, nullptr, st.v[D_F].d(), 0.0));
  GPL_TRY(lik_pseudodata_la(ctx, st.lik, n, st.v[D_F].d(), st.y.d(), st.v[D_Z].d(), st.v[D_V].d()));
  GPL_TRY(vec_op(ctx, VOP_RSQRT_INV, n, st.v[D_V].d(), nullptr, st.v[D_S].d()));
  const double t0 = now();
  GPL_TRY(dist_factor(st, &logdet));
  const double t1 = now();
  GPL_TRY(dist_alpha(st));
  GPL_TRY(lik_loglik_sum(ctx, st.lik, n, st.v[D_FNEW].d(), st.y.d(), 0));
  GPL_TRY(reduce(ctx, RED_DOT, n, st.v[D_FNEW].d(), st.v[D_ALPHA].d(), 1));
  GPL_TRY(reduce(ctx, RED_SUMSQ, n, st.v[D_ALPHA].d(), nullptr, 2));
  GPL_TRY(fetch_scalars(ctx, 3, sc));
  const double t2 = now();
  *energy = -(sc[0] - 0.5 * sc[1] - 0.5 * logdet);                            // R/laplace.R:33-38
  const double alpha_sq = sc[2];
  double t3 = t2, t4 = t2, t5 = t2;
  if (want_grad) {
    GPL_TRY(st.ch.x_sweep());
    t3 = now();
    GPL_TRY(st.ch.y_sweep());
    t4 = now();
    // local block traces: sum_ij (alpha_i alpha_j - s_i Binv_ij s_j) dK_ij/dlog(theta_p), and
    // sum_j s_j^2 Binv_jj, accumulated on the device, then ONE all-reduce of |theta| + 1 numbers
    DevBuf acc, one;
    GPL_TRY(acc.alloc(ctx, sizeof(double) * (MAX_GRAD_PARAMS + 8)));
    GPL_TRY(one.alloc(ctx, sizeof(double) * (MAX_GRAD_PARAMS + 8)));
    GPL_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double) * (MAX_GRAD_PARAMS + 8), ctx->stream));
    for (int64_t j = L.first_owned_from(0); j < L.nblk; j += L.world) {
      const int64_t j0 = L.col0(j), w = L.width(j);
      const double* blk = d->X + L.lc0[(size_t)j] * L.ld;
      if (ncf > 0) {
        GPL_TRY(grad_trace(ctx, st.hp.plan, st.Z.d(), n, st.Zaux.d(), n, blk, L.ld, st.v[D_S].d(),
                           st.v[D_ALPHA].d(), ncf, gp, one.d(), j0, w));
        acc_vec_kernel<<<1, 64, 0, ctx->stream>>>(ncf, one.d(), acc.d());
        GPL_LAUNCH_CHECK(ctx);
      }
      diag_dot_s2_kernel<<<1, 1024, 0, ctx->stream>>>(blk + j0, L.ld, (int)w, st.v[D_S].d() + j0,
                                                      acc.d() + MAX_GRAD_PARAMS);
      GPL_LAUNCH_CHECK(ctx);
    }
    GPL_TRY(st.ops.allreduce_sum(dist::S_MAIN, acc.d(), MAX_GRAD_PARAMS + 1));
    double hg[MAX_GRAD_PARAMS + 1];
    GPL_CUDA(cudaMemcpyAsync(hg, acc.p, sizeof(double) * (MAX_GRAD_PARAMS + 1), cudaMemcpyDeviceToHost,
                             ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < ncf; ++q) grad[pos[q]] = -0.5 * hg[q];
    const double sigma2 = lik->param * lik->param;
    for (int j = 0; j < nparam; ++j)
      if (param_factor[j] < 0) grad[j] = -0.5 * 2.0 * sigma2 * (alpha_sq - hg[MAX_GRAD_PARAMS]);
    t5 = now();
  }
  if (h_alpha) {
    GPL_CUDA(cudaMemcpyAsync(h_alpha, st.v[D_ALPHA].d(), sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    GPL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (tm) {
    memset(tm, 0, sizeof(*tm));
    tm->kbuild_s = t0;                 // upload + features (+ allocation on the first call)
    tm->factor_s = t1 - t0;            // K build in the block-cyclic layout + factorisation
    tm->solve_s = t2 - t1;
    tm->trtri_s = t3 - t2;
    tm->lauum_s = t4 - t3;
    tm->trace_s = t5 - t4;
    tm->total_s = now();
    tm->nccl_bytes = d->nccl_bytes;
  }
  return GPL_OK;
}

}  // namespace gpl

using namespace gpl;

extern "C" {

int gpl_dist_unique_id(void* id128) {
  if (!id128) return GPL_ERR_BAD_ARG;
  NcclApi* api = nccl_api();
  if (!api) return GPL_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == GPL_DIST_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) {
    set_error("ncclGetUniqueId: %s", api->GetErrorString(r));
    return GPL_ERR_NCCL;
  }
  memcpy(id128, &id, sizeof(id));
  return GPL_OK;
}

int gpl_dist_init(gpl_ctx* h, const void* id128, int32_t rank, int32_t world, gpl_dist** out) {
  if (!h || !out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) {
    set_error("gpl_dist_init: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  *out = nullptr;
  Ctx* ctx = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaSetDevice(ctx->device));
  Dist* d = new Dist();
  d->ctx = ctx;
  d->rank = rank;
  d->world = world;
  if (world > 1) {
    d->api = nccl_api();
    if (!d->api) { delete d; return GPL_ERR_NCCL; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    // The panel broadcasts overlap the trailing updates, and a rank that reaches a broadcast early SPINS
    // in NCCL's kernel until the slowest rank joins: every CTA NCCL may use is an SM taken from the
    // FP64 GEMM for that long.  A few CTAs are plenty (a bare 800 MB broadcast runs at 600+ GB/s with the
    // default 16; the schedule needs ~60 GB/s), so the communicator is capped (GPL_NCCL_MAX_CTAS, 0 = NCCL default).
    int max_ctas = getenv("GPL_NCCL_MAX_CTAS") ? atoi(getenv("GPL_NCCL_MAX_CTAS")) : 4;
    ncclResult_t r;
    if (max_ctas > 0 && d->api->CommInitRankConfig) {
      ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
      cfg.maxCTAs = max_ctas;
      cfg.minCTAs = 1;
      r = d->api->CommInitRankConfig(&d->comm, world, id, rank, &cfg);
    } else {
      r = d->api->CommInitRank(&d->comm, world, id, rank);
    }
    if (r != ncclSuccess) {
      set_error("ncclCommInitRank: %s", d->api->GetErrorString(r));
      delete d;
      return GPL_ERR_NCCL;
    }
    // second communicator without the cap; its id travels over the first one
    if (max_ctas > 0 && d->api->CommInitRankConfig) {
      ncclUniqueId id2;
      if (rank == 0) {
        r = d->api->GetUniqueId(&id2);
        if (r != ncclSuccess) { set_error("ncclGetUniqueId: %s", d->api->GetErrorString(r)); delete d; return GPL_ERR_NCCL; }
      }
      void* dbuf = nullptr;
      GPL_CUDA(cudaMalloc(&dbuf, sizeof(id2)));
      if (rank == 0) GPL_CUDA(cudaMemcpy(dbuf, &id2, sizeof(id2), cudaMemcpyHostToDevice));
      r = d->api->Broadcast(dbuf, dbuf, sizeof(id2), ncclChar, 0, d->comm, 0);
      if (r != ncclSuccess) { set_error("ncclBroadcast (id): %s", d->api->GetErrorString(r)); delete d; return GPL_ERR_NCCL; }
      GPL_CUDA(cudaStreamSynchronize(0));
      GPL_CUDA(cudaMemcpy(&id2, dbuf, sizeof(id2), cudaMemcpyDeviceToHost));
      GPL_CUDA(cudaFree(dbuf));
      r = d->api->CommInitRank(&d->comm_wide, world, id2, rank);
      if (r != ncclSuccess) { set_error("ncclCommInitRank (wide): %s", d->api->GetErrorString(r)); delete d; return GPL_ERR_NCCL; }
    }
  }
  int lo = 0, hi = 0;
  GPL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  GPL_CUDA(cudaStreamCreateWithPriority(&d->comm_stream, cudaStreamNonBlocking, hi));
  GPL_CUDA(cudaStreamCreateWithPriority(&d->aux_stream, cudaStreamNonBlocking, hi));
  for (int i = 0; i < dist::EV_COUNT; ++i) GPL_CUDA(cudaEventCreateWithFlags(&d->ev[i], cudaEventDisableTiming));
  GPL_CUDA(cudaMalloc(&d->d_info, sizeof(int) * 4));
  GPL_CUDA(cudaMalloc(&d->d_acc, sizeof(double) * 8));
  *out = reinterpret_cast<gpl_dist*>(d);
  return GPL_OK;
}

int gpl_dist_destroy(gpl_dist* h) {
  Dist* d = reinterpret_cast<Dist*>(h);
  if (!d) return GPL_OK;
  cudaSetDevice(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  cudaStreamSynchronize(d->comm_stream);
  if (d->comm_wide && d->api) d->api->CommDestroy(d->comm_wide);
  if (d->comm && d->api) d->api->CommDestroy(d->comm);
  for (double* p : {d->A, d->X, d->tmp, d->dinv, d->d_acc})
    if (p) cudaFree(p);
  for (double* p : d->panel)
    if (p) cudaFree(p);
  if (d->d_info) cudaFree(d->d_info);
  for (int i = 0; i < dist::EV_COUNT; ++i)
    if (d->ev[i]) cudaEventDestroy(d->ev[i]);
  if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
  if (d->aux_stream) cudaStreamDestroy(d->aux_stream);
  delete d;
  return GPL_OK;
}

int gpl_dist_release(gpl_dist* h) {
  Dist* d = reinterpret_cast<Dist*>(h);
  if (!d) return GPL_OK;
  GPL_CUDA(cudaSetDevice(d->ctx->device));
  GPL_CUDA(cudaStreamSynchronize(d->ctx->stream));
  GPL_CUDA(cudaStreamSynchronize(d->comm_stream));
  for (double** p : {&d->A, &d->X, &d->tmp, &d->dinv}) {
    if (*p) GPL_CUDA(cudaFree(*p));
    *p = nullptr;
  }
  for (int i = 0; i < dist::MAX_WINDOW; ++i) {
    if (d->panel[i]) GPL_CUDA(cudaFree(d->panel[i]));
    d->panel[i] = nullptr;
    d->capP[i] = 0;
  }
  d->capA = d->capX = d->capT = d->capD = 0;
  return GPL_OK;
}

int gpl_dist_bcast_bench(gpl_dist* h, int64_t count, int32_t reps, double* seconds) {
  Dist* d = reinterpret_cast<Dist*>(h);
  if (!d || count <= 0 || reps <= 0 || !seconds) {
    set_error("gpl_dist_bcast_bench: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  GPL_CUDA(cudaSetDevice(d->ctx->device));
  GPL_TRY(ensure(&d->panel[0], &d->capP[0], (size_t)count));
  cudaEvent_t e0, e1;
  GPL_CUDA(cudaEventCreate(&e0));
  GPL_CUDA(cudaEventCreate(&e1));
  for (int r = -2; r < reps; ++r) {                      // two warm-up rounds, then `reps` timed ones
    if (r == 0) GPL_CUDA(cudaEventRecord(e0, d->comm_stream));
    if (d->world > 1)
      GPL_NCCL(d, d->api->Broadcast(d->panel[0], d->panel[0], (size_t)count, ncclDouble, (r + 2) % d->world, d->comm,
                                    d->comm_stream));
  }
  GPL_CUDA(cudaEventRecord(e1, d->comm_stream));
  GPL_CUDA(cudaStreamSynchronize(d->comm_stream));
  float ms = 0.f;
  GPL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *seconds = ms * 1e-3 / reps;
  return GPL_OK;
}

int gpl_dist_rank(const gpl_dist* h) { return h ? reinterpret_cast<const Dist*>(h)->rank : -1; }
int gpl_dist_world(const gpl_dist* h) { return h ? reinterpret_cast<const Dist*>(h)->world : 0; }

int gpl_dist_fit_full(gpl_dist* h, const gpl_cf_model* cf, const gpl_lik* lik, const gpl_approx* approx,
                      const double* x, int64_t n, int32_t d, const double* y, const double* trials,
                      const double* offset, double jitter, int64_t nb, gpl_fit_out* out, double* fmean,
                      double* z, double* V, double* alpha, double* fvar, double* cavity_mean,
                      double* cavity_var, gpl_dist_timings* timings) {
  if (!h) {
    set_error("gpl_dist_fit_full: no communicator");
    return GPL_ERR_BAD_ARG;
  }
  return dist_fit_full_impl(reinterpret_cast<Dist*>(h), cf, lik, approx, x, n, d, y, trials, offset, jitter, nb,
                            out, fmean, z, V, alpha, fvar, cavity_mean, cavity_var, timings);
}

int gpl_dist_energy_grad(gpl_dist* h, const gpl_cf_model* cf, const gpl_lik* lik, const double* x, int64_t n,
                         int32_t d, const double* y, double jitter, int64_t nb, int32_t nparam,
                         const int32_t* param_factor, const int32_t* param_slot, double* energy, double* grad,
                         double* alpha, gpl_dist_timings* timings) {
  if (!h) {
    set_error("gpl_dist_energy_grad: no communicator");
    return GPL_ERR_BAD_ARG;
  }
  return dist_energy_grad_impl(reinterpret_cast<Dist*>(h), cf, lik, x, n, d, y, jitter, nb, nparam, param_factor,
                               param_slot, energy, grad, alpha, timings);
}

}  // extern "C"
