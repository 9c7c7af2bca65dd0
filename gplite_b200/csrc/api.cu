// C-ABI entry points of libgplite_b200.so (include/gplite_b200.h): context management,
// covariance evaluation and the dense linear-algebra hooks.  Fits, energy/gradient and
// prediction live in fit.cu.
#include "cf.cuh"
#include "fit.cuh"

#include <stdarg.h>
#include <string.h>

namespace gpl {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cf_eval_dev(Ctx* ctx, const gpl_cf_model* cf, const double* d_x1, int64_t n1, int64_t ldx1,
                const double* d_x2, int64_t n2, int64_t ldx2, int d, bool same,
                uint64_t term_mask, bool diag_only, double jitter, double* d_out, int64_t ldk) {
  HostPlan hp;
  GPL_TRY(make_plan(cf, d, term_mask, &hp));
  const int F = hp.plan.nfeat;
  DevBuf z1, z2;                       // freed (stream-ordered) on every return path
  const int64_t ldz1 = n1 > 0 ? n1 : 1, ldz2 = n2 > 0 ? n2 : 1;
  if (F > 0) {
    GPL_TRY(z1.alloc(ctx, sizeof(double) * ldz1 * F));
    GPL_TRY(prep_features(ctx, hp, d_x1, n1, ldx1, z1.d(), ldz1));
    if (!same) {
      GPL_TRY(z2.alloc(ctx, sizeof(double) * ldz2 * F));
      GPL_TRY(prep_features(ctx, hp, d_x2, n2, ldx2, z2.d(), ldz2));
    }
  }
  if (diag_only) {
    const int64_t n = same ? n1 : (n1 < n2 ? n1 : n2);
    return cf_diag(ctx, hp.plan, z1.d(), ldz1, same ? z1.d() : z2.d(), same ? ldz1 : ldz2, n, jitter,
                   d_out);
  }
  return cf_pairwise(ctx, hp.plan, z1.d(), n1, ldz1, z2.d(), n2, ldz2, same, jitter, d_out, ldk);
}

}  // namespace gpl

using namespace gpl;

extern "C" {

int gpl_abi_version(void) { return GPL_ABI_VERSION; }

const char* gpl_last_error(void) { return g_err; }

int gpl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int gpl_ctx_create(int device, uint64_t stream, gpl_ctx** out) {
  if (!out) return GPL_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libgplite_b200 has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return GPL_ERR_CUDA;
  }
  GPL_CUDA(cudaSetDevice(device));
  Ctx* c = new Ctx();
  c->device = device;
  if (stream) {
    c->stream = (cudaStream_t)(uintptr_t)stream;
    c->own_stream = false;
  } else {
    GPL_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  {
    int lo = 0, hi = 0;
    GPL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    GPL_CUDA(cudaStreamCreateWithPriority(&c->panel_stream, cudaStreamNonBlocking, hi));
    GPL_CUDA(cudaEventCreateWithFlags(&c->ev_col, cudaEventDisableTiming));
    GPL_CUDA(cudaEventCreateWithFlags(&c->ev_diag, cudaEventDisableTiming));
    for (int i = 0; i < Ctx::NAUX; ++i) {
      GPL_CUDA(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
      GPL_CUDA(cudaEventCreateWithFlags(&c->aux_ev[i], cudaEventDisableTiming));
    }
  }
  cudaDeviceProp prop;
  GPL_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  GPL_CUDA(cudaMalloc(&c->d_info, sizeof(int32_t) * 4));
  GPL_CUDA(cudaMalloc(&c->d_scal, sizeof(double) * 256));
  GPL_CUDA(cudaMallocHost(&c->h_scal, sizeof(double) * 256));
  // keep freed async allocations cached in the pool: fits reuse multi-GB buffers
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = reinterpret_cast<gpl_ctx*>(c);
  return GPL_OK;
}

int gpl_ctx_destroy(gpl_ctx* h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return GPL_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->dinv) cudaFree(c->dinv);
  if (c->d_info) cudaFree(c->d_info);
  if (c->d_scal) cudaFree(c->d_scal);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->panel_stream) cudaStreamDestroy(c->panel_stream);
  if (c->ev_col) cudaEventDestroy(c->ev_col);
  if (c->ev_diag) cudaEventDestroy(c->ev_diag);
  for (int i = 0; i < Ctx::NAUX; ++i) {
    if (c->aux[i]) cudaStreamDestroy(c->aux[i]);
    if (c->aux_ev[i]) cudaEventDestroy(c->aux_ev[i]);
  }
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return GPL_OK;
}

int gpl_ctx_sync(gpl_ctx* h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaStreamSynchronize(c->stream));
  return GPL_OK;
}

int gpl_ctx_release_cache(gpl_ctx* h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaSetDevice(c->device));
  GPL_CUDA(cudaStreamSynchronize(c->stream));
  if (c->dinv) GPL_CUDA(cudaFree(c->dinv));
  c->dinv = nullptr;
  c->dinv_blocks = 0;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) GPL_CUDA(cudaMemPoolTrimTo(pool, 0));
  return GPL_OK;
}

int64_t gpl_ctx_launch_count(gpl_ctx* h) { return reinterpret_cast<Ctx*>(h)->launches; }

int gpl_ctx_profile(gpl_ctx* h, int32_t enable) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaStreamSynchronize(c->stream));
  c->profile = enable != 0;
  c->ev_used = 0;
  c->prof_flops = 0.0;
  c->prof_ms = 0.0;
  c->prof_launches = 0;
  c->prof_recs.clear();
  return GPL_OK;
}

int gpl_ctx_gemm_config(gpl_ctx* h, int32_t cfg) {
  if (cfg < 0 || cfg > 3) {
    set_error("gpl_ctx_gemm_config: cfg must be 0 (auto), 1 (128x128), 2 (64x64) or 3 (32x32)");
    return GPL_ERR_BAD_ARG;
  }
  reinterpret_cast<Ctx*>(h)->gemm_force = cfg;
  return GPL_OK;
}

int gpl_ctx_profile_read(gpl_ctx* h, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaStreamSynchronize(c->stream));
  // GPL_GEMM_LOG=path: append one CSV line per profiled launch (shape, tile, milliseconds)
  const char* log_path = getenv("GPL_GEMM_LOG");
  FILE* log = (log_path && c->ev_used) ? fopen(log_path, "a") : nullptr;
  for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
    float ms = 0.f;
    GPL_CUDA(cudaEventElapsedTime(&ms, c->ev_pool[i], c->ev_pool[i + 1]));
    c->prof_ms += ms;
    if (log && i / 2 < c->prof_recs.size()) {
      const Ctx::GemmRec& r = c->prof_recs[i / 2];
      fprintf(log, "%d,%d,%d,%d,%d,%d,%.5f\n", r.m, r.n, r.k, r.lower, r.kskip, r.tile, ms);
    }
  }
  if (log) fclose(log);
  c->ev_used = 0;
  c->prof_recs.clear();
  if (gemm_ms) *gemm_ms = c->prof_ms;
  if (gemm_flops) *gemm_flops = c->prof_flops;
  if (gemm_launches) *gemm_launches = c->prof_launches;
  return GPL_OK;
}

int gpl_dev_alloc(gpl_ctx* h, uint64_t bytes, void** dptr) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaSetDevice(c->device));
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    set_error("device allocation of %llu bytes failed", (unsigned long long)bytes);
    return GPL_ERR_NOMEM;
  }
  GPL_CUDA(e);
  return GPL_OK;
}

int gpl_dev_free(gpl_ctx* h, void* dptr) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaStreamSynchronize(c->stream));
  GPL_CUDA(cudaFree(dptr));
  return GPL_OK;
}

int gpl_memcpy_h2d(gpl_ctx* h, void* dst, const void* src, uint64_t bytes) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  GPL_CUDA(cudaStreamSynchronize(c->stream));
  return GPL_OK;
}

int gpl_memcpy_d2h(gpl_ctx* h, void* dst, const void* src, uint64_t bytes) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  GPL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  GPL_CUDA(cudaStreamSynchronize(c->stream));
  return GPL_OK;
}

int gpl_cf_eval_dev(gpl_ctx* h, const gpl_cf_model* cf, const double* d_x1, int64_t n1,
                    const double* d_x2, int64_t n2, int32_t d, int32_t same, uint64_t term_mask,
                    int32_t diag_only, double jitter, double* d_out, int64_t ldk) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (n1 < 0 || n2 < 0 || d <= 0 || !d_out || (!diag_only && ldk < n1)) {
    set_error("gpl_cf_eval_dev: bad dimensions");
    return GPL_ERR_BAD_ARG;
  }
  return cf_eval_dev(c, cf, d_x1, n1, n1 > 0 ? n1 : 1, same ? d_x1 : d_x2, same ? n1 : n2,
                     same ? (n1 > 0 ? n1 : 1) : (n2 > 0 ? n2 : 1), d, same != 0, term_mask,
                     diag_only != 0, jitter, d_out, ldk);
}

int gpl_cf_eval(gpl_ctx* h, const gpl_cf_model* cf, const double* x1, int64_t n1, const double* x2,
                int64_t n2, int32_t d, int32_t same, uint64_t term_mask, int32_t diag_only,
                double jitter, double* out) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (same) { x2 = x1; n2 = n1; }
  if (n1 < 0 || n2 < 0 || d <= 0) {
    set_error("gpl_cf_eval: bad dimensions");
    return GPL_ERR_BAD_ARG;
  }
  const int64_t nout = diag_only ? (n1 < n2 ? n1 : n2) : n1 * n2;
  if (nout == 0) return GPL_OK;
  DevBuf dx1, dx2, dk;
  cudaStream_t s = c->stream;
  GPL_TRY(dx1.alloc(c, sizeof(double) * n1 * d));
  GPL_CUDA(cudaMemcpyAsync(dx1.p, x1, sizeof(double) * n1 * d, cudaMemcpyHostToDevice, s));
  if (!same) {
    GPL_TRY(dx2.alloc(c, sizeof(double) * n2 * d));
    GPL_CUDA(cudaMemcpyAsync(dx2.p, x2, sizeof(double) * n2 * d, cudaMemcpyHostToDevice, s));
  }
  GPL_TRY(dk.alloc(c, sizeof(double) * nout));
  GPL_TRY(cf_eval_dev(c, cf, dx1.d(), n1, n1, same ? dx1.d() : dx2.d(), n2, n2, d, same != 0, term_mask,
                      diag_only != 0, jitter, dk.d(), n1));
  GPL_CUDA(cudaMemcpyAsync(out, dk.p, sizeof(double) * nout, cudaMemcpyDeviceToHost, s));
  GPL_CUDA(cudaStreamSynchronize(s));
  return GPL_OK;
}

int gpl_cf_eval_block_dev(gpl_ctx* h, const gpl_cf_model* cf, const double* d_x1, int64_t n1,
                          int64_t ldx1, const double* d_x2, int64_t n2, int64_t ldx2, int32_t d,
                          double* d_out, int64_t ldk) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (n1 < 0 || n2 < 0 || d <= 0 || !d_out || ldk < n1 || ldx1 < n1 || ldx2 < n2) {
    set_error("gpl_cf_eval_block_dev: bad dimensions");
    return GPL_ERR_BAD_ARG;
  }
  return cf_eval_dev(c, cf, d_x1, n1, ldx1, d_x2, n2, ldx2, d, false, 0, false, 0.0, d_out, ldk);
}

static int native_cf(gpl_ctx* h, int kind, const double* x1, int64_t n1, const double* x2,
                     int64_t n2, int32_t d, double p0, double p1, double p2, int32_t diag_only,
                     double* out) {
  gpl_cf_factor f;
  memset(&f, 0, sizeof(f));
  f.kind = kind;
  f.p[0] = p0; f.p[1] = p1; f.p[2] = p2;
  gpl_cf_model m{1, 1, &f};
  return gpl_cf_eval(h, &m, x1, n1, x2, n2, d, 0, 0, diag_only, 0.0, out);
}

int gpl_cf_sexp_c(gpl_ctx* h, const double* x1, int64_t n1, const double* x2, int64_t n2,
                  int32_t d, double lscale, double magn, int32_t diag_only, double* out) {
  return native_cf(h, GPL_CF_SEXP, x1, n1, x2, n2, d, lscale, magn, 0, diag_only, out);
}
int gpl_cf_matern32_c(gpl_ctx* h, const double* x1, int64_t n1, const double* x2, int64_t n2,
                      int32_t d, double lscale, double magn, int32_t diag_only, double* out) {
  return native_cf(h, GPL_CF_MATERN32, x1, n1, x2, n2, d, lscale, magn, 0, diag_only, out);
}
int gpl_cf_matern52_c(gpl_ctx* h, const double* x1, int64_t n1, const double* x2, int64_t n2,
                      int32_t d, double lscale, double magn, int32_t diag_only, double* out) {
  return native_cf(h, GPL_CF_MATERN52, x1, n1, x2, n2, d, lscale, magn, 0, diag_only, out);
}
int gpl_cf_nn_c(gpl_ctx* h, const double* x1, int64_t n1, const double* x2, int64_t n2, int32_t d,
                double sigma0, double sigma, double magn, int32_t diag_only, double* out) {
  return native_cf(h, GPL_CF_NN, x1, n1, x2, n2, d, sigma0, sigma, magn, diag_only, out);
}

int gpl_dgemm_dev(gpl_ctx* h, int32_t transa, int32_t transb, int64_t m, int64_t n, int64_t k,
                  double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                  double beta, double* C, int64_t ldc, int32_t lower) {
  return dgemm(reinterpret_cast<Ctx*>(h), transa != 0, transb != 0, m, n, k, alpha, A, lda, B, ldb,
               beta, C, ldc, lower != 0);
}

int gpl_dpotrf_dev(gpl_ctx* h, int64_t n, double* A, int64_t lda, int32_t* info, double* logdet) {
  return dpotrf(reinterpret_cast<Ctx*>(h), n, A, lda, info, logdet);
}

int gpl_dtrsm_right_dev(gpl_ctx* h, int32_t trans, int64_t m, int64_t n, const double* L,
                        int64_t ldl, double* B, int64_t ldb) {
  return dtrsm_right(reinterpret_cast<Ctx*>(h), trans != 0, m, n, L, ldl, B, ldb, 0);
}

int gpl_dtrsm_right_dinv_dev(gpl_ctx* h, int32_t trans, int64_t m, int64_t n, const double* L,
                             int64_t ldl, const double* d_dinv, double* B, int64_t ldb) {
  return dtrsm_right_with(reinterpret_cast<Ctx*>(h), const_cast<double*>(d_dinv), trans != 0, m, n, L,
                          ldl, B, ldb);
}

int gpl_rowsumsq_dev(gpl_ctx* h, int64_t m, int64_t n, const double* A, int64_t lda, double* out) {
  return rowsumsq(reinterpret_cast<Ctx*>(h), m, n, A, lda, out);
}

int gpl_dpotri_dev(gpl_ctx* h, int64_t n, double* A, int64_t lda) {
  return dpotri(reinterpret_cast<Ctx*>(h), n, A, lda);
}

int gpl_dpotrf_get_dinv_dev(gpl_ctx* h, int64_t n, double* d_out) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  const int64_t nblk = (n + NB - 1) / NB;
  if (nblk > c->dinv_blocks || !d_out) {
    set_error("gpl_dpotrf_get_dinv_dev: no factorisation of that size cached");
    return GPL_ERR_BAD_ARG;
  }
  GPL_CUDA(cudaMemcpyAsync(d_out, c->dinv, sizeof(double) * nblk * NB * NB, cudaMemcpyDeviceToDevice,
                           c->stream));
  return GPL_OK;
}

int gpl_dtrsv_dev(gpl_ctx* h, int32_t trans, int64_t n, const double* L, int64_t ldl,
                  const double* d_dinv, double* b, double* x) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  return trsv_lower(c, d_dinv ? d_dinv : c->dinv, trans != 0, n, L, ldl, b, x);
}

int gpl_dgemv_dev(gpl_ctx* h, int32_t trans, int64_t m, int64_t n, const double* A, int64_t lda,
                  const double* x, double* y) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  return trans ? gemv_t(c, m, n, A, lda, x, y) : gemv_n(c, m, n, A, lda, x, y);
}

int gpl_scale_block_dev(gpl_ctx* h, int64_t m, int64_t n, double* A, int64_t lda, const double* r,
                        const double* cs, double alpha, double diag_add, int64_t diag_offset) {
  return scale_block(reinterpret_cast<Ctx*>(h), m, n, A, lda, r, cs, alpha, diag_add, diag_offset);
}

int gpl_lik_pseudodata_la_dev(gpl_ctx* h, const gpl_lik* lik, int64_t n, const double* f,
                              const double* y, const double* trials, const double* offset,
                              double* z, double* V) {
  if (!lik || n < 0 || !f || !y || !z || !V) {
    set_error("gpl_lik_pseudodata_la_dev: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  LikDev ld{lik->kind, lik->link, lik->param, trials, offset};
  return lik_pseudodata_la(reinterpret_cast<Ctx*>(h), ld, n, f, y, z, V);
}

int gpl_lik_loglik_dev(gpl_ctx* h, const gpl_lik* lik, int64_t n, const double* f, const double* y,
                       const double* trials, const double* offset, double* loglik_host) {
  if (!lik || n < 0 || !f || !y || !loglik_host) {
    set_error("gpl_lik_loglik_dev: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  Ctx* c = reinterpret_cast<Ctx*>(h);
  LikDev ld{lik->kind, lik->link, lik->param, trials, offset};
  GPL_TRY(lik_loglik_sum(c, ld, n, f, y, 0));
  return fetch_scalars(c, 1, loglik_host);
}

int gpl_lik_ep_sites_dev(gpl_ctx* h, const gpl_lik* lik, int64_t n, const double* fmean,
                         const double* fvar, const double* z_old, const double* P_old, const double* y,
                         const double* trials, const double* offset, double damping, int32_t quad_order,
                         const double* quad_x, const double* quad_w, double* z_new, double* V_new,
                         double* P_new, double* cavity_mean, double* cavity_var, double* logZ,
                         int32_t* flags) {
  if (!lik || n < 0 || !fmean || !fvar || !z_old || !P_old || !y || !quad_x || !quad_w || !z_new ||
      !V_new || !P_new || !cavity_mean || !cavity_var || !logZ || !flags) {
    set_error("gpl_lik_ep_sites_dev: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  Ctx* c = reinterpret_cast<Ctx*>(h);
  LikDev ld{lik->kind, lik->link, lik->param, trials, offset};
  EpSiteArgs ea;
  ea.fmean = fmean; ea.fvar = fvar; ea.z_old = z_old; ea.P_old = P_old; ea.y = y;
  ea.z_new = z_new; ea.V_new = V_new; ea.P_new = P_new;
  ea.cav_mean = cavity_mean; ea.cav_var = cavity_var; ea.logZ = logZ;
  ea.damping = damping;
  ea.quad_order = quad_order;
  ea.quad_x = quad_x; ea.quad_w = quad_w;
  DevBuf fl;
  GPL_TRY(fl.alloc(c, sizeof(int) * 4));
  GPL_TRY(lik_ep_sites(c, ld, n, ea, static_cast<int*>(fl.p)));
  int hf = 0;
  GPL_CUDA(cudaMemcpyAsync(&hf, fl.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GPL_CUDA(cudaStreamSynchronize(c->stream));
  *flags = hf;
  return GPL_OK;
}

int gpl_colsumsq_dev(gpl_ctx* h, int64_t m, int64_t n, const double* A, int64_t lda, double* out) {
  if (m < 0 || n < 0 || !A || !out) {
    set_error("gpl_colsumsq_dev: bad arguments");
    return GPL_ERR_BAD_ARG;
  }
  return colsumsq(reinterpret_cast<Ctx*>(h), m, n, A, lda, out);
}

}  // extern "C"
