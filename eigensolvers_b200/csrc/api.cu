// api.cu — C ABI of libcudavec: context, operator, BLAS-1 / tall-skinny / SpMV entry points.
// Host-side glue only; kernels live in kernels_vec.cuh / kernels_spmv.cuh.
#include <stdarg.h>
#include <unordered_map>
#include <vector>
#include "internal.h"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void cv_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *cv_last_error(void) { return g_err; }
extern "C" int cv_abi_version(void) { return CV_ABI_VERSION; }

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
extern "C" size_t cv_ctx_scratch_bytes(void) {
  return CV_N_COUNTERS * sizeof(unsigned) + (CV_N_SCALARS + CV_N_PARTIALS) * sizeof(double);
}

extern "C" int cv_ctx_create(int device, void *scratch_dev, size_t scratch_bytes, cv_ctx **out) {
  CV_REQUIRE(out && scratch_dev, "cv_ctx_create: null argument");
  CV_REQUIRE(scratch_bytes >= cv_ctx_scratch_bytes(), "cv_ctx_create: scratch too small (%zu < %zu)",
             scratch_bytes, cv_ctx_scratch_bytes());
  CV_REQUIRE(((uintptr_t)scratch_dev & 255) == 0, "cv_ctx_create: scratch must be 256-byte aligned");
  CV_CUDA(cudaSetDevice(device));
  cv_ctx *c = new cv_ctx();
  c->device = device;
  CV_CUDA(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
  char *p = static_cast<char *>(scratch_dev);
  c->counters = reinterpret_cast<unsigned *>(p);
  p += CV_N_COUNTERS * sizeof(unsigned);
  c->scalars = reinterpret_cast<double *>(p);
  p += CV_N_SCALARS * sizeof(double);
  c->partials = reinterpret_cast<double *>(p);
  CV_CUDA(cudaMemset(scratch_dev, 0, CV_N_COUNTERS * sizeof(unsigned) + CV_N_SCALARS * sizeof(double)));
  CV_CUDA(cudaMallocHost(&c->mailbox, (CV_N_SCALARS + 8) * sizeof(double)));
  c->host_flag = reinterpret_cast<unsigned long long *>(c->mailbox + CV_N_SCALARS);
  *c->host_flag = 0ull;
  c->host_seq = 0ull;
  c->prepushed_x = nullptr;
  c->prepushed_op = nullptr;
  c->push_early = getenv("EIGB200_PUSH_EARLY") ? atoi(getenv("EIGB200_PUSH_EARLY")) != 0 : true;
  c->slab_mode = getenv("EIGB200_SLAB") ? atoi(getenv("EIGB200_SLAB")) : 1;
  c->snake = getenv("EIGB200_SNAKE") ? atoi(getenv("EIGB200_SNAKE")) : 1;
  c->precond_dinv = nullptr;
  c->precond_z = c->precond_t = nullptr;
  c->launches = 0;
  c->prof = nullptr;
  c->reorth_eta = getenv("EIGB200_REORTH_ETA") ? atof(getenv("EIGB200_REORTH_ETA")) : 0.1;
  c->defer_reduce = false;
  c->comm = nullptr;
  c->peer = nullptr;
  c->rank = 0;
  c->world = 1;
  *out = c;
  return CV_OK;
}

extern "C" int cv_ctx_destroy(cv_ctx *ctx) {
  if (!ctx) return CV_OK;
  if (ctx->comm || ctx->peer) cv_comm_finalize(ctx);
  if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
  if (ctx->prof) {
    if (ctx->prof->created)
      for (int i = 0; i < CV_PROF_POOL; ++i) {
        cudaEventDestroy(ctx->prof->start[i]);
        cudaEventDestroy(ctx->prof->stop[i]);
      }
    delete ctx->prof;
  }
  delete ctx;
  return CV_OK;
}

extern "C" int cv_ctx_launch_count(cv_ctx *ctx, uint64_t *count) {
  CV_REQUIRE(ctx && count, "cv_ctx_launch_count: null argument");
  *count = ctx->launches;
  return CV_OK;
}

extern "C" int cv_ctx_set_recycle(cv_ctx *ctx, int enable) {
  CV_REQUIRE(ctx, "cv_ctx_set_recycle: null context");
  if (ctx->recycle.enabled != (enable != 0)) ctx->recycle.valid = false;
  ctx->recycle.enabled = enable != 0;
  return CV_OK;
}

extern "C" int cv_ctx_set_reorth_eta(cv_ctx *ctx, double eta) {
  CV_REQUIRE(ctx && eta >= 0.0, "cv_ctx_set_reorth_eta: bad argument");
  ctx->reorth_eta = eta;
  return CV_OK;
}

extern "C" int cv_ctx_set_option(cv_ctx *ctx, const char *name, double value) {
  CV_REQUIRE(ctx && name, "cv_ctx_set_option: null argument");
  if (!strcmp(name, "reorth_eta")) {
    CV_REQUIRE(value >= 0.0, "cv_ctx_set_option: reorth_eta must be >= 0");
    ctx->reorth_eta = value;
  } else if (!strcmp(name, "slab_mode")) {
    ctx->slab_mode = value != 0.0;
  } else if (!strcmp(name, "snake")) {
    ctx->snake = value != 0.0;
  } else if (!strcmp(name, "push_early")) {
    ctx->push_early = value != 0.0;
  } else {
    cv_set_error("cv_ctx_set_option: unknown option '%s'", name);
    return CV_ERR_ARG;
  }
  return CV_OK;
}

extern "C" int cv_ctx_trace_read(cv_ctx *ctx, double *out16, int reset) {
  CV_REQUIRE(ctx && out16, "cv_ctx_trace_read: null argument");
  CV_CUDA(cudaDeviceSynchronize());
  CV_CUDA(cudaMemcpy(out16, ctx->scalars + CV_S_TRACE, 16 * sizeof(double), cudaMemcpyDeviceToHost));
  if (reset) CV_CUDA(cudaMemset(ctx->scalars + CV_S_TRACE, 0, 16 * sizeof(double)));
  return CV_OK;
}

extern "C" int cv_ctx_sm_count(cv_ctx *ctx, int *sms) {
  CV_REQUIRE(ctx && sms, "cv_ctx_sm_count: null argument");
  *sms = ctx->sms;
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// kernel-class timing (CUDA events on the launching stream)
// ------------------------------------------------------------------------------------------
static void prof_drain(cv_prof_state *p) {
  for (int i = 0; i < p->used; ++i) {
    cudaEventSynchronize(p->stop[i]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->start[i], p->stop[i]) == cudaSuccess) {
      p->ms[p->cls[i]] += ms;
      p->count[p->cls[i]]++;
    }
  }
  p->used = 0;
}

cv_prof_scope::cv_prof_scope(cv_ctx *c, int cls, cudaStream_t s, double alg_bytes) : ctx(c), st(s), slot(-1) {
  cv_prof_state *p = c->prof;
  if (!p || !p->enabled) return;
  if (p->used == CV_PROF_POOL) prof_drain(p);
  slot = p->used++;
  p->cls[slot] = cls;
  p->bytes[cls] += alg_bytes;
  cudaEventRecord(p->start[slot], st);
}

cv_prof_scope::~cv_prof_scope() {
  if (slot >= 0) cudaEventRecord(ctx->prof->stop[slot], st);
}

extern "C" int cv_ctx_profile(cv_ctx *ctx, int enable) {
  CV_REQUIRE(ctx, "cv_ctx_profile: null context");
  if (!ctx->prof) ctx->prof = new cv_prof_state();
  cv_prof_state *p = ctx->prof;
  if (enable && !p->created) {
    for (int i = 0; i < CV_PROF_POOL; ++i) {
      CV_CUDA(cudaEventCreate(&p->start[i]));
      CV_CUDA(cudaEventCreate(&p->stop[i]));
    }
    p->created = true;
  }
  if (!enable && p->enabled) prof_drain(p);
  p->enabled = enable != 0;
  return CV_OK;
}

extern "C" int cv_ctx_profile_read(cv_ctx *ctx, int n_classes, double *ms, uint64_t *count, double *bytes) {
  CV_REQUIRE(ctx && ms && count && bytes, "cv_ctx_profile_read: null argument");
  CV_REQUIRE(n_classes >= 1 && n_classes <= CV_PROF_CLASSES, "cv_ctx_profile_read: n_classes=%d outside 1..%d",
             n_classes, CV_PROF_CLASSES);
  for (int i = 0; i < n_classes; ++i) ms[i] = 0.0, count[i] = 0, bytes[i] = 0.0;
  cv_prof_state *p = ctx->prof;
  if (!p) return CV_OK;
  prof_drain(p);
  for (int i = 0; i < CV_PROF_CLASSES; ++i) {
    if (i < n_classes) {
      ms[i] = p->ms[i];
      count[i] = p->count[i];
      bytes[i] = p->bytes[i];
    }
    p->ms[i] = 0.0;
    p->count[i] = 0;
    p->bytes[i] = 0.0;
  }
  return CV_OK;
}

int cv_fetch_scalars(cv_ctx *ctx, int offset, int count, cudaStream_t st) {
  CV_CUDA(cudaMemcpyAsync(ctx->mailbox + offset, ctx->scalars + offset, sizeof(double) * count,
                          cudaMemcpyDeviceToHost, st));
  if (ctx->peer)
    CV_CUDA(cudaMemcpyAsync(ctx->mailbox + CV_S_ERR, ctx->scalars + CV_S_ERR, sizeof(double),
                            cudaMemcpyDeviceToHost, st));
  CV_CUDA(cudaStreamSynchronize(st));
  if (ctx->peer && ctx->mailbox[CV_S_ERR] != 0.0) {
    cv_set_error("peer-memory collective timed out waiting for another rank");
    return CV_ERR_COMM;
  }
  return CV_OK;
}

int cv_check_launch(cv_ctx *ctx, const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cv_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return CV_ERR_CUDA;
  }
  ctx->launches++;
  return CV_OK;
}

// Grid size of a persistent grid-stride launch: enough CTAs for the work, at most ONE resident
// wave of this particular kernel (SMs x CTAs/SM from the occupancy calculator), so that no
// partial second wave leaves SMs idle at the tail.
int cv_occ_grid(cv_ctx *ctx, const void *kernel, int64_t work_items, int items_per_cta) {
  static std::unordered_map<const void *, int> cache;
  int occ;
  auto it = cache.find(kernel);
  if (it == cache.end()) {
    occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CV_BLOCK, 0) != cudaSuccess || occ < 1) {
      cudaGetLastError();
      occ = 1;
    }
    cache[kernel] = occ;
  } else {
    occ = it->second;
  }
  int64_t need = (work_items + items_per_cta - 1) / items_per_cta;
  int64_t cap = (int64_t)ctx->sms * occ;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}
#define CV_KGRID(kf, work) cv_occ_grid(ctx, (const void *)(kf), (work), CV_BLOCK)

// ------------------------------------------------------------------------------------------
// BLAS-1 (internal launchers are shared with solvers.cu through internal.h)
// ------------------------------------------------------------------------------------------
static inline int vecW(int cplx, std::initializer_list<const void *> ptrs) {
  if (cplx) return 1;
  for (const void *p : ptrs)
    if (p && ((uintptr_t)p & 15)) return 1;
  return 2;
}

extern "C" int cv_copy(cv_ctx *ctx, int64_t n, int cplx, const void *x, void *y, void *stream) {
  CV_REQUIRE(ctx && x && y && n >= 0, "cv_copy: bad argument");
  if (n == 0 || x == y) return CV_OK;
  CV_CUDA(cudaMemcpyAsync(y, x, (size_t)n * (cplx ? 16 : 8), cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  return CV_OK;
}

extern "C" int cv_scal(cv_ctx *ctx, int64_t n, int x_cplx, int y_cplx, double a_re, double a_im,
                       const void *x, void *y, void *stream) {
  CV_REQUIRE(ctx && x && y && n >= 0, "cv_scal: bad argument");
  CV_REQUIRE(y_cplx || !x_cplx, "cv_scal: complex input needs complex output");
  CV_REQUIRE(y_cplx || a_im == 0.0, "cv_scal: complex factor needs complex output");
  if (n == 0) return CV_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!x_cplx && !y_cplx) {
    int W = vecW(0, {x, y});
    if (W == 2) {
      auto kf = k_scal_rr<2>;
      kf<<<CV_KGRID(kf, n / 2 + 1), CV_BLOCK, 0, st>>>(n, a_re, (const double *)x, (double *)y);
    } else {
      auto kf = k_scal_rr<1>;
      kf<<<CV_KGRID(kf, n), CV_BLOCK, 0, st>>>(n, a_re, (const double *)x, (double *)y);
    }
  } else {
    cplx a = make_cplx(a_re, a_im);
    if (x_cplx) {
      auto kf = k_scal<cplx, cplx, cplx>;
      kf<<<CV_KGRID(kf, n), CV_BLOCK, 0, st>>>(n, a, (const cplx *)x, (cplx *)y);
    } else {
      auto kf = k_scal<double, cplx, cplx>;
      kf<<<CV_KGRID(kf, n), CV_BLOCK, 0, st>>>(n, a, (const double *)x, (cplx *)y);
    }
  }
  return cv_check_launch(ctx, "scal");
}

extern "C" int cv_real(cv_ctx *ctx, int64_t n, const void *x, double *y, void *stream) {
  CV_REQUIRE(ctx && x && y && n >= 0, "cv_real: bad argument");
  if (n == 0) return CV_OK;
  k_real<<<CV_KGRID(k_real, n), CV_BLOCK, 0, (cudaStream_t)stream>>>(n, (const cplx *)x, y);
  return cv_check_launch(ctx, "real");
}

extern "C" int cv_conj(cv_ctx *ctx, int64_t n, const void *x, void *y, void *stream) {
  CV_REQUIRE(ctx && x && y && n >= 0, "cv_conj: bad argument");
  if (n == 0) return CV_OK;
  k_conj<<<CV_KGRID(k_conj, n), CV_BLOCK, 0, (cudaStream_t)stream>>>(n, (const cplx *)x, (cplx *)y);
  return cv_check_launch(ctx, "conj");
}

// device-side dot into ctx->scalars[slot .. slot+NRED)
int cv_dot_dev(cv_ctx *ctx, int64_t n, int cplx_, int conj, const void *x, const void *y, int slot,
               cudaStream_t st) {
  int W = vecW(cplx_, {x, y});
  double *out = ctx->scalars + slot;
  cv_prof_scope prof(ctx, 3, st);
#define LAUNCH_DOT(T, WW, CJ)                                                                          \
  do {                                                                                                 \
    auto kf = k_dot<T, WW, CJ>;                                                                        \
    kf<<<CV_KGRID(kf, n / WW + 1), CV_BLOCK, 0, st>>>(n, (const T *)x, (const T *)y, ctx->partials,     \
                                                      ctx->counters, out);                            \
  } while (0)
  if (cplx_) {
    if (conj)
      LAUNCH_DOT(cplx, 1, true);
    else
      LAUNCH_DOT(cplx, 1, false);
  } else if (W == 2) {
    LAUNCH_DOT(double, 2, false);
  } else {
    LAUNCH_DOT(double, 1, false);
  }
#undef LAUNCH_DOT
  CV_TRY(cv_check_launch(ctx, "dot"));
  return cv_reduce_ranks(ctx, slot, cplx_ ? 2 : 1, st);
}

int cv_nrm2sq_dev(cv_ctx *ctx, int64_t n, int cplx_, const void *x, int slot, cudaStream_t st) {
  int W = vecW(cplx_, {x});
  double *out = ctx->scalars + slot;
  cv_prof_scope prof(ctx, 3, st);
#define LAUNCH_NRM(T, WW)                                                                            \
  do {                                                                                               \
    auto kf = k_nrm2sq<T, WW>;                                                                       \
    kf<<<CV_KGRID(kf, n / WW + 1), CV_BLOCK, 0, st>>>(n, (const T *)x, ctx->partials, ctx->counters, out); \
  } while (0)
  if (cplx_)
    LAUNCH_NRM(cplx, 1);
  else if (W == 2)
    LAUNCH_NRM(double, 2);
  else
    LAUNCH_NRM(double, 1);
#undef LAUNCH_NRM
  CV_TRY(cv_check_launch(ctx, "nrm2sq"));
  return cv_reduce_ranks(ctx, slot, 1, st);
}

// x *= 1/sqrt(scalars[slot]); mode 1 keeps x when the factor is not finite
int cv_scale_dev(cv_ctx *ctx, int64_t n, int cplx_, void *x, int slot, int mode, cudaStream_t st) {
  int W = vecW(cplx_, {x});
  const double *s = ctx->scalars + slot;
  cv_prof_scope prof(ctx, 3, st);
#define LAUNCH_SC(T, WW, MD)                                                   \
  do {                                                                         \
    auto kf = k_scale_dev<T, WW, MD>;                                          \
    kf<<<CV_KGRID(kf, n / WW + 1), CV_BLOCK, 0, st>>>(n, (T *)x, s);            \
  } while (0)
  if (cplx_) {
    if (mode) LAUNCH_SC(cplx, 1, 1); else LAUNCH_SC(cplx, 1, 0);
  } else if (W == 2) {
    if (mode) LAUNCH_SC(double, 2, 1); else LAUNCH_SC(double, 2, 0);
  } else {
    if (mode) LAUNCH_SC(double, 1, 1); else LAUNCH_SC(double, 1, 0);
  }
#undef LAUNCH_SC
  return cv_check_launch(ctx, "scale_dev");
}

extern "C" int cv_dot(cv_ctx *ctx, int64_t n, int cplx_, int conj, const void *x, const void *y,
                      double *out2_host, void *stream) {
  CV_REQUIRE(ctx && x && y && out2_host && n >= 0, "cv_dot: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  out2_host[0] = out2_host[1] = 0.0;
  if (n == 0 && ctx->world == 1) return CV_OK;
  CV_TRY(cv_dot_dev(ctx, n, cplx_, conj, x, y, CV_S_TMP, st));
  CV_TRY(cv_fetch_scalars(ctx, CV_S_TMP, 2, st));
  out2_host[0] = ctx->mailbox[CV_S_TMP];
  out2_host[1] = cplx_ ? ctx->mailbox[CV_S_TMP + 1] : 0.0;
  return CV_OK;
}

extern "C" int cv_nrm2(cv_ctx *ctx, int64_t n, int cplx_, const void *x, double *out_host,
                       void *stream) {
  CV_REQUIRE(ctx && x && out_host && n >= 0, "cv_nrm2: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  *out_host = 0.0;
  if (n == 0 && ctx->world == 1) return CV_OK;
  CV_TRY(cv_nrm2sq_dev(ctx, n, cplx_, x, CV_S_TMP, st));
  CV_TRY(cv_fetch_scalars(ctx, CV_S_TMP, 1, st));
  *out_host = sqrt(ctx->mailbox[CV_S_TMP]);
  return CV_OK;
}

extern "C" int cv_normalize(cv_ctx *ctx, int64_t n, int cplx_, void *x, double *norm_host,
                            void *stream) {
  CV_REQUIRE(ctx && x && n >= 0, "cv_normalize: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 && ctx->world == 1) {
    if (norm_host) *norm_host = 0.0;
    return CV_OK;
  }
  CV_TRY(cv_nrm2sq_dev(ctx, n, cplx_, x, CV_S_TMP, st));
  CV_TRY(cv_scale_dev(ctx, n, cplx_, x, CV_S_TMP, 0, st));
  if (norm_host) {
    CV_TRY(cv_fetch_scalars(ctx, CV_S_TMP, 1, st));
    *norm_host = sqrt(ctx->mailbox[CV_S_TMP]);
  }
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// linear combinations
// ------------------------------------------------------------------------------------------
template <typename TV, typename TC, typename TY, int W, bool NORM>
static int launch_lincomb_nc(cv_ctx *ctx, const LcParams &p, int ncol, int grid, double *out_norm,
                             cudaStream_t st) {
  (void)grid;
  // algorithmic bytes: (m + ncol) vectors of n elements
  cv_prof_scope prof(ctx, 5, st, (double)(p.m + ncol) * (double)p.n * (double)sizeof(TY));
#define LC(NC)                                                                                     \
  do {                                                                                             \
    auto kf = k_lincomb<TV, TC, TY, W, NC, NORM>;                                                  \
    kf<<<CV_KGRID(kf, p.n / W + 1), CV_BLOCK, 0, st>>>(p, ctx->partials, ctx->counters, out_norm);  \
  } while (0)
  switch (ncol) {
    case 1: LC(1); break;
    case 2: LC(2); break;
    case 3: LC(3); break;
    case 4: LC(4); break;
    case 5: LC(5); break;
    case 6: LC(6); break;
    case 7: LC(7); break;
    case 8: LC(8); break;
    default:
      cv_set_error("lincomb: internal chunk of %d columns", ncol);
      return CV_ERR_ARG;
  }
#undef LC
  return cv_check_launch(ctx, "lincomb");
}

// One launch: ncol <= 4 outputs, m*ncol*(c_cplx?2:1) <= CV_MAX_COEF.  coef row-major m x ldc.
int cv_lincomb_launch(cv_ctx *ctx, int64_t n, int v_cplx, int c_cplx, int m, const void *const *v,
                      int ncol, const double *coef, int ldc, int col0, void *const *y, int norm_slot,
                      cudaStream_t st) {
  LcParams p;
  p.m = m;
  p.ncol = ncol;
  p.n = n;
  const int cs = c_cplx ? 2 : 1;
  for (int j = 0; j < m; ++j) {
    p.v[j] = v[j];
    for (int k = 0; k < ncol; ++k)
      for (int c = 0; c < cs; ++c) p.coef[(j * ncol + k) * cs + c] = coef[((size_t)j * ldc + col0 + k) * cs + c];
  }
  int W = v_cplx || c_cplx ? 1 : 2;
  for (int j = 0; j < m && W == 2; ++j)
    if ((uintptr_t)v[j] & 15) W = 1;
  for (int k = 0; k < ncol; ++k) {
    p.y[k] = y[k];
    if ((uintptr_t)y[k] & 15) W = 1;
  }
  int grid = cv_grid_for(ctx, (n + W - 1) / W, CV_BLOCK);
  double *out_norm = norm_slot >= 0 ? ctx->scalars + norm_slot : nullptr;
  int rc;
  if (!v_cplx && !c_cplx) {
    if (W == 2)
      rc = norm_slot >= 0 ? launch_lincomb_nc<double, double, double, 2, true>(ctx, p, ncol, grid, out_norm, st)
                          : launch_lincomb_nc<double, double, double, 2, false>(ctx, p, ncol, grid, out_norm, st);
    else
      rc = norm_slot >= 0 ? launch_lincomb_nc<double, double, double, 1, true>(ctx, p, ncol, grid, out_norm, st)
                          : launch_lincomb_nc<double, double, double, 1, false>(ctx, p, ncol, grid, out_norm, st);
  } else if (v_cplx && c_cplx) {
    rc = norm_slot >= 0 ? launch_lincomb_nc<cplx, cplx, cplx, 1, true>(ctx, p, ncol, grid, out_norm, st)
                        : launch_lincomb_nc<cplx, cplx, cplx, 1, false>(ctx, p, ncol, grid, out_norm, st);
  } else if (v_cplx && !c_cplx) {
    rc = norm_slot >= 0 ? launch_lincomb_nc<cplx, double, cplx, 1, true>(ctx, p, ncol, grid, out_norm, st)
                        : launch_lincomb_nc<cplx, double, cplx, 1, false>(ctx, p, ncol, grid, out_norm, st);
  } else {
    rc = norm_slot >= 0 ? launch_lincomb_nc<double, cplx, cplx, 1, true>(ctx, p, ncol, grid, out_norm, st)
                        : launch_lincomb_nc<double, cplx, cplx, 1, false>(ctx, p, ncol, grid, out_norm, st);
  }
  CV_TRY(rc);
  if (norm_slot >= 0) CV_TRY(cv_reduce_ranks(ctx, norm_slot, ncol, st));
  return CV_OK;
}

extern "C" int cv_lincomb(cv_ctx *ctx, int64_t n, int v_cplx, int c_cplx, int m,
                          const void *const *v_ptrs, int ncol, const double *coef_host,
                          void *const *y_ptrs, void *stream) {
  CV_REQUIRE(ctx && v_ptrs && coef_host && y_ptrs, "cv_lincomb: null argument");
  CV_REQUIRE(m >= 1 && ncol >= 1 && n >= 0, "cv_lincomb: bad shape");
  if (n == 0) return CV_OK;
  const int cs = c_cplx ? 2 : 1;
  // Any number of inputs: they are consumed in chunks of <= MCH; from the second chunk on the
  // partial results re-enter as extra inputs with unit coefficients (a thread reads every input of
  // an element before it writes the outputs of that element, so in-place accumulation is safe).
  const int MCH = 96;
  int chunk = c_cplx || v_cplx ? 4 : 8;  // outputs per pass over the inputs
  const int m_launch_max = (m <= MCH ? m : MCH + chunk);
  while (chunk > 1 && m_launch_max * chunk * cs > CV_MAX_COEF) chunk >>= 1;
  CV_REQUIRE((m <= MCH ? m : MCH + chunk) * chunk * cs <= CV_MAX_COEF, "cv_lincomb: too many coefficients");
  CV_REQUIRE(m <= MCH || v_cplx || !c_cplx, "cv_lincomb: complex coefficients on more than %d real inputs", MCH);
  std::vector<const void *> vp;
  std::vector<double> cf;
  for (int c0 = 0; c0 < ncol; c0 += chunk) {
    const int nc = ncol - c0 < chunk ? ncol - c0 : chunk;
    for (int j0 = 0; j0 < m; j0 += MCH) {
      const int mj = m - j0 < MCH ? m - j0 : MCH;
      const int extra = j0 > 0 ? nc : 0;
      vp.assign(mj + extra, nullptr);
      cf.assign((size_t)(mj + extra) * nc * cs, 0.0);
      for (int j = 0; j < mj; ++j) {
        vp[j] = v_ptrs[j0 + j];
        for (int k = 0; k < nc; ++k)
          for (int c = 0; c < cs; ++c) cf[((size_t)j * nc + k) * cs + c] = coef_host[((size_t)(j0 + j) * ncol + c0 + k) * cs + c];
      }
      for (int k = 0; k < extra; ++k) {
        vp[mj + k] = y_ptrs[c0 + k];
        cf[((size_t)(mj + k) * nc + k) * cs] = 1.0;
      }
      // partial results have the OUTPUT type: complex outputs re-enter as complex inputs
      const int in_cplx = v_cplx;
      CV_TRY(cv_lincomb_launch(ctx, n, in_cplx, c_cplx, mj + extra, vp.data(), nc, cf.data(), nc, 0, y_ptrs + c0,
                               -1, (cudaStream_t)stream));
    }
  }
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// tall-skinny product into ctx->scalars[slot ...): layout ((i*b + k)*NRED + c)
// ------------------------------------------------------------------------------------------
template <typename T, int W, bool CONJ>
static int launch_tsdot(cv_ctx *ctx, const TsParams &p, int slot, const double *gate, cudaStream_t st) {
  const double eta2 = ctx->reorth_eta * ctx->reorth_eta;
  constexpr int NR = Num<T>::NRED;
  const int64_t np = (p.n + W - 1) / W;
  double *out = ctx->scalars + slot;
  // MI vectors per CTA slab: 16 for a single right-hand side, 8 for 2, 4 for 3..4
  int MI = p.b == 1 ? 16 : (p.b == 2 ? 8 : 4);
  int ny = (p.m + MI - 1) / MI;
  int64_t per_slab = (int64_t)MI * p.b * NR;
  {
    // algorithmic bytes: V once, W once per slab of MI vectors
    cv_prof_scope prof(ctx, 1, st, (double)(p.m + p.b * ny) * (double)p.n * (double)sizeof(T));
    // one resident wave in total: the y-slabs share the SMs
#define TSD(MI_, B_)                                                                               \
  do {                                                                                             \
    auto kf = k_tsdot<T, W, CONJ, MI_, B_>;                                                        \
    int wave = CV_KGRID(kf, (int64_t)1 << 40);                                                     \
    int gx = wave / ny;                                                                            \
    if (gx < 1) gx = 1;                                                                            \
    int64_t need = (np + CV_BLOCK - 1) / CV_BLOCK;                                                 \
    if (gx > need) gx = (int)need;                                                                 \
    while ((int64_t)gx * per_slab * ny > (int64_t)CV_N_PARTIALS && gx > 1) gx >>= 1;               \
    dim3 grid(gx, ny);                                                                             \
    kf<<<grid, CV_BLOCK, 0, st>>>(p, gate, eta2, ctx->partials, ctx->counters, out);               \
  } while (0)
    switch (p.b) {
      case 1: TSD(16, 1); break;
      case 2: TSD(8, 2); break;
      case 3: TSD(4, 3); break;
      case 4: TSD(4, 4); break;
      default:
        cv_set_error("tsdot: b=%d outside 1..4", p.b);
        return CV_ERR_ARG;
    }
#undef TSD
  }
  CV_TRY(cv_check_launch(ctx, "tsdot"));
  return cv_reduce_ranks(ctx, slot, p.m * p.b * NR, st);
}

int cv_tsdot_dev(cv_ctx *ctx, int64_t n, int cplx_, int conj, int m, const void *const *v, int b,
                 const void *const *w, int slot, cudaStream_t st, int gate_slot) {
  const double *gate = gate_slot >= 0 ? ctx->scalars + gate_slot : nullptr;
  CV_REQUIRE(m >= 1 && m <= CV_MAX_PTRS && b >= 1 && b <= 4, "tsdot: m=%d b=%d out of range", m, b);
  CV_REQUIRE(m * b * (cplx_ ? 2 : 1) <= CV_MAX_RED, "tsdot: too many values");
  CV_REQUIRE(CV_MAX_PTRS / 4 <= (int)CV_N_COUNTERS, "tsdot: counters");
  TsParams p;
  p.m = m;
  p.b = b;
  p.n = n;
  int W = cplx_ ? 1 : 2;
  for (int i = 0; i < m; ++i) {
    p.v[i] = v[i];
    if ((uintptr_t)v[i] & 15) W = 1;
  }
  for (int k = 0; k < b; ++k) {
    p.w[k] = w[k];
    if ((uintptr_t)w[k] & 15) W = 1;
  }
  if (cplx_) return conj ? launch_tsdot<cplx, 1, true>(ctx, p, slot, gate, st) : launch_tsdot<cplx, 1, false>(ctx, p, slot, gate, st);
  // real: conjugation is the identity
  return W == 2 ? launch_tsdot<double, 2, false>(ctx, p, slot, gate, st) : launch_tsdot<double, 1, false>(ctx, p, slot, gate, st);
}

extern "C" int cv_tsdot(cv_ctx *ctx, int64_t n, int cplx_, int conj, int m, const void *const *v_ptrs,
                        int b, const void *const *w_ptrs, double *out_host, void *stream) {
  CV_REQUIRE(ctx && v_ptrs && w_ptrs && out_host && n >= 0, "cv_tsdot: bad argument");
  CV_REQUIRE(m >= 1 && b >= 1, "cv_tsdot: m=%d b=%d out of range", m, b);
  cudaStream_t st = (cudaStream_t)stream;
  const int nr = cplx_ ? 2 : 1;
  // any m, any b: V in chunks of CV_MAX_PTRS vectors, right-hand sides in chunks of 4; results are
  // assembled as out[(i*b + k)*nr + c]
  for (int i0 = 0; i0 < m; i0 += CV_MAX_PTRS) {
    const int mm = m - i0 < CV_MAX_PTRS ? m - i0 : CV_MAX_PTRS;
    for (int k0 = 0; k0 < b; k0 += 4) {
      int bb = b - k0 < 4 ? b - k0 : 4;
      CV_TRY(cv_tsdot_dev(ctx, n, cplx_, conj, mm, v_ptrs + i0, bb, w_ptrs + k0, CV_S_TS, st));
      CV_TRY(cv_fetch_scalars(ctx, CV_S_TS, mm * bb * nr, st));
      for (int i = 0; i < mm; ++i)
        for (int k = 0; k < bb; ++k)
          for (int c = 0; c < nr; ++c)
            out_host[((size_t)(i0 + i) * b + k0 + k) * nr + c] = ctx->mailbox[CV_S_TS + (i * bb + k) * nr + c];
    }
  }
  return CV_OK;
}

// w -= V h  (h = ctx->scalars[h_slot ..)), optional |w|^2 into norm_slot
int cv_tsupdate_dev(cv_ctx *ctx, int64_t n, int cplx_, int m, const void *const *v, int h_slot,
                    void *w, int norm_slot, cudaStream_t st, int gate_slot) {
  const double *gate = gate_slot >= 0 ? ctx->scalars + gate_slot : nullptr;
  CV_REQUIRE(m >= 1 && m <= CV_MAX_PTRS, "tsupdate: m=%d out of range", m);
  TsParams p;
  p.m = m;
  p.b = 1;
  p.n = n;
  int W = cplx_ ? 1 : 2;
  for (int i = 0; i < m; ++i) {
    p.v[i] = v[i];
    if ((uintptr_t)v[i] & 15) W = 1;
  }
  if ((uintptr_t)w & 15) W = 1;
  const double *h = ctx->scalars + h_slot;
  double *on = norm_slot >= 0 ? ctx->scalars + norm_slot : nullptr;
  size_t sh = sizeof(double) * m * (cplx_ ? 2 : 1);
  const double eta2 = ctx->reorth_eta * ctx->reorth_eta;
  {
    cv_prof_scope prof(ctx, 2, st, (double)(m + 2) * (double)n * (cplx_ ? 16.0 : 8.0));
#define TSU(T, WW, NM)                                                                             \
  do {                                                                                             \
    auto kf = k_tsupdate<T, WW, NM>;                                                               \
    kf<<<CV_KGRID(kf, n / WW + 1), CV_BLOCK, sh, st>>>(p, h, gate, eta2, ctx->rank == 0 ? 1.0 : 0.0, \
                                                       (T *)w, ctx->partials, ctx->counters, on);  \
  } while (0)
    if (cplx_) {
      if (on) TSU(cplx, 1, true); else TSU(cplx, 1, false);
    } else if (W == 2) {
      if (on) TSU(double, 2, true); else TSU(double, 2, false);
    } else {
      if (on) TSU(double, 1, true); else TSU(double, 1, false);
    }
#undef TSU
  }
  CV_TRY(cv_check_launch(ctx, "tsupdate"));
  if (on) CV_TRY(cv_reduce_ranks(ctx, norm_slot, 1, st));
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// fused Arnoldi orthogonalisation step (kernels_orth.cuh)
// ------------------------------------------------------------------------------------------
// Poll the mailbox flag the fused kernel releases once its scalars are in host memory.
int cv_wait_mailbox(cv_ctx *ctx, unsigned long long seq, cudaStream_t st) {
  volatile unsigned long long *flag = ctx->host_flag;
  unsigned long long spins = 0;
  while (*flag < seq) {
    if ((++spins & 0xFFFFull) == 0) {
      cudaError_t e = cudaStreamQuery(st);
      if (e == cudaSuccess) {
        if (*flag >= seq) break;
        cv_set_error("fused Arnoldi step finished without publishing its results");
        return CV_ERR_CUDA;
      }
      if (e != cudaErrorNotReady) {
        cv_set_error("fused Arnoldi step failed: %s", cudaGetErrorString(e));
        return CV_ERR_CUDA;
      }
    }
  }
  __sync_synchronize();
  if (ctx->peer && ctx->mailbox[CV_S_ERR] != 0.0) {
    cv_set_error("peer-memory collective timed out waiting for another rank");
    return CV_ERR_COMM;
  }
  return CV_OK;
}

// One launch: h = basis^H w, w -= basis h (twice if needed), w /= |w|, halo push of the new w.
// Scalar slots (doubles in ctx->scalars, mirrored to the mailbox): s_flag, s_flag+1 = |w'|^2,
// s_flag+2..4 = SpMV dots, s_flag+5.. = h1, s_h2.. = h2.  *fused = false when this configuration
// has to take the separate-kernel path (NCCL transport).
int cv_orth_step_dev(cv_ctx *ctx, cv_op *op, int64_t n, int cplx_, int m, const void *const *basis, void *w,
                     int s_flag, int s_h2, int s_lag, double eta, cudaStream_t st, bool *fused) {
  *fused = false;
  static const int use_fused = getenv("EIGB200_FUSED") ? atoi(getenv("EIGB200_FUSED")) : 1;
  if (!use_fused || (ctx->world > 1 && !ctx->peer)) return CV_OK;
  CV_REQUIRE(m >= 1 && m <= CV_MAX_PTRS, "orth_step: m=%d out of range", m);
  OrthArgs a;
  a.p.m = m;
  a.p.b = 1;
  a.p.n = n;
  int W = cplx_ ? 1 : 2;
  for (int i = 0; i < m; ++i) {
    a.p.v[i] = basis[i];
    if ((uintptr_t)basis[i] & 15) W = 1;
  }
  if ((uintptr_t)w & 15) W = 1;
  a.p.w[0] = w;
  a.partials = ctx->partials;
  a.bar = ctx->counters + CV_COUNTER_BAR;
  a.scal = ctx->scalars;
  a.s_flag = s_flag;
  a.s_nrm = s_flag + 1;
  a.s_w = s_flag + 2;
  a.s_h1 = s_flag + 5;
  a.s_h2 = s_h2;
  a.s_lag = s_lag;
  a.eta2 = eta * eta;
  a.me = ctx->rank;
  a.world = ctx->world;
  a.err = ctx->scalars + CV_S_ERR;
  for (int p = 0; p < CV_MAX_WORLD; ++p) a.pp.win[p] = nullptr;
  a.push.nseg = 0;
  a.push.nflag = 0;
  a.push.seq = 0;
  a.push.ticket = ctx->counters + CV_COUNTER_PUSH;
  a.push_early = 0;
  a.slab_mode = ctx->slab_mode;
  a.snake = ctx->snake;
  if (ctx->world > 1) {
    a.pp = *cv_peer_ptrs(ctx);
    const bool dia = cv_op_banded(op);
    const bool halo = dia ? (op->lo_len > 0 || op->hi_len > 0) : (op->n_halo > 0 || (!op->send_off.empty() && op->send_off.back() > 0));
    if (halo && op->peer_halo) {
      CV_TRY(cv_peer_plan_exchange(ctx, op, cplx_ != 0, &a.push, nullptr));
      ctx->prepushed_x = w;
      ctx->prepushed_op = op;
      if (dia && ctx->push_early) a.push_early = 1;  // contiguous ranges: pushed by phase B as produced
    }
  }
  a.trace = ctx->scalars + CV_S_TRACE;
  static const int publish_late = getenv("EIGB200_PUBLISH_LATE") ? atoi(getenv("EIGB200_PUBLISH_LATE")) : 0;
  static const int coop = getenv("EIGB200_COOP") ? atoi(getenv("EIGB200_COOP")) : 1;
  a.publish_late = publish_late;
  a.host_mb = ctx->mailbox;
  a.host_flag = ctx->host_flag;
  a.host_seq = ++ctx->host_seq;
  const void *kf = cplx_ ? (const void *)k_orth_step<cplx, 1>
                         : (W == 2 ? (const void *)k_orth_step<double, 2> : (const void *)k_orth_step<double, 1>);
  const size_t sh = sizeof(double) * m * (cplx_ ? 2 : 1);
  // co-resident CTAs of this kernel with its largest dynamic shared-memory request
  static std::unordered_map<const void *, int> occ_cache;
  auto it = occ_cache.find(kf);
  if (it == occ_cache.end()) {
    int occ = 0;
    CV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf, CV_BLOCK, sizeof(double) * 2 * CV_MAX_PTRS));
    CV_REQUIRE(occ >= 1, "orth_step: kernel does not fit on an SM");
    it = occ_cache.emplace(kf, occ).first;
  }
  const int cap = it->second * ctx->sms;
  const int mi_max = cplx_ ? ORTH_MI<cplx>::value : ORTH_MI<double>::value;
  const int ny = (m + mi_max - 1) / mi_max;
  int64_t need = (n / W + CV_BLOCK - 1) / CV_BLOCK;
  if (need < ny) need = ny;
  const int grid = (int)(need < cap ? need : cap);
  CV_REQUIRE(grid >= ny, "orth_step: grid %d smaller than %d slabs", grid, ny);
  void *params[1] = {(void *)&a};
  {
    // algorithmic bytes of the first pass: basis twice (dots, update), w read twice and written once;
    // a second Gram-Schmidt pass (rare) is added by the solver once the mailbox says it ran
    cv_prof_scope prof(ctx, 1, st, (double)(2 * m + 3) * (double)n * (cplx_ ? 16.0 : 8.0));
    if (coop) {
      CV_CUDA(cudaLaunchCooperativeKernel(kf, dim3(grid), dim3(CV_BLOCK), params, sh, st));
    } else {
      CV_CUDA(cudaLaunchKernel(kf, dim3(grid), dim3(CV_BLOCK), params, sh, st));
    }
  }
  CV_TRY(cv_check_launch(ctx, "orth_step"));
  *fused = true;
  return cv_wait_mailbox(ctx, a.host_seq, st);
}

// ------------------------------------------------------------------------------------------
// Gram-Schmidt against a set (reference semantics, numpyVector.py:121-145)
// ------------------------------------------------------------------------------------------
// scalar slots of the chain form a ring: step i only reads the slot of step i-1
constexpr int CV_GS_RING = 64;
static inline int gs_slot(int i, int NR) { return CV_S_GS + 2 * NR * (i % CV_GS_RING); }

template <typename T, int W>
static int gs_chain(cv_ctx *ctx, int64_t n, const T *x_in, int m, const void *const *q, T *x_out,
                    cudaStream_t st) {
  constexpr int NR = Num<T>::NRED;
  auto kf = k_mgs_step<T, W>;
  int grid = CV_KGRID(kf, n / W + 1);
  // algorithmic bytes: every step streams x (read + write), the previous and the current q
  cv_prof_scope prof(ctx, 4, st, (double)(4 * m + 2) * (double)n * (double)sizeof(T));
  // step i: subtract projection on q[i-1] (coefficients from slot i-1), dots with q[i]
  for (int i = 0; i <= m; ++i) {
    const T *src = (i == 0) ? x_in : x_out;
    const T *qp = (i == 0) ? nullptr : static_cast<const T *>(q[i - 1]);
    const T *qc = (i == m) ? nullptr : static_cast<const T *>(q[i]);
    const double *tprev = ctx->scalars + gs_slot(i > 0 ? i - 1 : 0, NR);
    double *tout = ctx->scalars + gs_slot(i, NR);
    k_mgs_step<T, W><<<grid, CV_BLOCK, 0, st>>>(n, src, x_out, qp, tprev, qc, ctx->partials,
                                                ctx->counters, tout);
    CV_TRY(cv_check_launch(ctx, "mgs_step"));
    CV_TRY(cv_reduce_ranks(ctx, gs_slot(i, NR), 2 * NR, st));
  }
  return CV_OK;
}

extern "C" int cv_gs_against_set(cv_ctx *ctx, int64_t n, int cplx_, const void *x_in, int m,
                                 const void *const *q_ptrs, double lindep, void *x_out, int *status,
                                 double *innerprod_host, void *stream) {
  CV_REQUIRE(ctx && x_in && x_out && status && innerprod_host, "cv_gs_against_set: null argument");
  CV_REQUIRE(m >= 0, "cv_gs_against_set: m=%d out of range", m);
  CV_REQUIRE(m == 0 || q_ptrs, "cv_gs_against_set: null q_ptrs");
  CV_REQUIRE(x_in != x_out, "cv_gs_against_set: out of place only");
  static_assert(CV_S_GS + 4 * CV_GS_RING <= CV_S_TS, "Gram-Schmidt scalar ring overlaps the tall-skinny slots");
  cudaStream_t st = (cudaStream_t)stream;
  const int NR = cplx_ ? 2 : 1;
  int W = cplx_ ? 1 : 2;
  if (((uintptr_t)x_in | (uintptr_t)x_out) & 15) W = 1;
  for (int i = 0; i < m; ++i)
    if ((uintptr_t)q_ptrs[i] & 15) W = 1;
  if (cplx_)
    CV_TRY((gs_chain<cplx, 1>(ctx, n, (const cplx *)x_in, m, q_ptrs, (cplx *)x_out, st)));
  else if (W == 2)
    CV_TRY((gs_chain<double, 2>(ctx, n, (const double *)x_in, m, q_ptrs, (double *)x_out, st)));
  else
    CV_TRY((gs_chain<double, 1>(ctx, n, (const double *)x_in, m, q_ptrs, (double *)x_out, st)));
  const int slot = gs_slot(m, NR);  // innerprod = x.x
  CV_TRY(cv_fetch_scalars(ctx, slot, NR, st));
  innerprod_host[0] = ctx->mailbox[slot];
  innerprod_host[1] = cplx_ ? ctx->mailbox[slot + 1] : 0.0;
  // numpyVector.py:141: `innerprod > lindep` (for complex data numpy compares lexicographically;
  // the real part decides unless it ties)
  bool ok = innerprod_host[0] > lindep || (innerprod_host[0] == lindep && innerprod_host[1] > 0.0);
  if (!ok) {
    *status = 1;
    return CV_OK;
  }
  *status = 0;
  if (cplx_) {
    k_div_csqrt<<<CV_KGRID(k_div_csqrt, n), CV_BLOCK, 0, st>>>(n, (cplx *)x_out, ctx->scalars + slot);
    return cv_check_launch(ctx, "div_csqrt");
  }
  return cv_scale_dev(ctx, n, 0, x_out, slot, 0, st);
}

// ------------------------------------------------------------------------------------------
// operator
// ------------------------------------------------------------------------------------------
extern "C" int cv_op_create_csr(cv_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                                const int64_t *indptr_dev, const int32_t *indices_dev,
                                const double *data_dev, cv_op **out) {
  CV_REQUIRE(ctx && out && indptr_dev, "cv_op_create_csr: null argument");
  CV_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "cv_op_create_csr: negative size");
  CV_REQUIRE(n_cols < ((int64_t)1 << 31), "cv_op_create_csr: column indices are int32");
  CV_REQUIRE(nnz == 0 || (indices_dev && data_dev), "cv_op_create_csr: null arrays");
  static uint64_t next_id = 0;
  cv_op *op = new cv_op();
  op->id = ++next_id;
  op->n_rows = n_rows;
  op->n_cols = n_cols;
  op->nnz = nnz;
  op->indptr = indptr_dev;
  op->indices = indices_dev;
  op->data = data_dev;
  op->fmt = CV_FMT_CSR;
  double mean = n_rows ? (double)nnz / (double)n_rows : 0.0;
  op->csr_group = mean <= 4 ? 2 : mean <= 8 ? 4 : mean <= 16 ? 8 : mean <= 48 ? 16 : 32;
  *out = op;
  return CV_OK;
}

extern "C" int cv_op_create_kron(cv_ctx *ctx, int64_t n_rows, int64_t row0, int ndim, const int32_t *dims, int nterm,
                                 const int32_t *terms7, const double *coef, const double *tab_val_dev,
                                 const int32_t *tab_col_dev, int tab_len, const double *dtab_dev,
                                 const int32_t *dtab_off, int dtab_len, int64_t max_offset, int64_t nnz_equiv,
                                 cv_op **out) {
  CV_REQUIRE(ctx && out && dims && dtab_dev && dtab_off, "cv_op_create_kron: null argument");
  CV_REQUIRE(ndim >= 1 && ndim <= KR_MAX_DIM, "cv_op_create_kron: ndim=%d outside 1..%d", ndim, KR_MAX_DIM);
  CV_REQUIRE(nterm >= 0 && nterm <= KR_MAX_TERMS, "cv_op_create_kron: %d product terms, at most %d", nterm, KR_MAX_TERMS);
  CV_REQUIRE(nterm == 0 || (terms7 && coef && tab_val_dev && tab_col_dev), "cv_op_create_kron: null term tables");
  CV_REQUIRE(tab_len >= 0 && tab_len <= KR_MAX_TAB && dtab_len >= 1 && dtab_len <= KR_MAX_DTAB,
             "cv_op_create_kron: tables of %d / %d entries exceed %d / %d", tab_len, dtab_len, KR_MAX_TAB, KR_MAX_DTAB);
  CV_REQUIRE(n_rows >= 0 && row0 >= 0 && max_offset >= 0, "cv_op_create_kron: negative size");
  static uint64_t next_kron_id = (uint64_t)1 << 40;
  cv_op *op = new cv_op();
  cv_op::Kron &q = op->kron;
  long long N = 1;
  for (int d = ndim - 1; d >= 0; --d) {
    CV_REQUIRE(dims[d] >= 1 && dims[d] <= 256, "cv_op_create_kron: dims[%d]=%d outside 1..256", d, dims[d]);
    q.dims[d] = dims[d];
    q.stride[d] = N;
    N *= dims[d];
    int sh = 32;
    while (((unsigned long long)1 << (sh - 32)) < (unsigned long long)dims[d]) ++sh;
    q.shift[d] = sh;
    q.magic[d] = (((unsigned long long)1 << sh) / (unsigned long long)dims[d]) + 1ull;
    q.dtab_off[d] = dtab_off[d];
    CV_REQUIRE(dtab_off[d] >= 0 && dtab_off[d] + dims[d] <= dtab_len, "cv_op_create_kron: diagonal table of mode %d out of range", d);
  }
  for (int d = ndim; d < KR_MAX_DIM; ++d) q.dims[d] = 1, q.stride[d] = 0, q.magic[d] = 0, q.shift[d] = 0, q.dtab_off[d] = 0;
  CV_REQUIRE(N < ((long long)1 << 31), "cv_op_create_kron: product basis of %lld states exceeds int32 rows", N);
  CV_REQUIRE(row0 + n_rows <= N, "cv_op_create_kron: rows [%lld, %lld) outside the product basis", (long long)row0,
             (long long)(row0 + n_rows));
  for (int t = 0; t < nterm; ++t) {
    KronTerm &k = q.term[t];
    k.mode_a = terms7[7 * t + 0];
    k.mode_b = terms7[7 * t + 1];
    k.tab_a = terms7[7 * t + 2];
    k.tab_b = terms7[7 * t + 3];
    k.w_a = terms7[7 * t + 4];
    k.w_b = terms7[7 * t + 5];
    k.coef = coef[t];
    CV_REQUIRE(k.mode_a >= 0 && k.mode_a < ndim && k.mode_b >= -1 && k.mode_b < ndim && k.mode_b != k.mode_a,
               "cv_op_create_kron: term %d has modes (%d, %d)", t, k.mode_a, k.mode_b);
    CV_REQUIRE(k.w_a >= 1 && k.tab_a >= 0 && k.tab_a + k.w_a * dims[k.mode_a] <= tab_len, "cv_op_create_kron: term %d table A out of range", t);
    CV_REQUIRE(k.mode_b < 0 || (k.w_b >= 1 && k.tab_b >= 0 && k.tab_b + k.w_b * dims[k.mode_b] <= tab_len),
               "cv_op_create_kron: term %d table B out of range", t);
  }
  q.ndim = ndim;
  q.nterm = nterm;
  q.tab_val = tab_val_dev;
  q.tab_col = tab_col_dev;
  q.dtab = dtab_dev;
  q.tab_len = tab_len;
  q.dtab_len = dtab_len;
  op->id = ++next_kron_id;
  op->n_rows = n_rows;
  op->n_cols = n_rows;
  op->nnz = nnz_equiv;
  op->row0 = row0;
  op->lo_len = (int)max_offset;
  op->hi_len = (int)max_offset;
  op->fmt = CV_FMT_KRON;
  *out = op;
  return CV_OK;
}

extern "C" int cv_op_destroy(cv_op *op) {
  delete op;
  return CV_OK;
}

extern "C" int cv_op_sell_widths(cv_ctx *ctx, cv_op *op, int32_t *widths_dev, void *stream) {
  CV_REQUIRE(ctx && op && widths_dev, "cv_op_sell_widths: null argument");
  int64_t ns = (op->n_rows + 31) / 32;
  if (ns == 0) return CV_OK;
  k_sell_widths<<<cv_grid_for(ctx, ns, CV_WARPS), CV_BLOCK, 0, (cudaStream_t)stream>>>(
      op->n_rows, ns, op->indptr, widths_dev);
  return cv_check_launch(ctx, "sell_widths");
}

extern "C" int cv_op_attach_sell(cv_ctx *ctx, cv_op *op, const int64_t *slice_ptr_dev,
                                 int64_t padded_nnz, int32_t *sell_col_dev, double *sell_val_dev,
                                 void *stream) {
  CV_REQUIRE(ctx && op && slice_ptr_dev, "cv_op_attach_sell: null argument");
  CV_REQUIRE(!op->data_im, "cv_op_attach_sell: a complex-valued operator stays in CSR storage");
  CV_REQUIRE(padded_nnz == 0 || (sell_col_dev && sell_val_dev), "cv_op_attach_sell: null storage");
  CV_REQUIRE(padded_nnz % 64 == 0, "cv_op_attach_sell: slice widths must be even (padded_nnz %% 64 == 0)");
  CV_REQUIRE(((uintptr_t)sell_val_dev & 15) == 0 && ((uintptr_t)sell_col_dev & 7) == 0,
             "cv_op_attach_sell: value / column storage must be 16- / 8-byte aligned");
  int64_t ns = (op->n_rows + 31) / 32;
  op->n_slices = ns;
  op->slice_ptr = slice_ptr_dev;
  op->sell_col = sell_col_dev;
  op->sell_val = sell_val_dev;
  op->padded_nnz = padded_nnz;
  if (ns > 0) {
    k_sell_fill<<<cv_grid_for(ctx, ns, CV_WARPS), CV_BLOCK, 0, (cudaStream_t)stream>>>(
        op->n_rows, op->n_cols, ns, op->indptr, op->indices, op->data, slice_ptr_dev, sell_col_dev,
        sell_val_dev);
    CV_TRY(cv_check_launch(ctx, "sell_fill"));
  }
  op->fmt = CV_FMT_SELL;
  return CV_OK;
}

extern "C" int cv_op_attach_dia(cv_ctx *ctx, cv_op *op, int n_diag, const int32_t *offsets_host,
                                const int32_t *col_global_dev, int64_t row0, double *dia_val_dev,
                                int64_t ld, int *ok_host, void *stream) {
  CV_REQUIRE(ctx && op && offsets_host && dia_val_dev && ok_host, "cv_op_attach_dia: null argument");
  CV_REQUIRE(!op->data_im, "cv_op_attach_dia: a complex-valued operator stays in CSR storage");
  CV_REQUIRE(n_diag >= 1 && n_diag <= CV_MAX_DIAG, "cv_op_attach_dia: n_diag=%d outside 1..%d", n_diag, CV_MAX_DIAG);
  CV_REQUIRE(ld >= op->n_rows, "cv_op_attach_dia: leading dimension smaller than the row count");
  for (int d = 1; d < n_diag; ++d)
    CV_REQUIRE(offsets_host[d] > offsets_host[d - 1], "cv_op_attach_dia: offsets must be strictly ascending");
  cudaStream_t st = (cudaStream_t)stream;
  // offsets and the error flag go through the scalar scratch area (reinterpreted as ints)
  int *d_off = reinterpret_cast<int *>(ctx->scalars + CV_S_TS);
  int *d_bad = d_off + CV_MAX_DIAG;
  int h_buf[CV_MAX_DIAG + 1];
  for (int d = 0; d < n_diag; ++d) h_buf[d] = offsets_host[d];
  for (int d = n_diag; d <= CV_MAX_DIAG; ++d) h_buf[d] = 0;
  CV_CUDA(cudaMemcpyAsync(d_off, h_buf, sizeof(h_buf), cudaMemcpyHostToDevice, st));
  CV_CUDA(cudaStreamSynchronize(st));  // h_buf is a stack array
  if (op->n_rows > 0) {
    k_dia_fill<<<cv_grid_for(ctx, op->n_rows, CV_BLOCK), CV_BLOCK, 0, st>>>(
        op->n_rows, row0, op->indptr, col_global_dev ? col_global_dev : op->indices, op->data, n_diag, d_off,
        dia_val_dev, ld, d_bad);
    CV_TRY(cv_check_launch(ctx, "dia_fill"));
  }
  int bad = 0;
  CV_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  CV_CUDA(cudaStreamSynchronize(st));
  *ok_host = bad ? 0 : 1;
  if (bad) return CV_OK;
  op->n_diag = n_diag;
  int mn = 0, mx = 0;
  for (int d = 0; d < n_diag; ++d) {
    op->dia_off[d] = offsets_host[d];
    mn = offsets_host[d] < mn ? offsets_host[d] : mn;
    mx = offsets_host[d] > mx ? offsets_host[d] : mx;
  }
  op->dia_val = dia_val_dev;
  op->dia_ld = ld;
  op->lo_len = -mn;
  op->hi_len = mx;
  op->row0 = row0;
  op->fmt = CV_FMT_DIA;
  return CV_OK;
}

extern "C" int cv_op_set_imag(cv_ctx *ctx, cv_op *op, const double *data_im_dev) {
  CV_REQUIRE(ctx && op, "cv_op_set_imag: null argument");
  CV_REQUIRE(op->fmt == CV_FMT_CSR && !op->slice_ptr && !op->dia_val,
             "cv_op_set_imag: complex values are carried by plain CSR storage only");
  CV_REQUIRE(op->nnz == 0 || data_im_dev, "cv_op_set_imag: null array");
  op->data_im = op->nnz ? data_im_dev : nullptr;
  return CV_OK;
}

extern "C" int cv_op_set_format(cv_op *op, int fmt) {
  CV_REQUIRE(op, "cv_op_set_format: null operator");
  CV_REQUIRE(!op->data_im || fmt == CV_FMT_CSR, "cv_op_set_format: a complex-valued operator stays in CSR storage");
  CV_REQUIRE(op->fmt != CV_FMT_KRON || fmt == CV_FMT_KRON, "cv_op_set_format: a matrix-free operator has no stored format");
  CV_REQUIRE(fmt == CV_FMT_CSR || (fmt == CV_FMT_SELL && op->slice_ptr) || (fmt == CV_FMT_DIA && op->dia_val) ||
                 (fmt == CV_FMT_KRON && op->fmt == CV_FMT_KRON),
             "cv_op_set_format: format %d not available", fmt);
  op->fmt = fmt;
  return CV_OK;
}

extern "C" int cv_op_info(cv_op *op, int64_t *n_rows, int64_t *nnz, int64_t *padded_nnz, int *fmt) {
  CV_REQUIRE(op, "cv_op_info: null operator");
  if (n_rows) *n_rows = op->n_rows;
  if (nnz) *nnz = op->nnz;
  if (padded_nnz) *padded_nnz = op->padded_nnz;
  if (fmt) *fmt = op->fmt;
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// SpMV dispatch
// ------------------------------------------------------------------------------------------
template <typename T, bool HALO, bool EPI, bool DOTS>
static int launch_spmv_fmt(cv_ctx *ctx, cv_op *op, const SpmvArgs<T> &a, cudaStream_t st) {
  // Bytes the STORED format must move per launch (x read once, y written once, + u1 when fused):
  //   DIA  8*D*ld                       (no index stream)
  //   SELL 12*padded_nnz + 8*n_slices
  //   CSR  12*nnz + 8*(n_rows+1)
  // next to SURVEY 8d's format-independent CSR-equivalent figure 12*nnz + 20*N (36*N complex).
  const double vecb = (double)sizeof(T) * (double)op->n_rows * (2.0 + (EPI && a.u1 ? 1.0 : 0.0));
  const double matb = op->fmt == CV_FMT_DIA    ? 8.0 * (double)op->n_diag * (double)op->dia_ld
                      : op->fmt == CV_FMT_KRON ? 0.0  // matrix-free: tables of a few KB in shared memory
                      : op->fmt == CV_FMT_SELL ? 12.0 * (double)op->padded_nnz + 8.0 * (double)op->n_slices
                                               : 12.0 * (double)op->nnz + 8.0 * (double)(op->n_rows + 1);
  cv_prof_add_bytes(ctx, 6, 12.0 * (double)op->nnz + (sizeof(T) == 16 ? 36.0 : 20.0) * (double)op->n_rows);
  cv_prof_scope prof(ctx, 0, st, matb + vecb);
  if (op->fmt == CV_FMT_DIA) {
    DiaArgs<T> d;
    d.s = a;
    d.dia_val = op->dia_val;
    d.ld = op->dia_ld;
    d.n_diag = op->n_diag;
    for (int k = 0; k < CV_MAX_DIAG; ++k) d.off[k] = k < op->n_diag ? op->dia_off[k] : 0;
    d.halo_lo = static_cast<const T *>(op->peer_halo ? op->halo_lo_cur : op->halo_lo);
    d.halo_hi = static_cast<const T *>(op->peer_halo ? op->halo_hi_cur : op->halo_hi);
    d.lo_len = op->lo_len;
    d.hi_len = op->hi_len;
    static const int pair_rows = getenv("EIGB200_DIA_PAIR") ? atoi(getenv("EIGB200_DIA_PAIR")) : 1;
    bool paired = false;
    if constexpr (sizeof(T) == 8) {
      // two rows per thread with 128-bit loads (real vectors, 16-byte aligned x, even leading dimension)
      if (pair_rows && (((uintptr_t)a.x | (uintptr_t)op->dia_val) & 15) == 0 && (op->dia_ld & 1) == 0 && op->n_rows >= 2) {
        auto kf = k_spmv_dia2<HALO, EPI, DOTS>;
        int wave = cv_occ_grid(ctx, (const void *)kf, (int64_t)1 << 40, CV_BLOCK);
        int64_t need = ((op->n_rows + 1) / 2 + CV_BLOCK - 1) / CV_BLOCK;
        int grid = (int)((!DOTS && need <= 16 * (int64_t)wave) || need < wave ? need : wave);
        kf<<<grid, CV_BLOCK, 0, st>>>(d);
        paired = true;
      }
    }
    if (!paired) {
      auto kf = k_spmv_dia<T, HALO, EPI, DOTS>;
      int wave = cv_occ_grid(ctx, (const void *)kf, (int64_t)1 << 40, CV_BLOCK);
      int64_t need = (op->n_rows + CV_BLOCK - 1) / CV_BLOCK;
      // with fused dots every CTA ends in a reduction epilogue (~3 us of latency): keep ONE
      // persistent wave; without, small problems run one row per thread for load balance
      int grid = (int)((!DOTS && need <= 16 * (int64_t)wave) || need < wave ? need : wave);
      kf<<<grid, CV_BLOCK, 0, st>>>(d);
    }
  } else if (op->fmt == CV_FMT_KRON) {
    KronArgs<T> k;
    k.s = a;
    const cv_op::Kron &q = op->kron;
    k.ndim = q.ndim;
    k.nterm = q.nterm;
    for (int d = 0; d < KR_MAX_DIM; ++d) {
      k.dims[d] = q.dims[d];
      k.dtab_off[d] = q.dtab_off[d];
      k.stride[d] = q.stride[d];
      k.magic[d] = q.magic[d];
      k.shift[d] = q.shift[d];
    }
    for (int t = 0; t < q.nterm; ++t) k.term[t] = q.term[t];
    k.tab_val = q.tab_val;
    k.tab_col = q.tab_col;
    k.dtab = q.dtab;
    k.tab_len = q.tab_len;
    k.dtab_len = q.dtab_len;
    k.row0 = op->row0;
    k.halo_lo = static_cast<const T *>(op->peer_halo ? op->halo_lo_cur : op->halo_lo);
    k.halo_hi = static_cast<const T *>(op->peer_halo ? op->halo_hi_cur : op->halo_hi);
    k.lo_len = op->lo_len;
    k.hi_len = op->hi_len;
    auto kf = k_spmv_kron<T, HALO, EPI, DOTS>;
    int wave = cv_occ_grid(ctx, (const void *)kf, (int64_t)1 << 40, CV_BLOCK);
    int64_t need = (op->n_rows + CV_BLOCK - 1) / CV_BLOCK;
    int grid = (int)(need < wave ? need : wave);  // every CTA stages the tables: one persistent wave
    kf<<<grid, CV_BLOCK, 0, st>>>(k);
  } else if (op->fmt == CV_FMT_SELL) {
    auto kf = k_spmv_sell<T, HALO, EPI, DOTS>;
    int wave = cv_occ_grid(ctx, (const void *)kf, (int64_t)1 << 40, CV_WARPS);
    int64_t need = (op->n_slices + CV_WARPS - 1) / CV_WARPS;
    // few slices per warp: a static slice->warp map would quantise badly (3.3 slices per warp on
    // the 100^3 Laplacian = 18 % idle); launch one slice per warp and let the hardware CTA
    // scheduler balance.  Many slices per warp: one persistent wave.
    // many slices per warp, or fused dots (each CTA then ends in a ~3 us reduction epilogue,
    // measured 33 us extra with 7813 small CTAs): one persistent wave.
    int grid = (int)((!DOTS && need <= 16 * (int64_t)wave) || need < wave ? need : wave);
    kf<<<grid, CV_BLOCK, 0, st>>>(a);
  } else {
#define CSR(G)                                                                    \
  {                                                                               \
    auto kf = k_spmv_csr<T, G, HALO, EPI, DOTS>;                                  \
    int grid = cv_occ_grid(ctx, (const void *)kf, op->n_rows, CV_BLOCK / G);      \
    kf<<<grid, CV_BLOCK, 0, st>>>(a);                                             \
  }
    switch (op->csr_group) {
      case 2: CSR(2); break;
      case 4: CSR(4); break;
      case 8: CSR(8); break;
      case 16: CSR(16); break;
      default: CSR(32); break;
    }
#undef CSR
  }
  return cv_check_launch(ctx, "spmv");
}

template <typename T>
static int launch_spmv_t(cv_ctx *ctx, cv_op *op, int mode, T sigma, const T *x, T *y, double alpha,
                         double beta1, const T *u1, bool epi, int dots_slot, cudaStream_t st) {
  SpmvArgs<T> a;
  a.n_rows = op->n_rows;
  a.n_slices = op->n_slices;
  a.slice_ptr = op->slice_ptr;
  a.sell_col = op->sell_col;
  a.sell_val = op->sell_val;
  a.indptr = op->indptr;
  a.indices = op->indices;
  a.data = op->data;
  a.data_im = op->data_im;
  CV_REQUIRE(!op->data_im || (sizeof(T) == 16 && op->fmt == CV_FMT_CSR),
             "spmv: a complex-valued operator acts on complex vectors, in CSR storage");
  a.x = x;
  a.halo = static_cast<const T *>(op->halobuf);
  a.n_local_cols = op->n_halo > 0 ? op->n_cols - op->n_halo : op->n_cols;
  a.y = y;
  a.mode = mode;
  a.sigma = sigma;
  a.alpha = alpha;
  a.beta1 = beta1;
  a.u1 = u1;
  a.partials = ctx->partials;
  a.counter = ctx->counters;
  a.out = dots_slot >= 0 ? ctx->scalars + dots_slot : nullptr;
  a.wait = HaloWait{nullptr, 0u, 0ull, nullptr, nullptr};
  const bool dia = cv_op_banded(op);
  // a rank takes part in the exchange when it RECEIVES (n_halo > 0) or only SENDS (structurally
  // one-sided couplings, an empty row block): the same predicate as cv_orth_step_dev
  const bool halo = dia ? (ctx->world > 1 && (op->lo_len > 0 || op->hi_len > 0))
                        : (op->n_halo > 0 || (!op->send_off.empty() && op->send_off.back() > 0));
  const bool dots = dots_slot >= 0;
  if (halo) {
    if (dia)
      CV_TRY(cv_halo_exchange_dia(ctx, op, sizeof(T) == 16, x, st));
    else
      CV_TRY(cv_halo_exchange(ctx, op, sizeof(T) == 16, x, st));
    if (op->peer_halo) {
      a.halo = static_cast<const T *>(op->halo_cur);  // parity of this exchange
      a.wait = op->wait;                              // the kernel polls its sources' flags itself
    }
  }
  int rc;
  if (op->n_rows == 0) {  // nothing to compute here, but the collectives above/below must still happen
    if (dots) {
      CV_CUDA(cudaMemsetAsync(ctx->scalars + dots_slot, 0, 3 * sizeof(double), st));
      CV_TRY(cv_reduce_ranks(ctx, dots_slot, 3, st));
    }
    return CV_OK;
  }
#define GO(H, E, D) rc = launch_spmv_fmt<T, H, E, D>(ctx, op, a, st)
  if (halo) {
    if (epi) { if (dots) GO(true, true, true); else GO(true, true, false); }
    else     { if (dots) GO(true, false, true); else GO(true, false, false); }
  } else {
    if (epi) { if (dots) GO(false, true, true); else GO(false, true, false); }
    else     { if (dots) GO(false, false, true); else GO(false, false, false); }
  }
#undef GO
  CV_TRY(rc);
  if (dots) CV_TRY(cv_reduce_ranks(ctx, dots_slot, 3, st));
  return CV_OK;
}

int cv_spmv_dev(cv_ctx *ctx, cv_op *op, int cplx_, int mode, double sre, double sim, const void *x,
                void *y, double alpha, double beta1, const void *u1, bool epi, int dots_slot,
                cudaStream_t st) {
  CV_REQUIRE(mode >= 0 && mode <= 2, "spmv: mode %d", mode);
  CV_REQUIRE(x != y, "spmv: x and y must not alias");
  if (op->n_rows == 0 && ctx->world == 1) return CV_OK;
  if (cplx_)
    return launch_spmv_t<cplx>(ctx, op, mode, make_cplx(sre, sim), (const cplx *)x, (cplx *)y, alpha,
                               beta1, (const cplx *)u1, epi, dots_slot, st);
  CV_REQUIRE(sim == 0.0, "spmv: complex shift needs complex vectors");
  return launch_spmv_t<double>(ctx, op, mode, sre, (const double *)x, (double *)y, alpha, beta1,
                               (const double *)u1, epi, dots_slot, st);
}

extern "C" int cv_spmv(cv_ctx *ctx, cv_op *op, int cplx_, int mode, double sigma_re, double sigma_im,
                       const void *x, void *y, void *stream) {
  CV_REQUIRE(ctx && op && x && y, "cv_spmv: null argument");
  return cv_spmv_dev(ctx, op, cplx_, mode, sigma_re, sigma_im, x, y, 1.0, 0.0, nullptr, false, -1,
                     (cudaStream_t)stream);
}

extern "C" int cv_spmv_dots(cv_ctx *ctx, cv_op *op, int cplx_, int mode, double sigma_re,
                            double sigma_im, const void *x, void *y, double *out3_host, void *stream) {
  CV_REQUIRE(ctx && op && x && y, "cv_spmv_dots: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CV_TRY(cv_spmv_dev(ctx, op, cplx_, mode, sigma_re, sigma_im, x, y, 1.0, 0.0, nullptr, false,
                     CV_S_TMP, st));
  if (out3_host) {
    CV_TRY(cv_fetch_scalars(ctx, CV_S_TMP, 3, st));
    for (int i = 0; i < 3; ++i) out3_host[i] = ctx->mailbox[CV_S_TMP + i];
  }
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// extension of the overlap / operator matrices by one column (numpyVector.py:205-238)
// ------------------------------------------------------------------------------------------
extern "C" int cv_extend_columns(cv_ctx *ctx, cv_op *op, int64_t n, int cplx_, int m,
                                 const void *const *v_ptrs, void *ket_tmp, double *s_col_host,
                                 double *h_col_host, void *stream) {
  CV_REQUIRE(ctx && v_ptrs && m >= 1, "cv_extend_columns: bad argument");
  CV_REQUIRE(s_col_host || h_col_host, "cv_extend_columns: nothing to compute");
  CV_REQUIRE(!h_col_host || (op && ket_tmp), "cv_extend_columns: operator column needs op and ket_tmp");
  cudaStream_t st = (cudaStream_t)stream;
  const int nr = cplx_ ? 2 : 1;
  const void *w[2];
  int b = 0;
  if (s_col_host) w[b++] = v_ptrs[m - 1];
  if (h_col_host) {
    CV_REQUIRE(op->n_rows == n, "cv_extend_columns: operator has %lld rows, vectors %lld", (long long)op->n_rows, (long long)n);
    CV_TRY(cv_spmv_dev(ctx, op, cplx_, CV_SPMV_PLAIN, 0.0, 0.0, v_ptrs[m - 1], ket_tmp, 1.0, 0.0,
                       nullptr, false, -1, st));
    w[b++] = ket_tmp;
  }
  for (int i0 = 0; i0 < m; i0 += CV_MAX_PTRS) {  // any m: V in chunks of CV_MAX_PTRS vectors
    const int mm = m - i0 < CV_MAX_PTRS ? m - i0 : CV_MAX_PTRS;
    CV_TRY(cv_tsdot_dev(ctx, n, cplx_, 1, mm, v_ptrs + i0, b, w, CV_S_TS, st));
    CV_TRY(cv_fetch_scalars(ctx, CV_S_TS, mm * b * nr, st));
    for (int i = 0; i < mm; ++i) {
      int k = 0;
      if (s_col_host) {
        for (int c = 0; c < nr; ++c) s_col_host[(i0 + i) * nr + c] = ctx->mailbox[CV_S_TS + (i * b + k) * nr + c];
        ++k;
      }
      if (h_col_host)
        for (int c = 0; c < nr; ++c) h_col_host[(i0 + i) * nr + c] = ctx->mailbox[CV_S_TS + (i * b + k) * nr + c];
    }
  }
  return CV_OK;
}
