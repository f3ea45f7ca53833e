// common.cuh — shared device/host helpers of libcudavec (sm_100a only).
//
// Everything on this path is fp64 and HBM-bound, so the helpers here are about
// three things only: wide coalesced loads with streaming cache hints, deterministic two-stage
// reductions (warp shuffle -> block -> ordered sum over blocks by the last block to arrive),
// and a complex type whose memory layout equals numpy's complex128.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <vector>
#include "../../include/cudavec.h"

// ------------------------------------------------------------------------------------------
// error plumbing (no exceptions across the C ABI)
// ------------------------------------------------------------------------------------------
void cv_set_error(const char *fmt, ...);

#define CV_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      cv_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return CV_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define CV_TRY(call)            \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != CV_OK) return rc__; \
  } while (0)

#define CV_REQUIRE(cond, ...)     \
  do {                            \
    if (!(cond)) {                \
      cv_set_error(__VA_ARGS__);  \
      return CV_ERR_ARG;          \
    }                             \
  } while (0)

// ------------------------------------------------------------------------------------------
// launch geometry: B200 has 148 SMs; streaming kernels run as persistent grid-stride grids of
// 148 x 8 CTAs x 256 threads (= 2048 resident threads per SM, one full wave).
// ------------------------------------------------------------------------------------------
constexpr int CV_BLOCK = 256;
constexpr int CV_WARPS = CV_BLOCK / 32;
constexpr int CV_CTAS_PER_SM = 8;
constexpr int CV_MAX_GRID = 148 * CV_CTAS_PER_SM * 2;  // upper bound used to size scratch
constexpr int CV_MAX_PTRS = 128;                       // vectors per tall-skinny call
constexpr int CV_MAX_COEF = 256;                       // doubles of coefficients per launch
constexpr int CV_MAX_RED = 1024;                       // reduction values per launch

// scratch layout handed over by the host language (torch tensor), see cv_ctx_create
constexpr size_t CV_N_COUNTERS = 64;                     // unsigned tickets
constexpr size_t CV_N_SCALARS = 8192;                    // doubles: reduction results / coefficients (lock-step solves: one block each)
constexpr size_t CV_N_PARTIALS = (size_t)1 << 21;        // doubles: per-CTA partial sums (16 MiB)

struct cv_comm_state;  // NCCL state, comm.cu
struct cv_peer_state;  // peer-memory transport (CUDA IPC windows over NVLink), peer.cu

// peer-memory transport limits: one NVSwitch node
constexpr int CV_MAX_WORLD = 8;
constexpr int CV_AR_DEPTH = 4;       // all-reduce slots in flight (2 suffice, see peer.cu)
constexpr int CV_AR_MAX = 1152;      // doubles per all-reduce message
constexpr int CV_COUNTER_PUSH = 63;  // ticket of the halo push kernel inside ctx->counters
constexpr int CV_COUNTER_BAR = 60;   // {arrivals, generation} of the fused Arnoldi step's grid barrier

// optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline)
// classes: 0 fused SpMV, 1 fused Arnoldi step / tall-skinny dot, 2 tall-skinny update, 3 other
// vector kernels, 4 Gram-Schmidt against a set, 5 linear combinations, 6 (no launches) the SpMV's
// CSR-equivalent bytes 12 nnz + 20 N next to class 0's bytes of the format actually stored, 7 spare.
// `bytes` accumulates the ALGORITHMIC bytes of the timed launches (SURVEY 8d formulas, stated at
// each launch site) so that GB/s = bytes / ms is reproducible from one bench line.
constexpr int CV_PROF_CLASSES = 8;
constexpr int CV_PROF_POOL = 2048;  // event pairs kept in flight before they are drained
struct cv_prof_state {
  bool enabled = false;
  cudaEvent_t start[CV_PROF_POOL], stop[CV_PROF_POOL];
  int cls[CV_PROF_POOL];
  int used = 0;
  bool created = false;
  double ms[CV_PROF_CLASSES] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t count[CV_PROF_CLASSES] = {0, 0, 0, 0, 0, 0, 0, 0};
  double bytes[CV_PROF_CLASSES] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// GCROT recycling across solves (solvers.cu): which (c,u) ring slots of the solver workspace hold
// vectors that are valid for which operator / shift
struct cv_op;
struct cv_recycle_state {
  bool enabled = false, valid = false;
  uint64_t op_id = 0;  // cv_op::id (monotonic), not the pointer: a new operator may reuse the address
  int cplx = 0, mode = 0, m = 0, k = 0;
  double sre = 0.0, sim = 0.0;
  int64_t n = 0;
  const void *work = nullptr;
  std::vector<int> cu_slots, free_slots;
};

struct cv_ctx {
  int device;
  int sms;
  unsigned *counters;   // device, zeroed
  double *scalars;      // device
  double *partials;     // device
  double *mailbox;      // pinned host mirror of `scalars`
  uint64_t launches;
  cv_comm_state *comm;  // null when world == 1
  cv_peer_state *peer;  // non-null: collectives go through peer memory instead of NCCL
  int rank, world;
  cv_prof_state *prof;
  // GCROT re-orthogonalises only when the first Gram-Schmidt pass leaves less than eta of the
  // norm (Daniel-Gragg-Kaufman-Stewart); eta = 1/sqrt(2) is the classic value, 0.1 keeps the
  // basis orthogonal to ~10 eps while skipping the second pass in all but cancelling steps
  double reorth_eta;
  bool defer_reduce;  // solvers batch several reductions into one all-reduce
  // fused Arnoldi step: results arrive in the mailbox followed by a released sequence flag
  unsigned long long *host_flag;  // pinned, device-visible
  unsigned long long host_seq;
  // the vector whose halo phase C of the fused step has already pushed (next SpMV skips its push)
  const void *prepushed_x;
  const void *prepushed_op;
  cv_recycle_state recycle;
  bool push_early;  // fused step pushes unnormalised halo rows from phase B (EIGB200_PUSH_EARLY=0: off)
  int slab_mode;    // dot-phase work split of the fused Arnoldi step (kernels_orth.cuh SlabMap; EIGB200_SLAB)
  int snake;        // update phase walks the rows downwards (EIGB200_SNAKE)
  // diagonal right preconditioner of the current cv_solve_precond call (null: none) and its two scratch vectors
  const void *precond_dinv;
  void *precond_z, *precond_t;
};

// RAII bracket: records an event pair around the launches issued inside its scope
struct cv_prof_scope {
  cv_ctx *ctx;
  cudaStream_t st;
  int slot;
  cv_prof_scope(cv_ctx *c, int cls, cudaStream_t s, double alg_bytes = 0.0);
  ~cv_prof_scope();
};
inline void cv_prof_add_bytes(cv_ctx *ctx, int cls, double b) {
  if (ctx->prof && ctx->prof->enabled) ctx->prof->bytes[cls] += b;
}

inline int cv_grid_for(const cv_ctx *ctx, int64_t work_items, int items_per_cta) {
  int64_t need = (work_items + items_per_cta - 1) / items_per_cta;
  int64_t cap = (int64_t)ctx->sms * CV_CTAS_PER_SM;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

int cv_fetch_scalars(cv_ctx *ctx, int offset, int count, cudaStream_t st);  // D2H + sync
int cv_reduce_ranks(cv_ctx *ctx, int offset, int count, cudaStream_t st);   // NCCL sum (world>1)

// ------------------------------------------------------------------------------------------
// complex128 with numpy layout
// ------------------------------------------------------------------------------------------
struct __align__(16) cplx {
  double re, im;
};

__host__ __device__ inline cplx make_cplx(double r, double i) {
  cplx c;
  c.re = r;
  c.im = i;
  return c;
}

template <typename T>
struct Num;

template <>
struct Num<double> {
  static constexpr int NRED = 1;
  __host__ __device__ static inline double zero() { return 0.0; }
  __host__ __device__ static inline double make(double r, double) { return r; }
  __device__ static inline double mul(double a, double b) { return a * b; }
  __device__ static inline double add(double a, double b) { return a + b; }
  __device__ static inline double sub(double a, double b) { return a - b; }
  __device__ static inline double conj(double a) { return a; }
  __device__ static inline double abs2(double a) { return a * a; }
  __device__ static inline double scale(double a, double s) { return a * s; }
  // acc += a*b
  __device__ static inline void fma(double &acc, double a, double b) { acc = ::fma(a, b, acc); }
  // acc += conj(a)*b
  __device__ static inline void fmac(double &acc, double a, double b) { acc = ::fma(a, b, acc); }
  // acc += s*b with real s
  __device__ static inline void fmar(double &acc, double s, double b) { acc = ::fma(s, b, acc); }
  __device__ static inline void to_red(double a, double *out) { out[0] = a; }
  __device__ static inline double from_red(const double *in) { return in[0]; }
};

template <>
struct Num<cplx> {
  static constexpr int NRED = 2;
  __host__ __device__ static inline cplx zero() { return make_cplx(0.0, 0.0); }
  __host__ __device__ static inline cplx make(double r, double i) { return make_cplx(r, i); }
  __device__ static inline cplx mul(cplx a, cplx b) {
    return make_cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
  }
  __device__ static inline cplx add(cplx a, cplx b) { return make_cplx(a.re + b.re, a.im + b.im); }
  __device__ static inline cplx sub(cplx a, cplx b) { return make_cplx(a.re - b.re, a.im - b.im); }
  __device__ static inline cplx conj(cplx a) { return make_cplx(a.re, -a.im); }
  __device__ static inline double abs2(cplx a) { return a.re * a.re + a.im * a.im; }
  __device__ static inline cplx scale(cplx a, double s) { return make_cplx(a.re * s, a.im * s); }
  __device__ static inline void fma(cplx &acc, cplx a, cplx b) {
    acc.re = ::fma(a.re, b.re, acc.re);
    acc.re = ::fma(-a.im, b.im, acc.re);
    acc.im = ::fma(a.re, b.im, acc.im);
    acc.im = ::fma(a.im, b.re, acc.im);
  }
  __device__ static inline void fmac(cplx &acc, cplx a, cplx b) {  // conj(a)*b
    acc.re = ::fma(a.re, b.re, acc.re);
    acc.re = ::fma(a.im, b.im, acc.re);
    acc.im = ::fma(a.re, b.im, acc.im);
    acc.im = ::fma(-a.im, b.re, acc.im);
  }
  __device__ static inline void fmar(cplx &acc, double s, cplx b) {
    acc.re = ::fma(s, b.re, acc.re);
    acc.im = ::fma(s, b.im, acc.im);
  }
  __device__ static inline void to_red(cplx a, double *out) {
    out[0] = a.re;
    out[1] = a.im;
  }
  __device__ static inline cplx from_red(const double *in) { return make_cplx(in[0], in[1]); }
};

// mixed helpers used by kernels templated on both a vector type and a coefficient type
__device__ inline void cfma(double &acc, double c, double v) { acc = ::fma(c, v, acc); }
__device__ inline void cfma(cplx &acc, double c, cplx v) { Num<cplx>::fmar(acc, c, v); }
__device__ inline void cfma(cplx &acc, cplx c, cplx v) { Num<cplx>::fma(acc, c, v); }
__device__ inline void cfma(cplx &acc, cplx c, double v) {
  acc.re = ::fma(c.re, v, acc.re);
  acc.im = ::fma(c.im, v, acc.im);
}

// ------------------------------------------------------------------------------------------
// cache-hinted loads/stores.  Matrix values/indices and one-shot vector streams bypass L1
// (ld.global.nc.L1::no_allocate) so that L1 stays available to the gathered x; results that
// the next kernel re-reads go through the normal path so they can stay in the 126 MB L2.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream2(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(v.x), "=d"(v.y)
               : "l"(p));
  return v;
}
__device__ __forceinline__ cplx ld_stream(const cplx *p) {
  double2 v = ld_stream2(reinterpret_cast<const double2 *>(p));
  return make_cplx(v.x, v.y);
}
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int2 ld_stream2(const int2 *p) {
  int2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];"
               : "=r"(v.x), "=r"(v.y)
               : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream4(const int4 *p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
// gathered x: read-only path, allocate in L1 (neighbouring rows hit the same lines)
__device__ __forceinline__ double ld_gather(const double *p) { return __ldg(p); }
__device__ __forceinline__ cplx ld_gather(const cplx *p) {
  double2 v = __ldg(reinterpret_cast<const double2 *>(p));
  return make_cplx(v.x, v.y);
}
// plain (coherent) loads for data written earlier by other kernels and possibly L2 resident
__device__ __forceinline__ double ld_plain(const double *p) { return *p; }
__device__ __forceinline__ cplx ld_plain(const cplx *p) {
  double2 v = *reinterpret_cast<const double2 *>(p);
  return make_cplx(v.x, v.y);
}
// L2-only loads (ld.global.cg): data another SM may have rewritten earlier in the SAME kernel
// (fused multi-phase kernels) must not be served from a stale L1 line
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ cplx ld_cg(const cplx *p) {
  double2 v = __ldcg(reinterpret_cast<const double2 *>(p));
  return make_cplx(v.x, v.y);
}
__device__ __forceinline__ void st_plain(double *p, double v) { *p = v; }
__device__ __forceinline__ void st_plain(cplx *p, cplx v) {
  *reinterpret_cast<double2 *>(p) = make_double2(v.re, v.im);
}

// ------------------------------------------------------------------------------------------
// deterministic grid reduction
//
// Each CTA reduces NV doubles (warp shuffles, then across its 8 warps through shared memory),
// stores them at partials[(v)*nblk + blockIdx.x], fences and takes a ticket.  The CTA that
// draws the last ticket sums every value over all CTAs in a fixed order (warp w owns values
// w, w+8, ..; lanes stride over CTAs; xor-shuffle tree) and writes out[v].  The result does
// not depend on which CTA came last, so repeated runs are bit-identical.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ void grid_reduce(double (&vals)[NV], double *partials, unsigned *counter,
                                            double *out, int nblk_total, int blk_linear) {
  __shared__ double s_red[CV_WARPS][NV];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double r = warp_sum(vals[v]);
    if (lane == 0) s_red[warp][v] = r;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < CV_WARPS; ++w) r += s_red[w][threadIdx.x];
    partials[(size_t)threadIdx.x * nblk_total + blk_linear] = r;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(counter, 1u);
    s_last = (t == (unsigned)nblk_total - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int v = warp; v < NV; v += CV_WARPS) {
    const double *p = partials + (size_t)v * nblk_total;
    double r = 0.0;
    for (int b = lane; b < nblk_total; b += 32) r += __ldcg(p + b);
    r = warp_sum(r);
    if (lane == 0) out[v] = r;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// runtime-sized variant (tall-skinny products): values live in shared memory, s_vals[nv]
// already summed over the CTA.
__device__ __forceinline__ void grid_reduce_dyn(const double *s_vals, int nv, double *partials,
                                                unsigned *counter, double *out, int nblk_total,
                                                int blk_linear) {
  __shared__ bool s_last_dyn;
  for (int v = threadIdx.x; v < nv; v += blockDim.x)
    partials[(size_t)v * nblk_total + blk_linear] = s_vals[v];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = atomicAdd(counter, 1u);
    s_last_dyn = (t == (unsigned)nblk_total - 1u);
  }
  __syncthreads();
  if (!s_last_dyn) return;
  __threadfence();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int v = warp; v < nv; v += CV_WARPS) {
    const double *p = partials + (size_t)v * nblk_total;
    double r = 0.0;
    for (int b = lane; b < nblk_total; b += 32) r += __ldcg(p + b);
    r = warp_sum(r);
    if (lane == 0) out[v] = r;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// ------------------------------------------------------------------------------------------
// peer-memory transport, device side (host side and protocol notes: peer.cu)
// ------------------------------------------------------------------------------------------
constexpr unsigned long long CV_PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

struct PeerWindow {  // one per rank, in CUDA-IPC exported device memory, mapped by every peer
  unsigned long long halo_flag[CV_MAX_WORLD];  // [src] halo sequence published by src
  unsigned long long ar_seq;                   // local all-reduce counter (owner only)
  unsigned long long pad[23];
  // all-reduce mailboxes, LL protocol (as NCCL's low-latency protocol): every double travels as ONE
  // 16-byte store {lo32, seq32, hi32, seq32}; each 8-byte half carries its own tag, so the
  // receiver needs no fence and no separate flag: it polls the slot until both tags match.
  uint4 ar_ll[CV_AR_DEPTH][CV_MAX_WORLD][CV_AR_MAX];  // [slot][src][value]
};
struct PeerPtrs {
  PeerWindow *win[CV_MAX_WORLD];
};
struct PushSeg {
  void *dst;           // peer memory
  const int32_t *idx;  // null: contiguous range starting at src_start
  int64_t src_start;
  int64_t count;       // elements
};
struct PushArgs {
  PushSeg seg[2 * CV_MAX_WORLD];
  int nseg;
  unsigned long long *flag_dst[CV_MAX_WORLD];  // halo_flag[me] in each destination's window
  int nflag;
  unsigned long long seq;
  unsigned *ticket;
};
// what a sharded SpMV waits for before it reads its halo buffers
struct HaloWait {
  const unsigned long long *flags;  // own window's halo_flag
  unsigned mask;                    // source ranks
  unsigned long long seq;
  double *err;
  const double *scale_sq;  // non-null: halo entries were pushed UNNORMALISED and are multiplied by
                           // 1/sqrt(*scale_sq) on the fly, like the owner did with its rows.  Used by the
                           // two-barrier form of the fused Arnoldi step (rows pushed before the norm
                           // was known); the single-barrier form pushes normalised rows and leaves it null
};
__device__ __forceinline__ double halo_scale(const HaloWait &w) {
  if (!w.scale_sq) return 1.0;
  const double f = 1.0 / sqrt(__ldcg(w.scale_sq));
  return isfinite(f) ? f : 1.0;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= seq; false when the bounded wait expires (a peer died)
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long seq) {
  if (ld_acquire_sys(flag) >= seq) return true;
  const unsigned long long t0 = global_ns();
  for (;;) {
    for (int i = 0; i < 64; ++i)
      if (ld_acquire_sys(flag) >= seq) return true;
    if (global_ns() - t0 > CV_PEER_TIMEOUT_NS) return false;
  }
}
// CTA prologue of a sharded SpMV: the halo of this exchange has landed
__device__ __forceinline__ void halo_wait_cta(const HaloWait &w) {
  if (w.mask == 0u) return;
  if (threadIdx.x < CV_MAX_WORLD && ((w.mask >> threadIdx.x) & 1u))
    if (!wait_flag(w.flags + threadIdx.x, w.seq)) *w.err = 1.0;
  __syncthreads();
}

__device__ __forceinline__ void st_ll(uint4 *p, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"((unsigned)b), "r"(tag),
               "r"((unsigned)(b >> 32)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ bool ld_ll(const uint4 *p, unsigned tag, double &v) {
  unsigned lo, f1, hi, f2;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(p) : "memory");
  if (f1 != tag || f2 != tag) return false;
  v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
  return true;
}

// All-reduce (sum, in place) of buf[0..count), count <= CV_AR_MAX, executed by ALL threads of ONE
// CTA.  Thread t sends value t to every peer's mailbox (one tagged 16-byte store over NVLink per
// peer), then polls its own mailboxes for the other ranks' value t and sums in rank order, so all
// ranks compute bit-identical sums.  One one-way NVLink latency, no fences.  Mailbox slot reuse:
// slot = seq % CV_AR_DEPTH is rewritten at seq + DEPTH, which a rank can only reach after every
// rank has contributed to seq + DEPTH - 1, i.e. finished reading seq.
__device__ __forceinline__ void cta_peer_allreduce(const PeerPtrs &pp, int me, int world, double *buf,
                                                   int count, double *err) {
  __shared__ int s_bad;
  PeerWindow *own = pp.win[me];
  __syncthreads();
  const unsigned long long seq = own->ar_seq + 1;
  if (threadIdx.x == 0) s_bad = 0;
  const unsigned tag = (unsigned)seq;
  const int slot = (int)(seq % CV_AR_DEPTH);
  for (int t = threadIdx.x; t < count; t += blockDim.x) {
    const double v = __ldcg(buf + t);
    for (int p = 0; p < world; ++p)
      if (p != me) st_ll(&pp.win[p]->ar_ll[slot][me][t], v, tag);
  }
  for (int t = threadIdx.x; t < count; t += blockDim.x) {
    const double mine = __ldcg(buf + t);
    double s = 0.0;
    for (int q = 0; q < world; ++q) {
      double v = mine;
      if (q != me) {
        const uint4 *src = &own->ar_ll[slot][q][t];
        if (!ld_ll(src, tag, v)) {
          const unsigned long long t0 = global_ns();
          bool ok = false;
          while (!ok) {
            for (int i = 0; i < 64 && !ok; ++i) ok = ld_ll(src, tag, v);
            if (!ok && global_ns() - t0 > CV_PEER_TIMEOUT_NS) break;
          }
          if (!ok) {
            s_bad = 1;
            v = 0.0;
          }
        }
      }
      s += v;
    }
    buf[t] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    own->ar_seq = seq;
    if (s_bad) *err = 1.0;
  }
  __threadfence();
  __syncthreads();
}

// Push ranges of x into the peers' halo buffers (all threads of the grid), then the last CTA to
// finish raises the peers' flags.
template <typename T>
__device__ __forceinline__ void grid_halo_push(const PushArgs &a, const T *__restrict__ x) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int s = 0; s < a.nseg; ++s) {
    const PushSeg &g = a.seg[s];
    T *dst = static_cast<T *>(g.dst);
    if (g.idx) {
      for (int64_t i = tid; i < g.count; i += stride) st_plain(dst + i, ld_cg(x + g.idx[i]));
    } else {
      const T *src = x + g.src_start;
      if (sizeof(T) == 8 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        const int64_t n2 = g.count >> 1;
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        int64_t i = tid;
        for (; i + 3 * stride < n2; i += 4 * stride) {
          double2 v0 = __ldcg(s2 + i), v1 = __ldcg(s2 + i + stride), v2 = __ldcg(s2 + i + 2 * stride),
                  v3 = __ldcg(s2 + i + 3 * stride);
          d2[i] = v0;
          d2[i + stride] = v1;
          d2[i + 2 * stride] = v2;
          d2[i + 3 * stride] = v3;
        }
        for (; i < n2; i += stride) d2[i] = __ldcg(s2 + i);
        if ((g.count & 1) && tid == 0) st_plain(dst + g.count - 1, ld_cg(src + g.count - 1));
      } else {
        for (int64_t i = tid; i < g.count; i += stride) st_plain(dst + i, ld_cg(src + i));
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last_push;
  if (threadIdx.x == 0) s_last_push = atomicAdd(a.ticket, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (!s_last_push) return;
  __threadfence_system();
  if (threadIdx.x < a.nflag) st_release_sys(a.flag_dst[threadIdx.x], a.seq);
  if (threadIdx.x == 0) *a.ticket = 0u;
}

// scalar-slot map inside ctx->scalars (doubles).  Solvers use [CV_S_SOLVER, ...).
constexpr int CV_S_TMP = 0;        // 16 doubles: dot/norm results of the BLAS-1 entry points
constexpr int CV_S_GS = 16;        // 4*CV_MAX_PTRS doubles: MGS coefficients
constexpr int CV_S_TS = 1024;      // CV_MAX_RED doubles: tall-skinny results
constexpr int CV_S_SOLVER = 2048;  // solver scalars (h columns, norms, ...)
constexpr int CV_S_TRACE = (int)CV_N_SCALARS - 32;  // 16 doubles: phase times of the fused Arnoldi step
constexpr int CV_S_ERR = (int)CV_N_SCALARS - 1;  // raised (1.0) by a peer kernel whose bounded spin expired
