// peer.cu — NVLink/NVSwitch peer-memory transport for the row-sharded mode.
//
// The two collectives of the hot path are tiny and latency-bound: a sum of a few (<= ~130)
// doubles after every reduction kernel, and the halo of x before every SpMV.  An NCCL call costs
// 15-40 us of launch + protocol latency each; one Arnoldi step at 8 GPUs has ~300 us of useful
// work, so three NCCL calls per step cap the strong scaling.  Here every rank owns a small
// *window* and its halo buffers in cudaMalloc memory exported with CUDA IPC; peers map them and
//   * all-reduce: one 256-thread kernel stores its values into every peer's window (P2P stores
//     over NVLink), publishes a sequence flag with st.release.sys, spins on the flags of the
//     other ranks in its OWN window and sums the contributions in rank order — all ranks obtain
//     bit-identical sums (the drivers branch on them) in ~one NVLink round trip;
//   * halo: the owner PUSHES the entries a peer needs straight into that peer's halo buffer
//     (double-buffered by sequence parity), then publishes a flag; the receiver spins on the
//     flags of its sources before its SpMV reads the buffer.  No staging buffer, no rendezvous.
// Every spin is bounded (CV_PEER_TIMEOUT_NS); on expiry an error word is raised that
// cv_fetch_scalars turns into CV_ERR_COMM on the host, so a lost peer cannot hang the GPU.
//
// NCCL (comm.cu) stays as the fall-back transport when IPC mapping is unavailable.
#include "internal.h"

constexpr unsigned long long CV_PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

struct PeerWindow {
  unsigned long long ar_flag[CV_AR_DEPTH][CV_MAX_WORLD];
  unsigned long long halo_flag[CV_MAX_WORLD];
  unsigned long long pad[24];
  double ar_data[CV_AR_DEPTH][CV_MAX_WORLD][CV_AR_MAX];
};

struct PeerPtrs {
  PeerWindow *win[CV_MAX_WORLD];
};

extern "C" size_t cv_peer_window_bytes(void) { return sizeof(PeerWindow); }

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
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= seq; false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long seq) {
  if (ld_acquire_sys(flag) >= seq) return true;
  const unsigned long long t0 = global_ns();
  for (;;) {
    for (int i = 0; i < 64; ++i)
      if (ld_acquire_sys(flag) >= seq) return true;
    if (global_ns() - t0 > CV_PEER_TIMEOUT_NS) return false;
  }
}

// ------------------------------------------------------------------------------------------
// all-reduce (sum) of `count` <= CV_AR_MAX doubles, in place, one CTA
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_peer_allreduce(const __grid_constant__ PeerPtrs pp, int me, int world,
                                                        double *buf, int count, unsigned long long seq,
                                                        double *err) {
  const int slot = (int)(seq % CV_AR_DEPTH);
  for (int t = threadIdx.x; t < count; t += blockDim.x) {
    const double v = buf[t];
    for (int p = 0; p < world; ++p) pp.win[p]->ar_data[slot][me][t] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if (threadIdx.x < world) {
    st_release_sys(&pp.win[threadIdx.x]->ar_flag[slot][me], seq);
    if (!wait_flag(&pp.win[me]->ar_flag[slot][threadIdx.x], seq)) s_bad = 1;
  }
  __syncthreads();
  __threadfence_system();
  const PeerWindow *own = pp.win[me];
  for (int t = threadIdx.x; t < count; t += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < world; ++q) s += ld_relaxed_sys(&own->ar_data[slot][q][t]);
    buf[t] = s;
  }
  if (threadIdx.x == 0 && s_bad) *err = 1.0;
}

// ------------------------------------------------------------------------------------------
// halo push: copy (or gather) ranges of x into peers' halo buffers, then raise their flags
// ------------------------------------------------------------------------------------------
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

template <typename T>
__global__ void __launch_bounds__(CV_BLOCK) k_halo_push(const __grid_constant__ PushArgs a, const T *__restrict__ x) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int s = 0; s < a.nseg; ++s) {
    const PushSeg &g = a.seg[s];
    T *dst = static_cast<T *>(g.dst);
    if (g.idx) {
      for (int64_t i = tid; i < g.count; i += stride) st_plain(dst + i, ld_gather(x + g.idx[i]));
    } else {
      const T *src = x + g.src_start;
      if (sizeof(T) == 8 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        const int64_t n2 = g.count >> 1;
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        int64_t i = tid;
        for (; i + 3 * stride < n2; i += 4 * stride) {
          double2 v0 = s2[i], v1 = s2[i + stride], v2 = s2[i + 2 * stride], v3 = s2[i + 3 * stride];
          d2[i] = v0;
          d2[i + stride] = v1;
          d2[i + 2 * stride] = v2;
          d2[i + 3 * stride] = v3;
        }
        for (; i < n2; i += stride) d2[i] = s2[i];
        if ((g.count & 1) && tid == 0) st_plain(dst + g.count - 1, ld_plain(src + g.count - 1));
      } else {
        for (int64_t i = tid; i < g.count; i += stride) st_plain(dst + i, ld_plain(src + i));
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < a.nflag) st_release_sys(a.flag_dst[threadIdx.x], a.seq);
  if (threadIdx.x == 0) *a.ticket = 0u;
}

// wait until every source rank in `mask` has published halo sequence `seq`
__global__ void k_halo_wait(const unsigned long long *flags, unsigned mask, unsigned long long seq, double *err) {
  if (threadIdx.x < CV_MAX_WORLD && ((mask >> threadIdx.x) & 1u))
    if (!wait_flag(flags + threadIdx.x, seq)) *err = 1.0;
  __threadfence_system();
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct cv_peer_state {
  PeerPtrs pp;
  unsigned long long ar_seq = 0, halo_seq = 0;
};

extern "C" int cv_peer_alloc(cv_ctx *ctx, size_t bytes, void **ptr_dev, void *handle64) {
  CV_REQUIRE(ctx && ptr_dev && handle64 && bytes > 0, "cv_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CV_CUDA(cudaSetDevice(ctx->device));
  void *p = nullptr;
  CV_CUDA(cudaMalloc(&p, bytes));
  CV_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    cv_set_error("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    return CV_ERR_COMM;
  }
  memcpy(handle64, &h, 64);
  *ptr_dev = p;
  return CV_OK;
}

extern "C" int cv_peer_open(cv_ctx *ctx, const void *handle64, void **ptr_dev) {
  CV_REQUIRE(ctx && handle64 && ptr_dev, "cv_peer_open: bad argument");
  CV_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr_dev, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cv_set_error("cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e));
    return CV_ERR_COMM;
  }
  return CV_OK;
}

extern "C" int cv_peer_close(cv_ctx *ctx, void *ptr_dev) {
  (void)ctx;
  if (ptr_dev) cudaIpcCloseMemHandle(ptr_dev);
  return CV_OK;
}

extern "C" int cv_peer_free(cv_ctx *ctx, void *ptr_dev) {
  (void)ctx;
  if (ptr_dev) cudaFree(ptr_dev);
  return CV_OK;
}

extern "C" int cv_comm_attach_peers(cv_ctx *ctx, void *const *window_ptrs) {
  CV_REQUIRE(ctx && window_ptrs, "cv_comm_attach_peers: null argument");
  CV_REQUIRE(ctx->world > 1 && ctx->world <= CV_MAX_WORLD, "cv_comm_attach_peers: world=%d outside 2..%d", ctx->world,
             CV_MAX_WORLD);
  cv_peer_state *s = new cv_peer_state();
  for (int p = 0; p < CV_MAX_WORLD; ++p) s->pp.win[p] = nullptr;
  for (int p = 0; p < ctx->world; ++p) {
    CV_REQUIRE(window_ptrs[p], "cv_comm_attach_peers: null window for rank %d", p);
    s->pp.win[p] = static_cast<PeerWindow *>(window_ptrs[p]);
  }
  delete ctx->peer;
  ctx->peer = s;
  return CV_OK;
}

int cv_peer_detach(cv_ctx *ctx) {
  delete ctx->peer;
  ctx->peer = nullptr;
  return CV_OK;
}

extern "C" int cv_comm_transport(cv_ctx *ctx, int *transport) {
  CV_REQUIRE(ctx && transport, "cv_comm_transport: null argument");
  *transport = ctx->world == 1 ? 0 : (ctx->peer ? 2 : 1);
  return CV_OK;
}

int cv_peer_allreduce(cv_ctx *ctx, double *buf_dev, int count, cudaStream_t st) {
  cv_peer_state *s = ctx->peer;
  for (int c0 = 0; c0 < count; c0 += CV_AR_MAX) {
    const int c = count - c0 < CV_AR_MAX ? count - c0 : CV_AR_MAX;
    k_peer_allreduce<<<1, 256, 0, st>>>(s->pp, ctx->rank, ctx->world, buf_dev + c0, c, ++s->ar_seq,
                                        ctx->scalars + CV_S_ERR);
    CV_TRY(cv_check_launch(ctx, "peer_allreduce"));
  }
  return CV_OK;
}

// ---- halo plans ---------------------------------------------------------------------------
extern "C" int cv_op_set_halo_peers(cv_ctx *ctx, cv_op *op, void *const *halo_base, const int64_t *parity_stride_bytes,
                                    const int64_t *dst_off_elems) {
  CV_REQUIRE(ctx && op && halo_base && parity_stride_bytes && dst_off_elems, "cv_op_set_halo_peers: null argument");
  CV_REQUIRE(ctx->peer, "cv_op_set_halo_peers: peer transport not attached");
  CV_REQUIRE(!op->send_off.empty(), "cv_op_set_halo_peers: call cv_op_set_halo first");
  op->peer_base.assign(halo_base, halo_base + ctx->world);
  op->peer_stride.assign(parity_stride_bytes, parity_stride_bytes + ctx->world);
  op->peer_dst_off.assign(dst_off_elems, dst_off_elems + ctx->world);
  op->peer_halo = true;
  return CV_OK;
}

extern "C" int cv_op_set_dia_halo_peers(cv_ctx *ctx, cv_op *op, void *const *halo_base) {
  CV_REQUIRE(ctx && op && halo_base, "cv_op_set_dia_halo_peers: null argument");
  CV_REQUIRE(ctx->peer, "cv_op_set_dia_halo_peers: peer transport not attached");
  CV_REQUIRE(op->n_global > 0, "cv_op_set_dia_halo_peers: call cv_op_set_dia_halo first");
  op->peer_base.assign(halo_base, halo_base + ctx->world);
  op->peer_halo = true;
  return CV_OK;
}

// bytes of one parity of the DIA halo allocation: [lo | hi], each padded to 256 bytes
static inline size_t dia_lo_bytes(const cv_op *op) { return (((size_t)op->lo_len * 16) + 255) & ~(size_t)255; }
static inline size_t dia_hi_bytes(const cv_op *op) { return (((size_t)op->hi_len * 16) + 255) & ~(size_t)255; }
extern "C" size_t cv_op_dia_halo_bytes(cv_op *op) {
  return op ? 2 * (dia_lo_bytes(op) + dia_hi_bytes(op)) + 256 : 0;
}

template <typename T>
static int launch_push(cv_ctx *ctx, const PushArgs &a, int64_t total, const void *x, cudaStream_t st) {
  int64_t need = (total + 4 * CV_BLOCK - 1) / (4 * CV_BLOCK);
  int64_t cap = (int64_t)ctx->sms * 4;
  int grid = (int)(need < 1 ? 1 : (need < cap ? need : cap));
  k_halo_push<T><<<grid, CV_BLOCK, 0, st>>>(a, static_cast<const T *>(x));
  return cv_check_launch(ctx, "halo_push");
}

static int push_and_wait(cv_ctx *ctx, PushArgs &a, int64_t total, unsigned src_mask, bool cplx_, const void *x,
                         cudaStream_t st) {
  cv_peer_state *s = ctx->peer;
  a.seq = s->halo_seq;
  a.ticket = ctx->counters + CV_COUNTER_PUSH;
  if (a.nflag > 0 || a.nseg > 0) {
    if (cplx_)
      CV_TRY(launch_push<cplx>(ctx, a, total, x, st));
    else
      CV_TRY(launch_push<double>(ctx, a, total, x, st));
  }
  if (src_mask) {
    k_halo_wait<<<1, 32, 0, st>>>(s->pp.win[ctx->rank]->halo_flag, src_mask, s->halo_seq, ctx->scalars + CV_S_ERR);
    CV_TRY(cv_check_launch(ctx, "halo_wait"));
  }
  return CV_OK;
}

// general (SELL / CSR) halo: gather owned entries straight into the peers' halo buffers
int cv_halo_exchange_peer(cv_ctx *ctx, cv_op *op, bool cplx_, const void *x, cudaStream_t st) {
  cv_peer_state *s = ctx->peer;
  const unsigned long long seq = ++s->halo_seq;
  const int par = (int)(seq & 1ull);
  const size_t eb = cplx_ ? 16 : 8;
  PushArgs a;
  a.nseg = 0;
  a.nflag = 0;
  unsigned src_mask = 0;
  int64_t total = 0;
  for (int p = 0; p < ctx->world; ++p) {
    if (p == ctx->rank) continue;
    const int64_t ns = op->send_off[p + 1] - op->send_off[p];
    const int64_t nr = op->recv_off[p + 1] - op->recv_off[p];
    if (ns > 0) {
      PushSeg &g = a.seg[a.nseg++];
      g.dst = static_cast<char *>(op->peer_base[p]) + (size_t)par * op->peer_stride[p] + (size_t)op->peer_dst_off[p] * eb;
      g.idx = op->send_idx + op->send_off[p];
      g.src_start = 0;
      g.count = ns;
      total += ns;
      a.flag_dst[a.nflag++] = &s->pp.win[p]->halo_flag[ctx->rank];
    }
    if (nr > 0) src_mask |= 1u << p;
  }
  op->halo_cur = static_cast<char *>(op->peer_base[ctx->rank]) + (size_t)par * op->peer_stride[ctx->rank];
  return push_and_wait(ctx, a, total, src_mask, cplx_, x, st);
}

// DIA halo: contiguous ranges of x into the peers' lower / upper band buffers
int cv_halo_exchange_dia_peer(cv_ctx *ctx, cv_op *op, bool cplx_, const void *x, cudaStream_t st) {
  cv_peer_state *s = ctx->peer;
  const unsigned long long seq = ++s->halo_seq;
  const int par = (int)(seq & 1ull);
  const size_t eb = cplx_ ? 16 : 8;
  const size_t lo_b = dia_lo_bytes(op), par_b = lo_b + dia_hi_bytes(op);
  PushArgs a;
  a.nseg = 0;
  a.nflag = 0;
  unsigned src_mask = 0, dst_mask = 0;
  int64_t total = 0;
  for (const auto &r : op->dia_send) {
    CV_REQUIRE(a.nseg < 2 * CV_MAX_WORLD, "DIA halo: too many send ranges");
    PushSeg &g = a.seg[a.nseg++];
    g.dst = static_cast<char *>(op->peer_base[r.peer]) + (size_t)par * par_b + (r.band ? lo_b : 0) + (size_t)r.dst_off * eb;
    g.idx = nullptr;
    g.src_start = r.start;
    g.count = r.count;
    total += r.count;
    if (!((dst_mask >> r.peer) & 1u)) {
      dst_mask |= 1u << r.peer;
      a.flag_dst[a.nflag++] = &s->pp.win[r.peer]->halo_flag[ctx->rank];
    }
  }
  for (const auto &r : op->dia_recv_lo) src_mask |= 1u << r.peer;
  for (const auto &r : op->dia_recv_hi) src_mask |= 1u << r.peer;
  char *own = static_cast<char *>(op->peer_base[ctx->rank]) + (size_t)par * par_b;
  op->halo_lo_cur = own;
  op->halo_hi_cur = own + lo_b;
  return push_and_wait(ctx, a, total, src_mask, cplx_, x, st);
}
