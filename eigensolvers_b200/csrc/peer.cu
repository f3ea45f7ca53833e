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

extern "C" size_t cv_peer_window_bytes(void) { return sizeof(PeerWindow); }

// stand-alone all-reduce (reductions outside the fused Arnoldi step): one CTA
__global__ void __launch_bounds__(256) k_peer_allreduce(const __grid_constant__ PeerPtrs pp, int me, int world,
                                                        double *buf, int count, double *err) {
  cta_peer_allreduce(pp, me, world, buf, count, err);
}

// stand-alone halo push (SpMV whose input was not produced by the fused Arnoldi step)
template <typename T>
__global__ void __launch_bounds__(CV_BLOCK) k_halo_push(const __grid_constant__ PushArgs a, const T *__restrict__ x) {
  grid_halo_push<T>(a, x);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct cv_peer_state {
  PeerPtrs pp;
  unsigned long long halo_seq = 0;  // all ranks issue the same sequence of halo exchanges
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
    k_peer_allreduce<<<1, 256, 0, st>>>(s->pp, ctx->rank, ctx->world, buf_dev + c0, c, ctx->scalars + CV_S_ERR);
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

const PeerPtrs *cv_peer_ptrs(cv_ctx *ctx) { return ctx->peer ? &ctx->peer->pp : nullptr; }

// Plan the next halo exchange of `op`: advances the exchange sequence, selects the parity buffers
// the next SpMV reads (op->halo_*_cur), fills the push description (what THIS rank stores into its
// peers) and the wait description (which flags the SpMV polls).  The caller launches the push
// (k_halo_push, or phase C of the fused Arnoldi step).
int cv_peer_plan_exchange(cv_ctx *ctx, cv_op *op, bool cplx_, PushArgs *a, int64_t *total_out) {
  cv_peer_state *s = ctx->peer;
  CV_REQUIRE(s && op->peer_halo, "peer halo exchange without peer buffers");
  const unsigned long long seq = ++s->halo_seq;   // flags: one sequence for all operators (all ranks issue the same)
  const int par = (int)(++op->halo_count & 1ull);  // buffers: double-buffered PER OPERATOR, so two operators
                                                   // used alternately (H, Hsolve) still alternate parities
  const size_t eb = cplx_ ? 16 : 8;
  a->nseg = 0;
  a->nflag = 0;
  a->seq = seq;
  a->ticket = ctx->counters + CV_COUNTER_PUSH;
  // Flags travel in BOTH directions between any two ranks that exchange data in EITHER direction:
  // a sender may only start exchange k+2 (which overwrites the parity buffer of exchange k at its
  // peer) after that peer has published exchange k+1, i.e. after the peer's SpMV k has finished
  // (stream order).  With flags only along the data this held for symmetric patterns alone; a
  // structurally one-sided coupling (or an asymmetric band) would let the sender run ahead.
  unsigned send_mask = 0, recv_mask = 0;
  int64_t total = 0;
  if (cv_op_banded(op)) {
    const size_t lo_b = dia_lo_bytes(op), par_b = lo_b + dia_hi_bytes(op);
    for (const auto &r : op->dia_send) {
      CV_REQUIRE(a->nseg < 2 * CV_MAX_WORLD, "DIA halo: too many send ranges");
      PushSeg &g = a->seg[a->nseg++];
      g.dst = static_cast<char *>(op->peer_base[r.peer]) + (size_t)par * par_b + (r.band ? lo_b : 0) +
              (size_t)r.dst_off * eb;
      g.idx = nullptr;
      g.src_start = r.start;
      g.count = r.count;
      total += r.count;
      send_mask |= 1u << r.peer;
    }
    for (const auto &r : op->dia_recv_lo) recv_mask |= 1u << r.peer;
    for (const auto &r : op->dia_recv_hi) recv_mask |= 1u << r.peer;
    char *own = static_cast<char *>(op->peer_base[ctx->rank]) + (size_t)par * par_b;
    op->halo_lo_cur = own;
    op->halo_hi_cur = own + lo_b;
  } else {
    for (int p = 0; p < ctx->world; ++p) {
      if (p == ctx->rank) continue;
      const int64_t ns = op->send_off[p + 1] - op->send_off[p];
      const int64_t nr = op->recv_off[p + 1] - op->recv_off[p];
      if (ns > 0) {
        PushSeg &g = a->seg[a->nseg++];
        g.dst = static_cast<char *>(op->peer_base[p]) + (size_t)par * op->peer_stride[p] +
                (size_t)op->peer_dst_off[p] * eb;
        g.idx = op->send_idx + op->send_off[p];
        g.src_start = 0;
        g.count = ns;
        total += ns;
        send_mask |= 1u << p;
      }
      if (nr > 0) recv_mask |= 1u << p;
    }
    op->halo_cur = static_cast<char *>(op->peer_base[ctx->rank]) + (size_t)par * op->peer_stride[ctx->rank];
  }
  const unsigned src_mask = send_mask | recv_mask;
  for (int p = 0; p < ctx->world; ++p)
    if ((src_mask >> p) & 1u) a->flag_dst[a->nflag++] = &s->pp.win[p]->halo_flag[ctx->rank];
  op->wait.flags = s->pp.win[ctx->rank]->halo_flag;
  op->wait.mask = src_mask;
  op->wait.seq = seq;
  op->wait.err = ctx->scalars + CV_S_ERR;
  op->wait.scale_sq = nullptr;
  if (total_out) *total_out = total;
  return CV_OK;
}

// plan + stand-alone push kernel; the SpMV that follows waits for the peers' flags in its prologue
int cv_halo_exchange_peer(cv_ctx *ctx, cv_op *op, bool cplx_, const void *x, cudaStream_t st) {
  if (ctx->prepushed_x == x && ctx->prepushed_op == op) {  // phase C of the Arnoldi step already pushed x
    ctx->prepushed_x = nullptr;
    return CV_OK;
  }
  ctx->prepushed_x = nullptr;
  PushArgs a;
  int64_t total = 0;
  CV_TRY(cv_peer_plan_exchange(ctx, op, cplx_, &a, &total));
  if (a.nseg > 0 || a.nflag > 0) {
    if (cplx_)
      CV_TRY(launch_push<cplx>(ctx, a, total, x, st));
    else
      CV_TRY(launch_push<double>(ctx, a, total, x, st));
  }
  return CV_OK;
}
