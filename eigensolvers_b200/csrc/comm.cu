// comm.cu — row-sharded mode: NCCL plumbing (one process per GPU, NVLink/NVSwitch), the halo
// exchange of x that precedes every sharded SpMV, the scalar all-reduce that follows every
// reduction, and the host-side integer routines that build the partition and halo maps
// (bit-exact against the scipy slicing oracle; the reference has no partitioner).
//
// NCCL is resolved at run time from the libnccl.so.2 that torch already loaded in this
// process, so the library loads (and single-GPU mode works) without NCCL on the link line.
#include <dlfcn.h>
#include <algorithm>
#include "internal.h"

typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclFloat64_ = 8 };  // ncclDataType_t: ncclDouble
enum { ncclSum_ = 0 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.handle) return CV_OK;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // torch's copy, if already mapped
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW);
  if (!h) {
    cv_set_error("NCCL not found: %s", dlerror());
    return CV_ERR_COMM;
  }
#define SYM(field, name)                                  \
  *(void **)(&g_nccl.field) = dlsym(h, name);             \
  if (!g_nccl.field) {                                    \
    cv_set_error("NCCL symbol %s missing", name);         \
    return CV_ERR_COMM;                                   \
  }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.handle = h;
  return CV_OK;
}

#define CV_NCCL(call)                                                                         \
  do {                                                                                        \
    int r__ = (call);                                                                         \
    if (r__ != ncclSuccess_) {                                                                \
      cv_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));  \
      return CV_ERR_COMM;                                                                     \
    }                                                                                         \
  } while (0)

struct cv_comm_state {
  ncclComm_t comm = nullptr;
};

extern "C" int cv_comm_unique_id(void *id128) {
  CV_REQUIRE(id128, "cv_comm_unique_id: null argument");
  CV_TRY(nccl_load());
  ncclUniqueId id;
  CV_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return CV_OK;
}

extern "C" int cv_comm_init(cv_ctx *ctx, const void *id128, int rank, int world) {
  CV_REQUIRE(ctx && id128 && world >= 1 && rank >= 0 && rank < world, "cv_comm_init: bad argument");
  ctx->rank = rank;
  ctx->world = world;
  if (world == 1) return CV_OK;
  CV_TRY(nccl_load());
  CV_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  cv_comm_state *s = new cv_comm_state();
  CV_NCCL(g_nccl.CommInitRank(&s->comm, world, id, rank));
  ctx->comm = s;
  return CV_OK;
}

extern "C" int cv_comm_init_peer_only(cv_ctx *ctx, int rank, int world) {
  CV_REQUIRE(ctx && world >= 1 && rank >= 0 && rank < world, "cv_comm_init_peer_only: bad argument");
  CV_REQUIRE(world <= CV_MAX_WORLD, "cv_comm_init_peer_only: world=%d exceeds %d", world, CV_MAX_WORLD);
  ctx->rank = rank;
  ctx->world = world;
  return CV_OK;
}

extern "C" int cv_comm_finalize(cv_ctx *ctx) {
  if (ctx && ctx->comm) {
    if (ctx->comm->comm) g_nccl.CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
  }
  if (ctx && ctx->peer) cv_peer_detach(ctx);
  if (ctx) {
    ctx->world = 1;
    ctx->rank = 0;
  }
  return CV_OK;
}

extern "C" int cv_comm_allreduce(cv_ctx *ctx, double *buf_dev, int count, void *stream) {
  CV_REQUIRE(ctx && buf_dev && count >= 0, "cv_comm_allreduce: bad argument");
  if (ctx->world == 1 || count == 0) return CV_OK;
  if (ctx->peer) return cv_peer_allreduce(ctx, buf_dev, count, (cudaStream_t)stream);
  CV_REQUIRE(ctx->comm, "cv_comm_allreduce: neither a peer-memory window nor an NCCL communicator is attached");
  CV_NCCL(g_nccl.AllReduce(buf_dev, buf_dev, (size_t)count, ncclFloat64_, ncclSum_, ctx->comm->comm,
                           (cudaStream_t)stream));
  return CV_OK;
}

int cv_reduce_ranks(cv_ctx *ctx, int offset, int count, cudaStream_t st) {
  if (ctx->world == 1 || ctx->defer_reduce) return CV_OK;
  return cv_comm_allreduce(ctx, ctx->scalars + offset, count, (void *)st);
}

// ------------------------------------------------------------------------------------------
// halo exchange: pack owned entries -> grouped send/recv over NVLink -> halo buffer
// ------------------------------------------------------------------------------------------
extern "C" int cv_op_set_halo(cv_ctx *ctx, cv_op *op, int64_t n_halo, const int32_t *send_idx_dev,
                              const int64_t *send_off, const int64_t *recv_off, void *sendbuf_dev,
                              void *halobuf_dev) {
  CV_REQUIRE(ctx && op && send_off && recv_off, "cv_op_set_halo: null argument");
  CV_REQUIRE(n_halo >= 0 && n_halo <= op->n_cols, "cv_op_set_halo: n_halo out of range");
  op->n_halo = n_halo;
  op->send_idx = send_idx_dev;
  op->send_off.assign(send_off, send_off + ctx->world + 1);
  op->recv_off.assign(recv_off, recv_off + ctx->world + 1);
  CV_REQUIRE(op->recv_off[ctx->world] == n_halo, "cv_op_set_halo: recv offsets do not sum to n_halo");
  op->sendbuf = sendbuf_dev;
  op->halobuf = halobuf_dev;
  return CV_OK;
}

int cv_halo_exchange(cv_ctx *ctx, cv_op *op, bool cplx_, const void *x, cudaStream_t st) {
  if (op->n_halo == 0 && (op->send_off.empty() || op->send_off.back() == 0)) return CV_OK;
  if (op->peer_halo) return cv_halo_exchange_peer(ctx, op, cplx_, x, st);
  CV_REQUIRE(ctx->world > 1 && ctx->comm, "halo exchange without a communicator");
  const int64_t n_send = op->send_off[ctx->world];
  if (n_send > 0) {
    int grid = cv_grid_for(ctx, n_send, CV_BLOCK);
    if (cplx_)
      k_pack<cplx><<<grid, CV_BLOCK, 0, st>>>(n_send, op->send_idx, (const cplx *)x, (cplx *)op->sendbuf);
    else
      k_pack<double><<<grid, CV_BLOCK, 0, st>>>(n_send, op->send_idx, (const double *)x, (double *)op->sendbuf);
    CV_TRY(cv_check_launch(ctx, "pack"));
  }
  const size_t w = cplx_ ? 2 : 1;  // doubles per element
  CV_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < ctx->world; ++p) {
    if (p == ctx->rank) continue;
    int64_t ns = op->send_off[p + 1] - op->send_off[p];
    int64_t nr = op->recv_off[p + 1] - op->recv_off[p];
    if (ns > 0)
      CV_NCCL(g_nccl.Send((const double *)op->sendbuf + op->send_off[p] * w, (size_t)ns * w, ncclFloat64_,
                          p, ctx->comm->comm, st));
    if (nr > 0)
      CV_NCCL(g_nccl.Recv((double *)op->halobuf + op->recv_off[p] * w, (size_t)nr * w, ncclFloat64_, p,
                          ctx->comm->comm, st));
  }
  CV_NCCL(g_nccl.GroupEnd());
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// DIA halo: the band below / above the owned rows arrives as contiguous ranges (no pack kernel)
// ------------------------------------------------------------------------------------------
// Pure host integer routine: the exchange plan of the DIA band halo for `rank`.  A rank needs the
// rows [r0 - lo_len, r0) (lower band) and [r1, r1 + hi_len) (upper band), clipped to [0, N); every
// contiguous piece owned by one peer is one range.  send: what this rank stores into its peers
// ({peer, first local row, count, band of the PEER it lands in (0 lower / 1 upper), slot inside that
// band buffer}); recv: what arrives here ({peer, band, first slot, count}).
static void dia_plan(const int64_t *offsets, int P, int me, int64_t lo_len, int64_t hi_len,
                     std::vector<cv_op::Range> &send, std::vector<cv_op::Range> &recv_lo,
                     std::vector<cv_op::Range> &recv_hi) {
  const int64_t N = offsets[P], r0 = offsets[me], r1 = offsets[me + 1];
  send.clear();
  recv_lo.clear();
  recv_hi.clear();
  auto overlap = [](int64_t a0, int64_t a1, int64_t b0, int64_t b1, int64_t &s, int64_t &c) {
    s = a0 > b0 ? a0 : b0;
    int64_t e = a1 < b1 ? a1 : b1;
    c = e > s ? e - s : 0;
  };
  for (int p = 0; p < P; ++p) {
    if (p == me) continue;
    const int64_t p0 = offsets[p], p1 = offsets[p + 1];
    int64_t s, c;
    // what p needs from my rows: first the part of p's lower band, then of its upper band
    overlap(p0 - lo_len < 0 ? 0 : p0 - lo_len, p0, r0, r1, s, c);
    if (c > 0) send.push_back({p, s - r0, c, 0, s - (p0 - lo_len)});
    overlap(p1, p1 + hi_len > N ? N : p1 + hi_len, r0, r1, s, c);
    if (c > 0) send.push_back({p, s - r0, c, 1, s - p1});
    // what I need from p's rows
    overlap(r0 - lo_len < 0 ? 0 : r0 - lo_len, r0, p0, p1, s, c);
    if (c > 0) recv_lo.push_back({p, s - (r0 - lo_len), c, 0, 0});
    overlap(r1, r1 + hi_len > N ? N : r1 + hi_len, p0, p1, s, c);
    if (c > 0) recv_hi.push_back({p, s - r1, c, 1, 0});
  }
}

extern "C" int cv_dia_halo_plan(const int64_t *offsets, int world, int rank, int64_t lo_len, int64_t hi_len,
                                int cap, int *n_send, int64_t *send5, int *n_recv, int64_t *recv4) {
  CV_REQUIRE(offsets && n_send && send5 && n_recv && recv4, "cv_dia_halo_plan: null argument");
  CV_REQUIRE(world >= 1 && rank >= 0 && rank < world && lo_len >= 0 && hi_len >= 0 && cap >= 0,
             "cv_dia_halo_plan: bad argument");
  std::vector<cv_op::Range> send, rlo, rhi;
  dia_plan(offsets, world, rank, lo_len, hi_len, send, rlo, rhi);
  CV_REQUIRE((int)send.size() <= cap && (int)(rlo.size() + rhi.size()) <= cap, "cv_dia_halo_plan: capacity %d too small", cap);
  *n_send = (int)send.size();
  for (size_t i = 0; i < send.size(); ++i) {
    send5[5 * i + 0] = send[i].peer;
    send5[5 * i + 1] = send[i].start;
    send5[5 * i + 2] = send[i].count;
    send5[5 * i + 3] = send[i].band;
    send5[5 * i + 4] = send[i].dst_off;
  }
  int k = 0;
  for (const auto *list : {&rlo, &rhi})
    for (const auto &r : *list) {
      recv4[4 * k + 0] = r.peer;
      recv4[4 * k + 1] = r.band;
      recv4[4 * k + 2] = r.start;
      recv4[4 * k + 3] = r.count;
      ++k;
    }
  *n_recv = k;
  return CV_OK;
}

extern "C" int cv_op_set_dia_halo(cv_ctx *ctx, cv_op *op, const int64_t *offsets, void *halo_lo_dev,
                                  void *halo_hi_dev) {
  CV_REQUIRE(ctx && op && offsets, "cv_op_set_dia_halo: null argument");
  CV_REQUIRE(op->dia_val || op->fmt == CV_FMT_KRON, "cv_op_set_dia_halo: operator has neither DIA storage nor a Kronecker form");
  CV_REQUIRE((op->lo_len == 0 || halo_lo_dev) && (op->hi_len == 0 || halo_hi_dev), "cv_op_set_dia_halo: null halo buffer");
  op->halo_lo = halo_lo_dev;
  op->halo_hi = halo_hi_dev;
  op->n_global = offsets[ctx->world];
  dia_plan(offsets, ctx->world, ctx->rank, op->lo_len, op->hi_len, op->dia_send, op->dia_recv_lo, op->dia_recv_hi);
  return CV_OK;
}

int cv_halo_exchange_dia(cv_ctx *ctx, cv_op *op, bool cplx_, const void *x, cudaStream_t st) {
  if (op->peer_halo) return cv_halo_exchange_peer(ctx, op, cplx_, x, st);
  CV_REQUIRE(ctx->world > 1 && ctx->comm, "halo exchange without a communicator");
  CV_REQUIRE(op->n_global > 0, "DIA operator has no exchange plan (cv_op_set_dia_halo)");
  const size_t w = cplx_ ? 2 : 1;
  CV_NCCL(g_nccl.GroupStart());
  // per peer the matching order is: sends (for its lower band, then upper band) / receives (my
  // lower band, then upper band); both sides enumerate peers and bands in the same order
  for (int p = 0; p < ctx->world; ++p) {
    if (p == ctx->rank) continue;
    for (const auto &r : op->dia_send)
      if (r.peer == p)
        CV_NCCL(g_nccl.Send((const double *)x + r.start * w, (size_t)r.count * w, ncclFloat64_, p, ctx->comm->comm, st));
    for (const auto &r : op->dia_recv_lo)
      if (r.peer == p)
        CV_NCCL(g_nccl.Recv((double *)op->halo_lo + r.start * w, (size_t)r.count * w, ncclFloat64_, p, ctx->comm->comm, st));
    for (const auto &r : op->dia_recv_hi)
      if (r.peer == p)
        CV_NCCL(g_nccl.Recv((double *)op->halo_hi + r.start * w, (size_t)r.count * w, ncclFloat64_, p, ctx->comm->comm, st));
  }
  CV_NCCL(g_nccl.GroupEnd());
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// host-side integer routines: row partition and halo maps
// ------------------------------------------------------------------------------------------
extern "C" int cv_partition_rows(int64_t n, int world, int64_t *offsets) {
  CV_REQUIRE(n >= 0 && world >= 1 && offsets, "cv_partition_rows: bad argument");
  for (int p = 0; p <= world; ++p) offsets[p] = (int64_t)(((__int128)p * n) / world);
  return CV_OK;
}

static void halo_collect(const int64_t *indptr, const int32_t *indices, int64_t row0, int64_t row1,
                         std::vector<int32_t> &halo) {
  halo.clear();
  for (int64_t k = indptr[row0]; k < indptr[row1]; ++k) {
    int32_t c = indices[k];
    if (c < row0 || c >= row1) halo.push_back(c);
  }
  std::sort(halo.begin(), halo.end());
  halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
}

extern "C" int cv_halo_count(const int64_t *indptr, const int32_t *indices, int64_t row0, int64_t row1,
                             int64_t *n_halo) {
  CV_REQUIRE(indptr && n_halo && row0 >= 0 && row1 >= row0, "cv_halo_count: bad argument");
  std::vector<int32_t> halo;
  halo_collect(indptr, indices, row0, row1, halo);
  *n_halo = (int64_t)halo.size();
  return CV_OK;
}

extern "C" int cv_halo_build(const int64_t *indptr, const int32_t *indices, int64_t row0, int64_t row1,
                             const int64_t *offsets, int world, int64_t n_halo, int32_t *halo_cols,
                             int32_t *halo_owner, int64_t *local_indptr, int32_t *local_indices) {
  CV_REQUIRE(indptr && offsets && local_indptr && row0 >= 0 && row1 >= row0 && world >= 1,
             "cv_halo_build: bad argument");
  std::vector<int32_t> halo;
  halo_collect(indptr, indices, row0, row1, halo);
  CV_REQUIRE((int64_t)halo.size() == n_halo, "cv_halo_build: n_halo=%lld but %zu halo columns found",
             (long long)n_halo, halo.size());
  for (int64_t h = 0; h < n_halo; ++h) {
    halo_cols[h] = halo[h];
    // owner p: offsets[p] <= c < offsets[p+1]   (np.searchsorted(offsets, c, 'right') - 1)
    const int64_t *ub = std::upper_bound(offsets, offsets + world + 1, (int64_t)halo[h]);
    halo_owner[h] = (int32_t)(ub - offsets) - 1;
  }
  const int64_t nloc = row1 - row0, base = indptr[row0];
  for (int64_t r = 0; r <= nloc; ++r) local_indptr[r] = indptr[row0 + r] - base;
  for (int64_t k = indptr[row0]; k < indptr[row1]; ++k) {
    int32_t c = indices[k];
    if (c >= row0 && c < row1) {
      local_indices[k - base] = (int32_t)(c - row0);
    } else {
      int64_t pos = std::lower_bound(halo.begin(), halo.end(), c) - halo.begin();
      local_indices[k - base] = (int32_t)(nloc + pos);
    }
  }
  return CV_OK;
}
