// internal.h — declarations shared by the translation units of libcudavec
#pragma once
#include <vector>
#include "common.cuh"
#include "kernels_vec.cuh"
#include "kernels_spmv.cuh"
#include "kernels_dia.cuh"
#include "kernels_kron.cuh"
#include "kernels_orth.cuh"
#include "kernels_batch.cuh"

struct cv_op {
  uint64_t id = 0;  // unique per created operator (recycling state is keyed on it)
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  // CSR (borrowed device arrays)
  const int64_t *indptr = nullptr;
  const int32_t *indices = nullptr;
  const double *data = nullptr;
  const double *data_im = nullptr;  // imaginary parts of a complex-valued H (CSR only, cv_op_set_imag)
  int csr_group = 8;
  // SELL-32 (borrowed device arrays)
  int64_t n_slices = 0, padded_nnz = 0;
  const int64_t *slice_ptr = nullptr;
  const int32_t *sell_col = nullptr;
  const double *sell_val = nullptr;
  int fmt = CV_FMT_CSR;
  // DIA (borrowed device array)
  int n_diag = 0;
  int dia_off[CV_MAX_DIAG];
  const double *dia_val = nullptr;
  int64_t dia_ld = 0;
  int lo_len = 0, hi_len = 0;          // band below / above the owned rows (sharded mode)
  void *halo_lo = nullptr, *halo_hi = nullptr;
  // start: local row (send) / buffer slot (recv); band/dst_off: where a SENT range lands at the
  // peer (0 = its lower band buffer, 1 = upper) — used by the peer-memory push
  struct Range { int peer; int64_t start, count; int band; int64_t dst_off; };
  std::vector<Range> dia_send, dia_recv_lo, dia_recv_hi;
  int64_t row0 = 0, n_global = 0;
  // matrix-free Kronecker-sum operator (kernels_kron.cuh): geometry + terms; tables are borrowed device arrays
  struct Kron {
    int ndim = 0, nterm = 0, tab_len = 0, dtab_len = 0;
    int dims[KR_MAX_DIM], dtab_off[KR_MAX_DIM], shift[KR_MAX_DIM];
    long long stride[KR_MAX_DIM];
    unsigned long long magic[KR_MAX_DIM];
    KronTerm term[KR_MAX_TERMS];
    const double *tab_val = nullptr, *dtab = nullptr;
    const int *tab_col = nullptr;
  } kron;
  // row-sharded mode: columns >= n_cols - n_halo address the halo buffer
  int64_t n_halo = 0;
  const int32_t *send_idx = nullptr;
  std::vector<int64_t> send_off, recv_off;
  void *sendbuf = nullptr, *halobuf = nullptr;
  // peer-memory transport: halo buffers live in IPC-exported allocations, double-buffered by the
  // parity of the exchange sequence number; *_cur point at the parity the next SpMV reads
  bool peer_halo = false;
  unsigned long long halo_count = 0;     // exchanges of THIS operator so far (parity of its halo buffers)
  std::vector<void *> peer_base;         // [world] base of every rank's halo allocation (own at [rank])
  std::vector<int64_t> peer_stride;      // [world] bytes between the two parities (general halo)
  std::vector<int64_t> peer_dst_off;     // [world] element offset of MY block inside peer p's halo
  void *halo_cur = nullptr, *halo_lo_cur = nullptr, *halo_hi_cur = nullptr;
  HaloWait wait = {nullptr, 0u, 0ull, nullptr, nullptr};  // what the next sharded SpMV polls (mask 0: nothing)
};

// formats whose off-block x entries live in the two contiguous band buffers (halo_lo / halo_hi)
inline bool cv_op_banded(const cv_op *op) { return op->fmt == CV_FMT_DIA || op->fmt == CV_FMT_KRON; }

int cv_check_launch(cv_ctx *ctx, const char *what);

// device-resident variants: results land in ctx->scalars[slot ...) (already summed over ranks)
int cv_dot_dev(cv_ctx *ctx, int64_t n, int cplx_, int conj, const void *x, const void *y, int slot,
               cudaStream_t st);
int cv_nrm2sq_dev(cv_ctx *ctx, int64_t n, int cplx_, const void *x, int slot, cudaStream_t st);
int cv_scale_dev(cv_ctx *ctx, int64_t n, int cplx_, void *x, int slot, int mode, cudaStream_t st);
int cv_occ_grid(cv_ctx *ctx, const void *kernel, int64_t work_items, int items_per_cta);
int cv_lincomb_launch(cv_ctx *ctx, int64_t n, int v_cplx, int c_cplx, int m, const void *const *v,
                      int ncol, const double *coef, int ldc, int col0, void *const *y, int norm_slot,
                      cudaStream_t st);
int cv_tsdot_dev(cv_ctx *ctx, int64_t n, int cplx_, int conj, int m, const void *const *v, int b,
                 const void *const *w, int slot, cudaStream_t st, int gate_slot = -1);
int cv_tsupdate_dev(cv_ctx *ctx, int64_t n, int cplx_, int m, const void *const *v, int h_slot,
                    void *w, int norm_slot, cudaStream_t st, int gate_slot = -1);
int cv_spmv_dev(cv_ctx *ctx, cv_op *op, int cplx_, int mode, double sre, double sim, const void *x,
                void *y, double alpha, double beta1, const void *u1, bool epi, int dots_slot,
                cudaStream_t st);
int cv_halo_exchange(cv_ctx *ctx, cv_op *op, bool cplx_, const void *x, cudaStream_t st);
int cv_halo_exchange_dia(cv_ctx *ctx, cv_op *op, bool cplx_, const void *x, cudaStream_t st);
int cv_halo_exchange_peer(cv_ctx *ctx, cv_op *op, bool cplx_, const void *x, cudaStream_t st);
int cv_peer_plan_exchange(cv_ctx *ctx, cv_op *op, bool cplx_, PushArgs *a, int64_t *total_out);
const PeerPtrs *cv_peer_ptrs(cv_ctx *ctx);
int cv_wait_mailbox(cv_ctx *ctx, unsigned long long seq, cudaStream_t st);
int cv_orth_step_dev(cv_ctx *ctx, cv_op *op, int64_t n, int cplx_, int m, const void *const *basis, void *w,
                     int s_flag, int s_h2, int s_lag, double eta, cudaStream_t st, bool *fused);
int cv_peer_detach(cv_ctx *ctx);
int cv_peer_allreduce(cv_ctx *ctx, double *buf_dev, int count, cudaStream_t st);
