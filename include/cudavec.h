/*
 * cudavec.h — C ABI of libcudavec.so, the B200 (sm_100a) device back-end behind
 * `CudaVector`, the drop-in replacement for the reference's NumpyVector hot path.
 *
 * The reference (chem-rano/eigensolvers) has no FFI: its plugin boundary is the Python
 * ABC `AbstractVector` (abstractVector.py:15-169).  Every entry point below is the
 * device-side body of one call site of that ABC as implemented by numpyVector.py; the
 * Python class eigensolvers_b200/cudaVector.py binds them with ctypes (INTEGRATION.md
 * shows the binding a maintainer of the reference would add).
 *
 * Conventions
 *   - plain C types only; every function returns CV_OK (0) or a CV_ERR_* code and
 *     cv_last_error() then holds a message.  No exceptions cross the ABI.
 *   - all `*_dev` / vector pointers are DEVICE pointers (torch.Tensor.data_ptr()); the
 *     library never allocates vector memory, the host language owns lifetimes.
 *   - `cplx` flags: 0 = float64, 1 = complex128 stored interleaved (re,im) like numpy.
 *   - `n` counts ELEMENTS (complex elements when cplx=1).
 *   - scalar results are returned to HOST pointers; such calls synchronise `stream`.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - distributed mode (cv_comm_*): vectors are the local row shard; every scalar result
 *     is summed over ranks before it is returned, so all ranks see identical scalars.
 */
#ifndef CUDAVEC_H
#define CUDAVEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CV_ABI_VERSION 2

#define CV_OK 0
#define CV_ERR_CUDA 1        /* a CUDA runtime call failed                              */
#define CV_ERR_ARG 2         /* invalid argument                                        */
#define CV_ERR_UNSUPPORTED 3 /* combination not implemented                             */
#define CV_ERR_NUMERIC 4     /* non-finite value met inside a solver (scipy LinAlgError) */
#define CV_ERR_COMM 5        /* NCCL / peer-memory transport failure                    */

/* spmv modes */
#define CV_SPMV_PLAIN 0   /* y = H x            numpyVector.py:98-100 (applyOp)          */
#define CV_SPMV_SHIFT 1   /* y = sigma x - H x  numpyVector.py:152 (solve, Green's fn)   */
#define CV_SPMV_RSHIFT 2  /* y = H x - sigma x  numpyVector.py:154 (reverseGF)           */

/* operator storage formats */
#define CV_FMT_CSR 0
#define CV_FMT_SELL 1     /* sliced ELL, slice height 32 */
#define CV_FMT_DIA 2      /* one dense value stream per distinct column offset (banded structure) */
#define CV_FMT_KRON 3     /* matrix-free sum of Kronecker products of small 1-D matrices           */

/* linear solvers (numpyVector.py:160-163) */
#define CV_SOLVER_GCROTMK 0
#define CV_SOLVER_MINRES 1

typedef struct cv_ctx cv_ctx; /* per-device context: reduction scratch, pinned mailbox, counters */
typedef struct cv_op cv_op;   /* device-resident Hamiltonian                                     */

int cv_abi_version(void);
const char *cv_last_error(void);

/* ---- context ------------------------------------------------------------------------ */
/* Bytes of device scratch the caller must hand to cv_ctx_create (partial sums, scalars). */
size_t cv_ctx_scratch_bytes(void);
int cv_ctx_create(int device, void *scratch_dev, size_t scratch_bytes, cv_ctx **out);
int cv_ctx_destroy(cv_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int cv_ctx_launch_count(cv_ctx *ctx, uint64_t *count);
int cv_ctx_sm_count(cv_ctx *ctx, int *sms);
/* GCROT recycling (scipy's CU= argument, _gcrotmk.py:227-236, which the reference leaves unused):
 * when enabled, the (c,u) pairs a solve leaves in the workspace are reused by the next cv_solve
 * with the same operator, shift, vector type, workspace and (m,k).  Off by default.            */
int cv_ctx_set_recycle(cv_ctx *ctx, int enable);
/* GCROT's Arnoldi step repeats its Gram-Schmidt pass when the first one left less than eta of the
 * vector's norm (Daniel-Gragg-Kaufman-Stewart).  Default 0.1; 0 = never, > 1 = always twice.    */
int cv_ctx_set_reorth_eta(cv_ctx *ctx, double eta);
/* Tuning switches by name: "reorth_eta" (as above), "slab_mode" (dot-phase work split of the fused
 * Arnoldi step: 0 full slabs + remainder, 1 even slabs with balanced load batches), "push_early". */
int cv_ctx_set_option(cv_ctx *ctx, const char *name, double value);
/* Accumulated phase times (ns, as seen by CTA 0) of the fused Arnoldi-step kernel: [0] dots,
 * [1] barrier + all-reduce, [2] update/normalise/push, [3] barrier before a second pass,
 * [5] launches, [6] passes, [7] halo flags / late push.  Synchronises the device.             */
int cv_ctx_trace_read(cv_ctx *ctx, double *out16, int reset);
/* Optional kernel timing with CUDA events on the launching stream, per kernel class:
 *   0 fused SpMV   1 fused Arnoldi step / tall-skinny dot   2 tall-skinny update   3 other vector
 *   kernels   4 Gram-Schmidt against a set   5 linear combinations   6 (bytes only) the SpMV's
 *   CSR-equivalent bytes 12 nnz + 20 N, next to class 0's bytes of the stored format   7 spare.
 * cv_ctx_profile_read synchronises, returns for the first n_classes (<= 8) classes the accumulated
 * milliseconds, launch counts and ALGORITHMIC bytes (SURVEY 8d formulas) of the timed launches,
 * and resets the accumulators: achieved GB/s = bytes / ms.                                      */
int cv_ctx_profile(cv_ctx *ctx, int enable);
int cv_ctx_profile_read(cv_ctx *ctx, int n_classes, double *ms, uint64_t *count, double *bytes);

/* ---- distributed mode (row-sharded H) -------------------------------------- */
/* 128-byte NCCL unique id, created on rank 0 and broadcast by the host language. */
int cv_comm_unique_id(void *id128);
int cv_comm_init(cv_ctx *ctx, const void *id128, int rank, int world);
/* Row-sharded mode WITHOUT an NCCL communicator: every collective must then go through the
 * peer-memory transport (cv_comm_attach_peers).  For ranks that share one device (NCCL refuses
 * duplicate GPUs; CUDA IPC works between processes on the same device) — used by the 2-process
 * parity tests that run on a single GPU.                                                     */
int cv_comm_init_peer_only(cv_ctx *ctx, int rank, int world);
int cv_comm_finalize(cv_ctx *ctx);
/* Sum `count` doubles in place over all ranks (device buffer); no-op for world 1. */
int cv_comm_allreduce(cv_ctx *ctx, double *buf_dev, int count, void *stream);

/* Peer-memory transport (csrc/peer.cu): inside one NVSwitch node the scalar all-reduce and the
 * halo exchange run over CUDA-IPC-mapped peer memory (P2P stores + sequence flags over NVLink)
 * instead of NCCL calls.  The host language allocates exportable device memory with
 * cv_peer_alloc (handle64 = the 64-byte cudaIpcMemHandle_t to ship to the other ranks, e.g. with
 * torch.distributed.all_gather_object), maps the other ranks' allocations with cv_peer_open and
 * registers (a) one window of cv_peer_window_bytes() per rank with cv_comm_attach_peers
 * (window_ptrs[p] = rank p's window as mapped in THIS process, own allocation at [rank]) and
 * (b) per operator the halo buffers (cv_op_set_halo_peers / cv_op_set_dia_halo_peers).
 * cv_comm_transport: 0 = single GPU, 1 = NCCL, 2 = peer memory.                               */
size_t cv_peer_window_bytes(void);
int cv_peer_alloc(cv_ctx *ctx, size_t bytes, void **ptr_dev, void *handle64);
int cv_peer_open(cv_ctx *ctx, const void *handle64, void **ptr_dev);
int cv_peer_close(cv_ctx *ctx, void *ptr_dev);
int cv_peer_free(cv_ctx *ctx, void *ptr_dev);
int cv_comm_attach_peers(cv_ctx *ctx, void *const *window_ptrs /* world */);
int cv_comm_transport(cv_ctx *ctx, int *transport);

/* Host-side integer routines (bit-exact vs. the scipy slicing oracle).
 * cv_partition_rows: offsets[p] = floor(p*n/P), p = 0..P.
 * cv_halo_count / cv_halo_build: for the row block [row0,row1) of a CSR matrix with GLOBAL
 * column indices, the sorted unique off-block columns (the halo), their owner rank, and the
 * local CSR whose columns are renumbered to [0,nloc) (owned) ++ [nloc, nloc+nhalo) (halo).   */
int cv_partition_rows(int64_t n, int world, int64_t *offsets /* world+1 */);
int cv_halo_count(const int64_t *indptr, const int32_t *indices, int64_t row0, int64_t row1,
                  int64_t *n_halo);
int cv_halo_build(const int64_t *indptr, const int32_t *indices, int64_t row0, int64_t row1,
                  const int64_t *offsets, int world, int64_t n_halo,
                  int32_t *halo_cols /* n_halo, sorted global ids */,
                  int32_t *halo_owner /* n_halo */,
                  int64_t *local_indptr /* row1-row0+1 */,
                  int32_t *local_indices /* nnz of the block */);
/* Register the exchange plan of this rank's operator: for each peer the owned local rows to
 * send (device int32 list, concatenated, send_off[world+1]) and the halo slots received
 * (contiguous per owner, recv_off[world+1]).                                               */
int cv_op_set_halo(cv_ctx *ctx, cv_op *op, int64_t n_halo, const int32_t *send_idx_dev,
                   const int64_t *send_off /* host, world+1 */,
                   const int64_t *recv_off /* host, world+1 */, void *sendbuf_dev,
                   void *halobuf_dev /* each 16*max(n_send,n_halo) bytes */);

/* Peer-memory halo of a general (SELL/CSR) operator, after cv_op_set_halo: halo_base[p] = rank
 * p's halo allocation (2 parities of parity_stride_bytes[p] each; 16 bytes per halo slot),
 * dst_off_elems[p] = first slot of MY block inside rank p's halo (p's recv_off[my rank]).      */
int cv_op_set_halo_peers(cv_ctx *ctx, cv_op *op, void *const *halo_base /* world */,
                         const int64_t *parity_stride_bytes /* world */,
                         const int64_t *dst_off_elems /* world */);

/* ---- operator ------------------------------------------------------------------------- */
/* Borrow a CSR matrix already resident on the device (int64 indptr, int32 column indices,
 * float64 values; numpyVector.py:100 takes scipy.sparse / ndarray H).                      */
int cv_op_create_csr(cv_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                     const int64_t *indptr_dev, const int32_t *indices_dev,
                     const double *data_dev, cv_op **out);
int cv_op_destroy(cv_op *op);
/* sliced-ELL (SELL-32x2): (1) per-slice widths (max row length of each 32-row slice) into
 * widths_dev, (2) the caller rounds them up to EVEN numbers, scans 32*width into slice_ptr (int64,
 * n_slices+1, element offsets) and supplies storage (values 16-byte, columns 8-byte aligned),
 * (3) the fill kernel lays every slice out in column PAIRS: entry (row r, column 2p+e) at
 * slice_ptr[s] + (p*32 + r)*2 + e, so the SpMV reads two values per 128-bit load.            */
int cv_op_sell_widths(cv_ctx *ctx, cv_op *op, int32_t *widths_dev, void *stream);
int cv_op_attach_sell(cv_ctx *ctx, cv_op *op, const int64_t *slice_ptr_dev, int64_t padded_nnz,
                      int32_t *sell_col_dev, double *sell_val_dev, void *stream);
/* Diagonal storage for operators whose entries all lie on n_diag (<= 64) distinct offsets
 * col-row (sorted ascending, host array).  dia_val_dev is caller storage of n_diag*ld doubles,
 * zero-filled; the fill kernel scatters the CSR values into it.  col_global_dev (may be NULL =
 * the operator's own column array) holds GLOBAL column ids and row0 the first global row, so the
 * call also works for a rank's row block.  *ok_host = 0 if some entry is off the given diagonals
 * (the operator then keeps its previous format).                                               */
int cv_op_attach_dia(cv_ctx *ctx, cv_op *op, int n_diag, const int32_t *offsets_host,
                     const int32_t *col_global_dev, int64_t row0, double *dia_val_dev, int64_t ld,
                     int *ok_host, void *stream);
/* Row-sharded DIA: the band below/above the owned block is received into two contiguous buffers
 * (16*lo_len and 16*hi_len bytes, lo_len = max(0,-min offset), hi_len = max(0,max offset)) by
 * contiguous range sends derived from the partition `offsets` (host, world+1).                   */
int cv_op_set_dia_halo(cv_ctx *ctx, cv_op *op, const int64_t *offsets, void *halo_lo_dev,
                       void *halo_hi_dev);
/* The same plan as a pure host routine (integer outputs for the bit-exact check against the numpy
 * oracle): send5[i] = {peer, first local row, count, band at the peer (0 lower, 1 upper), slot in
 * that band buffer}, recv4[i] = {peer, band, first slot, count} (lower band first); cap = capacity
 * of both arrays in ranges.                                                                      */
int cv_dia_halo_plan(const int64_t *offsets, int world, int rank, int64_t lo_len, int64_t hi_len,
                     int cap, int *n_send, int64_t *send5, int *n_recv, int64_t *recv4);
/* Peer-memory halo of a DIA operator, after cv_op_set_dia_halo: every rank allocates
 * cv_op_dia_halo_bytes(op) with cv_peer_alloc (same layout on all ranks: two parities of
 * [lower band | upper band]); halo_base[p] = rank p's allocation as mapped here.              */
size_t cv_op_dia_halo_bytes(cv_op *op);
int cv_op_set_dia_halo_peers(cv_ctx *ctx, cv_op *op, void *const *halo_base /* world */);
/* Matrix-free sum-of-products operator  H = sum_s coef_s (x)_d h_{d,s}  on the product basis with
 * `dims` (last mode fastest) — the form of the reference's physics Hamiltonians before assembly
 * (unittests/test_lanczosBlockTTNS.py:21-35).  Nothing of the N x N matrix is stored.  This rank
 * applies the rows [row0, row0 + n_rows).  Single-factor DIAGONAL terms arrive pre-summed per mode in
 * dtab_dev (dtab_off[d] + digit); every other factor is an ELL table of width w: from entry `tab`,
 * row n holds w (value, element offset = (column - n) * stride of the mode) slots in tab_val_dev /
 * tab_col_dev (padding: value 0, offset 0).  terms7[7*t..] = {mode_a, mode_b
 * (-1: single factor), tab_a, tab_b, w_a, w_b, 0}, coef[t].  max_offset = largest |column - row| (the
 * band a row-sharded rank needs from its neighbours; halo set-up as for DIA: cv_op_set_dia_halo,
 * cv_op_set_dia_halo_peers), nnz_equiv = non-zeros of the equivalent CSR (bookkeeping only).      */
int cv_op_create_kron(cv_ctx *ctx, int64_t n_rows, int64_t row0, int ndim, const int32_t *dims, int nterm,
                      const int32_t *terms7, const double *coef, const double *tab_val_dev,
                      const int32_t *tab_col_dev, int tab_len, const double *dtab_dev,
                      const int32_t *dtab_off, int dtab_len, int64_t max_offset, int64_t nnz_equiv,
                      cv_op **out);
/* Work split of the fused Arnoldi-step kernels (k_orth_step / k_orth_step_batch), computed on the host by
 * the same functions the kernels call -- pure host routines for the CPU tests (no reference counterpart:
 * SciPy's Arnoldi step, _gcrotmk.py:112-141, is a sequential loop).
 * cv_orth_slab_plan: the m basis vectors are cut into ny slabs; slab s holds vectors [i0[s], i0[s+1]) and is
 *   reduced by CTAs [start[s], start[s+1]) of a grid of `grid` CTAs (arrays of ny+1 entries; mode = the
 *   "slab_mode" option).
 * cv_orth_batch_plan: the same for up to 4 lock-step problems of one launch with basis sizes m[q] (bit q of active_mask:
 *   problem q takes part in this pass): slab s belongs to problem slab_q[s], holds slab_mi[s] vectors from
 *   slab_i0[s], CTAs [slab_start[s], slab_start[s+1]); the update phase of problem q runs on CTAs
 *   [prob_start[q], prob_start[q+1]).  Arrays sized for 32 slabs + 1 and nprob + 1.                       */
int cv_orth_slab_plan(int m, int grid, int cplx, int mode, int *ny_out, int *start_out, int *i0_out);
int cv_orth_batch_plan(int nprob, const int *m, unsigned active_mask, int grid, int cplx, int *nslab_out,
                       int *slab_q, int *slab_i0, int *slab_mi, int *slab_start, int *prob_start);

/* Complex-valued (Hermitian) H -- numpyVector.py:98-100 applies whatever `other @ array` accepts:
 * entry k of the CSR operator becomes data_dev[k] + i * data_im_dev[k] (same sparsity, device array
 * borrowed).  Such an operator acts on complex vectors only and stays in CSR storage.                */
int cv_op_set_imag(cv_ctx *ctx, cv_op *op, const double *data_im_dev);
int cv_op_set_format(cv_op *op, int fmt);
int cv_op_info(cv_op *op, int64_t *n_rows, int64_t *nnz, int64_t *padded_nnz, int *fmt);

/* y = H x | sigma x - H x | H x - sigma x  (numpyVector.py:100,152,154).  H real; x,y real
 * or complex (FEAST: complex sigma and x, feast.py:90).                                     */
int cv_spmv(cv_ctx *ctx, cv_op *op, int cplx, int mode, double sigma_re, double sigma_im,
            const void *x, void *y, void *stream);
/* Same pass, plus the partial dots the Krylov solvers need: out[0..1] = <x|y> (conjugated),
 * out[2] = <y|y>.  This is the "fused shifted SpMV" of the headline metric.                 */
int cv_spmv_dots(cv_ctx *ctx, cv_op *op, int cplx, int mode, double sigma_re, double sigma_im,
                 const void *x, void *y, double *out3_host, void *stream);

/* ---- BLAS-1 family (numpyVector.py:57-96) -------------------------------------------- */
int cv_copy(cv_ctx *ctx, int64_t n, int cplx, const void *x, void *y, void *stream);
/* y = a x.  x real & a real -> y real; x real & a complex (y_cplx=1) -> y complex; x complex
 * -> y complex.  __mul__/__rmul__/__truediv__ (numpyVector.py:57-64).                        */
int cv_scal(cv_ctx *ctx, int64_t n, int x_cplx, int y_cplx, double a_re, double a_im,
            const void *x, void *y, void *stream);
int cv_real(cv_ctx *ctx, int64_t n, const void *x_cplx, double *y, void *stream); /* :83-84 */
int cv_conj(cv_ctx *ctx, int64_t n, const void *x_cplx, void *y_cplx, void *stream); /* :86-87 */
/* out[0..1] = sum conj?(x) * y  (np.vdot / np.dot, numpyVector.py:89-93).                   */
int cv_dot(cv_ctx *ctx, int64_t n, int cplx, int conj, const void *x, const void *y,
           double *out2_host, void *stream);
int cv_nrm2(cv_ctx *ctx, int64_t n, int cplx, const void *x, double *out_host, void *stream);
/* in place x /= ||x||  (numpyVector.py:76-78); returns the norm that was divided out.       */
int cv_normalize(cv_ctx *ctx, int64_t n, int cplx, void *x, double *norm_host, void *stream);

/* ---- linear combinations / tall-skinny products ---------------------------------------- */
/* Y_k = sum_j coef[j*ncol + k] V_j, k < ncol: one pass over the m inputs for all outputs
 * (linearCombination numpyVector.py:105-119 as driven by basisTransformation
 * util_funcs.py:208-231).  coef is a HOST array, row-major m x ncol, interleaved re/im when
 * c_cplx.  Outputs must not alias inputs.  Any m (inputs beyond 96 are accumulated in further
 * passes) and any ncol (8 real / 4 complex outputs per pass).                              */
int cv_lincomb(cv_ctx *ctx, int64_t n, int v_cplx, int c_cplx, int m, const void *const *v_ptrs,
               int ncol, const double *coef_host, void *const *y_ptrs, void *stream);
/* C[i*b + k] = sum conj?(V_i) W_k : overlapMatrix / matrixRepresentation / pick
 * (numpyVector.py:180-203, util_funcs.py:321-322).  out is HOST, re/im interleaved if cplx.
 * Any m and b (chunks of 128 vectors x 4 right-hand sides per launch).                      */
int cv_tsdot(cv_ctx *ctx, int64_t n, int cplx, int conj, int m, const void *const *v_ptrs,
             int b, const void *const *w_ptrs, double *out_host, void *stream);

/* Sequential modified Gram-Schmidt of x against qs with UNCONJUGATED products and division by
 * q.q, then the LINDEP test x.x > lindep and normalisation (numpyVector.py:121-145).
 * status: 0 = ok (x_out normalised), 1 = linearly dependent (reference returns None).
 * innerprod_host[0..1] = x.x after projection.  Any m (the reference has no cap either).    */
int cv_gs_against_set(cv_ctx *ctx, int64_t n, int cplx, const void *x_in, int m,
                      const void *const *q_ptrs, double lindep, void *x_out, int *status,
                      double *innerprod_host, void *stream);

/* New column of the overlap and operator matrices when v_last joins the list
 * (extendOverlapMatrix :223-238, extendMatrixRepresentation :205-221):
 * s_col[i] = <v_i|v_last>, h_col[i] = <v_i|H v_last>, i < m (v_ptrs includes v_last as its
 * last entry).  ket_tmp is caller scratch of n elements.  One SpMV + one pass over V.       */
int cv_extend_columns(cv_ctx *ctx, cv_op *op, int64_t n, int cplx, int m,
                      const void *const *v_ptrs, void *ket_tmp, double *s_col_host,
                      double *h_col_host, void *stream);

/* ---- approximate shifted solve (numpyVector.py:147-178) ------------------------------- */
typedef struct cv_solve_stats {
  int info;          /* scipy's second return value: 0 converged, >0 not converged          */
  int n_matvec;      /* operator applications                                               */
  int n_outer;       /* GCROT outer iterations / MINRES iterations                          */
  int n_sync;        /* host synchronisations                                               */
  int n_reorth;      /* Arnoldi steps that needed the second Gram-Schmidt pass              */
  double resid;      /* last residual norm estimate                                         */
  double b_norm;
  double orth_loss;  /* max | |v_j|^2 - 1 | seen by the Arnoldi health monitor               */
  int n_safe;        /* 1 if the solve switched to the classic re-orthogonalisation threshold */
  int n_recycled;    /* recycled (c,u) pairs this solve started from                          */
} cv_solve_stats;

/* One fused Arnoldi orthogonalisation step (GCROT's inner loop, _gcrotmk.py:112-141) on caller
 * vectors: h = basis^H w (classical Gram-Schmidt, repeated when the first pass kept less than eta of
 * |w|), w <- (w - basis h)/|w'|, one kernel.  ww = |w|^2 of this rank's rows (the fused SpMV supplies
 * it inside cv_solve).  out_host: [0] 1.0 if the second pass ran, [1] |w'|^2, [2..] h (m values,
 * re/im interleaved if cplx).  op may be NULL on a single GPU (sharded: the operator whose halo the
 * step pushes).  For tests and micro-benchmarks; cv_solve launches the same kernel.              */
int cv_arnoldi_step(cv_ctx *ctx, cv_op *op, int64_t n, int cplx, int m, const void *const *basis,
                    void *w, double ww, double eta, double *out_host, void *stream);

size_t cv_solve_workspace_bytes(int64_t n, int cplx, int solver, int m, int k);
/* Solve (sigma I - H) x = b (reverse: (H - sigma I) x = b) with the device restatement of
 * scipy.sparse.linalg.gcrotmk (GCROT(m,k), _gcrotmk.py:187-506) or minres (minres.py:13-379).
 * x0 may be NULL.  maxiter counts GCROT outer iterations / MINRES iterations exactly as the
 * reference passes `linearIter`.  Non-convergence is reported through stats->info; the host
 * layer raises like numpyVector.py:175-177.                                                */
int cv_solve(cv_ctx *ctx, cv_op *op, int cplx, int solver, int reverse, double sigma_re,
             double sigma_im, const void *b, const void *x0, void *x_out, double rtol,
             double atol, int maxiter, int m, int k, void *work_dev, size_t work_bytes,
             cv_solve_stats *stats, void *stream);

/* cv_solve with a DIAGONAL RIGHT PRECONDITIONER M = diag(dinv) — SciPy's `M=` argument of gcrotmk
 * (_gcrotmk.py:100 z = M v), which the reference leaves unused: z_j = M v_j, w = A z_j, the solution
 * update uses the z_j; the residual test is on the true residual b - A x, as in cv_solve.  dinv_dev has
 * n elements of the vectors' type (e.g. 1/(sigma - H_ii): Jacobi).  GCROT only; work_dev must hold
 * cv_solve_workspace_bytes() + two more vectors (each n elements rounded up to 256 bytes).          */
int cv_solve_precond(cv_ctx *ctx, cv_op *op, int cplx, int solver, int reverse, double sigma_re,
                     double sigma_im, const void *b, const void *x0, void *x_out, double rtol,
                     double atol, int maxiter, int m, int k, const void *dinv_dev, void *work_dev,
                     size_t work_bytes, cv_solve_stats *stats, void *stream);

/* LOCK-STEP solves: nrhs (<= 8) independent systems (sigma_q I - H) x_q = b_q with the same operator
 * (the nBlock solves of one block-Lanczos step, inexact_Lanczos.py:319-320; the m0 solves of a FEAST
 * quadrature node, feast.py:190-201) advance one GCROT Arnoldi step at a time together: the matrix
 * is read once per step for all of them (DIA storage) and one fused orthogonalisation kernel with one
 * grid barrier and one host message serves all problems.  Per solve: same recurrences, stopping
 * rules and return codes as cv_solve (stats[q]).  GCROT only, single GPU, m + 2k + 2 <= 64;
 * work_dev holds nrhs workspaces of cv_solve_workspace_bytes() each.  x0 may be NULL (or hold NULLs). */
int cv_solve_batch(cv_ctx *ctx, cv_op *op, int cplx, int nrhs, int reverse, const double *sigma_re,
                   const double *sigma_im, const void *const *b, const void *const *x0,
                   void *const *x_out, double rtol, double atol, int maxiter, int m, int k,
                   void *work_dev, size_t work_bytes, cv_solve_stats *stats, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CUDAVEC_H */
