// kernels_spmv.cuh — the fused shifted sparse matrix-vector product
//
//     y = alpha * op(x) + beta1 * u1,   op(x) = H x | sigma x - H x | H x - sigma x
//     d_xy = <x|y>,  d_yy = <y|y>                       (partial dots, same pass)
//
// replacing `sigma*x - H@x` / `H@x - sigma*x` / `H@x` (numpyVector.py:152,154,100), the Lanczos
// three-term update of MINRES (minres.py:215-222) and the w-norm of GCROT's Arnoldi step
// (_gcrotmk.py:112-114) — each of which is a separate full pass with temporaries in the
// reference.  H is real (fp64 values, int32 columns); x,y are fp64 or complex128 (FEAST's
// complex shift: a complex x is two real right-hand sides sharing one pass over H).
//
// Algorithmic bytes per launch: 12*nnz + 20*N (fp64) / 12*nnz + 36*N (complex), SURVEY §8d.
//
// Two storage formats:
//   SELL-32  thread-per-row, 32-row slices stored column-major: the val/col streams of a warp
//            are perfectly coalesced 256 B / 128 B segments, x is gathered through L1/L2
//            (neighbouring rows of the product-basis / stencil Hamiltonians hit the same lines).
//   CSR      G threads per row (G = 2..32 by mean row length) with a shuffle reduction; the
//            general fallback (dense-as-CSR test matrices, irregular rows).
#pragma once
#include "common.cuh"

template <typename T>
struct SpmvArgs {
  int64_t n_rows;
  // SELL
  int64_t n_slices;
  const int64_t *slice_ptr;
  const int32_t *sell_col;
  const double *sell_val;
  // CSR
  const int64_t *indptr;
  const int32_t *indices;
  const double *data;
  const double *data_im;  // non-null: complex-valued matrix, entry k = data[k] + i data_im[k] (complex vectors only)
  // vectors
  const T *x;
  const T *halo;        // columns >= n_local_cols read halo[c - n_local_cols]
  int64_t n_local_cols;
  T *y;
  int mode;
  T sigma;
  double alpha, beta1;
  const T *u1;
  // reductions
  double *partials;
  unsigned *counter;
  double *out;  // {Re<x|y>, Im<x|y>, <y|y>}
  HaloWait wait;  // peer transport: flags of the ranks that push this SpMV's halo (mask 0: none)
};

template <typename T, bool HALO>
__device__ __forceinline__ T spmv_gather(const SpmvArgs<T> &a, int c) {
  if (HALO && c >= a.n_local_cols) return ld_gather(a.halo + (c - a.n_local_cols));
  return ld_gather(a.x + c);
}

// shift + epilogue for one row; returns y_row and accumulates the dots
template <typename T, bool EPI, bool DOTS>
__device__ __forceinline__ void spmv_finish_row(const SpmvArgs<T> &a, int64_t row, T hx, T &d_xy,
                                                double &d_yy) {
  T xr = Num<T>::zero();
  if (a.mode != CV_SPMV_PLAIN || DOTS) xr = ld_gather(a.x + row);
  T r;
  if (a.mode == CV_SPMV_PLAIN)
    r = hx;
  else if (a.mode == CV_SPMV_SHIFT)
    r = Num<T>::sub(Num<T>::mul(a.sigma, xr), hx);
  else
    r = Num<T>::sub(hx, Num<T>::mul(a.sigma, xr));
  if (EPI) {
    r = Num<T>::scale(r, a.alpha);
    if (a.u1) Num<T>::fmar(r, a.beta1, ld_plain(a.u1 + row));
  }
  st_plain(a.y + row, r);
  if (DOTS) {
    Num<T>::fmac(d_xy, xr, r);
    d_yy += Num<T>::abs2(r);
  }
}

template <typename T, bool DOTS>
__device__ __forceinline__ void spmv_reduce(const SpmvArgs<T> &a, T d_xy, double d_yy) {
  if (DOTS) {
    // uniform layout for real and complex data: out = {Re<x|y>, Im<x|y>, <y|y>}
    double vals[3] = {0.0, 0.0, d_yy};
    Num<T>::to_red(d_xy, vals);
    grid_reduce<3>(vals, a.partials, a.counter, a.out, gridDim.x, blockIdx.x);
  }
}

// ------------------------------------------------------------------------------------------
// SELL-32x2: one warp per 32-row slice, one thread per row.  Inside a slice the columns are stored
// in PAIRS: entry (row lane, column 2p+e) lives at base + (p*32 + lane)*2 + e, so a thread fetches
// two values with ONE 128-bit load and their two column indices with ONE 64-bit load, and a warp's
// loads are contiguous 512 B / 256 B segments.  (Scalar loads cost 3 load instructions per
// non-zero — value, index, gathered x — and the LSU issue rate, not DRAM, bounded the kernel:
// the one-row DIA kernel showed the same signature, profiles/r1_ncu_c3_full_kernels.csv.  Pairs
// bring it to 2 per non-zero.)  Slice widths are even (host rounds up); padding entries have
// value 0 and point at the row's own column.  x is gathered through L1/L2: neighbouring rows of
// stencil / product-basis Hamiltonians hit the same lines; general sparsity has no reusable tile
// to stage.  Four pairs per iteration: 8 streaming loads, then 8 gathers in flight per thread (the
// two-pair version ran at 70 % of DRAM peak with 14 % issue utilisation: latency bound, ncu r2).
// ------------------------------------------------------------------------------------------
template <typename T, bool HALO, bool EPI, bool DOTS>
__global__ void __launch_bounds__(CV_BLOCK, sizeof(T) == 8 ? 6 : 4)
    k_spmv_sell(const __grid_constant__ SpmvArgs<T> a) {
  if (HALO) halo_wait_cta(a.wait);
  const int lane = threadIdx.x & 31;
  // 32-bit slice/row counters (rows < 2^31 because column indices are int32): fewer registers
  const int warp0 = blockIdx.x * CV_WARPS + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * CV_WARPS;
  const int n_slices = (int)a.n_slices;
  T d_xy = Num<T>::zero();
  double d_yy = 0.0;
  for (int s = warp0; s < n_slices; s += nwarps) {
    const int64_t base = __ldg(a.slice_ptr + s);
    const int npair = (int)((__ldg(a.slice_ptr + s + 1) - base) >> 6);  // width / 2
    const double2 *vp = reinterpret_cast<const double2 *>(a.sell_val + base) + lane;
    const int2 *cp = reinterpret_cast<const int2 *>(a.sell_col + base) + lane;
    T acc0 = Num<T>::zero(), acc1 = Num<T>::zero();
    int p = 0;
    for (; p + 4 <= npair; p += 4) {  // 8 streaming loads, then 8 gathers in flight per thread
      int2 c[4];
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c[u] = ld_stream2(cp + (p + u) * 32);
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_stream2(vp + (p + u) * 32);
      T xa[4], xb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xa[u] = spmv_gather<T, HALO>(a, c[u].x);
        xb[u] = spmv_gather<T, HALO>(a, c[u].y);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        Num<T>::fmar(acc0, v[u].x, xa[u]);
        Num<T>::fmar(acc1, v[u].y, xb[u]);
      }
    }
    for (; p + 2 <= npair; p += 2) {
      const int2 c0 = ld_stream2(cp + (p + 0) * 32), c1 = ld_stream2(cp + (p + 1) * 32);
      const double2 v0 = ld_stream2(vp + (p + 0) * 32), v1 = ld_stream2(vp + (p + 1) * 32);
      const T x00 = spmv_gather<T, HALO>(a, c0.x), x01 = spmv_gather<T, HALO>(a, c0.y);
      const T x10 = spmv_gather<T, HALO>(a, c1.x), x11 = spmv_gather<T, HALO>(a, c1.y);
      Num<T>::fmar(acc0, v0.x, x00);
      Num<T>::fmar(acc1, v0.y, x01);
      Num<T>::fmar(acc0, v1.x, x10);
      Num<T>::fmar(acc1, v1.y, x11);
    }
    if (p < npair) {
      const int2 c0 = ld_stream2(cp + p * 32);
      const double2 v0 = ld_stream2(vp + p * 32);
      Num<T>::fmar(acc0, v0.x, spmv_gather<T, HALO>(a, c0.x));
      Num<T>::fmar(acc1, v0.y, spmv_gather<T, HALO>(a, c0.y));
    }
    const int row = s * 32 + lane;
    if (row < (int)a.n_rows) spmv_finish_row<T, EPI, DOTS>(a, row, Num<T>::add(acc0, acc1), d_xy, d_yy);
  }
  spmv_reduce<T, DOTS>(a, d_xy, d_yy);
}

// ------------------------------------------------------------------------------------------
// CSR: G lanes cooperate on one row.
// ------------------------------------------------------------------------------------------
template <typename T, int G, bool HALO, bool EPI, bool DOTS>
__global__ void __launch_bounds__(CV_BLOCK) k_spmv_csr(const __grid_constant__ SpmvArgs<T> a) {
  if (HALO) halo_wait_cta(a.wait);
  constexpr int RPB = CV_BLOCK / G;
  const int grp = threadIdx.x / G, gl = threadIdx.x % G;
  T d_xy = Num<T>::zero();
  double d_yy = 0.0;
  for (int64_t row0 = (int64_t)blockIdx.x * RPB; row0 < a.n_rows; row0 += (int64_t)gridDim.x * RPB) {
    const int64_t row = row0 + grp;
    const bool valid = row < a.n_rows;
    T acc = Num<T>::zero();
    if (valid) {
      int64_t rs = __ldg(a.indptr + row);
      const int64_t re = __ldg(a.indptr + row + 1);
      if constexpr (sizeof(T) == 16) {
        if (a.data_im != nullptr) {  // complex Hermitian H: split real / imaginary value streams
          for (int64_t k = rs + gl; k < re; k += G) {
            int c = ld_stream(a.indices + k);
            cplx v = make_cplx(ld_stream(a.data + k), ld_stream(a.data_im + k));
            Num<T>::fma(acc, v, spmv_gather<T, HALO>(a, c));
          }
          rs = re;
        }
      }
      for (int64_t k = rs + gl; k < re; k += G) {
        int c = ld_stream(a.indices + k);
        double v = ld_stream(a.data + k);
        Num<T>::fmar(acc, v, spmv_gather<T, HALO>(a, c));
      }
    }
    // reduce over the G lanes of the group (G divides 32, groups are lane-aligned)
    double r[Num<T>::NRED];
    Num<T>::to_red(acc, r);
#pragma unroll
    for (int c = 0; c < Num<T>::NRED; ++c)
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
    if (valid && gl == 0) spmv_finish_row<T, EPI, DOTS>(a, row, Num<T>::from_red(r), d_xy, d_yy);
  }
  spmv_reduce<T, DOTS>(a, d_xy, d_yy);
}

// ------------------------------------------------------------------------------------------
// SELL construction (one-time, per operator)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CV_BLOCK)
    k_sell_widths(int64_t n_rows, int64_t n_slices, const int64_t *__restrict__ indptr,
                  int32_t *__restrict__ widths) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * CV_WARPS + (threadIdx.x >> 5);
  for (int64_t s = warp0; s < n_slices; s += (int64_t)gridDim.x * CV_WARPS) {
    int64_t row = s * 32 + lane;
    int len = 0;
    if (row < n_rows) len = (int)(indptr[row + 1] - indptr[row]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) widths[s] = len;
  }
}

__global__ void __launch_bounds__(CV_BLOCK)
    k_sell_fill(int64_t n_rows, int64_t n_cols, int64_t n_slices, const int64_t *__restrict__ indptr,
                const int32_t *__restrict__ indices, const double *__restrict__ data,
                const int64_t *__restrict__ slice_ptr, int32_t *__restrict__ sell_col,
                double *__restrict__ sell_val) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * CV_WARPS + (threadIdx.x >> 5);
  for (int64_t s = warp0; s < n_slices; s += (int64_t)gridDim.x * CV_WARPS) {
    const int64_t base = slice_ptr[s];
    const int width = (int)((slice_ptr[s + 1] - base) >> 5);  // even: the caller rounds the widths up
    const int64_t row = s * 32 + lane;
    int64_t rs = 0, len = 0;
    if (row < n_rows) {
      rs = indptr[row];
      len = indptr[row + 1] - rs;
    }
    // padding points at a column that is certainly valid and already cached by this row
    int32_t pad_col = (int32_t)(row < n_rows ? (row < n_cols ? row : n_cols - 1) : 0);
    for (int j = 0; j < width; ++j) {
      int64_t dst = base + ((int64_t)(j >> 1) * 32 + lane) * 2 + (j & 1);  // pair-interleaved (k_spmv_sell)
      if (j < len) {
        sell_col[dst] = indices[rs + j];
        sell_val[dst] = data[rs + j];
      } else {
        sell_col[dst] = pad_col;
        sell_val[dst] = 0.0;
      }
    }
  }
}

// pack the owned entries a peer needs into a contiguous send buffer (halo exchange)
template <typename T>
__global__ void __launch_bounds__(CV_BLOCK)
    k_pack(int64_t n_send, const int32_t *__restrict__ idx, const T *__restrict__ x,
           T *__restrict__ sendbuf) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_send; i += stride)
    st_plain(sendbuf + i, ld_gather(x + idx[i]));
}
