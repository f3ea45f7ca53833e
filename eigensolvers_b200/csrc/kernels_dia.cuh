// kernels_dia.cuh — fused shifted SpMV for BANDED-STRUCTURE Hamiltonians in diagonal (DIA) storage.
//
// Product-basis and stencil Hamiltonians (the coupled-oscillator family, the 3-D Laplacian) have a
// handful of distinct column offsets col-row (25 and 7): every row uses the same offsets, rows at
// a basis edge simply miss some.  Storing one dense value stream per offset,
//     dia_val[d*ld + row] = H[row, row + off[d]]   (0 where the entry does not exist),
// removes the column indices altogether (8 instead of 12 bytes per stored entry, i.e. LESS DRAM
// traffic than the CSR-algorithmic 12*nnz + 20*N the roofline is quoted on) and, more
// importantly, removes the dependent load chain col -> x[col] that bounds the SELL kernel:
// the x address is row + off[d], known without touching the matrix, so all 2*D loads of a row
// are independent and perfectly coalesced (consecutive lanes read consecutive x and values).
//
// Row-sharded mode: x index row+off may fall below 0 or above n_loc; those entries live in two
// contiguous halo buffers (the band below / above the owned block), filled by plain contiguous
// ncclSend/ncclRecv ranges (no pack kernel).
#pragma once
#include "kernels_spmv.cuh"

constexpr int CV_MAX_DIAG = 64;

template <typename T>
struct DiaArgs {
  SpmvArgs<T> s;            // x, y, mode, sigma, epilogue and reduction fields
  const double *dia_val;    // [n_diag][ld]
  int64_t ld;
  int n_diag;
  int off[CV_MAX_DIAG];
  const T *halo_lo;         // x entries for local index in [-lo_len, 0): halo_lo[idx + lo_len]
  const T *halo_hi;         // x entries for local index in [n, n + hi_len): halo_hi[idx - n]
  int lo_len, hi_len;
};

// x entry at local index i (may be outside the owned block); out-of-band indices (only reached
// through zero padding values) are clamped to a valid address
template <typename T, bool HALO>
__device__ __forceinline__ T dia_x(const DiaArgs<T> &a, int i, int n, double hs) {
  if (HALO) {
    if (i < 0) {
      i += a.lo_len;
      return Num<T>::scale(ld_gather(a.halo_lo + (i < 0 ? 0 : i)), hs);
    }
    if (i >= n) {
      i -= n;
      return Num<T>::scale(ld_gather(a.halo_hi + (i >= a.hi_len ? a.hi_len - 1 : i)), hs);
    }
    return ld_gather(a.s.x + i);
  }
  i = i < 0 ? 0 : (i >= n ? n - 1 : i);
  return ld_gather(a.s.x + i);
}

template <typename T, bool HALO, bool EPI, bool DOTS>
__global__ void __launch_bounds__(CV_BLOCK, sizeof(T) == 8 ? (HALO ? 5 : 6) : (HALO ? 3 : 4))
    k_spmv_dia(const __grid_constant__ DiaArgs<T> a) {
  constexpr int DB = 8;  // diagonals per load batch
  if (HALO) halo_wait_cta(a.s.wait);
  const double hs = HALO ? halo_scale(a.s.wait) : 1.0;
  const int n = (int)a.s.n_rows;
  const int stride = gridDim.x * blockDim.x;
  T d_xy = Num<T>::zero();
  double d_yy = 0.0;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
    T acc0 = Num<T>::zero(), acc1 = Num<T>::zero();
    const double *vp = a.dia_val + row;
    for (int d0 = 0; d0 < a.n_diag; d0 += DB) {
      double v[DB];
      T xv[DB];
#pragma unroll
      for (int k = 0; k < DB; ++k) {
        const bool ok = d0 + k < a.n_diag;
        v[k] = ok ? ld_stream(vp + (int64_t)(d0 + k) * a.ld) : 0.0;
        xv[k] = ok ? dia_x<T, HALO>(a, row + a.off[d0 + k], n, hs) : Num<T>::zero();
      }
#pragma unroll
      for (int k = 0; k < DB; k += 2) {
        Num<T>::fmar(acc0, v[k], xv[k]);
        Num<T>::fmar(acc1, v[k + 1], xv[k + 1]);
      }
    }
    spmv_finish_row<T, EPI, DOTS>(a.s, row, Num<T>::add(acc0, acc1), d_xy, d_yy);
  }
  spmv_reduce<T, DOTS>(a.s, d_xy, d_yy);
}

// ------------------------------------------------------------------------------------------
// Real vectors, two rows per thread: the value stream of a diagonal and (for even offsets) the x
// entries of a row pair are ONE 128-bit load each, halving the load-instruction count.  The
// one-row kernel above issues 50 LDG.64 per row for 25 diagonals and ncu shows its LSU pipe 75 %
// busy with DRAM at 69 % (profiles/r1_ncu_c3_full_kernels.csv): it is instruction-issue bound
// before it is bandwidth bound.  Pairs that are misaligned (odd offset) or touch the halo/edge
// fall back to two 64-bit loads.  Requires x and dia_val 16-byte aligned (ld is a multiple of 32).
// ------------------------------------------------------------------------------------------
template <bool HALO>
__device__ __forceinline__ double2 dia_x2(const DiaArgs<double> &a, int i, int n, double hs) {
  if (!(i & 1) && i >= 0 && i + 1 < n) return __ldg(reinterpret_cast<const double2 *>(a.s.x + i));
  return make_double2(dia_x<double, HALO>(a, i, n, hs), dia_x<double, HALO>(a, i + 1, n, hs));
}

template <bool HALO, bool EPI, bool DOTS>
__global__ void __launch_bounds__(CV_BLOCK, 4) k_spmv_dia2(const __grid_constant__ DiaArgs<double> a) {
  constexpr int DB = 4;  // diagonals per load batch (x2 rows x (value + x) = 16 128-bit loads in flight)
  if (HALO) halo_wait_cta(a.s.wait);
  const double hs = HALO ? halo_scale(a.s.wait) : 1.0;
  const int n = (int)a.s.n_rows;
  const int npair = (n + 1) >> 1;
  const int stride = gridDim.x * blockDim.x;
  double d_xy = 0.0, d_yy = 0.0;
  for (int pr = blockIdx.x * blockDim.x + threadIdx.x; pr < npair; pr += stride) {
    const int row = pr << 1;
    const bool two = row + 1 < n;
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;  // (a*: row, b*: row+1) x two accumulators
    const double *vp = a.dia_val + row;
    for (int d0 = 0; d0 < a.n_diag; d0 += DB) {
      double2 v[DB], xv[DB];
#pragma unroll
      for (int k = 0; k < DB; ++k) {
        const bool ok = d0 + k < a.n_diag;
        // ld is a multiple of 32 and row is even: the pair of values is 16-byte aligned; the slot
        // of row+1 beyond n is zero padding inside the leading dimension
        v[k] = ok ? ld_stream2(reinterpret_cast<const double2 *>(vp + (int64_t)(d0 + k) * a.ld)) : make_double2(0.0, 0.0);
        xv[k] = ok ? dia_x2<HALO>(a, row + a.off[d0 + k], n, hs) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int k = 0; k < DB; k += 2) {
        a0 = fma(v[k].x, xv[k].x, a0);
        b0 = fma(v[k].y, xv[k].y, b0);
        a1 = fma(v[k + 1].x, xv[k + 1].x, a1);
        b1 = fma(v[k + 1].y, xv[k + 1].y, b1);
      }
    }
    spmv_finish_row<double, EPI, DOTS>(a.s, row, a0 + a1, d_xy, d_yy);
    if (two) spmv_finish_row<double, EPI, DOTS>(a.s, row + 1, b0 + b1, d_xy, d_yy);
  }
  spmv_reduce<double, DOTS>(a.s, d_xy, d_yy);
}

// ------------------------------------------------------------------------------------------
// DIA construction from CSR (one-time).  col_global[k] - (row0 + row) must be one of the n_diag
// offsets (sorted ascending, in shared memory); anything else raises *bad and the caller falls
// back to SELL.  dia_val must be zero-filled beforehand.
// ------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(CV_BLOCK)
    k_dia_fill(int64_t n_rows, int64_t row0, const int64_t *__restrict__ indptr,
               const int32_t *__restrict__ col_global, const double *__restrict__ data, int n_diag,
               const int *__restrict__ offsets, double *__restrict__ dia_val, int64_t ld, int *bad) {
  __shared__ int s_off[CV_MAX_DIAG];
  if (threadIdx.x < n_diag) s_off[threadIdx.x] = offsets[threadIdx.x];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += stride) {
    for (int64_t k = indptr[row]; k < indptr[row + 1]; ++k) {
      const int64_t off = (int64_t)col_global[k] - (row0 + row);
      int lo = 0, hi = n_diag - 1, pos = -1;
      while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (s_off[mid] == off) {
          pos = mid;
          break;
        }
        if (s_off[mid] < off)
          lo = mid + 1;
        else
          hi = mid - 1;
      }
      if (pos < 0) {
        atomicExch(bad, 1);
      } else {
        dia_val[(int64_t)pos * ld + row] += data[k];  // += : duplicate CSR entries sum like scipy
      }
    }
  }
}
