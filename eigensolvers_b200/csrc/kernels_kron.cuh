// kernels_kron.cuh — MATRIX-FREE fused shifted operator application for sum-of-products
// (Kronecker-sum) Hamiltonians
//
//     H = sum_s  c_s  (x)_d  h_{d,s}            h_{d,s} = small 1-D matrices, identity where absent
//
// the form the reference's physics Hamiltonians have before they are assembled
// (unittests/test_lanczosBlockTTNS.py:21-35 `operatorSumOfProduct`; examples/*.op files): mode
// energies sum_i w_i (n_i + 1/2) are single-factor diagonal terms, couplings c q_i q_j are
// two-factor terms with tridiagonal q.  SURVEY 8f.3.
//
// Nothing of the N x N matrix is stored: the value of every entry is a product of one or two
// 1-D table entries selected by the mixed-radix digits of the row index, and the column is
// row + sum_f (col_f - digit_f) * stride_f.  Per launch the kernel moves x once and y once
// (16 N bytes, fp64) instead of the 8*D*ld + 16 N of diagonal storage or 12 nnz + 20 N of CSR.
//
// Tables (device arrays, staged in shared memory by every CTA):
//   single-factor DIAGONAL terms are pre-summed on the host into one per-mode table
//       dtab[dtab_off[d] + n_d],   diag(row) = sum_d dtab[..]
//   every other factor is stored row-wise in ELL form with width w (<= its max non-zeros per row):
//       tab_val[first + n*w + j],  tab_col[first + n*w + j] = (column - n) * stride(mode)   (the
//       ELEMENT OFFSET the entry contributes to the gathered index; padding: value 0, offset 0)
// A term touches w_a (x w_b) entries of x per row: index = row + off_a (+ off_b), value =
// coef * val_a (* val_b); q_i q_j has exactly 2 x 2 entries per row = four independent gathers.
// (First version: column digits in the tables, 64-bit index arithmetic and clamped gathers — ~40
// instructions per gathered entry, instruction-bound at 1.03 ms for N = 2e7 against 0.68 ms of the
// DIA kernel that streams 4 GB of values; profiles/bench_lines/r2_dev_c3_2.json.)
//
// Row-sharded mode: like the DIA kernel, x entries below / above the owned block live in two
// contiguous band buffers (lo_len = hi_len = largest |column - row|) filled by the same halo
// push; rows are decoded from their GLOBAL index row0 + row.
#pragma once
#include "kernels_spmv.cuh"

constexpr int KR_MAX_DIM = 8;      // modes
constexpr int KR_MAX_TERMS = 40;   // product terms (after the diagonal single-factor ones were merged)
constexpr int KR_MAX_TAB = 1536;   // entries of all ELL tables together (shared memory: 16 B each)
constexpr int KR_MAX_DTAB = 512;   // entries of the merged diagonal tables

struct KronTerm {
  int mode_a, mode_b;  // mode_b = -1: single factor
  int tab_a, tab_b;    // first entry of the factor's ELL table
  int w_a, w_b;        // ELL widths
  double coef;
};

template <typename T>
struct KronArgs {
  SpmvArgs<T> s;  // x, y, mode, sigma, epilogue and reduction fields
  int ndim, nterm;
  int dims[KR_MAX_DIM];
  int dtab_off[KR_MAX_DIM];
  long long stride[KR_MAX_DIM];
  unsigned long long magic[KR_MAX_DIM];  // floor(2^sh / dim) + 1: n / dim == (n * magic) >> sh for n < 2^31
  int shift[KR_MAX_DIM];
  KronTerm term[KR_MAX_TERMS];
  const double *tab_val;
  const int *tab_col;  // element offsets, see above
  const double *dtab;
  int tab_len, dtab_len;
  long long row0;
  const T *halo_lo, *halo_hi;
  int lo_len, hi_len;
};

// x entry at local index i.  Without a halo every index the tables can produce is inside [0, n)
// (columns are valid digits; padding entries have offset 0), so the gather needs no clamp.
template <typename T, bool HALO>
__device__ __forceinline__ T kron_x(const KronArgs<T> &a, int i, int n, double hs) {
  if (HALO) {
    if (i < 0) {
      i += a.lo_len;
      return Num<T>::scale(ld_gather(a.halo_lo + (i < 0 ? 0 : i)), hs);
    }
    if (i >= n) {
      i -= n;
      return Num<T>::scale(ld_gather(a.halo_hi + (i >= a.hi_len ? a.hi_len - 1 : i)), hs);
    }
  }
  return ld_gather(a.s.x + i);
}

struct __align__(16) KronEntry {  // one ELL slot in shared memory: a single 128-bit LDS
  double val;
  int off;  // (column digit - row digit) * stride of the factor's mode, in elements
  int pad;
};

// sum over the product terms for one row; CHECK = the gathered index may leave [0, n) (halo bands)
template <typename T, bool CHECK>
__device__ __forceinline__ T kron_row_terms(const KronArgs<T> &a, const KronEntry *s_tab, int row, int n, double hs,
                                            unsigned long long digits, T acc0) {
  T acc1 = Num<T>::zero();
  for (int t = 0; t < a.nterm; ++t) {
    const KronTerm &k = a.term[t];
    const int na = (int)((digits >> (8 * k.mode_a)) & 0xFFull);
    const int ra = k.tab_a + na * k.w_a;
    if (k.mode_b < 0) {
      for (int ja = 0; ja < k.w_a; ja += 2) {
        const bool two = ja + 1 < k.w_a;
        const KronEntry e0 = s_tab[ra + ja], e1 = s_tab[ra + (two ? ja + 1 : ja)];
        const T x0 = kron_x<T, CHECK>(a, row + e0.off, n, hs), x1 = kron_x<T, CHECK>(a, row + e1.off, n, hs);
        Num<T>::fmar(acc0, k.coef * e0.val, x0);
        Num<T>::fmar(acc1, two ? k.coef * e1.val : 0.0, x1);
      }
      continue;
    }
    const int nb = (int)((digits >> (8 * k.mode_b)) & 0xFFull);
    const int rb = k.tab_b + nb * k.w_b;
    if (k.w_a == 2 && k.w_b == 2) {  // q_i q_j: exactly 2 x 2 entries per row, four independent gathers
      const KronEntry a0 = s_tab[ra], a1 = s_tab[ra + 1], b0 = s_tab[rb], b1 = s_tab[rb + 1];
      const double va0 = k.coef * a0.val, va1 = k.coef * a1.val;
      const T x00 = kron_x<T, CHECK>(a, row + a0.off + b0.off, n, hs), x01 = kron_x<T, CHECK>(a, row + a0.off + b1.off, n, hs);
      const T x10 = kron_x<T, CHECK>(a, row + a1.off + b0.off, n, hs), x11 = kron_x<T, CHECK>(a, row + a1.off + b1.off, n, hs);
      Num<T>::fmar(acc0, va0 * b0.val, x00);
      Num<T>::fmar(acc1, va0 * b1.val, x01);
      Num<T>::fmar(acc0, va1 * b0.val, x10);
      Num<T>::fmar(acc1, va1 * b1.val, x11);
      continue;
    }
    for (int ja = 0; ja < k.w_a; ++ja) {
      const KronEntry ea = s_tab[ra + ja];
      const double va = k.coef * ea.val;
      for (int jb = 0; jb < k.w_b; jb += 2) {
        const bool two = jb + 1 < k.w_b;
        const KronEntry b0 = s_tab[rb + jb], b1 = s_tab[rb + (two ? jb + 1 : jb)];
        const T x0 = kron_x<T, CHECK>(a, row + ea.off + b0.off, n, hs), x1 = kron_x<T, CHECK>(a, row + ea.off + b1.off, n, hs);
        Num<T>::fmar(acc0, va * b0.val, x0);
        Num<T>::fmar(acc1, two ? va * b1.val : 0.0, x1);
      }
    }
  }
  return Num<T>::add(acc0, acc1);
}

template <typename T, bool HALO, bool EPI, bool DOTS>
__global__ void __launch_bounds__(CV_BLOCK, sizeof(T) == 8 ? 5 : 3) k_spmv_kron(const __grid_constant__ KronArgs<T> a) {
  __shared__ KronEntry s_tab[KR_MAX_TAB];
  __shared__ double s_dtab[KR_MAX_DTAB];
  for (int i = threadIdx.x; i < a.tab_len; i += blockDim.x) {
    KronEntry e;
    e.val = a.tab_val[i];
    e.off = a.tab_col[i];
    e.pad = 0;
    s_tab[i] = e;
  }
  for (int i = threadIdx.x; i < a.dtab_len; i += blockDim.x) s_dtab[i] = a.dtab[i];
  if (HALO) halo_wait_cta(a.s.wait);  // ends in __syncthreads()
  else __syncthreads();
  const double hs = HALO ? halo_scale(a.s.wait) : 1.0;
  const int n = (int)a.s.n_rows;
  const int stride = gridDim.x * blockDim.x;
  T d_xy = Num<T>::zero();
  double d_yy = 0.0;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
    // mixed-radix digits of the global row index, 8 bits each, last mode fastest
    unsigned long long digits = 0ull;
    double diag = 0.0;
    {
      unsigned rem = (unsigned)(a.row0 + row);
#pragma unroll
      for (int d = KR_MAX_DIM - 1; d >= 0; --d) {
        if (d < a.ndim) {
          const unsigned q = (unsigned)(((unsigned long long)rem * a.magic[d]) >> a.shift[d]);
          const unsigned dg = rem - q * (unsigned)a.dims[d];
          rem = q;
          digits |= (unsigned long long)dg << (8 * d);
          diag += s_dtab[a.dtab_off[d] + (int)dg];
        }
      }
    }
    const T acc0 = Num<T>::scale(ld_gather(a.s.x + row), diag);
    T hx;
    // rows whose whole coupling band lies inside the owned block take the path without range checks
    // (the kernel is instruction bound: two compares per gathered entry cost 40 % on 2 GPUs)
    if (!HALO || (row >= a.lo_len && row < n - a.hi_len))
      hx = kron_row_terms<T, false>(a, s_tab, row, n, hs, digits, acc0);
    else
      hx = kron_row_terms<T, true>(a, s_tab, row, n, hs, digits, acc0);
    spmv_finish_row<T, EPI, DOTS>(a.s, row, hx, d_xy, d_yy);
  }
  spmv_reduce<T, DOTS>(a.s, d_xy, d_yy);
}
