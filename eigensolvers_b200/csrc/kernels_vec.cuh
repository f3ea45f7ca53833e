// kernels_vec.cuh — length-N vector kernels (BLAS-1 family, tall-skinny products, fused
// Krylov updates).  All are HBM-bound streaming kernels: 16-byte loads, persistent
// grid-stride grids (148 SMs x 8 CTAs x 256 threads), deterministic reductions.
//
// T is double or cplx.  Real vectors are processed W=2 elements per thread (one 128-bit
// load); complex vectors W=1 (one complex128 = 128 bits).  The host wrapper falls back to
// W=1 for real vectors whose pointers are not 16-byte aligned.
#pragma once
#include "common.cuh"

template <typename T, int W>
struct Pack {
  T e[W];
};

template <typename T, int W>
__device__ __forceinline__ Pack<T, W> pk_zero() {
  Pack<T, W> p;
#pragma unroll
  for (int k = 0; k < W; ++k) p.e[k] = Num<T>::zero();
  return p;
}

// STREAM=true: one-shot data (L1 no-allocate); false: plain loads
template <typename T, int W, bool STREAM>
__device__ __forceinline__ Pack<T, W> pk_load(const T *base, int64_t ip, int64_t n) {
  Pack<T, W> p;
  if constexpr (W == 1) {
    p.e[0] = STREAM ? ld_stream(base + ip) : ld_plain(base + ip);
  } else {
    static_assert(W == 2 && sizeof(T) == 8, "W=2 is for double only");
    int64_t i0 = ip * 2;
    if (i0 + 2 <= n) {
      double2 v = STREAM ? ld_stream2(reinterpret_cast<const double2 *>(base + i0))
                         : *reinterpret_cast<const double2 *>(base + i0);
      p.e[0] = v.x;
      p.e[1] = v.y;
    } else {
      p.e[0] = (i0 < n) ? base[i0] : 0.0;
      p.e[1] = 0.0;
    }
  }
  return p;
}

template <typename T, int W>
__device__ __forceinline__ void pk_store(T *base, int64_t ip, int64_t n, const Pack<T, W> &p) {
  if constexpr (W == 1) {
    st_plain(base + ip, p.e[0]);
  } else {
    int64_t i0 = ip * 2;
    if (i0 + 2 <= n) {
      *reinterpret_cast<double2 *>(base + i0) = make_double2(p.e[0], p.e[1]);
    } else if (i0 < n) {
      base[i0] = p.e[0];
    }
  }
}

__device__ __forceinline__ int64_t n_packs(int64_t n, int W) { return (n + W - 1) / W; }

// ------------------------------------------------------------------------------------------
// y = a * x   (TX in {double,cplx}, TA in {double,cplx}, TY = promoted type)
// numpyVector.py:57-64 (__mul__, __rmul__, __truediv__)
// ------------------------------------------------------------------------------------------
template <typename TX, typename TA, typename TY>
__global__ void __launch_bounds__(CV_BLOCK) k_scal(int64_t n, TA a, const TX *__restrict__ x,
                                                   TY *__restrict__ y) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    TY acc = Num<TY>::zero();
    cfma(acc, a, ld_stream(x + i));
    st_plain(y + i, acc);
  }
}

// real-by-real scaling with 128-bit accesses
template <int W>
__global__ void __launch_bounds__(CV_BLOCK) k_scal_rr(int64_t n, double a,
                                                      const double *__restrict__ x,
                                                      double *__restrict__ y) {
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<double, W> p = pk_load<double, W, true>(x, ip, n);
#pragma unroll
    for (int k = 0; k < W; ++k) p.e[k] *= a;
    pk_store<double, W>(y, ip, n, p);
  }
}

// y = Re(x)  /  y = conj(x)      numpyVector.py:83-87
__global__ void __launch_bounds__(CV_BLOCK) k_real(int64_t n, const cplx *__restrict__ x,
                                                   double *__restrict__ y) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = ld_stream(x + i).re;
}
__global__ void __launch_bounds__(CV_BLOCK) k_conj(int64_t n, const cplx *__restrict__ x,
                                                   cplx *__restrict__ y) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    cplx v = ld_stream(x + i);
    st_plain(y + i, make_cplx(v.re, -v.im));
  }
}

// ------------------------------------------------------------------------------------------
// out[0..NRED) = sum conj?(x) y     numpyVector.py:89-93 (vdot / dot)
// ------------------------------------------------------------------------------------------
template <typename T, int W, bool CONJ>
__global__ void __launch_bounds__(CV_BLOCK) k_dot(int64_t n, const T *__restrict__ x,
                                                  const T *__restrict__ y, double *partials,
                                                  unsigned *counter, double *out) {
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  T acc = Num<T>::zero();
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> a = pk_load<T, W, false>(x, ip, n);
    Pack<T, W> b = pk_load<T, W, false>(y, ip, n);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      if (CONJ)
        Num<T>::fmac(acc, a.e[k], b.e[k]);
      else
        Num<T>::fma(acc, a.e[k], b.e[k]);
    }
  }
  double vals[Num<T>::NRED];
  Num<T>::to_red(acc, vals);
  grid_reduce<Num<T>::NRED>(vals, partials, counter, out, gridDim.x, blockIdx.x);
}

// out[0] = sum |x|^2
template <typename T, int W>
__global__ void __launch_bounds__(CV_BLOCK) k_nrm2sq(int64_t n, const T *__restrict__ x,
                                                     double *partials, unsigned *counter,
                                                     double *out) {
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> a = pk_load<T, W, false>(x, ip, n);
#pragma unroll
    for (int k = 0; k < W; ++k) acc += Num<T>::abs2(a.e[k]);
  }
  double vals[1] = {acc};
  grid_reduce<1>(vals, partials, counter, out, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// x *= f(s) with s a DEVICE scalar produced by an earlier reduction (no host round trip).
//   MODE 0: f = 1/sqrt(s)   (normalise by a squared norm)
//   MODE 1: f = 1/sqrt(s) if that is finite, else 1   (scipy _fgmres: "if isfinite(alpha)")
// optional second vector x2 gets the same factor (GCROT scales cx and ux together).
// ------------------------------------------------------------------------------------------
template <typename T, int W, int MODE>
__global__ void __launch_bounds__(CV_BLOCK) k_scale_dev(int64_t n, T *__restrict__ x,
                                                        T *__restrict__ x2,
                                                        const double *__restrict__ s) {
  double f = 1.0 / sqrt(__ldcg(s));
  if (MODE == 1 && !isfinite(f)) f = 1.0;
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> a = pk_load<T, W, false>(x, ip, n);
#pragma unroll
    for (int k = 0; k < W; ++k) a.e[k] = Num<T>::scale(a.e[k], f);
    pk_store<T, W>(x, ip, n, a);
    if (x2) {
      Pack<T, W> b = pk_load<T, W, false>(x2, ip, n);
#pragma unroll
      for (int k = 0; k < W; ++k) b.e[k] = Num<T>::scale(b.e[k], f);
      pk_store<T, W>(x2, ip, n, b);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Linear combinations  Y_k = sum_j c[j,k] V_j   (k < NC outputs, one pass over the m inputs)
// numpyVector.py:105-119 through util_funcs.py:208-231; GCROT's ux / cx (_gcrotmk.py:430-447).
// Pointers and coefficients travel in kernel parameter space (constant bank, broadcast reads).
// Optional: out_norm[k] = sum |Y_k|^2 (fused norm of the result).
// ------------------------------------------------------------------------------------------
struct LcParams {
  const void *v[CV_MAX_PTRS];
  void *y[8];
  double coef[CV_MAX_COEF];  // [j*ncol + k], (re,im) pairs when the coefficient type is complex
  int m, ncol;
  int64_t n;
};

template <typename TC>
__device__ __forceinline__ TC lc_coef(const LcParams &p, int j, int k);
template <>
__device__ __forceinline__ double lc_coef<double>(const LcParams &p, int j, int k) {
  return p.coef[j * p.ncol + k];
}
template <>
__device__ __forceinline__ cplx lc_coef<cplx>(const LcParams &p, int j, int k) {
  return make_cplx(p.coef[2 * (j * p.ncol + k)], p.coef[2 * (j * p.ncol + k) + 1]);
}

template <typename TV, typename TC, typename TY, int W, int NC, bool NORM>
__global__ void __launch_bounds__(CV_BLOCK)
    k_lincomb(const __grid_constant__ LcParams p, double *partials, unsigned *counter,
              double *out_norm) {
  const int64_t n = p.n;
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  double nrm[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) nrm[k] = 0.0;
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<TY, W> acc[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[k] = pk_zero<TY, W>();
#pragma unroll 4
    for (int j = 0; j < p.m; ++j) {
      Pack<TV, W> v = pk_load<TV, W, false>(static_cast<const TV *>(p.v[j]), ip, n);
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        TC c = lc_coef<TC>(p, j, k);
#pragma unroll
        for (int w = 0; w < W; ++w) cfma(acc[k].e[w], c, v.e[w]);
      }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      pk_store<TY, W>(static_cast<TY *>(p.y[k]), ip, n, acc[k]);
      if (NORM) {
#pragma unroll
        for (int w = 0; w < W; ++w) nrm[k] += Num<TY>::abs2(acc[k].e[w]);
      }
    }
  }
  if (NORM) grid_reduce<NC>(nrm, partials, counter, out_norm, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// Tall-skinny product  C[i,k] = sum conj?(V_i) W_k,  i < m, k < B
// overlapMatrix / matrixRepresentation / extend* (numpyVector.py:180-238), pick
// (util_funcs.py:321-322), GCROT's orthogonalisation coefficients (_gcrotmk.py:117-128).
// grid = (gx, ceil(m/MI)); CTA (bx,by) handles vectors by*MI .. by*MI+MI-1 for its rows, so W
// is re-read ceil(m/MI) times and V exactly once.
// ------------------------------------------------------------------------------------------
struct TsParams {
  const void *v[CV_MAX_PTRS];
  const void *w[4];
  int m, b;
  int64_t n;
};

template <typename T, int W, bool CONJ, int MI, int B>
__global__ void __launch_bounds__(CV_BLOCK)
    k_tsdot(const __grid_constant__ TsParams p, const double *__restrict__ gate, double *partials,
            unsigned *counters, double *out /* [m*B*NRED] as ((i*B + k)*NRED + c) */) {
  constexpr int NR = Num<T>::NRED;
  const int64_t n = p.n;
  const int i0 = blockIdx.y * MI;
  const int mi = min(MI, p.m - i0);
  // Selective re-orthogonalisation (Daniel-Gragg-Kaufman-Stewart): gate = {|w|^2 before,
  // |w'|^2 after the first projection}.  If the first pass removed less than half of the
  // squared norm there was no cancellation and the second pass is skipped: result = 0.
  if (gate && __ldcg(gate + 1) >= 0.5 * __ldcg(gate)) {
    if (blockIdx.x == 0)
      for (int v = threadIdx.x; v < mi * B * NR; v += blockDim.x) out[(size_t)i0 * B * NR + v] = 0.0;
    return;
  }
  T acc[MI][B];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int k = 0; k < B; ++k) acc[i][k] = Num<T>::zero();

  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> wv[B];
#pragma unroll
    for (int k = 0; k < B; ++k) wv[k] = pk_load<T, W, false>(static_cast<const T *>(p.w[k]), ip, n);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      if (i < mi) {
        Pack<T, W> vv = pk_load<T, W, false>(static_cast<const T *>(p.v[i0 + i]), ip, n);
#pragma unroll
        for (int k = 0; k < B; ++k)
#pragma unroll
          for (int w = 0; w < W; ++w) {
            if (CONJ)
              Num<T>::fmac(acc[i][k], vv.e[w], wv[k].e[w]);
            else
              Num<T>::fma(acc[i][k], vv.e[w], wv[k].e[w]);
          }
      }
    }
  }
  // CTA reduction of MI*B*NR doubles
  __shared__ double s_part[CV_WARPS][MI * B * NR];
  __shared__ double s_vals[MI * B * NR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int k = 0; k < B; ++k) {
      double r[NR];
      Num<T>::to_red(acc[i][k], r);
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        double s = warp_sum(r[c]);
        if (lane == 0) s_part[warp][(i * B + k) * NR + c] = s;
      }
    }
  __syncthreads();
  for (int v = threadIdx.x; v < MI * B * NR; v += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < CV_WARPS; ++w) s += s_part[w][v];
    s_vals[v] = s;
  }
  __syncthreads();
  const int nv = mi * B * NR;  // valid values of this y-slab are the first mi*B*NR
  grid_reduce_dyn(s_vals, nv, partials + (size_t)blockIdx.y * MI * B * NR * gridDim.x,
                  counters + blockIdx.y, out + (size_t)i0 * B * NR, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// w <- w - sum_j h[j] V_j   with h in DEVICE memory (result of k_tsdot); optional
// out_norm[0] = |w_new|^2.   GCROT orthogonalisation update (_gcrotmk.py:117-129).
// ------------------------------------------------------------------------------------------
template <typename T, int W, bool NORM>
__global__ void __launch_bounds__(CV_BLOCK)
    k_tsupdate(const __grid_constant__ TsParams p, const double *__restrict__ h,
               const double *__restrict__ gate, T *__restrict__ wvec, double *partials,
               unsigned *counter, double *out_norm) {
  extern __shared__ double s_h[];  // m * NRED doubles
  constexpr int NR = Num<T>::NRED;
  if (gate && __ldcg(gate + 1) >= 0.5 * __ldcg(gate)) {  // second pass skipped (see k_tsdot)
    if (NORM && blockIdx.x == 0 && threadIdx.x == 0) out_norm[0] = __ldcg(gate + 1);
    return;
  }
  for (int j = threadIdx.x; j < p.m * NR; j += blockDim.x) s_h[j] = __ldcg(h + j);
  __syncthreads();
  const int64_t n = p.n;
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  double nrm[1] = {0.0};
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> acc = pk_load<T, W, false>(wvec, ip, n);
#pragma unroll 4
    for (int j = 0; j < p.m; ++j) {
      Pack<T, W> v = pk_load<T, W, false>(static_cast<const T *>(p.v[j]), ip, n);
      T c = Num<T>::from_red(s_h + j * NR);
      T mc = Num<T>::sub(Num<T>::zero(), c);
#pragma unroll
      for (int w = 0; w < W; ++w) Num<T>::fma(acc.e[w], mc, v.e[w]);
    }
    pk_store<T, W>(wvec, ip, n, acc);
    if (NORM) {
#pragma unroll
      for (int w = 0; w < W; ++w) nrm[0] += Num<T>::abs2(acc.e[w]);
    }
  }
  if (NORM) grid_reduce<1>(nrm, partials, counter, out_norm, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// One step of the reference's sequential Gram-Schmidt (numpyVector.py:133-140), fused:
//   x_out = x_in - (t1_prev/t2_prev) q_prev      (skipped when q_prev == nullptr)
//   t1 = x_out . q_cur,  t2 = q_cur . q_cur      (UNCONJUGATED, numpyVector.py:135-136)
// For the final step q_cur == nullptr and t1 = x_out . x_out (innerprod, :140).
// Coefficients stay in device memory; the host reads only the final innerprod.
// ------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(CV_BLOCK)
    k_mgs_step(int64_t n, const T *__restrict__ x_in, T *__restrict__ x_out,
               const T *__restrict__ q_prev, const double *__restrict__ t_prev,
               const T *__restrict__ q_cur, double *partials, unsigned *counter, double *t_out) {
  constexpr int NR = Num<T>::NRED;
  T coef = Num<T>::zero();
  if (q_prev) {
    // c = t1/t2 (complex division for complex data)
    T t1 = Num<T>::from_red(t_prev), t2 = Num<T>::from_red(t_prev + NR);
    if constexpr (NR == 1) {
      coef = t1 / t2;
    } else {
      double d = t2.re * t2.re + t2.im * t2.im;
      coef = make_cplx((t1.re * t2.re + t1.im * t2.im) / d, (t1.im * t2.re - t1.re * t2.im) / d);
    }
  }
  T mc = Num<T>::sub(Num<T>::zero(), coef);
  T a1 = Num<T>::zero(), a2 = Num<T>::zero();
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> x = pk_load<T, W, false>(x_in, ip, n);
    if (q_prev) {
      Pack<T, W> qp = pk_load<T, W, false>(q_prev, ip, n);
#pragma unroll
      for (int w = 0; w < W; ++w) Num<T>::fma(x.e[w], mc, qp.e[w]);
    }
    if (q_prev || x_out != x_in) pk_store<T, W>(x_out, ip, n, x);
    if (q_cur) {
      Pack<T, W> qc = pk_load<T, W, false>(q_cur, ip, n);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        Num<T>::fma(a1, x.e[w], qc.e[w]);
        Num<T>::fma(a2, qc.e[w], qc.e[w]);
      }
    } else {
#pragma unroll
      for (int w = 0; w < W; ++w) Num<T>::fma(a1, x.e[w], x.e[w]);
    }
  }
  double vals[2 * NR];
  Num<T>::to_red(a1, vals);
  Num<T>::to_red(a2, vals + NR);
  grid_reduce<2 * NR>(vals, partials, counter, t_out, gridDim.x, blockIdx.x);
}

// x *= 1/sqrt(s) for complex s (np.sqrt of a complex innerprod, numpyVector.py:142).
// For real data MODE 0 of k_scale_dev is used instead.
__global__ void __launch_bounds__(CV_BLOCK)
    k_div_csqrt(int64_t n, cplx *__restrict__ x, const double *__restrict__ s) {
  // principal square root of s = (sr, si), then 1/sqrt
  double sr = __ldcg(s), si = __ldcg(s + 1);
  double mod = hypot(sr, si);
  double rr = sqrt(0.5 * (mod + sr)), ri = copysign(sqrt(0.5 * (mod - sr)), si);
  double d = rr * rr + ri * ri;
  cplx inv = make_cplx(rr / d, -ri / d);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    st_plain(x + i, Num<cplx>::mul(ld_plain(x + i), inv));
}

// ------------------------------------------------------------------------------------------
// y += a x (a: host scalar), optional out[0] = |y_new|^2 .   MINRES: y -= (alfa/beta) r2 and
// beta^2 = r2.r2 in one pass (minres.py:222-228).
// ------------------------------------------------------------------------------------------
template <typename T, int W, bool NORM>
__global__ void __launch_bounds__(CV_BLOCK)
    k_axpy_norm(int64_t n, T a, const T *__restrict__ x, T *__restrict__ y, double *partials,
                unsigned *counter, double *out) {
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  double nrm[1] = {0.0};
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> xv = pk_load<T, W, false>(x, ip, n);
    Pack<T, W> yv = pk_load<T, W, false>(y, ip, n);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      Num<T>::fma(yv.e[w], a, xv.e[w]);
      if (NORM) nrm[0] += Num<T>::abs2(yv.e[w]);
    }
    pk_store<T, W>(y, ip, n, yv);
  }
  if (NORM) grid_reduce<1>(nrm, partials, counter, out, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// GCROT residual/solution update with the coefficient taken from device memory
// (_gcrotmk.py:458-461):   gamma = g[0..NRED);  r -= gamma cx;  x += gamma ux;
// out[0] = |r_new|^2 (next outer iteration's beta).
// ------------------------------------------------------------------------------------------
template <typename T, int W>
__global__ void __launch_bounds__(CV_BLOCK)
    k_gcrot_update(int64_t n, const double *__restrict__ g, const T *__restrict__ cx,
                   const T *__restrict__ ux, T *__restrict__ r, T *__restrict__ x,
                   double *partials, unsigned *counter, double *out) {
  constexpr int NR = Num<T>::NRED;
  double gv[NR];
#pragma unroll
  for (int c = 0; c < NR; ++c) gv[c] = __ldcg(g + c);
  T gamma = Num<T>::from_red(gv);
  T mgamma = Num<T>::sub(Num<T>::zero(), gamma);
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  double nrm[1] = {0.0};
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> c = pk_load<T, W, false>(cx, ip, n);
    Pack<T, W> u = pk_load<T, W, false>(ux, ip, n);
    Pack<T, W> rv = pk_load<T, W, false>(r, ip, n);
    Pack<T, W> xv = pk_load<T, W, false>(x, ip, n);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      Num<T>::fma(rv.e[w], mgamma, c.e[w]);
      Num<T>::fma(xv.e[w], gamma, u.e[w]);
      nrm[0] += Num<T>::abs2(rv.e[w]);
    }
    pk_store<T, W>(r, ip, n, rv);
    pk_store<T, W>(x, ip, n, xv);
  }
  grid_reduce<1>(nrm, partials, counter, out, gridDim.x, blockIdx.x);
}

// scale cx, ux by 1/sqrt(s) (s = |cx|^2 on device) and compute gamma = <cx_scaled | r>
// (_gcrotmk.py:449-458), one pass.
template <typename T, int W>
__global__ void __launch_bounds__(CV_BLOCK)
    k_gcrot_scale_dot(int64_t n, const double *__restrict__ s, T *__restrict__ cx,
                      T *__restrict__ ux, const T *__restrict__ r, double *partials,
                      unsigned *counter, double *out) {
  const double f = 1.0 / sqrt(__ldcg(s));
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  T acc = Num<T>::zero();
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<T, W> c = pk_load<T, W, false>(cx, ip, n);
    Pack<T, W> u = pk_load<T, W, false>(ux, ip, n);
    Pack<T, W> rv = pk_load<T, W, false>(r, ip, n);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      c.e[w] = Num<T>::scale(c.e[w], f);
      u.e[w] = Num<T>::scale(u.e[w], f);
      Num<T>::fmac(acc, c.e[w], rv.e[w]);
    }
    pk_store<T, W>(cx, ip, n, c);
    pk_store<T, W>(ux, ip, n, u);
  }
  double vals[Num<T>::NRED];
  Num<T>::to_red(acc, vals);
  grid_reduce<Num<T>::NRED>(vals, partials, counter, out, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// MINRES direction/solution update (minres.py:259-265, 278), one pass:
//   w_new = (s*vsrc - oldeps*w1 - delta*w2) * denom ;  x += phi * w_new ; out[0] = |x|^2
// vsrc is the unnormalised Lanczos vector (v = s*vsrc is never materialised); w_new overwrites
// w1's storage (w1 is dead after this step), the host rotates the three buffers.
// ------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(CV_BLOCK)
    k_minres_update(int64_t n, double s, double oldeps, double delta, double denom, double phi,
                    const double *__restrict__ vsrc, double *__restrict__ w1,
                    const double *__restrict__ w2, double *__restrict__ x, double *partials,
                    unsigned *counter, double *out) {
  int64_t np = n_packs(n, W), stride = (int64_t)gridDim.x * blockDim.x;
  double nrm[1] = {0.0};
  for (int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += stride) {
    Pack<double, W> v = pk_load<double, W, false>(vsrc, ip, n);
    Pack<double, W> a = pk_load<double, W, false>(w1, ip, n);
    Pack<double, W> b = pk_load<double, W, false>(w2, ip, n);
    Pack<double, W> xv = pk_load<double, W, false>(x, ip, n);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      // same association as the reference expression (v - oldeps*w1 - delta*w2) * denom
      double wn = ((s * v.e[w] - oldeps * a.e[w]) - delta * b.e[w]) * denom;
      a.e[w] = wn;
      xv.e[w] = xv.e[w] + phi * wn;
      nrm[0] += xv.e[w] * xv.e[w];
    }
    pk_store<double, W>(w1, ip, n, a);
    pk_store<double, W>(x, ip, n, xv);
  }
  grid_reduce<1>(nrm, partials, counter, out, gridDim.x, blockIdx.x);
}
