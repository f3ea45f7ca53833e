// kernels_vec.cuh — length-N vector kernels (BLAS-1 family, tall-skinny products, fused
// Krylov updates).  All are HBM-bound streaming kernels: 16-byte loads, persistent
// grid-stride grids sized from the kernel's real occupancy, deterministic reductions.
//
// T is double or cplx.  Real vectors are processed W=2 elements per thread (one 128-bit
// load); complex vectors W=1 (one complex128 = 128 bits).  The host wrapper falls back to
// W=1 for real vectors whose pointers are not 16-byte aligned.
//
// Memory-level parallelism is explicit: every loop first issues ALL loads of an iteration
// into registers (branch-free, full packs only) and only then does arithmetic; the odd tail
// element of a W=2 run is handled once by thread 0 outside the loop.  (The first version
// bounds-checked inside the pack load; ptxas then serialised load -> DFMA -> load and the
// tall-skinny dot ran with ONE load in flight per thread: 4.6 TB/s, profiles/r1_notes.md.)
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

// full pack load, no bounds check.  STREAM=true: one-shot data (L1 no-allocate)
template <typename T, int W, bool STREAM>
__device__ __forceinline__ Pack<T, W> pk_ld(const T *base, int64_t ip) {
  Pack<T, W> p;
  if constexpr (W == 1) {
    p.e[0] = STREAM ? ld_stream(base + ip) : ld_plain(base + ip);
  } else {
    static_assert(W == 2 && sizeof(T) == 8, "W=2 is for double only");
    double2 v = STREAM ? ld_stream2(reinterpret_cast<const double2 *>(base) + ip)
                       : *(reinterpret_cast<const double2 *>(base) + ip);
    p.e[0] = v.x;
    p.e[1] = v.y;
  }
  return p;
}

template <typename T, int W>
__device__ __forceinline__ void pk_st(T *base, int64_t ip, const Pack<T, W> &p) {
  if constexpr (W == 1) {
    st_plain(base + ip, p.e[0]);
  } else {
    *(reinterpret_cast<double2 *>(base) + ip) = make_double2(p.e[0], p.e[1]);
  }
}

// loop bookkeeping shared by all kernels
struct VecLoop {
  int64_t npf;     // number of FULL packs
  int64_t tid;     // global thread id
  int64_t stride;  // threads in the grid
  bool tail;       // this thread owns the odd last element (W=2, n odd)
  int64_t itail;
};
template <int W>
__device__ __forceinline__ VecLoop vec_loop(int64_t n) {
  VecLoop L;
  L.npf = n / W;
  L.tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  L.stride = (int64_t)gridDim.x * blockDim.x;
  L.tail = (W == 2) && (n & 1) && L.tid == 0;
  L.itail = n - 1;
  return L;
}

// ------------------------------------------------------------------------------------------
// y = a * x   (TX in {double,cplx}, TA in {double,cplx}, TY = promoted type)
// numpyVector.py:57-64 (__mul__, __rmul__, __truediv__)
// ------------------------------------------------------------------------------------------
template <typename TX, typename TA, typename TY>
__global__ void __launch_bounds__(CV_BLOCK, 8) k_scal(int64_t n, TA a, const TX *__restrict__ x,
                                                      TY *__restrict__ y) {
  constexpr int U = 4;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = tid; i0 < n; i0 += U * stride) {
    TX v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = i0 + u * stride;
      v[u] = (i < n) ? ld_stream(x + i) : Num<TX>::zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = i0 + u * stride;
      if (i < n) {
        TY acc = Num<TY>::zero();
        cfma(acc, a, v[u]);
        st_plain(y + i, acc);
      }
    }
  }
}

// real-by-real scaling with 128-bit accesses
template <int W>
__global__ void __launch_bounds__(CV_BLOCK, 8) k_scal_rr(int64_t n, double a,
                                                         const double *__restrict__ x,
                                                         double *__restrict__ y) {
  constexpr int U = 4;
  const VecLoop L = vec_loop<W>(n);
  for (int64_t ip0 = L.tid; ip0 < L.npf; ip0 += U * L.stride) {
    Pack<double, W> p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
      p[u] = (ip < L.npf) ? pk_ld<double, W, true>(x, ip) : pk_zero<double, W>();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
#pragma unroll
      for (int k = 0; k < W; ++k) p[u].e[k] *= a;
      if (ip < L.npf) pk_st<double, W>(y, ip, p[u]);
    }
  }
  if (L.tail) y[L.itail] = a * x[L.itail];
}

// y = Re(x)  /  y = conj(x)      numpyVector.py:83-87
static __global__ void __launch_bounds__(CV_BLOCK, 8) k_real(int64_t n, const cplx *__restrict__ x,
                                                             double *__restrict__ y) {
  constexpr int U = 4;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = tid; i0 < n; i0 += U * stride) {
    cplx v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = i0 + u * stride;
      v[u] = (i < n) ? ld_stream(x + i) : make_cplx(0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = i0 + u * stride;
      if (i < n) y[i] = v[u].re;
    }
  }
}
static __global__ void __launch_bounds__(CV_BLOCK, 8) k_conj(int64_t n, const cplx *__restrict__ x,
                                                             cplx *__restrict__ y) {
  constexpr int U = 4;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = tid; i0 < n; i0 += U * stride) {
    cplx v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = i0 + u * stride;
      v[u] = (i < n) ? ld_stream(x + i) : make_cplx(0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = i0 + u * stride;
      if (i < n) st_plain(y + i, make_cplx(v[u].re, -v[u].im));
    }
  }
}

// ------------------------------------------------------------------------------------------
// out[0..NRED) = sum conj?(x) y     numpyVector.py:89-93 (vdot / dot)
// ------------------------------------------------------------------------------------------
template <typename T, int W, bool CONJ>
__global__ void __launch_bounds__(CV_BLOCK, 8) k_dot(int64_t n, const T *__restrict__ x,
                                                     const T *__restrict__ y, double *partials,
                                                     unsigned *counter, double *out) {
  constexpr int U = 4;
  const VecLoop L = vec_loop<W>(n);
  T acc = Num<T>::zero();
  for (int64_t ip0 = L.tid; ip0 < L.npf; ip0 += U * L.stride) {
    Pack<T, W> a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
      const bool ok = ip < L.npf;
      a[u] = ok ? pk_ld<T, W, false>(x, ip) : pk_zero<T, W>();
      b[u] = ok ? pk_ld<T, W, false>(y, ip) : pk_zero<T, W>();
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (CONJ)
          Num<T>::fmac(acc, a[u].e[k], b[u].e[k]);
        else
          Num<T>::fma(acc, a[u].e[k], b[u].e[k]);
      }
  }
  if (L.tail) Num<T>::fma(acc, x[L.itail], y[L.itail]);  // W=2 is real: conj is the identity
  double vals[Num<T>::NRED];
  Num<T>::to_red(acc, vals);
  grid_reduce<Num<T>::NRED>(vals, partials, counter, out, gridDim.x, blockIdx.x);
}

// out[0] = sum |x|^2
template <typename T, int W>
__global__ void __launch_bounds__(CV_BLOCK, 8) k_nrm2sq(int64_t n, const T *__restrict__ x,
                                                        double *partials, unsigned *counter,
                                                        double *out) {
  constexpr int U = 8;
  const VecLoop L = vec_loop<W>(n);
  double acc = 0.0;
  for (int64_t ip0 = L.tid; ip0 < L.npf; ip0 += U * L.stride) {
    Pack<T, W> a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
      a[u] = (ip < L.npf) ? pk_ld<T, W, false>(x, ip) : pk_zero<T, W>();
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < W; ++k) acc += Num<T>::abs2(a[u].e[k]);
  }
  if (L.tail) acc += Num<T>::abs2(x[L.itail]);
  double vals[1] = {acc};
  grid_reduce<1>(vals, partials, counter, out, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// x *= f(s) with s a DEVICE scalar produced by an earlier reduction (no host round trip).
//   MODE 0: f = 1/sqrt(s)   (normalise by a squared norm)
//   MODE 1: f = 1/sqrt(s) if that is finite, else 1   (scipy _fgmres: "if isfinite(alpha)")
// ------------------------------------------------------------------------------------------
template <typename T, int W, int MODE>
__global__ void __launch_bounds__(CV_BLOCK, 8) k_scale_dev(int64_t n, T *__restrict__ x,
                                                           const double *__restrict__ s) {
  constexpr int U = 4;
  double f = 1.0 / sqrt(__ldcg(s));
  if (MODE == 1 && !isfinite(f)) f = 1.0;
  const VecLoop L = vec_loop<W>(n);
  for (int64_t ip0 = L.tid; ip0 < L.npf; ip0 += U * L.stride) {
    Pack<T, W> a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
      a[u] = (ip < L.npf) ? pk_ld<T, W, false>(x, ip) : pk_zero<T, W>();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
#pragma unroll
      for (int k = 0; k < W; ++k) a[u].e[k] = Num<T>::scale(a[u].e[k], f);
      if (ip < L.npf) pk_st<T, W>(x, ip, a[u]);
    }
  }
  if (L.tail) x[L.itail] = Num<T>::scale(x[L.itail], f);
}

// ------------------------------------------------------------------------------------------
// y = d (.) x   element-wise (complex product for complex data; y may alias x).  Diagonal right
// preconditioner of the shifted solves: z_j = M v_j with M = diag(1 / (sigma - H_ii)) (SciPy's
// `M=` argument of gcrotmk, _gcrotmk.py:100 `z = rpsolve(v)`).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(CV_BLOCK, 8) k_diag_mul(int64_t n, const T *__restrict__ d, const T *x, T *y) {
  constexpr int U = 4;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = tid; i0 < n; i0 += U * stride) {
    T dv[U], xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      dv[u] = (i < n) ? ld_plain(d + i) : Num<T>::zero();
      xv[u] = (i < n) ? ld_plain(x + i) : Num<T>::zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n) st_plain(y + i, Num<T>::mul(dv[u], xv[u]));
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
  constexpr int JB = 4;  // inputs loaded per batch
  const int64_t n = p.n;
  const VecLoop L = vec_loop<W>(n);
  double nrm[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) nrm[k] = 0.0;
  for (int64_t ip = L.tid; ip < L.npf; ip += L.stride) {
    Pack<TY, W> acc[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[k] = pk_zero<TY, W>();
    for (int j0 = 0; j0 < p.m; j0 += JB) {
      Pack<TV, W> v[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj)
        v[jj] = (j0 + jj < p.m) ? pk_ld<TV, W, false>(static_cast<const TV *>(p.v[j0 + jj]), ip)
                                : pk_zero<TV, W>();
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        if (j0 + jj < p.m) {
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            TC c = lc_coef<TC>(p, j0 + jj, k);
#pragma unroll
            for (int w = 0; w < W; ++w) cfma(acc[k].e[w], c, v[jj].e[w]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      pk_st<TY, W>(static_cast<TY *>(p.y[k]), ip, acc[k]);
      if (NORM) {
#pragma unroll
        for (int w = 0; w < W; ++w) nrm[k] += Num<TY>::abs2(acc[k].e[w]);
      }
    }
  }
  if (L.tail) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      TY acc = Num<TY>::zero();
      for (int j = 0; j < p.m; ++j) cfma(acc, lc_coef<TC>(p, j, k), static_cast<const TV *>(p.v[j])[L.itail]);
      static_cast<TY *>(p.y[k])[L.itail] = acc;
      if (NORM) nrm[k] += Num<TY>::abs2(acc);
    }
  }
  if (NORM) grid_reduce<NC>(nrm, partials, counter, out_norm, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// Tall-skinny product  C[i,k] = sum conj?(V_i) W_k,  i < m, k < B
// overlapMatrix / matrixRepresentation / extend* (numpyVector.py:180-238), pick
// (util_funcs.py:321-322), GCROT's orthogonalisation coefficients (_gcrotmk.py:117-128).
// grid = (gx, ceil(m/MI)); CTA (bx,by) handles vectors by*MI .. by*MI+MI-1 for its rows, so W
// is re-read ceil(m/MI) times and V exactly once.  The MI loads of an iteration are issued
// LB at a time into registers before any arithmetic.
// ------------------------------------------------------------------------------------------
struct TsParams {
  const void *v[CV_MAX_PTRS];
  const void *w[4];
  int m, b;
  int64_t n;
};

template <typename T, int W, bool CONJ, int MI, int B>
__global__ void __launch_bounds__(CV_BLOCK, 3)
    k_tsdot(const __grid_constant__ TsParams p, const double *__restrict__ gate, double eta2,
            double *partials, unsigned *counters,
            double *out /* [m*B*NRED] as ((i*B + k)*NRED + c) */) {
  constexpr int NR = Num<T>::NRED;
  constexpr int LB = MI < 8 ? MI : 8;  // loads per batch
  const int64_t n = p.n;
  const int i0 = blockIdx.y * MI;
  const int mi = min(MI, p.m - i0);
  // Selective re-orthogonalisation (Daniel-Gragg-Kaufman-Stewart): gate = {|w|^2 before,
  // |w'|^2 after the first projection}.  If the first pass kept at least eta^2 of the squared
  // norm there was no harmful cancellation and the second pass is skipped: result = 0.
  if (gate && __ldcg(gate + 1) >= eta2 * __ldcg(gate)) {
    if (blockIdx.x == 0)
      for (int v = threadIdx.x; v < mi * B * NR; v += blockDim.x) out[(size_t)i0 * B * NR + v] = 0.0;
    return;
  }
  T acc[MI][B];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int k = 0; k < B; ++k) acc[i][k] = Num<T>::zero();

  const VecLoop L = vec_loop<W>(n);
  for (int64_t ip = L.tid; ip < L.npf; ip += L.stride) {
    Pack<T, W> wv[B];
#pragma unroll
    for (int k = 0; k < B; ++k) wv[k] = pk_ld<T, W, false>(static_cast<const T *>(p.w[k]), ip);
#pragma unroll
    for (int ib = 0; ib < MI; ib += LB) {
      Pack<T, W> vv[LB];
#pragma unroll
      for (int l = 0; l < LB; ++l)
        vv[l] = (ib + l < mi) ? pk_ld<T, W, false>(static_cast<const T *>(p.v[i0 + ib + l]), ip)
                              : pk_zero<T, W>();
#pragma unroll
      for (int l = 0; l < LB; ++l)
#pragma unroll
        for (int k = 0; k < B; ++k)
#pragma unroll
          for (int w = 0; w < W; ++w) {
            if (CONJ)
              Num<T>::fmac(acc[ib + l][k], vv[l].e[w], wv[k].e[w]);
            else
              Num<T>::fma(acc[ib + l][k], vv[l].e[w], wv[k].e[w]);
          }
    }
  }
  if (L.tail) {
    for (int i = 0; i < mi; ++i)
      for (int k = 0; k < B; ++k)
        Num<T>::fma(acc[i][k], static_cast<const T *>(p.v[i0 + i])[L.itail],
                    static_cast<const T *>(p.w[k])[L.itail]);
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
               const double *__restrict__ gate, double eta2, double skip_scale,
               T *__restrict__ wvec, double *partials, unsigned *counter, double *out_norm) {
  extern __shared__ double s_h[];  // m * NRED doubles
  constexpr int NR = Num<T>::NRED;
  constexpr int JB = 8;
  if (gate && __ldcg(gate + 1) >= eta2 * __ldcg(gate)) {  // second pass skipped (see k_tsdot)
    // the value is already summed over ranks; the all-reduce that follows must not count it
    // once per rank: rank 0 contributes it (skip_scale 1), the others contribute 0
    if (NORM && blockIdx.x == 0 && threadIdx.x == 0) out_norm[0] = skip_scale * __ldcg(gate + 1);
    return;
  }
  for (int j = threadIdx.x; j < p.m * NR; j += blockDim.x) s_h[j] = -__ldcg(h + j);
  __syncthreads();
  const int64_t n = p.n;
  const VecLoop L = vec_loop<W>(n);
  double nrm[1] = {0.0};
  for (int64_t ip = L.tid; ip < L.npf; ip += L.stride) {
    Pack<T, W> acc = pk_ld<T, W, false>(wvec, ip);
    for (int j0 = 0; j0 < p.m; j0 += JB) {
      Pack<T, W> v[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj)
        v[jj] = (j0 + jj < p.m) ? pk_ld<T, W, false>(static_cast<const T *>(p.v[j0 + jj]), ip)
                                : pk_zero<T, W>();
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        if (j0 + jj < p.m) {
          T mc = Num<T>::from_red(s_h + (j0 + jj) * NR);
#pragma unroll
          for (int w = 0; w < W; ++w) Num<T>::fma(acc.e[w], mc, v[jj].e[w]);
        }
      }
    }
    pk_st<T, W>(wvec, ip, acc);
    if (NORM) {
#pragma unroll
      for (int w = 0; w < W; ++w) nrm[0] += Num<T>::abs2(acc.e[w]);
    }
  }
  if (L.tail) {
    T acc = wvec[L.itail];
    for (int j = 0; j < p.m; ++j)
      Num<T>::fma(acc, Num<T>::from_red(s_h + j * NR), static_cast<const T *>(p.v[j])[L.itail]);
    wvec[L.itail] = acc;
    if (NORM) nrm[0] += Num<T>::abs2(acc);
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
__global__ void __launch_bounds__(CV_BLOCK, 6)
    k_mgs_step(int64_t n, const T *__restrict__ x_in, T *__restrict__ x_out,
               const T *__restrict__ q_prev, const double *__restrict__ t_prev,
               const T *__restrict__ q_cur, double *partials, unsigned *counter, double *t_out) {
  constexpr int NR = Num<T>::NRED;
  constexpr int U = 2;
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
  const T mc = Num<T>::sub(Num<T>::zero(), coef);
  T a1 = Num<T>::zero(), a2 = Num<T>::zero();
  const bool store = q_prev || x_out != x_in;
  const VecLoop L = vec_loop<W>(n);
  for (int64_t ip0 = L.tid; ip0 < L.npf; ip0 += U * L.stride) {
    Pack<T, W> x[U], qp[U], qc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
      const bool ok = ip < L.npf;
      x[u] = ok ? pk_ld<T, W, false>(x_in, ip) : pk_zero<T, W>();
      qp[u] = (ok && q_prev) ? pk_ld<T, W, false>(q_prev, ip) : pk_zero<T, W>();
      qc[u] = (ok && q_cur) ? pk_ld<T, W, false>(q_cur, ip) : pk_zero<T, W>();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        Num<T>::fma(x[u].e[w], mc, qp[u].e[w]);
        if (q_cur) {
          Num<T>::fma(a1, x[u].e[w], qc[u].e[w]);
          Num<T>::fma(a2, qc[u].e[w], qc[u].e[w]);
        } else {
          Num<T>::fma(a1, x[u].e[w], x[u].e[w]);
        }
      }
      if (store && ip < L.npf) pk_st<T, W>(x_out, ip, x[u]);
    }
  }
  if (L.tail) {
    T x = x_in[L.itail];
    if (q_prev) Num<T>::fma(x, mc, q_prev[L.itail]);
    if (store) x_out[L.itail] = x;
    if (q_cur) {
      Num<T>::fma(a1, x, q_cur[L.itail]);
      Num<T>::fma(a2, q_cur[L.itail], q_cur[L.itail]);
    } else {
      Num<T>::fma(a1, x, x);
    }
  }
  double vals[2 * NR];
  Num<T>::to_red(a1, vals);
  Num<T>::to_red(a2, vals + NR);
  grid_reduce<2 * NR>(vals, partials, counter, t_out, gridDim.x, blockIdx.x);
}

// x *= 1/sqrt(s) for complex s (np.sqrt of a complex innerprod, numpyVector.py:142).
// For real data MODE 0 of k_scale_dev is used instead.
static __global__ void __launch_bounds__(CV_BLOCK)
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
__global__ void __launch_bounds__(CV_BLOCK, 8)
    k_axpy_norm(int64_t n, T a, const T *__restrict__ x, T *__restrict__ y, double *partials,
                unsigned *counter, double *out) {
  constexpr int U = 2;
  const VecLoop L = vec_loop<W>(n);
  double nrm[1] = {0.0};
  for (int64_t ip0 = L.tid; ip0 < L.npf; ip0 += U * L.stride) {
    Pack<T, W> xv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
      const bool ok = ip < L.npf;
      xv[u] = ok ? pk_ld<T, W, false>(x, ip) : pk_zero<T, W>();
      yv[u] = ok ? pk_ld<T, W, false>(y, ip) : pk_zero<T, W>();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t ip = ip0 + u * L.stride;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        Num<T>::fma(yv[u].e[w], a, xv[u].e[w]);
        if (NORM) nrm[0] += Num<T>::abs2(yv[u].e[w]);
      }
      if (ip < L.npf) pk_st<T, W>(y, ip, yv[u]);
    }
  }
  if (L.tail) {
    T v = y[L.itail];
    Num<T>::fma(v, a, x[L.itail]);
    y[L.itail] = v;
    if (NORM) nrm[0] += Num<T>::abs2(v);
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
  const T gamma = Num<T>::from_red(gv);
  const T mgamma = Num<T>::sub(Num<T>::zero(), gamma);
  const VecLoop L = vec_loop<W>(n);
  double nrm[1] = {0.0};
  for (int64_t ip = L.tid; ip < L.npf; ip += L.stride) {
    Pack<T, W> c = pk_ld<T, W, false>(cx, ip);
    Pack<T, W> u = pk_ld<T, W, false>(ux, ip);
    Pack<T, W> rv = pk_ld<T, W, false>(r, ip);
    Pack<T, W> xv = pk_ld<T, W, false>(x, ip);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      Num<T>::fma(rv.e[w], mgamma, c.e[w]);
      Num<T>::fma(xv.e[w], gamma, u.e[w]);
      nrm[0] += Num<T>::abs2(rv.e[w]);
    }
    pk_st<T, W>(r, ip, rv);
    pk_st<T, W>(x, ip, xv);
  }
  if (L.tail) {
    T rv = r[L.itail], xv = x[L.itail];
    Num<T>::fma(rv, mgamma, cx[L.itail]);
    Num<T>::fma(xv, gamma, ux[L.itail]);
    r[L.itail] = rv;
    x[L.itail] = xv;
    nrm[0] += Num<T>::abs2(rv);
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
  const VecLoop L = vec_loop<W>(n);
  T acc = Num<T>::zero();
  for (int64_t ip = L.tid; ip < L.npf; ip += L.stride) {
    Pack<T, W> c = pk_ld<T, W, false>(cx, ip);
    Pack<T, W> u = pk_ld<T, W, false>(ux, ip);
    Pack<T, W> rv = pk_ld<T, W, false>(r, ip);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      c.e[w] = Num<T>::scale(c.e[w], f);
      u.e[w] = Num<T>::scale(u.e[w], f);
      Num<T>::fmac(acc, c.e[w], rv.e[w]);
    }
    pk_st<T, W>(cx, ip, c);
    pk_st<T, W>(ux, ip, u);
  }
  if (L.tail) {
    T c = Num<T>::scale(cx[L.itail], f), u = Num<T>::scale(ux[L.itail], f);
    cx[L.itail] = c;
    ux[L.itail] = u;
    Num<T>::fmac(acc, c, r[L.itail]);
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
  const VecLoop L = vec_loop<W>(n);
  double nrm[1] = {0.0};
  for (int64_t ip = L.tid; ip < L.npf; ip += L.stride) {
    Pack<double, W> v = pk_ld<double, W, false>(vsrc, ip);
    Pack<double, W> a = pk_ld<double, W, false>(w1, ip);
    Pack<double, W> b = pk_ld<double, W, false>(w2, ip);
    Pack<double, W> xv = pk_ld<double, W, false>(x, ip);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      // same association as the reference expression (v - oldeps*w1 - delta*w2) * denom
      double wn = ((s * v.e[w] - oldeps * a.e[w]) - delta * b.e[w]) * denom;
      a.e[w] = wn;
      xv.e[w] = xv.e[w] + phi * wn;
      nrm[0] += xv.e[w] * xv.e[w];
    }
    pk_st<double, W>(w1, ip, a);
    pk_st<double, W>(x, ip, xv);
  }
  if (L.tail) {
    const int64_t i = L.itail;
    double wn = ((s * vsrc[i] - oldeps * w1[i]) - delta * w2[i]) * denom;
    w1[i] = wn;
    x[i] = x[i] + phi * wn;
    nrm[0] += x[i] * x[i];
  }
  grid_reduce<1>(nrm, partials, counter, out, gridDim.x, blockIdx.x);
}
