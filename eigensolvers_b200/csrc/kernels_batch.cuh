// kernels_batch.cuh — LOCK-STEP kernels for several independent shifted solves with the same operator
// (the nBlock solves of one block-Lanczos step, inexact_Lanczos.py:319-320; the m0 solves of one
// FEAST quadrature node, feast.py:190-201).  Each solve keeps its own Krylov basis, so the basis
// traffic is what it is; what the batch saves is
//   * the matrix stream: k_spmv_dia_nb reads every stored value ONCE for NB right-hand sides;
//   * the fixed cost of an Arnoldi step: k_orth_step_batch does the dots of all problems, ONE grid
//     barrier, all updates, and hands ONE mailbox message to the host — instead of a launch, a
//     barrier and a host round trip per problem.  At N = 1e6 (BASELINE config 2) those fixed costs
//     are a third of a step.
// Single GPU only (no halo push, no cross-rank reduction): BASELINE config 2 is a one-GPU config
// and node-per-GPU FEAST replicates H.  Sharded runs use the one-problem kernels.
#pragma once
#include "kernels_orth.cuh"
#include "kernels_dia.cuh"

constexpr int CV_MAX_BATCH = 4;    // problems per lock-step launch
constexpr int CV_BATCH_PTRS = 64;  // basis vectors per problem in a batched launch (GCROT(20,20) needs 62)

// ------------------------------------------------------------------------------------------
// y_q = op(x_q), q < NB, + per-problem dots {Re<x|y>, Im<x|y>, <y|y>} at out[3q..3q+3)
// ------------------------------------------------------------------------------------------
template <typename T, int NB>
struct DiaBatchArgs {
  const double *dia_val;
  int64_t ld;
  int n_diag, n_rows, mode;
  int off[CV_MAX_DIAG];
  const T *x[NB];
  T *y[NB];
  T sigma[NB];
  double *partials;
  unsigned *counter;
  double *out;
};

template <typename T, int NB>
__global__ void __launch_bounds__(CV_BLOCK, 3) k_spmv_dia_nb(const __grid_constant__ DiaBatchArgs<T, NB> a) {
  constexpr int DB = 4;  // diagonals per load batch: DB values + DB*NB x entries in flight
  const int n = a.n_rows;
  const int stride = gridDim.x * blockDim.x;
  T d_xy[NB];
  double d_yy[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) d_xy[q] = Num<T>::zero(), d_yy[q] = 0.0;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
    T acc[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) acc[q] = Num<T>::zero();
    const double *vp = a.dia_val + row;
    for (int d0 = 0; d0 < a.n_diag; d0 += DB) {
      double v[DB];
      T xv[DB][NB];
#pragma unroll
      for (int k = 0; k < DB; ++k) {
        const bool ok = d0 + k < a.n_diag;
        v[k] = ok ? ld_stream(vp + (int64_t)(d0 + k) * a.ld) : 0.0;
        int i = row + (ok ? a.off[d0 + k] : 0);
        i = i < 0 ? 0 : (i >= n ? n - 1 : i);  // out-of-range only through zero padding values
#pragma unroll
        for (int q = 0; q < NB; ++q) xv[k][q] = ld_gather(a.x[q] + i);
      }
#pragma unroll
      for (int k = 0; k < DB; ++k)
#pragma unroll
        for (int q = 0; q < NB; ++q) Num<T>::fmar(acc[q], v[k], xv[k][q]);
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const T xr = ld_gather(a.x[q] + row);
      T r;
      if (a.mode == CV_SPMV_PLAIN)
        r = acc[q];
      else if (a.mode == CV_SPMV_SHIFT)
        r = Num<T>::sub(Num<T>::mul(a.sigma[q], xr), acc[q]);
      else
        r = Num<T>::sub(acc[q], Num<T>::mul(a.sigma[q], xr));
      st_plain(a.y[q] + row, r);
      Num<T>::fmac(d_xy[q], xr, r);
      d_yy[q] += Num<T>::abs2(r);
    }
  }
  double vals[3 * NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    vals[3 * q] = 0.0;
    vals[3 * q + 1] = 0.0;
    Num<T>::to_red(d_xy[q], vals + 3 * q);
    vals[3 * q + 2] = d_yy[q];
  }
  grid_reduce<3 * NB>(vals, a.partials, a.counter, a.out, gridDim.x, blockIdx.x);
}

// ------------------------------------------------------------------------------------------
// Fused Arnoldi step of up to CV_MAX_BATCH independent problems (single GPU).  Per problem the
// arithmetic is that of k_orth_step; the slots (relative to the problem's sbase) are the ones the
// one-problem solver uses.
// ------------------------------------------------------------------------------------------
struct OrthProb {
  const void *v[CV_BATCH_PTRS];  // basis [C,V]
  void *w;                       // vector being orthogonalised (in/out)
  int m;
  int sbase;    // first scalar slot of this problem's block
  int s_w_src;  // where the batched SpMV left this problem's {Re<x|y>, Im<x|y>, <y|y>}
  double eta2;
};
struct OrthBatchArgs {
  OrthProb prob[CV_MAX_BATCH];
  int nprob;
  int64_t n;
  // slot offsets inside a problem's block (solvers.cu)
  int o_flag, o_nrm, o_w, o_h1, o_h2, o_lag;
  double *partials;  // per problem: pstride doubles
  int64_t pstride;
  unsigned *bar;     // [0] arrivals, [1] generation
  unsigned *ticket;
  double *scal;
  int slab_mode, snake;
  double *host_mb;
  unsigned long long *host_flag;
  unsigned long long host_seq;
};

// Work split of one pass over all problems that take part in it.  Phase A: every (problem, slab of
// <= MI vectors) pair gets a contiguous range of CTAs proportional to its loads per row (mi + 1);
// phase B: every problem gets a range proportional to (m + 2).  All problems stream CONCURRENTLY —
// one ramp-up and one tail per phase for the whole batch, and at small N (1e6 rows, 4 problems) a
// CTA's stream is 4x longer than when the whole grid works on one problem after the other.
constexpr int CV_BATCH_SLABS = CV_MAX_BATCH * (CV_BATCH_PTRS / 8);
struct BatchMap {
  int nslab;
  int q[CV_BATCH_SLABS], i0[CV_BATCH_SLABS], mi[CV_BATCH_SLABS];
  int start[CV_BATCH_SLABS + 1];   // phase A: first CTA of each slab
  int bstart[CV_MAX_BATCH + 1];    // phase B: first CTA of each problem (inactive problems: empty range)
};
__host__ __device__ inline void batch_map_build(const OrthBatchArgs &a, unsigned active_mask, int G, int MI, BatchMap &M) {
  int ns = 0, wsum = 0, bsum = 0;
  for (int q = 0; q < a.nprob; ++q) {
    if (!((active_mask >> q) & 1u)) continue;
    const int m = a.prob[q].m;
    const int ny = (m + MI - 1) / MI;
    const int base = m / ny, rem = m % ny;
    int at = 0;
    for (int by = 0; by < ny; ++by) {
      M.q[ns] = q;
      M.i0[ns] = at;
      M.mi[ns] = base + (by < rem ? 1 : 0);
      at += M.mi[ns];
      wsum += M.mi[ns] + 1;
      ++ns;
    }
    bsum += m + 2;
  }
  M.nslab = ns;
  int acc = 0;
  for (int s = 0; s < ns; ++s) {
    M.start[s] = acc;
    int g = (int)(((long long)G * (M.mi[s] + 1)) / wsum);
    if (g < 1) g = 1;
    const int remaining = ns - 1 - s;
    if (acc + g > G - remaining) g = G - remaining - acc;
    if (s == ns - 1) g = G - acc;
    acc += g;
  }
  M.start[ns] = G;
  acc = 0;
  int left = 0;
  for (int q = 0; q < a.nprob; ++q) left += (active_mask >> q) & 1u;
  for (int q = 0; q < a.nprob; ++q) {
    M.bstart[q] = acc;
    if (!((active_mask >> q) & 1u)) continue;
    --left;
    int g = (int)(((long long)G * (a.prob[q].m + 2)) / bsum);
    if (g < 1) g = 1;
    if (acc + g > G - left) g = G - left - acc;
    if (left == 0) g = G - acc;
    acc += g;
  }
  M.bstart[a.nprob] = G;
}

template <typename T, int W>
__global__ void __launch_bounds__(CV_BLOCK, 3) k_orth_step_batch(const __grid_constant__ OrthBatchArgs a) {
  constexpr int NR = Num<T>::NRED;
  constexpr int MI = ORTH_MI<T>::value, LB = 8, JB = 8;
  extern __shared__ double s_h[];  // max_q m_q * NR doubles
  __shared__ double s_part[CV_WARPS][MI * NR + 1];
  __shared__ double s_vals[MI * NR + 1];
  __shared__ BatchMap s_map;
  __shared__ int s_state[CV_MAX_BATCH];  // after the barrier of a pass: 1 = this problem needs pass 2
  __shared__ unsigned s_mask;
  const int G = gridDim.x, c = blockIdx.x;
  const int64_t n = a.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t npf = n / W;
  const bool tail_mine = (W == 2) && (n & 1);
  double *p_ww = a.partials + (size_t)MI * NR * G;  // pass 2: [q][512] partial |w'|^2 of the problem's first slab
  double *p_nx = p_ww + CV_MAX_BATCH * 512;         // [q][512] explicit |v_new|^2 partials of the problem's phase-B CTAs
  unsigned final_mask1 = 0;                         // problems that were final in pass 1 (for the tail)

  for (int pass = 1; pass <= 2; ++pass) {
    // problems taking part in this pass, and the work split over them
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned mask = 0;
      for (int q = 0; q < a.nprob; ++q)
        if (pass == 1 || __ldcg(a.scal + a.prob[q].sbase + a.o_flag) != 0.0) mask |= 1u << q;
      s_mask = mask;
      batch_map_build(a, mask, G, MI, s_map);
    }
    __syncthreads();
    const unsigned mask = s_mask;
    // ---------------- phase A: this CTA's (problem, slab) ----------------------------------------
    {
      int sl_i = 0;
      while (sl_i + 1 < s_map.nslab && c >= s_map.start[sl_i + 1]) ++sl_i;
      const int q = s_map.q[sl_i];
      const OrthProb &P = a.prob[q];
      const int bx = c - s_map.start[sl_i];
      const int gx = s_map.start[sl_i + 1] - s_map.start[sl_i];
      const int i0 = s_map.i0[sl_i];
      const int mi = s_map.mi[sl_i];
      const int half = (mi > LB) ? (mi + 1) / 2 : mi;
      const bool want_ww = pass == 2 && i0 == 0;
      const T *wvec = static_cast<const T *>(P.w);
      T acc[MI];
      double ww = 0.0;
#pragma unroll
      for (int i = 0; i < MI; ++i) acc[i] = Num<T>::zero();
      for (int64_t ip = (int64_t)bx * blockDim.x + threadIdx.x; ip < npf; ip += (int64_t)gx * blockDim.x) {
        const Pack<T, W> wv = pk_ld_cg<T, W>(wvec, ip);
#pragma unroll
        for (int b = 0; b < MI / LB; ++b) {
          Pack<T, W> vv[LB];
#pragma unroll
          for (int l = 0; l < LB; ++l) {
            const int idx = b ? half + l : l;
            const bool ok = b ? idx < mi : l < half;
            vv[l] = ok ? pk_ld<T, W, false>(static_cast<const T *>(P.v[i0 + idx]), ip) : pk_zero<T, W>();
          }
#pragma unroll
          for (int l = 0; l < LB; ++l)
#pragma unroll
            for (int w = 0; w < W; ++w) Num<T>::fmac(acc[b * LB + l], vv[l].e[w], wv.e[w]);
        }
        if (want_ww) {
#pragma unroll
          for (int w = 0; w < W; ++w) ww += Num<T>::abs2(wv.e[w]);
        }
      }
      if (tail_mine && bx == 0 && threadIdx.x == 0) {
        const T wt = ld_cg(wvec + (n - 1));
#pragma unroll
        for (int sl = 0; sl < MI; ++sl) {
          const int idx = sl < LB ? sl : half + (sl - LB);
          const bool ok = sl < LB ? sl < half : idx < mi;
          if (ok) Num<T>::fmac(acc[sl], static_cast<const T *>(P.v[i0 + idx])[n - 1], wt);
        }
        if (want_ww) ww += Num<T>::abs2(wt);
      }
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        double r[NR];
        Num<T>::to_red(acc[i], r);
#pragma unroll
        for (int k = 0; k < NR; ++k) {
          const double s = warp_sum(r[k]);
          if (lane == 0) s_part[warp][i * NR + k] = s;
        }
      }
      {
        const double s = warp_sum(ww);
        if (lane == 0) s_part[warp][MI * NR] = s;
      }
      __syncthreads();
      for (int v = threadIdx.x; v < MI * NR + 1; v += blockDim.x) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < CV_WARPS; ++w) s += s_part[w][v];
        s_vals[v] = s;
      }
      __syncthreads();
      double *pslab = a.partials + (size_t)MI * NR * s_map.start[sl_i];
      for (int v = threadIdx.x; v < mi * NR; v += blockDim.x) {
        const int i = v / NR, k = v - i * NR;
        const int sl = i < half ? i : LB + (i - half);
        pslab[(size_t)v * gx + bx] = s_vals[sl * NR + k];
      }
      if (want_ww && threadIdx.x == 0) p_ww[q * 512 + bx] = s_vals[MI * NR];
    }
    // ---------------- ONE barrier: the last CTA finishes every problem's reduction ----------------
    grid_barrier_with(a.bar, [&]() {
      bool any_again = false;
      for (int q = 0; q < a.nprob; ++q) {
        if (threadIdx.x == 0) s_state[q] = 0;
        if (!((mask >> q) & 1u)) continue;  // uniform over the CTA
        const OrthProb &P = a.prob[q];
        const int m = P.m;
        const int s_h_out = P.sbase + (pass == 1 ? a.o_h1 : a.o_h2);
        int first = 0;  // first slab of this problem
        while (s_map.q[first] != q) ++first;
        for (int v = warp; v < m * NR; v += CV_WARPS) {
          int sl_i = first;
          while (sl_i + 1 < s_map.nslab && s_map.q[sl_i + 1] == q && v / NR >= s_map.i0[sl_i + 1]) ++sl_i;
          const int local = v - s_map.i0[sl_i] * NR;
          const int gx = s_map.start[sl_i + 1] - s_map.start[sl_i];
          const double *pp = a.partials + (size_t)MI * NR * s_map.start[sl_i] + (size_t)local * gx;
          double r = ordered_lane_sum(pp, gx, lane);
          r = warp_sum(r);
          if (lane == 0) a.scal[s_h_out + v] = r;
        }
        if (pass == 2 && warp == 0) {
          double r = ordered_lane_sum(p_ww + q * 512, s_map.start[first + 1] - s_map.start[first], lane);
          r = warp_sum(r);
          if (lane == 0) a.scal[s_h_out + m * NR] = r;
        }
        if (pass == 1 && threadIdx.x < 3) a.scal[P.sbase + a.o_w + threadIdx.x] = __ldcg(a.scal + P.s_w_src + threadIdx.x);
        if (pass == 1 && threadIdx.x == 32) a.scal[P.sbase + a.o_lag] = __ldcg(a.scal + P.sbase + a.o_nrm);
        __threadfence();
        __syncthreads();
        if (warp == 0) {
          double qq = 0.0;
          for (int v = lane; v < m * NR; v += 32) {
            const double hv = __ldcg(a.scal + s_h_out + v);
            qq = fma(hv, hv, qq);
          }
          qq = warp_sum(qq);
          if (lane == 0) {
            const double base = pass == 1 ? __ldcg(a.scal + P.sbase + a.o_w + 2) : __ldcg(a.scal + s_h_out + m * NR);
            double t = base - qq;
            bool again = false;
            if (pass == 1) {
              again = !(t >= P.eta2 * base);
              a.scal[P.sbase + a.o_flag] = again ? 1.0 : 0.0;
            } else if (!(t > 0.0)) {
              t = 0.0;
            }
            if (!again) a.scal[P.sbase + a.o_nrm] = t;
            s_state[q] = again ? 1 : 0;
          }
        }
        __syncthreads();
        any_again = any_again || s_state[q] != 0;
      }
      if (pass == 2 || !any_again) {
        // every Hessenberg column is complete: ONE message to the host for all problems
        __threadfence();
        for (int q = 0; q < a.nprob; ++q) {
          const OrthProb &P = a.prob[q];
          const int cnt = a.o_h1 + P.m * NR - a.o_flag;
          for (int t = threadIdx.x; t < cnt; t += blockDim.x)
            a.host_mb[P.sbase + a.o_flag + t] = __ldcg(a.scal + P.sbase + a.o_flag + t);
          for (int t = threadIdx.x; t < P.m * NR; t += blockDim.x)
            a.host_mb[P.sbase + a.o_h2 + t] = __ldcg(a.scal + P.sbase + a.o_h2 + t);
          if (threadIdx.x == 0) a.host_mb[P.sbase + a.o_lag] = __ldcg(a.scal + P.sbase + a.o_lag);
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(a.host_flag, a.host_seq);
      }
    });
    // ---------------- phase B: this CTA's problem: w <- (w - [C,V] h) [/ |w'| when final] ----------
    bool any_again = false;
    for (int q = 0; q < a.nprob; ++q)
      if (((mask >> q) & 1u) && pass == 1 && __ldcg(a.scal + a.prob[q].sbase + a.o_flag) != 0.0) any_again = true;
    {
      int q = 0;
      while (q + 1 < a.nprob && c >= s_map.bstart[q + 1]) ++q;
      while (!((mask >> q) & 1u)) ++q;  // empty ranges of inactive problems share their start with the next one
      const OrthProb &P = a.prob[q];
      const int bc = c - s_map.bstart[q];
      int qn = q + 1;
      while (qn < a.nprob && !((mask >> qn) & 1u)) ++qn;
      const int bg = (qn < a.nprob ? s_map.bstart[qn] : G) - s_map.bstart[q];
      const bool final_pass = pass == 2 || __ldcg(a.scal + P.sbase + a.o_flag) == 0.0;
      if (pass == 1 && final_pass) final_mask1 |= 1u << q;
      const int m = P.m;
      const int s_h_out = P.sbase + (pass == 1 ? a.o_h1 : a.o_h2);
      T *wvec = static_cast<T *>(P.w);
      double f = 1.0;
      if (final_pass) {
        f = 1.0 / sqrt(__ldcg(a.scal + P.sbase + a.o_nrm));
        if (!isfinite(f)) f = 1.0;
      }
      for (int j = threadIdx.x; j < m * NR; j += blockDim.x) s_h[j] = -__ldcg(a.scal + s_h_out + j);
      __syncthreads();
      double nx = 0.0;
      // rows downwards: phase A's tail is still in L2 (see k_orth_step)
      const int64_t b_first = (int64_t)bc * blockDim.x + threadIdx.x, b_stride = (int64_t)bg * blockDim.x;
      const int64_t b_count = b_first < npf ? (npf - b_first + b_stride - 1) / b_stride : 0;
      for (int64_t kk = 0; kk < b_count; ++kk) {
        const int64_t ip = a.snake ? b_first + (b_count - 1 - kk) * b_stride : b_first + kk * b_stride;
        Pack<T, W> acc = pk_ld_cg<T, W>(wvec, ip);
        for (int j0 = 0; j0 < m; j0 += JB) {
          Pack<T, W> v[JB];
#pragma unroll
          for (int jj = 0; jj < JB; ++jj)
            v[jj] = (j0 + jj < m) ? pk_ld<T, W, false>(static_cast<const T *>(P.v[j0 + jj]), ip) : pk_zero<T, W>();
#pragma unroll
          for (int jj = 0; jj < JB; ++jj) {
            if (j0 + jj < m) {
              const T mc = Num<T>::from_red(s_h + (j0 + jj) * NR);
#pragma unroll
              for (int w = 0; w < W; ++w) Num<T>::fma(acc.e[w], mc, v[jj].e[w]);
            }
          }
        }
#pragma unroll
        for (int w = 0; w < W; ++w) {
          acc.e[w] = Num<T>::scale(acc.e[w], f);
          nx += Num<T>::abs2(acc.e[w]);
        }
        pk_st<T, W>(wvec, ip, acc);
      }
      if (tail_mine && bc == 0 && threadIdx.x == 0) {
        T acc = ld_cg(wvec + (n - 1));
        for (int j = 0; j < m; ++j)
          Num<T>::fma(acc, Num<T>::from_red(s_h + j * NR), static_cast<const T *>(P.v[j])[n - 1]);
        acc = Num<T>::scale(acc, f);
        wvec[n - 1] = acc;
        nx += Num<T>::abs2(acc);
      }
      if (final_pass) {  // explicit |v_new|^2 partial of this CTA (health monitor of the NEXT step)
        const double s = warp_sum(nx);
        __syncthreads();
        if (lane == 0) s_part[warp][0] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
          double r = 0.0;
#pragma unroll
          for (int w = 0; w < CV_WARPS; ++w) r += s_part[w][0];
          p_nx[q * 512 + bc] = r;
          if (bc == 0) p_nx[CV_MAX_BATCH * 512 + q] = (double)bg;  // how many partials this problem has
        }
      }
    }
    if (pass == 2 || !any_again) break;
    grid_barrier_with(a.bar, [&]() {});  // pass 2 reads rows other CTAs have just rewritten
  }
  (void)final_mask1;
  // ---------------- kernel tail: the last CTA to finish sums the explicit |v_new|^2 partials ---------
  __syncthreads();
  __shared__ bool s_last_b;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last_b = atomicAdd(a.ticket, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (s_last_b) {
    __threadfence();
    for (int q = warp; q < a.nprob; q += CV_WARPS) {
      const int cnt = (int)__ldcg(p_nx + CV_MAX_BATCH * 512 + q);
      double r = ordered_lane_sum(p_nx + q * 512, cnt, lane);
      r = warp_sum(r);
      if (lane == 0) a.scal[a.prob[q].sbase + a.o_nrm] = r;  // read as the lagged norm by the next step
    }
    if (threadIdx.x == 0) *a.ticket = 0u;
  }
}
