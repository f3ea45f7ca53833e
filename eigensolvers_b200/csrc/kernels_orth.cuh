// kernels_orth.cuh — ONE persistent kernel per Arnoldi step of GCROT's inner FGMRES
// (_gcrotmk.py:112-141): everything between two operator applications, with ONE grid barrier and
// ONE cross-GPU all-reduce on the critical path.
//
//   phase A   h = [C,V]^H w                          (tall-skinny dot, V read once)
//   barrier   the last CTA to arrive sums the per-CTA partials in a fixed order, all-reduces
//             {<x|y>, <y|y>, h} over the ranks through peer memory (one NVLink one-way latency) and
//             obtains the norm of the projected vector WITHOUT a second reduction:
//                 |w'|^2 = |w|^2 - sum |h_i|^2        ([C,V] orthonormal; |w|^2 comes fused out of the SpMV)
//             Daniel-Gragg-Kaufman-Stewart test: if |w'|^2 < eta^2 |w|^2 the subtraction cancelled
//             (and the formula lost digits) -> pass 2 below.  Otherwise the Hessenberg column is
//             complete: it is written to the HOST mailbox (mapped pinned memory) and a sequence flag
//             released, so the host builds its Givens rotation and queues the next SpMV while
//   phase B   w <- (w - [C,V] h) / |w'|  streams through HBM; rows a neighbour needs are stored
//             into its halo buffer as they are produced (NVLink traffic hidden behind the HBM
//             stream); the last CTA to finish raises the neighbours' halo flags.
//   pass 2    (rare) w' is not normalised, a grid barrier, then phase A on w' (h2 and |w'|^2 in
//             the same pass), barrier, phase B with h2 and |w''|^2 = |w'|^2 - sum |h2_i|^2.
//
// The separate-kernel version of the same step needed 8 launches, 2 copies and a stream
// synchronisation (tsdot, all-reduce, tsupdate, all-reduce, D2H, scale, halo push, halo wait);
// at 8 GPUs that fixed cost (~90 us) was a third of the step.
#pragma once
#include "kernels_vec.cuh"

struct OrthArgs {
  TsParams p;      // v[0..m) = basis [C,V];  w[0] = vector being orthogonalised (in/out);  n
  double *partials;
  unsigned *bar;   // [0] arrivals, [1] generation
  double *scal;    // ctx->scalars
  int s_flag;      // out: 1.0 if the second pass ran
  int s_nrm;       // out: |w'|^2 after the last pass (summed over ranks)
  int s_w;         // in: {Re<x|y>, Im<x|y>, <y|y>} partial dots of the SpMV (this rank), out: summed
  int s_h1, s_h2;  // out: projection coefficients of pass 1 / pass 2  (s_h1 == s_w + 3)
  int s_lag;       // out: explicitly summed |v_j|^2 of the PREVIOUS step's normalised vector (should be 1;
                   // its local partial arrives in scal[s_nrm] and rides in this step's all-reduce)
  double eta2;
  // peers (world == 1: unused)
  PeerPtrs pp;
  int me, world;
  double *err;
  PushArgs push;   // nseg == 0 && nflag == 0: nothing to push
  int push_early;  // all segments are contiguous ranges: phase B stores the new rows into the peers'
                   // halo buffers as it produces them; otherwise (gather lists) a late push follows
  int slab_mode;     // phase A work split, see SlabMap
  int snake;         // phase B walks the rows downwards (L2 reuse of phase A's tail)
  int publish_late;  // debug: hand the scalars to the host at the END of the kernel
  double *trace;   // [16] accumulated phase times of CTA 0 in ns (tools/orth_trace.py)
  // host mailbox
  double *host_mb;
  unsigned long long *host_flag;
  unsigned long long host_seq;
};

// Grid barrier with serial work: every CTA arrives; the last one to arrive runs `work` (all its
// threads), then releases the others.  Needs all CTAs co-resident (cooperative launch).
template <typename F, typename G>
__device__ __forceinline__ void grid_barrier_with(unsigned *bar, F &&work, G &&after_release) {
  __shared__ bool s_last_bar;
  __shared__ unsigned s_gen_bar;
  __syncthreads();
  if (threadIdx.x == 0) {
    s_gen_bar = ld_acquire_gpu(bar + 1);  // read BEFORE arriving: cannot advance until we arrive
    __threadfence();
    s_last_bar = atomicAdd(bar, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (s_last_bar) {
    __threadfence();
    work();
    __syncthreads();
    if (threadIdx.x == 0) {
      bar[0] = 0u;
      __threadfence();
      st_release_gpu(bar + 1, s_gen_bar + 1u);
    }
    after_release();  // off the critical path of the other CTAs
  } else if (threadIdx.x == 0) {
    while (ld_acquire_gpu(bar + 1) == s_gen_bar) {
    }
  }
  __syncthreads();
}
template <typename F>
__device__ __forceinline__ void grid_barrier_with(unsigned *bar, F &&work) {
  grid_barrier_with(bar, work, []() {});
}

template <typename T, int W>
__device__ __forceinline__ Pack<T, W> pk_ld_cg(const T *base, int64_t ip) {
  Pack<T, W> p;
  if constexpr (W == 1) {
    p.e[0] = ld_cg(base + ip);
  } else {
    double2 v = __ldcg(reinterpret_cast<const double2 *>(base) + ip);
    p.e[0] = v.x;
    p.e[1] = v.y;
  }
  return p;
}

// sum p[lane], p[lane+32], ... in that fixed order with all loads of a batch in flight at once
// (the serial section of a grid barrier: every microsecond here is paid by all CTAs)
__device__ __forceinline__ double ordered_lane_sum(const double *p, int count, int lane) {
  double r = 0.0;
  for (int b0 = lane; b0 < count; b0 += 32 * 8) {
    double t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = (b0 + 32 * k < count) ? __ldcg(p + b0 + 32 * k) : 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += t[k];
  }
  return r;
}

// store the freshly computed pack `ip` of w into every peer range that contains it
template <typename T, int W>
__device__ __forceinline__ bool push_pack(const PushArgs &a, int64_t ip, const Pack<T, W> &v) {
  bool any = false;
  for (int s = 0; s < a.nseg; ++s) {
    const PushSeg &g = a.seg[s];
    const int64_t e0 = ip * W - g.src_start;  // position of the pack's first element in the range
    if (e0 + W <= 0 || e0 >= g.count) continue;
    T *dst = static_cast<T *>(g.dst);
    if constexpr (W == 2) {
      if (e0 >= 0 && e0 + 1 < g.count && (((uintptr_t)(dst + e0)) & 15) == 0) {
        *reinterpret_cast<double2 *>(dst + e0) = make_double2(v.e[0], v.e[1]);
      } else {
        if (e0 >= 0 && e0 < g.count) st_plain(dst + e0, v.e[0]);
        if (e0 + 1 >= 0 && e0 + 1 < g.count) st_plain(dst + e0 + 1, v.e[1]);
      }
    } else {
      st_plain(dst + e0, v.e[0]);
    }
    any = true;
  }
  return any;
}

// Phase A splits the basis into slabs of <= MI vectors (the accumulators of one thread) and gives
// every slab a share of the CTAs PROPORTIONAL to its load count (mi vectors + w per row).
//   mode 0  full slabs of MI and a remainder (m = 28 -> 16 + 12, m = 33 -> 16 + 16 + 1), loads in
//           batches of LB: the remainder slab's threads keep few loads in flight (one vector + w at
//           m = 33), so its CTAs run far below the bandwidth their share assumes and the whole grid
//           waits for them at the barrier;
//   mode 1  slabs as even as possible (33 -> 11 + 11 + 11) and, inside a slab of more than LB
//           vectors, two load batches of equal size (11 -> 6 + 5 instead of 8 + 3): every thread of
//           the grid has the same, balanced number of loads in flight.
struct SlabMap {
  int ny;
  int start[CV_MAX_PTRS / 8 + 2];  // first CTA of slab by; start[ny] = G
  int i0[CV_MAX_PTRS / 8 + 2];     // first basis vector of slab by; i0[ny] = m
};
__host__ __device__ inline void orth_slab_map(int m, int G, int MI, int mode, SlabMap &s) {
  s.ny = (m + MI - 1) / MI;
  const int base = m / s.ny, rem = m % s.ny;
  int at = 0;
  for (int by = 0; by < s.ny; ++by) {
    s.i0[by] = at;
    at += mode ? base + (by < rem ? 1 : 0) : ((m - by * MI) < MI ? (m - by * MI) : MI);
  }
  s.i0[s.ny] = m;
  const int total = m + s.ny;
  int acc = 0;
  for (int by = 0; by < s.ny; ++by) {
    s.start[by] = acc;
    const int mi = s.i0[by + 1] - s.i0[by];
    int g = (int)(((long long)G * (mi + 1)) / total);
    if (g < 1) g = 1;
    const int remaining = s.ny - 1 - by;  // at least one CTA for each later slab
    if (acc + g > G - remaining) g = G - remaining - acc;
    if (by == s.ny - 1) g = G - acc;
    acc += g;
  }
  s.start[s.ny] = G;
}

template <typename T>
struct ORTH_MI {
  static constexpr int value = sizeof(T) == 16 ? 8 : 16;
};

#define ORTH_TRACE(slot)                                  \
  do {                                                    \
    if (a.trace && c == 0 && threadIdx.x == 0) {          \
      const unsigned long long t_now = global_ns();       \
      a.trace[slot] += (double)(t_now - t_prev);          \
      t_prev = t_now;                                     \
    }                                                     \
  } while (0)

template <typename T, int W>
__global__ void __launch_bounds__(CV_BLOCK, 3) k_orth_step(const __grid_constant__ OrthArgs a) {
  constexpr int NR = Num<T>::NRED;
  // accumulators per thread / loads per batch: 16 / 8 for real data; complex128 accumulators take
  // two registers pairs each, 16 of them spilled at the 80-register budget (3 CTAs/SM), so 8 / 8
  constexpr int MI = ORTH_MI<T>::value, LB = 8, JB = 8;
  extern __shared__ double s_h[];  // m * NR doubles
  __shared__ double s_part[CV_WARPS][MI * NR + 1];
  __shared__ double s_vals[MI * NR + 1];
  __shared__ bool s_again;
  const int G = gridDim.x, c = blockIdx.x;
  const int m = a.p.m;
  const int64_t n = a.p.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *wvec = static_cast<T *>(const_cast<void *>(a.p.w[0]));
  __shared__ SlabMap s_map;  // the launcher makes G >= ny
  if (threadIdx.x == 0) orth_slab_map(m, G, MI, a.slab_mode, s_map);
  __syncthreads();
  int my_by = 0;
  while (my_by + 1 < s_map.ny && c >= s_map.start[my_by + 1]) ++my_by;
  const int64_t npf = n / W;
  const bool tail_mine = (W == 2) && (n & 1);
  double *p_ww = a.partials + (size_t)MI * NR * G;  // pass 2: partial |w'|^2 of the by == 0 CTAs
  double *p_nx = p_ww + 2048;                       // explicit |v_new|^2 partials, one per CTA
  unsigned long long t_prev = 0;
  if (a.trace && c == 0 && threadIdx.x == 0) {
    t_prev = global_ns();
    a.trace[5] += 1.0;
  }
  bool pushed = false;

  for (int pass = 1; pass <= 2; ++pass) {
    const int s_h_out = pass == 1 ? a.s_h1 : a.s_h2;
    // ---------------- phase A: this CTA's slab of up to MI basis vectors against w ----------
    {
      // slabs of MI vectors, the last one partial (m = 28 -> 16+12).  Splitting the VECTORS evenly was
      // measured SLOWER at N = 2e6 (dots 89 us -> 115 us with batches of 8+3 loads per thread, 103 us
      // with even slabs AND even batches 7+7): the phase is limited by loads in flight per thread, so
      // the slabs keep full batches of 8 and the CTA shares are balanced instead (SlabMap)
      const int by = my_by, bx = c - s_map.start[my_by];
      const int gx = s_map.start[my_by + 1] - s_map.start[my_by];
      const int i0 = s_map.i0[by];
      const int mi = s_map.i0[by + 1] - i0;
      // two load batches: vectors [0, half) accumulate in acc[0..LB), vectors [half, mi) in acc[LB..2LB)
      // (static register indices; only the pointer selection depends on `half`)
      const int half = (a.slab_mode && mi > LB) ? (mi + 1) / 2 : (mi < LB ? mi : LB);
      const bool want_ww = pass == 2 && by == 0;
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
            vv[l] = ok ? pk_ld<T, W, false>(static_cast<const T *>(a.p.v[i0 + idx]), ip) : pk_zero<T, W>();
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
          if (ok) Num<T>::fmac(acc[sl], static_cast<const T *>(a.p.v[i0 + idx])[n - 1], wt);
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
      double *pslab = a.partials + (size_t)MI * NR * s_map.start[by];
      for (int v = threadIdx.x; v < mi * NR; v += blockDim.x) {
        const int i = v / NR, k = v - i * NR;
        const int sl = i < half ? i : LB + (i - half);  // accumulator slot of vector i
        pslab[(size_t)v * gx + bx] = s_vals[sl * NR + k];
      }
      if (want_ww && threadIdx.x == 0) p_ww[bx] = s_vals[MI * NR];
    }
    ORTH_TRACE(0);
    grid_barrier_with(a.bar, [&]() {
      for (int v = warp; v < m * NR; v += CV_WARPS) {
        int by = 0;
        while (v / NR >= s_map.i0[by + 1]) ++by;
        const int local = v - s_map.i0[by] * NR;
        const int gx = s_map.start[by + 1] - s_map.start[by];
        const double *pp = a.partials + (size_t)MI * NR * s_map.start[by] + (size_t)local * gx;
        double r = ordered_lane_sum(pp, gx, lane);
        r = warp_sum(r);
        if (lane == 0) a.scal[s_h_out + v] = r;
      }
      if (pass == 2 && warp == 0) {
        double r = ordered_lane_sum(p_ww, s_map.start[1], lane);
        r = warp_sum(r);
        if (lane == 0) a.scal[s_h_out + m * NR] = r;  // travels with h2 in one all-reduce
      }
      __threadfence();
      if (a.world > 1) {
        if (pass == 1)  // {lagged |v_j|^2, <x|y>, <y|y>, h}: s_nrm, s_w, s_h1 are contiguous
          cta_peer_allreduce(a.pp, a.me, a.world, a.scal + a.s_nrm, 4 + m * NR, a.err);
        else
          cta_peer_allreduce(a.pp, a.me, a.world, a.scal + s_h_out, m * NR + 1, a.err);
      }
      __syncthreads();
      if (pass == 1 && threadIdx.x == 0) a.scal[a.s_lag] = __ldcg(a.scal + a.s_nrm);
      __syncthreads();
      // |w'|^2 = |w|^2 - sum |h_i|^2
      if (warp == 0) {
        double q = 0.0;
        for (int v = lane; v < m * NR; v += 32) {
          const double hv = __ldcg(a.scal + s_h_out + v);
          q = fma(hv, hv, q);
        }
        q = warp_sum(q);
        if (lane == 0) {
          const double base = pass == 1 ? __ldcg(a.scal + a.s_w + 2) : __ldcg(a.scal + s_h_out + m * NR);
          double t = base - q;
          bool again = false;
          if (pass == 1) {
            again = !(t >= a.eta2 * base);
            a.scal[a.s_flag] = again ? 1.0 : 0.0;
          } else if (!(t > 0.0)) {
            t = 0.0;  // fully dependent on the basis: scipy's breakdown branch (hlast <= eps |w|)
          }
          if (!again) a.scal[a.s_nrm] = t;
          s_again = again;
        }
      }
      __syncthreads();
    }, [&]() {
      if (!s_again && !a.publish_late) {
        // the Hessenberg column is complete: hand it to the host while phase B runs
        const bool two = pass == 2;
        for (int t = threadIdx.x; t < a.s_h1 + m * NR - a.s_flag; t += blockDim.x)
          a.host_mb[a.s_flag + t] = __ldcg(a.scal + a.s_flag + t);
        if (two)
          for (int t = threadIdx.x; t < m * NR; t += blockDim.x) a.host_mb[a.s_h2 + t] = __ldcg(a.scal + a.s_h2 + t);
        if (threadIdx.x == 0) {
          a.host_mb[a.err - a.scal] = __ldcg(a.err);
          a.host_mb[a.s_lag] = __ldcg(a.scal + a.s_lag);
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(a.host_flag, a.host_seq);
      }
    });
    ORTH_TRACE(1);
    // ---------------- phase B: w <- (w - [C,V] h) [/ |w'| when final] ------------------------
    const bool final_pass = pass == 2 || __ldcg(a.scal + a.s_flag) == 0.0;
    double f = 1.0;
    if (final_pass) {
      f = 1.0 / sqrt(__ldcg(a.scal + a.s_nrm));
      if (!isfinite(f)) f = 1.0;  // scipy: normalise only "if isfinite(alpha)"
    }
    for (int j = threadIdx.x; j < m * NR; j += blockDim.x) s_h[j] = -__ldcg(a.scal + s_h_out + j);
    __syncthreads();
    double nx = 0.0;
    // Phase B walks the rows in the OPPOSITE direction of phase A ("snake"): phase A has just streamed
    // [C,V] and w in ascending order, so the last ~126 MB of it — the high rows of every vector — are
    // still in L2 when phase B starts there.  Nothing at 160 MB per vector; at 8-20 MB per vector
    // (BASELINE config 2, or config 3 on 8 GPUs) it serves a good part of phase B from L2.
    const int64_t b_first = (int64_t)c * blockDim.x + threadIdx.x, b_stride = (int64_t)G * blockDim.x;
    const int64_t b_count = b_first < npf ? (npf - b_first + b_stride - 1) / b_stride : 0;
    for (int64_t kk = 0; kk < b_count; ++kk) {
      const int64_t ip = a.snake ? b_first + (b_count - 1 - kk) * b_stride : b_first + kk * b_stride;
      Pack<T, W> acc = pk_ld_cg<T, W>(wvec, ip);
      for (int j0 = 0; j0 < m; j0 += JB) {
        Pack<T, W> v[JB];
#pragma unroll
        for (int jj = 0; jj < JB; ++jj)
          v[jj] = (j0 + jj < m) ? pk_ld<T, W, false>(static_cast<const T *>(a.p.v[j0 + jj]), ip) : pk_zero<T, W>();
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
      if (final_pass && a.push_early) pushed |= push_pack<T, W>(a.push, ip, acc);
    }
    if (tail_mine && c == 0 && threadIdx.x == 0) {
      T acc = ld_cg(wvec + (n - 1));
      for (int j = 0; j < m; ++j)
        Num<T>::fma(acc, Num<T>::from_red(s_h + j * NR), static_cast<const T *>(a.p.v[j])[n - 1]);
      acc = Num<T>::scale(acc, f);
      wvec[n - 1] = acc;
      nx += Num<T>::abs2(acc);
      if (final_pass && a.push_early) {
        Pack<T, 1> one;
        one.e[0] = acc;
        pushed |= push_pack<T, 1>(a.push, n - 1, one);
      }
    }
    if (final_pass) {
      const double s = warp_sum(nx);
      if (lane == 0) s_part[warp][0] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < CV_WARPS; ++w) r += s_part[w][0];
        p_nx[c] = r;
      }
    }
    ORTH_TRACE(2);
    if (a.trace && c == 0 && threadIdx.x == 0) a.trace[6] += 1.0;
    if (final_pass) break;
    grid_barrier_with(a.bar, [&]() {});  // pass 2 reads rows other CTAs have just rewritten
    ORTH_TRACE(3);
  }
  if (a.publish_late) {
    __syncthreads();
    __shared__ bool s_last_pub;
    if (threadIdx.x == 0) {
      __threadfence();
      s_last_pub = atomicAdd(a.bar + 2, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (s_last_pub) {
      const bool two = __ldcg(a.scal + a.s_flag) != 0.0;
      for (int t = threadIdx.x; t < a.s_h1 + m * NR - a.s_flag; t += blockDim.x)
        a.host_mb[a.s_flag + t] = __ldcg(a.scal + a.s_flag + t);
      if (two)
        for (int t = threadIdx.x; t < m * NR; t += blockDim.x) a.host_mb[a.s_h2 + t] = __ldcg(a.scal + a.s_h2 + t);
      if (threadIdx.x == 0) a.host_mb[a.err - a.scal] = __ldcg(a.err);
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        a.bar[2] = 0u;
        st_release_sys(a.host_flag, a.host_seq);
      }
    }
  }
  // ---------------- kernel tail: the LAST CTA to finish sums the explicit |v_new|^2 partials (health
  // monitor, read by the next step) and raises the neighbours' halo flags -----------------------
  const bool early = a.push_early && (a.push.nseg > 0 || a.push.nflag > 0);
  if (pushed) __threadfence_system();
  __syncthreads();
  __shared__ bool s_last_orth;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last_orth = atomicAdd(a.push.ticket, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (s_last_orth) {
    __threadfence();
    if (warp == 0) {
      double r = ordered_lane_sum(p_nx, G, lane);
      r = warp_sum(r);
      if (lane == 0) a.scal[a.s_nrm] = r;  // local partial; summed over ranks by the next step
    }
    if (early) {
      __threadfence_system();
      if (threadIdx.x < a.push.nflag) st_release_sys(a.push.flag_dst[threadIdx.x], a.push.seq);
    }
    if (threadIdx.x == 0) *a.push.ticket = 0u;
  }
  if (!a.push_early && (a.push.nseg > 0 || a.push.nflag > 0)) {
    grid_barrier_with(a.bar, [&]() {});
    grid_halo_push<T>(a.push, wvec);
  }
  ORTH_TRACE(7);
}
