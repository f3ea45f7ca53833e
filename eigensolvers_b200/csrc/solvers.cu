// solvers.cu — device-resident restatement of the two iterative solvers NumpyVector.solve
// delegates to (numpyVector.py:147-178):
//
//   * scipy.sparse.linalg.gcrotmk  — flexible GCROT(m,k), Hicken & Zingg 2010
//       (scipy/sparse/linalg/_isolve/_gcrotmk.py:16-183 `_fgmres`, :187-506 `gcrotmk`)
//   * scipy.sparse.linalg.minres   — Paige & Saunders 1975 (…/_isolve/minres.py:13-379)
//
// All length-N work runs in the fused kernels of kernels_vec.cuh / kernels_spmv.cuh; the host
// sees only scalars (one pinned-mailbox read per inner iteration) and keeps the tiny
// Hessenberg QR / Givens recurrences, exactly the split the reference has between SciPy's
// BLAS-1 calls and its Python control flow.
//
// Differences from SciPy that are deliberate:
//   - GCROT orthogonalises the new Arnoldi vector against [C, V] with classical Gram-Schmidt
//     applied twice (two tall-skinny passes) instead of (nc+j) sequential dot/axpy pairs; in
//     exact arithmetic both give the same Hessenberg column, and CGS2 keeps orthogonality to
//     machine precision.
//   - unpreconditioned: z_j == v_j, so Z is not stored.
#include <complex>
#include <vector>
#include <limits>
#include <unordered_map>
#include <algorithm>
#include "internal.h"

typedef std::complex<double> zc;

namespace {

// scalar slots (doubles) inside ctx->scalars / ctx->mailbox.  The slots one Arnoldi step reads back
// are contiguous — S_NRM | S_W | S_H1 — so that one all-reduce and one device-to-host copy move
// them together.
constexpr int S_CX = CV_S_SOLVER;          // |ux|^2, |cx|^2
constexpr int S_GAMMA = CV_S_SOLVER + 2;   // <cx|r> (NRED)
constexpr int S_BETA = CV_S_SOLVER + 4;    // |r|^2
constexpr int S_FLAG = CV_S_SOLVER + 6;    // fused Arnoldi step: 1.0 if the second Gram-Schmidt pass ran
constexpr int S_NRM = CV_S_SOLVER + 7;     // squared norm of the orthogonalised vector
constexpr int S_W = CV_S_SOLVER + 8;       // {Re<x|y>, Im<x|y>, <y|y>} of the fused SpMV
constexpr int S_H1 = CV_S_SOLVER + 11;     // first-pass projection coefficients
constexpr int S_H2 = S_H1 + 2 * CV_MAX_PTRS;
constexpr int S_LAG = S_H2 + 2 * CV_MAX_PTRS + 2;  // explicit |v_j|^2 of the previous step (health monitor)
constexpr int S_END = S_LAG + 1;
static_assert(S_END <= (int)CV_N_SCALARS, "solver scalar slots exceed the mailbox");

inline size_t vec_stride_bytes(int64_t n, int cplx_) {
  size_t b = (size_t)n * (cplx_ ? 16 : 8);
  return (b + 255) & ~(size_t)255;
}

struct Workspace {
  char *base;
  size_t stride;
  void *vec(int i) const { return base + stride * (size_t)i; }
};

// y = a*x with a real host scalar (used for v0 = r/beta)
int scal_real(cv_ctx *ctx, int64_t n, int cplx_, double a, const void *x, void *y, cudaStream_t st) {
  return cv_scal(ctx, n, cplx_, cplx_, a, 0.0, x, y, (void *)st);
}

// y = d (.) x (diagonal preconditioner application)
int diag_mul(cv_ctx *ctx, int64_t n, int cplx_, const void *d, const void *x, void *y, cudaStream_t st) {
  cv_prof_scope prof(ctx, 3, st, 3.0 * (double)n * (cplx_ ? 16.0 : 8.0));
  if (cplx_) {
    auto kf = k_diag_mul<cplx>;
    kf<<<cv_occ_grid(ctx, (const void *)kf, n, CV_BLOCK), CV_BLOCK, 0, st>>>(n, (const cplx *)d, (const cplx *)x, (cplx *)y);
  } else {
    auto kf = k_diag_mul<double>;
    kf<<<cv_occ_grid(ctx, (const void *)kf, n, CV_BLOCK), CV_BLOCK, 0, st>>>(n, (const double *)d, (const double *)x, (double *)y);
  }
  return cv_check_launch(ctx, "diag_mul");
}

// Separate-kernel form of one Arnoldi orthogonalisation step (NCCL transport): two tall-skinny
// passes, the SpMV's dots and the projection coefficients in ONE all-reduce, one copy to the host.
int arnoldi_orth_unfused(cv_ctx *ctx, int64_t n, int cplx_, int nc, int nb, std::vector<const void *> &basis,
                         void *w, std::vector<zc> &B, int ldb, int j, std::vector<zc> &hcur,
                         cv_solve_stats *stats, cudaStream_t st) {
  const int NR = cplx_ ? 2 : 1;
  double *mb = ctx->mailbox;
  const void *wp[1] = {w};
      ctx->defer_reduce = true;
      int rc_dot = cv_tsdot_dev(ctx, n, cplx_, 1, nb, basis.data(), 1, wp, S_H1, st);
      ctx->defer_reduce = false;
      CV_TRY(rc_dot);
      CV_TRY(cv_reduce_ranks(ctx, S_W, 3 + nb * NR, st));
      CV_TRY(cv_tsupdate_dev(ctx, n, cplx_, nb, basis.data(), S_H1, w, S_NRM, st));
      CV_TRY(cv_fetch_scalars(ctx, S_NRM, 4 + nb * NR, st));
      stats->n_sync++;
      for (int i = 0; i < nb; ++i) {
        zc h = cplx_ ? zc(mb[S_H1 + 2 * i], mb[S_H1 + 2 * i + 1]) : zc(mb[S_H1 + i], 0.0);
        if (i < nc)
          B[(size_t)i * ldb + j] = h;
        else
          hcur[i - nc] = h;
      }
      // Re-orthogonalise only if the projection cancelled the vector down to less than eta of
      // its norm (Daniel-Gragg-Kaufman-Stewart); otherwise orthogonality is already ~eps/eta.
      if (!(mb[S_NRM] >= ctx->reorth_eta * ctx->reorth_eta * mb[S_W + 2])) {
        CV_TRY(cv_tsdot_dev(ctx, n, cplx_, 1, nb, basis.data(), 1, wp, S_H2, st));
        CV_TRY(cv_tsupdate_dev(ctx, n, cplx_, nb, basis.data(), S_H2, w, S_NRM, st));
        CV_TRY(cv_fetch_scalars(ctx, S_NRM, 1, st));
        CV_TRY(cv_fetch_scalars(ctx, S_H2, nb * NR, st));
        stats->n_sync++;
        stats->n_reorth++;
        for (int i = 0; i < nb; ++i) {
          zc h2 = cplx_ ? zc(mb[S_H2 + 2 * i], mb[S_H2 + 2 * i + 1]) : zc(mb[S_H2 + i], 0.0);
          if (i < nc)
            B[(size_t)i * ldb + j] += h2;
          else
            hcur[i - nc] += h2;
        }
      }
      CV_TRY(cv_scale_dev(ctx, n, cplx_, w, S_NRM, 1, st));
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// GCROT(m,k)
// ------------------------------------------------------------------------------------------
int gcrotmk(cv_ctx *ctx, cv_op *op, int cplx_, int mode, double sre, double sim, const void *b,
            const void *x0, void *x, double rtol, double atol_in, int maxiter, int m, int k,
            const Workspace &ws, cv_solve_stats *stats, cudaStream_t st) {
  const int64_t n = op->n_rows;
  const int NR = cplx_ ? 2 : 1;
  const size_t ebytes = cplx_ ? 16 : 8;
  const double eps = std::numeric_limits<double>::epsilon();
  double *mb = ctx->mailbox;
  // right preconditioner M = diag(dinv) (SciPy's `M=`): z_j = M v_j, w = A z_j, ux = M (V y) - U by
  const void *dinv = ctx->precond_dinv;

  // workspace map: r | V[0..m+k] | C ring (k+1) | U ring (k+1)
  const int nV = m + k + 1;
  void *r = ws.vec(0);
  auto V = [&](int i) { return ws.vec(1 + i); };
  auto Cs = [&](int i) { return ws.vec(1 + nV + i); };
  auto Us = [&](int i) { return ws.vec(1 + nV + (k + 1) + i); };
  std::vector<int> cu_slots;  // ring order: oldest first (scipy `CU` list)
  std::vector<int> free_slots;
  // Recycling (scipy's CU= argument, _gcrotmk.py:227-236, opt-in through cv_ctx_set_recycle): the
  // (c,u) pairs a previous solve left in this workspace are valid for the same operator and shift
  cv_recycle_state &rs = ctx->recycle;
  const bool reuse = rs.enabled && rs.valid && rs.op_id == op->id && rs.cplx == cplx_ && rs.mode == mode && rs.sre == sre &&
                     rs.sim == sim && rs.n == n && rs.work == (const void *)ws.base && rs.m == m && rs.k == k;
  rs.valid = false;
  if (reuse) {
    cu_slots = rs.cu_slots;
    free_slots = rs.free_slots;
  } else {
    for (int i = k; i >= 0; --i) free_slots.push_back(i);
  }

  // x = x0 or 0;  r = b - A x
  if (x0) {
    if (x0 != x) CV_TRY(cv_copy(ctx, n, cplx_, x0, x, (void *)st));
    CV_TRY(cv_spmv_dev(ctx, op, cplx_, mode, sre, sim, x, r, -1.0, 1.0, b, true, -1, st));
    stats->n_matvec++;
  } else {
    CV_CUDA(cudaMemsetAsync(x, 0, (size_t)n * ebytes, st));
    CV_TRY(cv_copy(ctx, n, cplx_, b, r, (void *)st));
  }
  CV_TRY(cv_nrm2sq_dev(ctx, n, cplx_, b, S_BETA + 1, st));
  CV_TRY(cv_nrm2sq_dev(ctx, n, cplx_, r, S_BETA, st));
  CV_TRY(cv_fetch_scalars(ctx, S_BETA, 2, st));
  stats->n_sync++;
  const double b_norm = sqrt(mb[S_BETA + 1]);
  double beta = sqrt(mb[S_BETA]);
  stats->b_norm = b_norm;
  if (!std::isfinite(b_norm)) {
    cv_set_error("gcrotmk: RHS must contain only finite numbers");
    return CV_ERR_ARG;
  }
  const double atol = std::max(atol_in, rtol * b_norm);  // _get_atol_rtol
  if (b_norm == 0.0) {  // _gcrotmk.py:307-309: x = b
    CV_TRY(cv_copy(ctx, n, cplx_, b, x, (void *)st));
    stats->info = 0;
    return CV_OK;
  }

  std::vector<const void *> basis(CV_MAX_PTRS);
  if (!cu_slots.empty()) {
    // (1) Re-orthonormalise the recycled C (scipy does a pivoted QR here, _gcrotmk.py:317-371): the
    // vectors were built with the optimistic single-pass Gram-Schmidt, and small losses would add up
    // from solve to solve.  Cholesky QR is enough because C is orthonormal to ~1e-8 at worst:
    // G = C^H C = R^H R,  C <- C R^-1,  U <- U R^-1 (keeps C = A U).  The V slots are free scratch.
    int nc0 = (int)cu_slots.size();
    for (int c = 0; c < nc0; ++c) basis[c] = Cs(cu_slots[c]);
    std::vector<zc> G((size_t)nc0 * nc0);
    for (int k0 = 0; k0 < nc0; k0 += 4) {
      const int bb = std::min(4, nc0 - k0);
      const void *wp4[4];
      for (int q = 0; q < bb; ++q) wp4[q] = basis[k0 + q];
      CV_TRY(cv_tsdot_dev(ctx, n, cplx_, 1, nc0, basis.data(), bb, wp4, CV_S_TS, st));
      CV_TRY(cv_fetch_scalars(ctx, CV_S_TS, nc0 * bb * NR, st));
      stats->n_sync++;
      for (int i = 0; i < nc0; ++i)
        for (int q = 0; q < bb; ++q)
          G[(size_t)i * nc0 + k0 + q] = cplx_ ? zc(mb[CV_S_TS + (i * bb + q) * 2], mb[CV_S_TS + (i * bb + q) * 2 + 1])
                                              : zc(mb[CV_S_TS + i * bb + q], 0.0);
    }
    double dev = 0.0;
    for (int i = 0; i < nc0; ++i)
      for (int q = 0; q < nc0; ++q) dev = std::max(dev, std::abs(G[(size_t)i * nc0 + q] - zc(i == q ? 1.0 : 0.0)));
    bool ring_ok = std::isfinite(dev);
    if (ring_ok && dev > 1e-12) {
      // upper-triangular R with G = R^H R (Cholesky), then M = R^-1
      std::vector<zc> Rm((size_t)nc0 * nc0, zc(0)), Mi((size_t)nc0 * nc0, zc(0));
      for (int jc = 0; jc < nc0 && ring_ok; ++jc) {
        for (int i = 0; i <= jc; ++i) {
          zc sacc = G[(size_t)i * nc0 + jc];
          for (int t = 0; t < i; ++t) sacc -= std::conj(Rm[(size_t)t * nc0 + i]) * Rm[(size_t)t * nc0 + jc];
          if (i < jc) {
            Rm[(size_t)i * nc0 + jc] = sacc / Rm[(size_t)i * nc0 + i];
          } else {
            if (!(sacc.real() > 0.25)) ring_ok = false;  // far from orthonormal: do not trust the ring
            Rm[(size_t)i * nc0 + i] = std::sqrt(sacc.real());
          }
        }
      }
      if (ring_ok) {
        for (int jc = 0; jc < nc0; ++jc) {  // back-substitution, column by column: R M = I
          Mi[(size_t)jc * nc0 + jc] = 1.0 / Rm[(size_t)jc * nc0 + jc];
          for (int i = jc - 1; i >= 0; --i) {
            zc sacc = 0;
            for (int t = i + 1; t <= jc; ++t) sacc += Rm[(size_t)i * nc0 + t] * Mi[(size_t)t * nc0 + jc];
            Mi[(size_t)i * nc0 + jc] = -sacc / Rm[(size_t)i * nc0 + i];
          }
        }
        const int cs2 = cplx_ ? 2 : 1;
        std::vector<double> cf((size_t)nc0 * nc0 * cs2);
        for (int i = 0; i < nc0; ++i)
          for (int q = 0; q < nc0; ++q) {
            cf[((size_t)i * nc0 + q) * cs2] = Mi[(size_t)i * nc0 + q].real();
            if (cplx_) cf[((size_t)i * nc0 + q) * cs2 + 1] = Mi[(size_t)i * nc0 + q].imag();
          }
        const int chunk = cplx_ ? 4 : 8;
        for (int which = 0; which < 2; ++which) {
          std::vector<const void *> src(nc0);
          std::vector<void *> dst(nc0);
          for (int c = 0; c < nc0; ++c) {
            src[c] = which == 0 ? Cs(cu_slots[c]) : Us(cu_slots[c]);
            dst[c] = V(1 + c);
          }
          for (int c0 = 0; c0 < nc0; c0 += chunk) {
            const int ncol = std::min(chunk, nc0 - c0);
            CV_TRY(cv_lincomb_launch(ctx, n, cplx_, cplx_, nc0, src.data(), ncol, cf.data(), nc0, c0, dst.data() + c0, -1, st));
          }
          for (int c = 0; c < nc0; ++c) CV_TRY(cv_copy(ctx, n, cplx_, dst[c], const_cast<void *>(src[c]), (void *)st));
        }
      }
    }
    if (!ring_ok) {  // start from an empty ring
      cu_slots.clear();
      free_slots.clear();
      for (int i = k; i >= 0; --i) free_slots.push_back(i);
      nc0 = 0;
    }
    stats->orth_loss = 0.0;
    if (nc0 > 0) {
      // (2) x += U C^H r,  r -= C C^H r  (_gcrotmk.py:373-388)
      for (int c = 0; c < nc0; ++c) basis[c] = Cs(cu_slots[c]);
      const void *rp[1] = {r};
      CV_TRY(cv_tsdot_dev(ctx, n, cplx_, 1, nc0, basis.data(), 1, rp, S_H1, st));
      CV_TRY(cv_tsupdate_dev(ctx, n, cplx_, nc0, basis.data(), S_H1, r, S_BETA, st));
      CV_TRY(cv_fetch_scalars(ctx, S_H1, nc0 * NR, st));
      stats->n_sync++;
      for (int i = 0; i < nc0 * NR; ++i) mb[S_H2 + i] = -mb[S_H1 + i];
      CV_CUDA(cudaMemcpyAsync(ctx->scalars + S_H2, mb + S_H2, sizeof(double) * nc0 * NR, cudaMemcpyHostToDevice, st));
      for (int c = 0; c < nc0; ++c) basis[c] = Us(cu_slots[c]);
      CV_TRY(cv_tsupdate_dev(ctx, n, cplx_, nc0, basis.data(), S_H2, x, -1, st));
      CV_TRY(cv_fetch_scalars(ctx, S_BETA, 1, st));
      stats->n_sync++;
      beta = sqrt(mb[S_BETA]);
      stats->n_recycled = nc0;
    }
  }

  const int mlmax = m + k;
  std::vector<zc> Q((size_t)(mlmax + 2) * (mlmax + 2)), R((size_t)(mlmax + 2) * (mlmax + 1));
  std::vector<zc> B((size_t)(k + 1) * (mlmax + 1)), y(mlmax + 2), hy(mlmax + 2), by(k + 1), hcur(mlmax + 2);
  const int ldq = mlmax + 2, ldr = mlmax + 1, ldb = mlmax + 1;
  std::vector<double> coef(2 * 2 * CV_MAX_PTRS);

  int j_outer = 0;
  bool converged = false;
  double eta_now = ctx->reorth_eta;  // 0 = never re-orthogonalise (monitor off), >= 1/sqrt(2) = classic
  // tolerated loss of orthogonality before the switch: three orders below the requested
  // accuracy, between 1e-10 and 1e-7
  const double orth_tol = std::min(1e-7, std::max(1e-10, 1e-3 * std::max(rtol, b_norm > 0 ? atol_in / b_norm : 0.0)));
  for (j_outer = 0; j_outer < maxiter; ++j_outer) {
    const double beta_tol = std::max(atol, rtol * b_norm);
    if (beta <= beta_tol && (j_outer > 0 || !cu_slots.empty())) {
      // recompute the residual to avoid rounding error (_gcrotmk.py:384-387)
      CV_TRY(cv_spmv_dev(ctx, op, cplx_, mode, sre, sim, x, r, -1.0, 1.0, b, true, S_W, st));
      stats->n_matvec++;
      CV_TRY(cv_fetch_scalars(ctx, S_W, 3, st));
      stats->n_sync++;
      beta = sqrt(mb[S_W + 2]);
    }
    stats->resid = beta;
    if (beta <= beta_tol) {
      converged = true;
      break;
    }
    const int nc = (int)cu_slots.size();
    const int ml = m + std::max(k - nc, 0);
    const double atol_inner = std::max(atol, rtol * b_norm) / beta;

    // ---- FGMRES (Arnoldi with projection against C), _gcrotmk.py:93-168 ----
    CV_TRY(scal_real(ctx, n, cplx_, 1.0 / beta, r, V(0), st));
    for (int c = 0; c < nc; ++c) basis[c] = Cs(cu_slots[c]);
    std::fill(Q.begin(), Q.end(), zc(0));
    std::fill(R.begin(), R.end(), zc(0));
    std::fill(B.begin(), B.end(), zc(0));
    Q[0] = 1.0;
    int j = 0;
    bool breakdown = false;
    double res = NAN;
    for (j = 0; j < ml; ++j) {
      void *w = V(j + 1);
      ctx->defer_reduce = true;  // reduced together with the projection coefficients below
      const void *zj = V(j);
      if (dinv) {
        CV_TRY(diag_mul(ctx, n, cplx_, dinv, V(j), ctx->precond_z, st));
        zj = ctx->precond_z;
      }
      int rc_mv = cv_spmv_dev(ctx, op, cplx_, mode, sre, sim, zj, w, 1.0, 0.0, nullptr, false, S_W, st);
      ctx->defer_reduce = false;
      CV_TRY(rc_mv);
      stats->n_matvec++;
      basis[nc + j] = V(j);
      const int nb = nc + j + 1;
      // Classical Gram-Schmidt against [C, V], repeated only if the projection cancelled the
      // vector down to less than eta of its norm (Daniel-Gragg-Kaufman-Stewart), normalisation
      // and the halo push of the new basis vector: ONE fused persistent kernel whose scalars land
      // in the mailbox (kernels_orth.cuh).  The SpMV's dots travel in the same all-reduce.
      bool fused = false;
      CV_TRY(cv_orth_step_dev(ctx, op, n, cplx_, nb, basis.data(), w, S_FLAG, S_H2, S_LAG, eta_now, st, &fused));
      if (fused) {
        stats->n_sync++;
        // Health monitor: the previous step normalised v_j with |w'|^2 = |w|^2 - sum|h|^2, which is
        // exact only while [C,V] is orthonormal.  Its explicitly summed norm arrives one step late;
        // a deviation from 1 means the optimistic single-pass threshold is letting orthogonality
        // errors compound (each cancelling step amplifies them), so the rest of this solve uses the
        // classic "twice is enough" threshold 1/sqrt(2) (as PETSc's refine-if-needed).
        if (j > 0) {
          const double dev = fabs(mb[S_LAG] - 1.0);
          if (dev > stats->orth_loss) stats->orth_loss = dev;
          if (dev > orth_tol && eta_now > 0.0 && eta_now < 0.70710678118654752) {
            eta_now = 0.70710678118654752;
            stats->n_safe++;
          }
        }
        for (int i = 0; i < nb; ++i) {
          zc h = cplx_ ? zc(mb[S_H1 + 2 * i], mb[S_H1 + 2 * i + 1]) : zc(mb[S_H1 + i], 0.0);
          if (mb[S_FLAG] != 0.0) h += cplx_ ? zc(mb[S_H2 + 2 * i], mb[S_H2 + 2 * i + 1]) : zc(mb[S_H2 + i], 0.0);
          if (i < nc)
            B[(size_t)i * ldb + j] = h;
          else
            hcur[i - nc] = h;
        }
        if (mb[S_FLAG] != 0.0) {
          stats->n_reorth++;
          cv_prof_add_bytes(ctx, 1, (double)(2 * nb + 2) * (double)n * (cplx_ ? 16.0 : 8.0));  // second pass
        }
      } else {
        CV_TRY(arnoldi_orth_unfused(ctx, n, cplx_, nc, nb, basis, w, B, ldb, j, hcur, stats, st));
      }
      const double w_norm = sqrt(mb[S_W + 2]);
      const double hlast = sqrt(mb[S_NRM]);
      hcur[j + 1] = hlast;
      if (!(hlast > eps * w_norm)) breakdown = true;

      // ---- insert column j into H = Q R (Givens), _gcrotmk.py:149-157 ----
      // u = blockdiag(Q,1)^H hcur
      std::vector<zc> u(j + 2);
      for (int c = 0; c <= j; ++c) {
        zc s = 0;
        for (int i = 0; i <= j; ++i) s += std::conj(Q[(size_t)i * ldq + c]) * hcur[i];
        u[c] = s;
      }
      u[j + 1] = hcur[j + 1];
      for (int i = 0; i <= j + 1; ++i) Q[(size_t)i * ldq + (j + 1)] = 0, Q[(size_t)(j + 1) * ldq + i] = 0;
      Q[(size_t)(j + 1) * ldq + (j + 1)] = 1.0;
      {
        const zc a = u[j], bb = u[j + 1];
        const double na = std::abs(a), nbb = std::abs(bb);
        const double rho = std::hypot(na, nbb);
        double c;
        zc s;
        if (rho == 0.0 || !std::isfinite(rho)) {
          c = 1.0;
          s = 0.0;
        } else if (na == 0.0) {
          c = 0.0;
          s = std::conj(bb) / nbb;
        } else {
          c = na / rho;
          s = (a / na) * std::conj(bb) / rho;
        }
        const zc rjj = c * a + s * bb;
        for (int i = 0; i < j; ++i) R[(size_t)i * ldr + j] = u[i];
        R[(size_t)j * ldr + j] = rjj;
        R[(size_t)(j + 1) * ldr + j] = 0.0;
        for (int i = 0; i <= j + 1; ++i) {
          zc qa = Q[(size_t)i * ldq + j], qb = Q[(size_t)i * ldq + j + 1];
          Q[(size_t)i * ldq + j] = qa * c + qb * std::conj(s);
          Q[(size_t)i * ldq + j + 1] = -qa * s + qb * c;
        }
      }
      res = std::abs(Q[j + 1]);  // |Q[0,-1]|
      if (res < atol_inner || breakdown) break;
    }
    if (j == ml) j = ml - 1;  // python's loop variable after a full sweep
    if (!std::isfinite(R[(size_t)j * ldr + j].real()) || !std::isfinite(R[(size_t)j * ldr + j].imag())) {
      // scipy raises LinAlgError inside _fgmres and gcrotmk breaks out reporting failure
      break;
    }
    const int ncol = j + 1;
    // y = lstsq(R[:ncol,:ncol], conj(Q[0,:ncol])) * beta  — triangular solve, zero pivots give 0
    for (int i = ncol - 1; i >= 0; --i) {
      zc s = std::conj(Q[i]);
      for (int c = i + 1; c < ncol; ++c) s -= R[(size_t)i * ldr + c] * y[c];
      zc d = R[(size_t)i * ldr + i];
      y[i] = (std::abs(d) > 0.0) ? s / d : zc(0);
    }
    for (int i = 0; i < ncol; ++i) y[i] *= beta;
    // by = B y ; hy = Q (R y)
    for (int c = 0; c < nc; ++c) {
      zc s = 0;
      for (int i = 0; i < ncol; ++i) s += B[(size_t)c * ldb + i] * y[i];
      by[c] = s;
    }
    std::vector<zc> ry(ncol + 1);
    for (int i = 0; i <= ncol; ++i) {
      zc s = 0;
      for (int c = 0; c < ncol; ++c) s += R[(size_t)i * ldr + c] * y[c];
      ry[i] = s;
    }
    for (int i = 0; i <= ncol; ++i) {
      zc s = 0;
      for (int c = 0; c <= ncol; ++c) s += Q[(size_t)i * ldq + c] * ry[c];
      hy[i] = s;
    }
    // ux = Z y - U by,  cx = V hy  in ONE pass over [V_0..V_ncol, U_0..U_nc-1] (_gcrotmk.py:430-447)
    CV_REQUIRE(!free_slots.empty(), "gcrotmk: CU ring exhausted");
    const int slot_new = free_slots.back();
    {
      const int mt = ncol + 1 + nc;
      std::vector<const void *> src(mt);
      const int cs = cplx_ ? 2 : 1;
      std::fill(coef.begin(), coef.end(), 0.0);
      for (int i = 0; i <= ncol; ++i) {
        src[i] = V(i);
        zc cu = (i < ncol) ? y[i] : zc(0);
        coef[(i * 2 + 0) * cs] = cu.real();
        coef[(i * 2 + 1) * cs] = hy[i].real();
        if (cplx_) {
          coef[(i * 2 + 0) * cs + 1] = cu.imag();
          coef[(i * 2 + 1) * cs + 1] = hy[i].imag();
        }
      }
      for (int c = 0; c < nc; ++c) {
        src[ncol + 1 + c] = Us(cu_slots[c]);
        coef[((ncol + 1 + c) * 2 + 0) * cs] = -by[c].real();
        if (cplx_) coef[((ncol + 1 + c) * 2 + 0) * cs + 1] = -by[c].imag();
      }
      void *outs[2] = {Us(slot_new), Cs(slot_new)};
      if (!dinv) {
        CV_TRY(cv_lincomb_launch(ctx, n, cplx_, cplx_, mt, src.data(), 2, coef.data(), 2, 0, outs, S_CX, st));
      } else {
        // t = V y and cx = V hy in one pass over V; ux = M t - U by
        void *outs1[2] = {ctx->precond_t, Cs(slot_new)};
        CV_TRY(cv_lincomb_launch(ctx, n, cplx_, cplx_, ncol + 1, src.data(), 2, coef.data(), 2, 0, outs1, S_CX, st));
        CV_TRY(diag_mul(ctx, n, cplx_, dinv, ctx->precond_t, ctx->precond_t, st));
        std::vector<const void *> src2(nc + 1);
        std::vector<double> coef2((size_t)(nc + 1) * cs, 0.0);
        src2[0] = ctx->precond_t;
        coef2[0] = 1.0;
        for (int c = 0; c < nc; ++c) {
          src2[1 + c] = Us(cu_slots[c]);
          coef2[(1 + c) * cs] = -by[c].real();
          if (cplx_) coef2[(1 + c) * cs + 1] = -by[c].imag();
        }
        void *outs2[1] = {Us(slot_new)};
        CV_TRY(cv_lincomb_launch(ctx, n, cplx_, cplx_, nc + 1, src2.data(), 1, coef2.data(), 1, 0, outs2, -1, st));
      }
    }
    CV_TRY(cv_fetch_scalars(ctx, S_CX, 2, st));
    stats->n_sync++;
    const double cx_norm = sqrt(mb[S_CX + 1]);
    const double alpha = 1.0 / cx_norm;
    if (!std::isfinite(alpha)) continue;  // cannot update, skip (_gcrotmk.py:451-456)
    // cx, ux *= alpha; gamma = <cx|r>; r -= gamma cx; x += gamma ux; beta = |r|
    {
      const int W = cplx_ ? 1 : 2;
      int grid = cplx_ ? cv_occ_grid(ctx, (const void *)k_gcrot_update<cplx, 1>, n / W + 1, CV_BLOCK)
                       : cv_occ_grid(ctx, (const void *)k_gcrot_update<double, 2>, n / W + 1, CV_BLOCK);
      cv_prof_scope prof(ctx, 3, st, 11.0 * (double)n * (cplx_ ? 16.0 : 8.0));  // scale_dot 5 + update 6 vectors
      if (cplx_) {
        k_gcrot_scale_dot<cplx, 1><<<grid, CV_BLOCK, 0, st>>>(n, ctx->scalars + S_CX + 1, (cplx *)Cs(slot_new),
                                                             (cplx *)Us(slot_new), (const cplx *)r,
                                                             ctx->partials, ctx->counters, ctx->scalars + S_GAMMA);
      } else {
        k_gcrot_scale_dot<double, 2><<<grid, CV_BLOCK, 0, st>>>(n, ctx->scalars + S_CX + 1, (double *)Cs(slot_new),
                                                               (double *)Us(slot_new), (const double *)r,
                                                               ctx->partials, ctx->counters, ctx->scalars + S_GAMMA);
      }
      CV_TRY(cv_check_launch(ctx, "gcrot_scale_dot"));
      CV_TRY(cv_reduce_ranks(ctx, S_GAMMA, NR, st));
      if (cplx_) {
        k_gcrot_update<cplx, 1><<<grid, CV_BLOCK, 0, st>>>(n, ctx->scalars + S_GAMMA, (const cplx *)Cs(slot_new),
                                                          (const cplx *)Us(slot_new), (cplx *)r, (cplx *)x,
                                                          ctx->partials, ctx->counters, ctx->scalars + S_BETA);
      } else {
        k_gcrot_update<double, 2><<<grid, CV_BLOCK, 0, st>>>(n, ctx->scalars + S_GAMMA, (const double *)Cs(slot_new),
                                                            (const double *)Us(slot_new), (double *)r, (double *)x,
                                                            ctx->partials, ctx->counters, ctx->scalars + S_BETA);
      }
      CV_TRY(cv_check_launch(ctx, "gcrot_update"));
      CV_TRY(cv_reduce_ranks(ctx, S_BETA, 1, st));
    }
    // truncate oldest, append the new pair (_gcrotmk.py:463-499)
    free_slots.pop_back();
    while ((int)cu_slots.size() >= k && !cu_slots.empty()) {
      free_slots.push_back(cu_slots.front());
      cu_slots.erase(cu_slots.begin());
    }
    cu_slots.push_back(slot_new);
    CV_TRY(cv_fetch_scalars(ctx, S_BETA, 1, st));
    stats->n_sync++;
    beta = sqrt(mb[S_BETA]);
  }
  stats->n_outer = converged ? j_outer : std::min(j_outer + 1, maxiter);
  stats->info = converged ? 0 : (j_outer >= maxiter ? maxiter : j_outer + 1);
  if (rs.enabled) {  // leave the ring for the next solve with this operator and shift
    rs.valid = true;
    rs.op_id = op->id;
    rs.cplx = cplx_;
    rs.mode = mode;
    rs.sre = sre;
    rs.sim = sim;
    rs.n = n;
    rs.work = ws.base;
    rs.m = m;
    rs.k = k;
    rs.cu_slots = cu_slots;
    rs.free_slots = free_slots;
  }
  return CV_OK;
}

// ------------------------------------------------------------------------------------------
// MINRES (real symmetric), minres.py:98-379 with shift = 0, M = I, check = False
// ------------------------------------------------------------------------------------------
int minres(cv_ctx *ctx, cv_op *op, int mode, double sigma, const double *b, const double *x0,
           double *x, double rtol, int maxiter, const Workspace &ws, cv_solve_stats *stats,
           cudaStream_t st) {
  const int64_t n = op->n_rows;
  const double eps = std::numeric_limits<double>::epsilon();
  double *mb = ctx->mailbox;
  double *rbuf[3] = {(double *)ws.vec(0), (double *)ws.vec(1), (double *)ws.vec(2)};
  double *wbuf[2] = {(double *)ws.vec(3), (double *)ws.vec(4)};
  const int W = ((uintptr_t)x & 15) ? 1 : 2;
  const int grid = W == 2 ? cv_occ_grid(ctx, (const void *)k_minres_update<2>, n / W + 1, CV_BLOCK)
                          : cv_occ_grid(ctx, (const void *)k_minres_update<1>, n / W + 1, CV_BLOCK);

  double *r1 = rbuf[0], *r2 = rbuf[0], *ynew = rbuf[1], *spare = rbuf[2];
  if (x0) {
    if (x0 != x) CV_TRY(cv_copy(ctx, n, 0, x0, x, (void *)st));
    CV_TRY(cv_spmv_dev(ctx, op, 0, mode, sigma, 0.0, x, r1, -1.0, 1.0, b, true, -1, st));
    stats->n_matvec++;
  } else {
    CV_CUDA(cudaMemsetAsync(x, 0, (size_t)n * 8, st));
    CV_TRY(cv_copy(ctx, n, 0, b, r1, (void *)st));
  }
  CV_TRY(cv_nrm2sq_dev(ctx, n, 0, r1, S_W, st));
  CV_TRY(cv_nrm2sq_dev(ctx, n, 0, b, S_W + 1, st));
  CV_TRY(cv_fetch_scalars(ctx, S_W, 2, st));
  stats->n_sync++;
  double beta1 = mb[S_W];
  const double bnorm = sqrt(mb[S_W + 1]);
  stats->b_norm = bnorm;
  stats->info = 0;
  if (beta1 == 0.0) return CV_OK;  // x is x0 (minres.py:171-172)
  if (bnorm == 0.0) {
    CV_TRY(cv_copy(ctx, n, 0, b, x, (void *)st));
    return CV_OK;
  }
  beta1 = sqrt(beta1);
  CV_CUDA(cudaMemsetAsync(wbuf[0], 0, (size_t)n * 8, st));
  CV_CUDA(cudaMemsetAsync(wbuf[1], 0, (size_t)n * 8, st));
  double *w_older = wbuf[0], *w_old = wbuf[1];  // scipy's (w2, w) before the update

  double oldb = 0, beta = beta1, dbar = 0, epsln = 0, phibar = beta1, rhs1 = beta1, rhs2 = 0;
  double tnorm2 = 0, gmax = 0, gmin = std::numeric_limits<double>::max(), cs = -1, sn = 0;
  double qrnorm = beta1, Anorm = 0, Acond = 0, rnorm = 0, ynorm = 0;
  int istop = 0, itn = 0;
  (void)rhs1; (void)rhs2; (void)qrnorm; (void)Acond; (void)rnorm;

  while (itn < maxiter) {
    itn++;
    const double s = 1.0 / beta;
    // y = A (s r2) - (beta/oldb) r1 ; alfa = v.y = s <r2|y>       (minres.py:212-222)
    CV_TRY(cv_spmv_dev(ctx, op, 0, mode, sigma, 0.0, r2, ynew, s, itn >= 2 ? -(beta / oldb) : 0.0,
                       itn >= 2 ? r1 : nullptr, true, S_W, st));
    stats->n_matvec++;
    CV_TRY(cv_fetch_scalars(ctx, S_W, 1, st));
    stats->n_sync++;
    const double alfa = s * mb[S_W];
    // y -= (alfa/beta) r2 ; beta_new^2 = y.y                         (:223-228)
    if (W == 2)
      k_axpy_norm<double, 2, true><<<grid, CV_BLOCK, 0, st>>>(n, -(alfa / beta), r2, ynew, ctx->partials,
                                                             ctx->counters, ctx->scalars + S_NRM);
    else
      k_axpy_norm<double, 1, true><<<grid, CV_BLOCK, 0, st>>>(n, -(alfa / beta), r2, ynew, ctx->partials,
                                                             ctx->counters, ctx->scalars + S_NRM);
    CV_TRY(cv_check_launch(ctx, "axpy_norm"));
    CV_TRY(cv_reduce_ranks(ctx, S_NRM, 1, st));
    CV_TRY(cv_fetch_scalars(ctx, S_NRM, 1, st));
    stats->n_sync++;
    // rotate: r1 = r2; r2 = y
    double *vsrc = r2;  // v = s * vsrc, needed by the direction update below
    if (itn == 1) {
      r1 = r2;
      r2 = ynew;
      ynew = spare;  // rbuf[2]; rbuf[0] still holds r1
    } else {
      double *dead = r1;
      r1 = r2;
      r2 = ynew;
      ynew = dead;
    }
    oldb = beta;
    const double beta_sq = mb[S_NRM];
    if (beta_sq < 0) {
      cv_set_error("minres: non-symmetric matrix");
      return CV_ERR_NUMERIC;
    }
    beta = sqrt(beta_sq);
    tnorm2 += alfa * alfa + oldb * oldb + beta * beta;
    if (itn == 1 && beta / beta1 <= 10 * eps) istop = -1;

    // plane rotation (:242-257)
    const double oldeps = epsln;
    const double delta = cs * dbar + sn * alfa;
    const double gbar = sn * dbar - cs * alfa;
    epsln = sn * beta;
    dbar = -cs * beta;
    const double root = std::hypot(gbar, dbar);
    double gamma = std::hypot(gbar, beta);
    gamma = std::max(gamma, eps);
    cs = gbar / gamma;
    sn = beta / gamma;
    const double phi = cs * phibar;
    phibar = sn * phibar;

    // w = (v - oldeps w1 - delta w2)/gamma ; x += phi w ; ynorm = |x|   (:259-265, 278)
    const double denom = 1.0 / gamma;
    if (W == 2)
      k_minres_update<2><<<grid, CV_BLOCK, 0, st>>>(n, s, oldeps, delta, denom, phi, vsrc, w_older, w_old, x,
                                                   ctx->partials, ctx->counters, ctx->scalars + S_W);
    else
      k_minres_update<1><<<grid, CV_BLOCK, 0, st>>>(n, s, oldeps, delta, denom, phi, vsrc, w_older, w_old, x,
                                                   ctx->partials, ctx->counters, ctx->scalars + S_W);
    CV_TRY(cv_check_launch(ctx, "minres_update"));
    CV_TRY(cv_reduce_ranks(ctx, S_W, 1, st));
    std::swap(w_older, w_old);  // new w sits in the former w_older buffer
    CV_TRY(cv_fetch_scalars(ctx, S_W, 1, st));
    stats->n_sync++;
    ynorm = sqrt(mb[S_W]);

    gmax = std::max(gmax, gamma);
    gmin = std::min(gmin, gamma);
    const double z = rhs1 / gamma;
    rhs1 = rhs2 - delta * z;
    rhs2 = -epsln * z;

    // norms and stopping tests (:270-328)
    Anorm = sqrt(tnorm2);
    const double epsx = Anorm * ynorm * eps;
    qrnorm = phibar;
    rnorm = qrnorm;
    const double test1 = (ynorm == 0 || Anorm == 0) ? INFINITY : rnorm / (Anorm * ynorm);
    const double test2 = (Anorm == 0) ? INFINITY : root / Anorm;
    Acond = gmax / gmin;
    if (istop == 0) {
      const double t1 = 1 + test1, t2 = 1 + test2;
      if (t2 <= 1) istop = 2;
      if (t1 <= 1) istop = 1;
      if (itn >= maxiter) istop = 6;
      if (Acond >= 0.1 / eps) istop = 4;
      if (epsx >= beta1) istop = 3;
      if (test2 <= rtol) istop = 2;
      if (test1 <= rtol) istop = 1;
    }
    stats->resid = rnorm;
    if (istop != 0) break;
  }
  stats->n_outer = itn;
  stats->info = (istop == 6) ? maxiter : 0;
  return CV_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// LOCK-STEP GCROT(m,k): nrhs independent solves with the same operator advance one Arnoldi step at a
// time TOGETHER (kernels_batch.cuh): one SpMV reading the matrix once for all of them, one fused
// orthogonalisation kernel with one grid barrier, one mailbox message per step.  Every solve keeps
// its own state (the same recurrences, stopping rules and return codes as gcrotmk() above); solves
// that are between two inner cycles do their outer update with the one-problem kernels and rejoin.
// Single GPU, fused step only; no recycling.
// ------------------------------------------------------------------------------------------
constexpr int S_BLOCK = ((S_END - CV_S_SOLVER + 7) / 8) * 8;  // scalar slots of one problem
static_assert(CV_S_SOLVER + CV_MAX_BATCH * 2 * S_BLOCK <= CV_S_TRACE, "lock-step scalar blocks exceed the mailbox");

struct LockstepSolve {
  // problem
  int cplx_ = 0, mode = 0, m = 20, k = 20, maxiter = 0;
  double sre = 0, sim = 0, rtol = 0, atol_in = 0;
  const void *b = nullptr;
  void *x = nullptr;
  Workspace ws{nullptr, 0};
  cv_solve_stats *stats = nullptr;
  int so = 0;  // offset of this problem's scalar block relative to the one-problem slots
  // state
  enum Phase { INNER, DONE } phase = DONE;
  bool converged = false;
  double b_norm = 0, beta = 0, atol = 0, atol_inner = 0, eta_now = 0.1, orth_tol = 1e-10, res = NAN;
  int j_outer = 0, nc = 0, ml = 0, j = 0;
  bool breakdown = false;
  std::vector<int> cu_slots, free_slots;
  std::vector<zc> Q, R, B, y, hy, by, hcur;
  int ldq = 0, ldr = 0, ldb = 0;
  std::vector<const void *> basis;

  void *r() const { return ws.vec(0); }
  void *V(int i) const { return ws.vec(1 + i); }
  void *Cs(int i) const { return ws.vec(1 + (m + k + 1) + i); }
  void *Us(int i) const { return ws.vec(1 + (m + k + 1) + (k + 1) + i); }

  int start(cv_ctx *ctx, cv_op *op, const void *x0, cudaStream_t st) {
    const int64_t n = op->n_rows;
    const size_t ebytes = cplx_ ? 16 : 8;
    double *mb = ctx->mailbox;
    for (int i = k; i >= 0; --i) free_slots.push_back(i);
    if (x0) {
      if (x0 != x) CV_TRY(cv_copy(ctx, n, cplx_, x0, x, (void *)st));
      CV_TRY(cv_spmv_dev(ctx, op, cplx_, mode, sre, sim, x, r(), -1.0, 1.0, b, true, -1, st));
      stats->n_matvec++;
    } else {
      CV_CUDA(cudaMemsetAsync(x, 0, (size_t)n * ebytes, st));
      CV_TRY(cv_copy(ctx, n, cplx_, b, r(), (void *)st));
    }
    CV_TRY(cv_nrm2sq_dev(ctx, n, cplx_, b, S_BETA + so + 1, st));
    CV_TRY(cv_nrm2sq_dev(ctx, n, cplx_, r(), S_BETA + so, st));
    CV_TRY(cv_fetch_scalars(ctx, S_BETA + so, 2, st));
    stats->n_sync++;
    b_norm = sqrt(mb[S_BETA + so + 1]);
    beta = sqrt(mb[S_BETA + so]);
    stats->b_norm = b_norm;
    if (!std::isfinite(b_norm)) {
      cv_set_error("gcrotmk: RHS must contain only finite numbers");
      return CV_ERR_ARG;
    }
    atol = std::max(atol_in, rtol * b_norm);
    if (b_norm == 0.0) {
      CV_TRY(cv_copy(ctx, n, cplx_, b, x, (void *)st));
      converged = true;
      phase = DONE;
      return CV_OK;
    }
    const int mlmax = m + k;
    Q.assign((size_t)(mlmax + 2) * (mlmax + 2), zc(0));
    R.assign((size_t)(mlmax + 2) * (mlmax + 1), zc(0));
    B.assign((size_t)(k + 1) * (mlmax + 1), zc(0));
    y.assign(mlmax + 2, zc(0));
    hy.assign(mlmax + 2, zc(0));
    by.assign(k + 1, zc(0));
    hcur.assign(mlmax + 2, zc(0));
    ldq = mlmax + 2, ldr = mlmax + 1, ldb = mlmax + 1;
    basis.assign(CV_MAX_PTRS, nullptr);
    eta_now = ctx->reorth_eta;
    orth_tol = std::min(1e-7, std::max(1e-10, 1e-3 * std::max(rtol, b_norm > 0 ? atol_in / b_norm : 0.0)));
    j_outer = 0;
    return outer_begin(ctx, op, st);
  }

  // top of the outer loop (_gcrotmk.py:374-393): stopping test, then the start of an inner cycle
  int outer_begin(cv_ctx *ctx, cv_op *op, cudaStream_t st) {
    const int64_t n = op->n_rows;
    double *mb = ctx->mailbox;
    if (j_outer >= maxiter) {
      phase = DONE;
      return CV_OK;
    }
    const double beta_tol = std::max(atol, rtol * b_norm);
    if (beta <= beta_tol && j_outer > 0) {
      CV_TRY(cv_spmv_dev(ctx, op, cplx_, mode, sre, sim, x, r(), -1.0, 1.0, b, true, S_W + so, st));
      stats->n_matvec++;
      CV_TRY(cv_fetch_scalars(ctx, S_W + so, 3, st));
      stats->n_sync++;
      beta = sqrt(mb[S_W + so + 2]);
    }
    stats->resid = beta;
    if (beta <= beta_tol) {
      converged = true;
      phase = DONE;
      return CV_OK;
    }
    nc = (int)cu_slots.size();
    ml = m + std::max(k - nc, 0);
    atol_inner = std::max(atol, rtol * b_norm) / beta;
    CV_TRY(scal_real(ctx, n, cplx_, 1.0 / beta, r(), V(0), st));
    for (int c = 0; c < nc; ++c) basis[c] = Cs(cu_slots[c]);
    std::fill(Q.begin(), Q.end(), zc(0));
    std::fill(R.begin(), R.end(), zc(0));
    std::fill(B.begin(), B.end(), zc(0));
    Q[0] = 1.0;
    j = 0;
    breakdown = false;
    res = NAN;
    phase = INNER;
    return CV_OK;
  }

  // after the batched step: Hessenberg column from the mailbox, Givens update, inner stopping test.
  // Returns true when the inner cycle ended.
  bool step_consume(cv_ctx *ctx) {
    const double eps = std::numeric_limits<double>::epsilon();
    const double *mb = ctx->mailbox;
    const int nb = nc + j + 1;
    stats->n_matvec++;
    if (j > 0) {
      const double dev = fabs(mb[S_LAG + so] - 1.0);
      if (dev > stats->orth_loss) stats->orth_loss = dev;
      if (dev > orth_tol && eta_now > 0.0 && eta_now < 0.70710678118654752) {
        eta_now = 0.70710678118654752;
        stats->n_safe++;
      }
    }
    const bool two = mb[S_FLAG + so] != 0.0;
    for (int i = 0; i < nb; ++i) {
      zc h = cplx_ ? zc(mb[S_H1 + so + 2 * i], mb[S_H1 + so + 2 * i + 1]) : zc(mb[S_H1 + so + i], 0.0);
      if (two) h += cplx_ ? zc(mb[S_H2 + so + 2 * i], mb[S_H2 + so + 2 * i + 1]) : zc(mb[S_H2 + so + i], 0.0);
      if (i < nc)
        B[(size_t)i * ldb + j] = h;
      else
        hcur[i - nc] = h;
    }
    if (two) stats->n_reorth++;
    const double w_norm = sqrt(mb[S_W + so + 2]);
    const double hlast = sqrt(mb[S_NRM + so]);
    hcur[j + 1] = hlast;
    if (!(hlast > eps * w_norm)) breakdown = true;
    std::vector<zc> u(j + 2);
    for (int c = 0; c <= j; ++c) {
      zc sacc = 0;
      for (int i = 0; i <= j; ++i) sacc += std::conj(Q[(size_t)i * ldq + c]) * hcur[i];
      u[c] = sacc;
    }
    u[j + 1] = hcur[j + 1];
    for (int i = 0; i <= j + 1; ++i) Q[(size_t)i * ldq + (j + 1)] = 0, Q[(size_t)(j + 1) * ldq + i] = 0;
    Q[(size_t)(j + 1) * ldq + (j + 1)] = 1.0;
    {
      const zc a = u[j], bb = u[j + 1];
      const double na = std::abs(a), nbb = std::abs(bb);
      const double rho = std::hypot(na, nbb);
      double c;
      zc sg;
      if (rho == 0.0 || !std::isfinite(rho)) {
        c = 1.0;
        sg = 0.0;
      } else if (na == 0.0) {
        c = 0.0;
        sg = std::conj(bb) / nbb;
      } else {
        c = na / rho;
        sg = (a / na) * std::conj(bb) / rho;
      }
      const zc rjj = c * a + sg * bb;
      for (int i = 0; i < j; ++i) R[(size_t)i * ldr + j] = u[i];
      R[(size_t)j * ldr + j] = rjj;
      R[(size_t)(j + 1) * ldr + j] = 0.0;
      for (int i = 0; i <= j + 1; ++i) {
        zc qa = Q[(size_t)i * ldq + j], qb = Q[(size_t)i * ldq + j + 1];
        Q[(size_t)i * ldq + j] = qa * c + qb * std::conj(sg);
        Q[(size_t)i * ldq + j + 1] = -qa * sg + qb * c;
      }
    }
    res = std::abs(Q[j + 1]);
    if (res < atol_inner || breakdown) return true;
    if (j + 1 >= ml) return true;  // python's loop variable stays at ml-1 after a full sweep
    ++j;
    return false;
  }

  // end of an inner cycle: least squares, new (c,u) pair, residual/solution update (_gcrotmk.py:179,430-499)
  int outer_end(cv_ctx *ctx, cv_op *op, cudaStream_t st) {
    const int64_t n = op->n_rows;
    double *mb = ctx->mailbox;
    if (!std::isfinite(R[(size_t)j * ldr + j].real()) || !std::isfinite(R[(size_t)j * ldr + j].imag())) {
      phase = DONE;  // scipy: LinAlgError inside _fgmres, gcrotmk reports failure
      ++j_outer;
      return CV_OK;
    }
    const int ncol = j + 1;
    for (int i = ncol - 1; i >= 0; --i) {
      zc sacc = std::conj(Q[i]);
      for (int c = i + 1; c < ncol; ++c) sacc -= R[(size_t)i * ldr + c] * y[c];
      zc d = R[(size_t)i * ldr + i];
      y[i] = (std::abs(d) > 0.0) ? sacc / d : zc(0);
    }
    for (int i = 0; i < ncol; ++i) y[i] *= beta;
    for (int c = 0; c < nc; ++c) {
      zc sacc = 0;
      for (int i = 0; i < ncol; ++i) sacc += B[(size_t)c * ldb + i] * y[i];
      by[c] = sacc;
    }
    std::vector<zc> ry(ncol + 1);
    for (int i = 0; i <= ncol; ++i) {
      zc sacc = 0;
      for (int c = 0; c < ncol; ++c) sacc += R[(size_t)i * ldr + c] * y[c];
      ry[i] = sacc;
    }
    for (int i = 0; i <= ncol; ++i) {
      zc sacc = 0;
      for (int c = 0; c <= ncol; ++c) sacc += Q[(size_t)i * ldq + c] * ry[c];
      hy[i] = sacc;
    }
    CV_REQUIRE(!free_slots.empty(), "gcrotmk: CU ring exhausted");
    const int slot_new = free_slots.back();
    {
      const int mt = ncol + 1 + nc;
      std::vector<const void *> src(mt);
      const int cs = cplx_ ? 2 : 1;
      std::vector<double> coef((size_t)mt * 2 * cs, 0.0);
      for (int i = 0; i <= ncol; ++i) {
        src[i] = V(i);
        zc cu = (i < ncol) ? y[i] : zc(0);
        coef[(i * 2 + 0) * cs] = cu.real();
        coef[(i * 2 + 1) * cs] = hy[i].real();
        if (cplx_) {
          coef[(i * 2 + 0) * cs + 1] = cu.imag();
          coef[(i * 2 + 1) * cs + 1] = hy[i].imag();
        }
      }
      for (int c = 0; c < nc; ++c) {
        src[ncol + 1 + c] = Us(cu_slots[c]);
        coef[((ncol + 1 + c) * 2 + 0) * cs] = -by[c].real();
        if (cplx_) coef[((ncol + 1 + c) * 2 + 0) * cs + 1] = -by[c].imag();
      }
      void *outs[2] = {Us(slot_new), Cs(slot_new)};
      CV_TRY(cv_lincomb_launch(ctx, n, cplx_, cplx_, mt, src.data(), 2, coef.data(), 2, 0, outs, S_CX + so, st));
    }
    CV_TRY(cv_fetch_scalars(ctx, S_CX + so, 2, st));
    stats->n_sync++;
    const double alpha = 1.0 / sqrt(mb[S_CX + so + 1]);
    ++j_outer;
    if (std::isfinite(alpha)) {
      const int W = cplx_ ? 1 : 2;
      int grid = cplx_ ? cv_occ_grid(ctx, (const void *)k_gcrot_update<cplx, 1>, n / W + 1, CV_BLOCK)
                       : cv_occ_grid(ctx, (const void *)k_gcrot_update<double, 2>, n / W + 1, CV_BLOCK);
      cv_prof_scope prof(ctx, 3, st, 11.0 * (double)n * (cplx_ ? 16.0 : 8.0));
      double *sc = ctx->scalars;
      if (cplx_) {
        k_gcrot_scale_dot<cplx, 1><<<grid, CV_BLOCK, 0, st>>>(n, sc + S_CX + so + 1, (cplx *)Cs(slot_new), (cplx *)Us(slot_new),
                                                             (const cplx *)r(), ctx->partials, ctx->counters, sc + S_GAMMA + so);
        CV_TRY(cv_check_launch(ctx, "gcrot_scale_dot"));
        k_gcrot_update<cplx, 1><<<grid, CV_BLOCK, 0, st>>>(n, sc + S_GAMMA + so, (const cplx *)Cs(slot_new), (const cplx *)Us(slot_new),
                                                          (cplx *)r(), (cplx *)x, ctx->partials, ctx->counters, sc + S_BETA + so);
      } else {
        k_gcrot_scale_dot<double, 2><<<grid, CV_BLOCK, 0, st>>>(n, sc + S_CX + so + 1, (double *)Cs(slot_new), (double *)Us(slot_new),
                                                               (const double *)r(), ctx->partials, ctx->counters, sc + S_GAMMA + so);
        CV_TRY(cv_check_launch(ctx, "gcrot_scale_dot"));
        k_gcrot_update<double, 2><<<grid, CV_BLOCK, 0, st>>>(n, sc + S_GAMMA + so, (const double *)Cs(slot_new),
                                                            (const double *)Us(slot_new), (double *)r(), (double *)x, ctx->partials,
                                                            ctx->counters, sc + S_BETA + so);
      }
      CV_TRY(cv_check_launch(ctx, "gcrot_update"));
      free_slots.pop_back();
      while ((int)cu_slots.size() >= k && !cu_slots.empty()) {
        free_slots.push_back(cu_slots.front());
        cu_slots.erase(cu_slots.begin());
      }
      cu_slots.push_back(slot_new);
      CV_TRY(cv_fetch_scalars(ctx, S_BETA + so, 1, st));
      stats->n_sync++;
      beta = sqrt(mb[S_BETA + so]);
    }
    return outer_begin(ctx, op, st);
  }
};

// batched SpMV of the active problems: one pass over a DIA matrix for all of them, else one launch each
template <typename T>
int lockstep_spmv(cv_ctx *ctx, cv_op *op, int mode, int np, LockstepSolve *const *P, cudaStream_t st) {
  const bool one_pass = op->fmt == CV_FMT_DIA && np >= 2;
  if (!one_pass) {
    for (int q = 0; q < np; ++q)
      CV_TRY(cv_spmv_dev(ctx, op, sizeof(T) == 16, mode, P[q]->sre, P[q]->sim, P[q]->V(P[q]->j), P[q]->V(P[q]->j + 1), 1.0, 0.0,
                         nullptr, false, CV_S_TS + 3 * q, st));
    return CV_OK;
  }
  const double vecb = (double)sizeof(T) * (double)op->n_rows * 2.0 * np;
  cv_prof_add_bytes(ctx, 6, np * (12.0 * (double)op->nnz + (sizeof(T) == 16 ? 36.0 : 20.0) * (double)op->n_rows));
  cv_prof_scope prof(ctx, 0, st, 8.0 * (double)op->n_diag * (double)op->dia_ld + vecb);
#define GO_NB(NB)                                                                                   \
  {                                                                                                 \
    DiaBatchArgs<T, NB> a;                                                                          \
    a.dia_val = op->dia_val;                                                                        \
    a.ld = op->dia_ld;                                                                              \
    a.n_diag = op->n_diag;                                                                          \
    a.n_rows = (int)op->n_rows;                                                                     \
    a.mode = mode;                                                                                  \
    for (int d = 0; d < CV_MAX_DIAG; ++d) a.off[d] = d < op->n_diag ? op->dia_off[d] : 0;           \
    for (int q = 0; q < NB; ++q) {                                                                  \
      a.x[q] = static_cast<const T *>(P[q]->V(P[q]->j));                                            \
      a.y[q] = static_cast<T *>(P[q]->V(P[q]->j + 1));                                              \
      a.sigma[q] = Num<T>::make(P[q]->sre, P[q]->sim);                                              \
    }                                                                                               \
    a.partials = ctx->partials;                                                                     \
    a.counter = ctx->counters;                                                                      \
    a.out = ctx->scalars + CV_S_TS;                                                                 \
    auto kf = k_spmv_dia_nb<T, NB>;                                                                 \
    int wave = cv_occ_grid(ctx, (const void *)kf, (int64_t)1 << 40, CV_BLOCK);                      \
    int64_t need = (op->n_rows + CV_BLOCK - 1) / CV_BLOCK;                                          \
    kf<<<(int)(need < wave ? need : wave), CV_BLOCK, 0, st>>>(a);                                   \
  }
  switch (np) {
    case 2: GO_NB(2); break;
    case 3: GO_NB(3); break;
    default: GO_NB(4); break;
  }
#undef GO_NB
  return cv_check_launch(ctx, "spmv_dia_nb");
}

int lockstep_orth(cv_ctx *ctx, int64_t n, int cplx_, int np, LockstepSolve *const *P, cudaStream_t st) {
  OrthBatchArgs a;
  a.nprob = np;
  a.n = n;
  a.o_flag = S_FLAG;
  a.o_nrm = S_NRM;
  a.o_w = S_W;
  a.o_h1 = S_H1;
  a.o_h2 = S_H2;
  a.o_lag = S_LAG;
  int W = cplx_ ? 1 : 2, mmax = 1;
  for (int q = 0; q < np; ++q) {
    LockstepSolve &s = *P[q];
    const int nb = s.nc + s.j + 1;
    CV_REQUIRE(nb <= CV_BATCH_PTRS, "lock-step solve: %d basis vectors exceed %d", nb, CV_BATCH_PTRS);
    s.basis[s.nc + s.j] = s.V(s.j);
    for (int i = 0; i < nb; ++i) {
      a.prob[q].v[i] = s.basis[i];
      if ((uintptr_t)s.basis[i] & 15) W = 1;
    }
    a.prob[q].w = s.V(s.j + 1);
    a.prob[q].m = nb;
    a.prob[q].sbase = s.so;   // offsets o_* are the absolute one-problem slots: block base = so
    a.prob[q].s_w_src = CV_S_TS + 3 * q;
    a.prob[q].eta2 = s.eta_now * s.eta_now;
    if (nb > mmax) mmax = nb;
    cv_prof_add_bytes(ctx, 1, (double)(2 * nb + 3) * (double)n * (cplx_ ? 16.0 : 8.0));
  }
  const void *kf = cplx_ ? (const void *)k_orth_step_batch<cplx, 1>
                         : (W == 2 ? (const void *)k_orth_step_batch<double, 2> : (const void *)k_orth_step_batch<double, 1>);
  static std::unordered_map<const void *, int> occ_cache;
  auto it = occ_cache.find(kf);
  if (it == occ_cache.end()) {
    int occ = 0;
    CV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf, CV_BLOCK, sizeof(double) * 2 * CV_BATCH_PTRS));
    CV_REQUIRE(occ >= 1, "lock-step orth: kernel does not fit on an SM");
    it = occ_cache.emplace(kf, occ).first;
  }
  const int cap = it->second * ctx->sms;
  const int mi_max = cplx_ ? ORTH_MI<cplx>::value : ORTH_MI<double>::value;
  int64_t need = (n / W + CV_BLOCK - 1) / CV_BLOCK;
  if (need < CV_BATCH_SLABS) need = CV_BATCH_SLABS;
  const int grid = (int)(need < cap ? need : cap);
  a.pstride = 0;  // one shared region: [MI*NR*G slab partials | CV_MAX_BATCH*512 ww | CV_MAX_BATCH*512 nx | counts]
  CV_REQUIRE(grid <= 512, "lock-step orth: grid of %d CTAs exceeds the per-problem partial capacity", grid);
  CV_REQUIRE((size_t)mi_max * 2 * grid + 2 * CV_MAX_BATCH * 512 + 64 <= CV_N_PARTIALS, "lock-step orth: partial sums exceed the scratch area");
  int total_slabs = 0;
  for (int q = 0; q < np; ++q) total_slabs += (P[q]->nc + P[q]->j + 1 + mi_max - 1) / mi_max;
  CV_REQUIRE(grid >= total_slabs && grid >= np, "lock-step orth: grid of %d CTAs smaller than %d slabs", grid, total_slabs);
  a.partials = ctx->partials;
  a.bar = ctx->counters + CV_COUNTER_BAR;
  a.ticket = ctx->counters + CV_COUNTER_PUSH;
  a.scal = ctx->scalars;
  a.slab_mode = ctx->slab_mode;
  a.snake = ctx->snake;
  a.host_mb = ctx->mailbox;
  a.host_flag = ctx->host_flag;
  a.host_seq = ++ctx->host_seq;
  const size_t sh = sizeof(double) * mmax * (cplx_ ? 2 : 1);
  void *params[1] = {(void *)&a};
  {
    cv_prof_scope prof(ctx, 1, st);
    CV_CUDA(cudaLaunchCooperativeKernel(kf, dim3(grid), dim3(CV_BLOCK), params, sh, st));
  }
  CV_TRY(cv_check_launch(ctx, "orth_step_batch"));
  return cv_wait_mailbox(ctx, a.host_seq, st);
}

int gcrotmk_lockstep(cv_ctx *ctx, cv_op *op, int cplx_, int mode, int nrhs, const double *sre, const double *sim,
                     const void *const *b, const void *const *x0, void *const *x, double rtol, double atol, int maxiter, int m,
                     int k, char *work, size_t ws_one, cv_solve_stats *stats, cudaStream_t st) {
  const int64_t n = op->n_rows;
  std::vector<LockstepSolve> S(nrhs);
  for (int q = 0; q < nrhs; ++q) {
    LockstepSolve &s = S[q];
    s.cplx_ = cplx_;
    s.mode = mode;
    s.m = m;
    s.k = k;
    s.maxiter = maxiter;
    s.sre = sre[q];
    s.sim = sim[q];
    s.rtol = rtol;
    s.atol_in = atol;
    s.b = b[q];
    s.x = x[q];
    s.ws = Workspace{work + ws_one * (size_t)q, vec_stride_bytes(n, cplx_)};
    s.stats = &stats[q];
    s.so = S_BLOCK * (q % (2 * CV_MAX_BATCH));
    CV_TRY(s.start(ctx, op, x0 ? x0[q] : nullptr, st));
  }
  std::vector<LockstepSolve *> act;
  for (;;) {
    act.clear();
    for (auto &s : S)
      if (s.phase == LockstepSolve::INNER) act.push_back(&s);
    if (act.empty()) break;
    for (size_t g0 = 0; g0 < act.size(); g0 += CV_MAX_BATCH) {
      const int np = (int)std::min<size_t>(CV_MAX_BATCH, act.size() - g0);
      LockstepSolve *const *P = act.data() + g0;
      // problems of one group must use distinct scalar blocks: so = block of the problem's index mod 8 (nrhs <= 8)
      if (cplx_)
        CV_TRY(lockstep_spmv<cplx>(ctx, op, mode, np, P, st));
      else
        CV_TRY(lockstep_spmv<double>(ctx, op, mode, np, P, st));
      CV_TRY(lockstep_orth(ctx, n, cplx_, np, P, st));
      for (int q = 0; q < np; ++q) {
        P[q]->stats->n_sync++;
        if (P[q]->step_consume(ctx)) CV_TRY(P[q]->outer_end(ctx, op, st));
      }
    }
  }
  for (int q = 0; q < nrhs; ++q) {
    stats[q].n_outer = S[q].converged ? S[q].j_outer : std::min(S[q].j_outer, maxiter);
    stats[q].info = S[q].converged ? 0 : std::max(1, std::min(S[q].j_outer, maxiter));
  }
  return CV_OK;
}

// The fused Arnoldi step on caller-owned vectors (tests, micro-benchmarks): exactly the launch GCROT's
// inner loop issues after every operator application.
extern "C" int cv_arnoldi_step(cv_ctx *ctx, cv_op *op, int64_t n, int cplx_, int m, const void *const *basis,
                               void *w, double ww, double eta, double *out_host, void *stream) {
  CV_REQUIRE(ctx && basis && w && out_host, "cv_arnoldi_step: null argument");
  CV_REQUIRE(m >= 1 && m <= CV_MAX_PTRS, "cv_arnoldi_step: m=%d outside 1..%d", m, CV_MAX_PTRS);
  CV_REQUIRE(ctx->world == 1 || op, "cv_arnoldi_step: the sharded step needs the operator (halo push plan)");
  cudaStream_t st = (cudaStream_t)stream;
  const int NR = cplx_ ? 2 : 1;
  double *mb = ctx->mailbox;
  // the SpMV normally leaves {Re<x|y>, Im<x|y>, <y|y>} (this rank's part) in S_W..S_W+2
  mb[S_W] = mb[S_W + 1] = 0.0;
  mb[S_W + 2] = ww;
  mb[S_NRM] = 0.0;
  CV_CUDA(cudaMemcpyAsync(ctx->scalars + S_NRM, mb + S_NRM, 4 * sizeof(double), cudaMemcpyHostToDevice, st));
  bool fused = false;
  CV_TRY(cv_orth_step_dev(ctx, op, n, cplx_, m, basis, w, S_FLAG, S_H2, S_LAG, eta, st, &fused));
  if (!fused) {
    cv_set_error("cv_arnoldi_step: the fused step is disabled in this configuration (EIGB200_FUSED=0 or NCCL transport)");
    return CV_ERR_UNSUPPORTED;
  }
  CV_CUDA(cudaStreamSynchronize(st));
  out_host[0] = mb[S_FLAG];
  out_host[1] = mb[S_NRM];
  for (int i = 0; i < m * NR; ++i) out_host[2 + i] = mb[S_H1 + i] + (mb[S_FLAG] != 0.0 ? mb[S_H2 + i] : 0.0);
  return CV_OK;
}

extern "C" int cv_solve_precond(cv_ctx *ctx, cv_op *op, int cplx_, int solver, int reverse, double sigma_re,
                                double sigma_im, const void *b, const void *x0, void *x_out, double rtol, double atol,
                                int maxiter, int m, int k, const void *dinv_dev, void *work_dev, size_t work_bytes,
                                cv_solve_stats *stats, void *stream) {
  CV_REQUIRE(ctx && op && dinv_dev && work_dev, "cv_solve_precond: null argument");
  CV_REQUIRE(solver == CV_SOLVER_GCROTMK, "cv_solve_precond: a right preconditioner applies to GCROT only (MINRES needs an SPD one)");
  CV_REQUIRE(((uintptr_t)dinv_dev & 15) == 0, "cv_solve_precond: the diagonal must be 16-byte aligned");
  const int mm = m <= 0 ? 20 : m, kk = k <= 0 ? mm : k;
  const size_t base = cv_solve_workspace_bytes(op->n_rows, cplx_, solver, mm, kk);
  const size_t stride = vec_stride_bytes(op->n_rows, cplx_);
  CV_REQUIRE(work_bytes >= base + 2 * stride, "cv_solve_precond: workspace needs two more vectors than cv_solve_workspace_bytes()");
  ctx->precond_dinv = dinv_dev;
  ctx->precond_z = static_cast<char *>(work_dev) + base;
  ctx->precond_t = static_cast<char *>(work_dev) + base + stride;
  ctx->recycle.valid = false;
  const int rc = cv_solve(ctx, op, cplx_, solver, reverse, sigma_re, sigma_im, b, x0, x_out, rtol, atol, maxiter, m, k, work_dev,
                          base, stats, stream);
  ctx->precond_dinv = nullptr;
  return rc;
}

extern "C" int cv_solve_batch(cv_ctx *ctx, cv_op *op, int cplx_, int nrhs, int reverse, const double *sigma_re,
                              const double *sigma_im, const void *const *b, const void *const *x0, void *const *x_out,
                              double rtol, double atol, int maxiter, int m, int k, void *work_dev, size_t work_bytes,
                              cv_solve_stats *stats, void *stream) {
  CV_REQUIRE(ctx && op && sigma_re && sigma_im && b && x_out && work_dev && stats, "cv_solve_batch: null argument");
  CV_REQUIRE(nrhs >= 1 && nrhs <= 2 * CV_MAX_BATCH, "cv_solve_batch: nrhs=%d outside 1..%d", nrhs, 2 * CV_MAX_BATCH);
  CV_REQUIRE(ctx->world == 1, "cv_solve_batch: lock-step solves run on a single GPU (row-sharded runs solve one at a time)");
  CV_REQUIRE(op->n_rows == op->n_cols - op->n_halo, "cv_solve_batch: operator must be square");
  CV_REQUIRE(maxiter >= 0 && rtol >= 0 && atol >= 0, "cv_solve_batch: bad tolerances");
  if (m <= 0) m = 20;
  if (k <= 0) k = m;
  CV_REQUIRE(m + 2 * k + 2 <= CV_BATCH_PTRS, "cv_solve_batch: GCROT(m=%d,k=%d) exceeds %d basis vectors per problem", m, k,
             CV_BATCH_PTRS);
  const size_t ws_one = cv_solve_workspace_bytes(op->n_rows, cplx_, CV_SOLVER_GCROTMK, m, k);
  CV_REQUIRE(((uintptr_t)work_dev & 255) == 0 && (ws_one & 255) == 0, "cv_solve_batch: workspace must be 256-byte aligned");
  CV_REQUIRE(work_bytes >= ws_one * (size_t)nrhs, "cv_solve_batch: workspace too small");
  for (int q = 0; q < nrhs; ++q) {
    CV_REQUIRE(b[q] && x_out[q], "cv_solve_batch: null vector %d", q);
    CV_REQUIRE(((uintptr_t)b[q] & 15) == 0 && ((uintptr_t)x_out[q] & 15) == 0 && (!x0 || !x0[q] || ((uintptr_t)x0[q] & 15) == 0),
               "cv_solve_batch: vectors must be 16-byte aligned");
    memset(&stats[q], 0, sizeof(stats[q]));
  }
  ctx->recycle.valid = false;
  return gcrotmk_lockstep(ctx, op, cplx_, reverse ? CV_SPMV_RSHIFT : CV_SPMV_SHIFT, nrhs, sigma_re, sigma_im, b, x0, x_out, rtol,
                          atol, maxiter, m, k, static_cast<char *>(work_dev), ws_one, stats, (cudaStream_t)stream);
}

extern "C" size_t cv_solve_workspace_bytes(int64_t n, int cplx_, int solver, int m, int k) {
  size_t stride = vec_stride_bytes(n, cplx_);
  if (solver == CV_SOLVER_MINRES) return stride * 5;
  return stride * (size_t)(1 + (m + k + 1) + 2 * (k + 1));
}

extern "C" int cv_solve(cv_ctx *ctx, cv_op *op, int cplx_, int solver, int reverse, double sigma_re,
                        double sigma_im, const void *b, const void *x0, void *x_out, double rtol,
                        double atol, int maxiter, int m, int k, void *work_dev, size_t work_bytes,
                        cv_solve_stats *stats, void *stream) {
  CV_REQUIRE(ctx && op && b && x_out && work_dev && stats, "cv_solve: null argument");
  CV_REQUIRE(op->n_rows == op->n_cols - op->n_halo, "cv_solve: operator must be square");
  CV_REQUIRE(maxiter >= 0 && rtol >= 0, "cv_solve: bad tolerances");
  CV_REQUIRE(((uintptr_t)work_dev & 255) == 0, "cv_solve: workspace must be 256-byte aligned");
  CV_REQUIRE(((uintptr_t)b & 15) == 0 && ((uintptr_t)x_out & 15) == 0 && (!x0 || ((uintptr_t)x0 & 15) == 0),
             "cv_solve: vectors must be 16-byte aligned");
  memset(stats, 0, sizeof(*stats));
  const int mode = reverse ? CV_SPMV_RSHIFT : CV_SPMV_SHIFT;
  Workspace ws{static_cast<char *>(work_dev), vec_stride_bytes(op->n_rows, cplx_)};
  cudaStream_t st = (cudaStream_t)stream;
  if (solver == CV_SOLVER_GCROTMK) {
    if (m <= 0) m = 20;
    if (k <= 0) k = m;
    CV_REQUIRE(atol >= 0, "cv_solve: gcrotmk called with invalid atol=%g", atol);
    CV_REQUIRE(m + 2 * k + 2 <= CV_MAX_PTRS, "cv_solve: GCROT(m=%d,k=%d) exceeds %d basis vectors", m, k, CV_MAX_PTRS);
    CV_REQUIRE(work_bytes >= cv_solve_workspace_bytes(op->n_rows, cplx_, solver, m, k),
               "cv_solve: workspace too small");
    return gcrotmk(ctx, op, cplx_, mode, sigma_re, sigma_im, b, x0, x_out, rtol, atol, maxiter, m, k, ws,
                   stats, st);
  }
  if (solver == CV_SOLVER_MINRES) {
    if (cplx_ || sigma_im != 0.0) {
      cv_set_error("cv_solve: MINRES is implemented for real symmetric systems only");
      return CV_ERR_UNSUPPORTED;
    }
    CV_REQUIRE(work_bytes >= cv_solve_workspace_bytes(op->n_rows, 0, solver, 0, 0), "cv_solve: workspace too small");
    ctx->recycle.valid = false;  // MINRES overwrites the workspace the ring lives in
    return minres(ctx, op, mode, sigma_re, (const double *)b, (const double *)x0, (double *)x_out, rtol,
                  maxiter, ws, stats, st);
  }
  cv_set_error("cv_solve: unknown solver %d", solver);
  return CV_ERR_ARG;
}

// ------------------------------------------------------------------------------------------
// work-split plans of the fused Arnoldi-step kernels, evaluated on the HOST by the very functions the
// kernels call (orth_slab_map / batch_map_build are __host__ __device__): lets the CPU tests check the
// index logic of the hot kernel exhaustively (every basis vector in exactly one slab, every slab and
// every problem at least one CTA, CTA ranges tiling the grid).  No device call.
// ------------------------------------------------------------------------------------------
extern "C" int cv_orth_slab_plan(int m, int grid, int cplx_, int mode, int *ny_out, int *start_out, int *i0_out) {
  CV_REQUIRE(ny_out && start_out && i0_out, "cv_orth_slab_plan: null argument");
  CV_REQUIRE(m >= 1 && m <= CV_MAX_PTRS, "cv_orth_slab_plan: m=%d outside 1..%d", m, CV_MAX_PTRS);
  const int MI = cplx_ ? ORTH_MI<cplx>::value : ORTH_MI<double>::value;
  CV_REQUIRE(grid >= (m + MI - 1) / MI, "cv_orth_slab_plan: grid %d smaller than the %d slabs (the launcher refuses this too)", grid,
             (m + MI - 1) / MI);
  SlabMap s;
  orth_slab_map(m, grid, MI, mode, s);
  *ny_out = s.ny;
  for (int i = 0; i <= s.ny; ++i) {
    start_out[i] = s.start[i];
    i0_out[i] = s.i0[i];
  }
  return CV_OK;
}

extern "C" int cv_orth_batch_plan(int nprob, const int *m, unsigned active_mask, int grid, int cplx_, int *nslab_out,
                                  int *slab_q, int *slab_i0, int *slab_mi, int *slab_start, int *prob_start) {
  CV_REQUIRE(m && nslab_out && slab_q && slab_i0 && slab_mi && slab_start && prob_start, "cv_orth_batch_plan: null argument");
  CV_REQUIRE(nprob >= 1 && nprob <= CV_MAX_BATCH, "cv_orth_batch_plan: nprob=%d outside 1..%d", nprob, CV_MAX_BATCH);
  const int MI = cplx_ ? ORTH_MI<cplx>::value : ORTH_MI<double>::value;
  OrthBatchArgs a = {};
  a.nprob = nprob;
  int slabs = 0, active = 0;
  for (int q = 0; q < nprob; ++q) {
    CV_REQUIRE(m[q] >= 1 && m[q] <= CV_BATCH_PTRS, "cv_orth_batch_plan: m[%d]=%d outside 1..%d", q, m[q], CV_BATCH_PTRS);
    a.prob[q].m = m[q];
    if ((active_mask >> q) & 1u) {
      slabs += (m[q] + MI - 1) / MI;
      ++active;
    }
  }
  CV_REQUIRE(active >= 1 && grid >= slabs && grid >= active, "cv_orth_batch_plan: grid %d smaller than %d slabs", grid, slabs);
  BatchMap M;
  batch_map_build(a, active_mask, grid, MI, M);
  *nslab_out = M.nslab;
  for (int s = 0; s < M.nslab; ++s) {
    slab_q[s] = M.q[s];
    slab_i0[s] = M.i0[s];
    slab_mi[s] = M.mi[s];
    slab_start[s] = M.start[s];
  }
  slab_start[M.nslab] = M.start[M.nslab];
  for (int q = 0; q <= nprob; ++q) prob_start[q] = M.bstart[q];
  return CV_OK;
}
