"""Row-sharded mode worker (one process per rank) — launched by tests/test_gpu_multirank.py and by
tools/run_*gpu_checks.sh:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multirank_worker.py [case ...]

With at least `world` GPUs every rank takes its own device (NCCL rendezvous, peer-memory transport
over NVLink).  With fewer GPUs all ranks SHARE cuda:0: gloo rendezvous, the peer-memory transport
over same-device CUDA IPC (NCCL refuses duplicate devices) — slower (the ranks time-slice the GPU)
but it exercises exactly the sharded code paths: partition, halo plan, halo push + flags, LL
all-reduce, fused Arnoldi step with cross-rank reduction, sharded Gram-Schmidt.

Every rank builds the same small Hamiltonians on the host, shards them through the product path
and compares with scipy / the CPU oracle / the reference goldens on the full problem.
Cases: kernels, onesided, lanczos, lindep, feast (default set) and zherm (complex-valued H).  Prints PASS/FAIL per rank; exit code 1 on FAIL.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    share = torch.cuda.device_count() < world
    if share:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [a for a in sys.argv[1:] if not a.startswith("-")] or ["kernels", "onesided", "lanczos", "lindep", "feast"]
    from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, hamiltonians as hm
    from eigensolvers_b200.lanczos import inexactLanczosDiagonalization
    from oracle.numpy_vector import NumpyVectorOracle as NV
    rt = Runtime.get()
    ok = True

    def check(name, cond):
        nonlocal ok
        if not cond:
            ok = False
            print(f"[rank {rank}] FAIL {name}", flush=True)

    def run(H, guess, sigma, L, maxit, eConv, cls=None):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = inexactLanczosDiagonalization(H, guess, sigma, L, maxit, eConv, writeOut=False)
        warnings.resetwarnings()
        return out

    # ---- FEAST, nodes over ranks: needs the UNSHARDED runtime, so it runs before init_distributed
    if "feast" in cases:
        from eigensolvers_b200.contour import feastDiagonalization
        g = np.load(os.path.join(ROOT, "tests", "golden", "feast_osc.npz"))
        H, _ = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
        o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 2000, "linear_tol": 1e-2}}
        for mode in ("nodes", "tasks", "dynamic"):
            Y = [CudaVector(np.ascontiguousarray(g["Q"][:, i]), dict(o)) for i in range(4)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ev, vecs, st = feastDiagonalization(H, Y, 16, "legendre", float(g["eMin"]), float(g["eMax"]), 1e-8, 12,
                                                    writeOut=False, distribute=mode)
            warnings.resetwarnings()
            inside_ref = np.sort([e for e in g["ev"] if g["eMin"] < e < g["eMax"]])
            inside = np.sort([e for e in ev if g["eMin"] < e < g["eMax"]])
            check(f"feast {mode}: eigenvalues in the window", len(inside) == len(inside_ref)
                  and np.allclose(inside, inside_ref, rtol=0, atol=5e-6))
            prof = feastDiagonalization.last_profile
            check(f"feast {mode}: profile", prof["world"] == world and len(prof["iterations"]) >= 1
                  and len(prof["iterations"][0]["matvecs_per_node"]) == 8
                  and "reduction_seconds" in prof["iterations"][0])

    rt.init_distributed()
    assert rt.world == world and rt.rank == rank
    check("transport", rt.transport == "peer" if share else rt.transport in ("peer", "nccl"))
    rng = np.random.default_rng(0)

    if "kernels" in cases:
        for name, H in (("lap", hm.laplacian3d(21)), ("osc", hm.coupled_oscillators((8, 6, 5, 5, 4))[0])):
            n = H.shape[0]
            x = rng.standard_normal(n)
            y = rng.standard_normal(n)
            X, Y = CudaVector(x), CudaVector(y)
            check(f"{name} roundtrip", np.array_equal(X.array, x))
            r0, r1 = rt.local_range(n)
            check(f"{name} local_array", np.array_equal(X.local_array, x[r0:r1]))
            check(f"{name} dot", abs(X.vdot(Y) - x @ y) <= 1e-11 * np.sqrt(n))
            check(f"{name} norm", abs(X.norm() - np.linalg.norm(x)) <= 1e-12 * np.linalg.norm(x))
            sigma = 0.9 if name == "lap" else 4.6
            for fmt in ("csr", "sell", "dia"):
                op = DeviceOperator.from_host(H, fmt=fmt)
                check(f"{name} {fmt} halo>0", op.n_halo > 0)
                check(f"{name} {fmt} format", op.format == fmt)
                check(f"{name} {fmt} spmv", np.allclose(X.applyOp(op).array, H @ x, rtol=1e-12, atol=1e-12))
                # fused Arnoldi step with this format's halo push (gather lists for csr/sell, ranges for dia)
                oo = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
                wf = CudaVector.solve(op, CudaVector(x, dict(oo)), sigma).array
                resf = np.linalg.norm(x - (sigma * wf - H @ wf)) / np.linalg.norm(x)
                check(f"{name} {fmt} gcrotmk residual {resf:.2e}", resf < 1e-8)
            if name == "osc":                                    # matrix-free Kronecker form, sharded (band halo as DIA)
                from eigensolvers_b200 import KroneckerSumOperator
                kop = KroneckerSumOperator.coupled_oscillators((8, 6, 5, 5, 4))
                check("osc kron halo>0", kop.n_halo > 0)
                check("osc kron spmv", np.allclose(X.applyOp(kop).array, H @ x, rtol=1e-12, atol=1e-12))
                wk = CudaVector.solve(kop, CudaVector(x, dict(oo)), sigma).array
                resk = np.linalg.norm(x - (sigma * wk - H @ wk)) / np.linalg.norm(x)
                check(f"osc kron gcrotmk residual {resk:.2e}", resk < 1e-8)
                # Jacobi-preconditioned solve on the sharded operators (local diagonal slice, SpMV input is the
                # preconditioned vector, so the halo is pushed by the SpMV itself instead of the Arnoldi step)
                for pop, pname in ((op, "dia"), (kop, "kron")):
                    po = {"linearSystemArgs": dict(oo["linearSystemArgs"], preconditioner="jacobi")}
                    mv0 = rt.stats["matvecs"]
                    wp = CudaVector.solve(pop, CudaVector(x, po), sigma).array
                    resp = np.linalg.norm(x - (sigma * wp - H @ wp)) / np.linalg.norm(x)
                    check(f"osc {pname} jacobi gcrotmk residual {resp:.2e}", resp < 1e-8)
                    check(f"osc {pname} jacobi uses fewer applications", rt.stats["matvecs"] - mv0 < 200)
            op2 = DeviceOperator.from_local_rows(H[r0:r1], n)   # row-block construction == slicing the full matrix
            check(f"{name} local rows", np.allclose(X.applyOp(op2).array, H @ x, rtol=1e-12, atol=1e-12))
            z = x + 1j * y                                       # complex vectors (FEAST)
            check(f"{name} complex spmv", np.allclose(CudaVector(z).applyOp(op).array, H @ z, rtol=1e-12, atol=1e-12))
            qs = [rng.standard_normal(n) for _ in range(4)]
            gq = CudaVector.orthogonalize_against_set(X, [CudaVector(q) for q in qs]).array
            gr = NV.orthogonalize_against_set(NV(x.copy()), [NV(q.copy()) for q in qs]).array
            check(f"{name} gs", np.allclose(gq, gr, rtol=1e-10, atol=1e-12))
            o = {"linearSystemArgs": {"linearSolver": "minres", "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
            wm = CudaVector.solve(op, CudaVector(x, dict(o)), sigma).array
            res = np.linalg.norm(x - (sigma * wm - H @ wm)) / np.linalg.norm(x)
            check(f"{name} minres residual {res:.2e}", res < 1e-5)
            zs = sigma + 0.05j                                   # complex shift: complex Arnoldi step + complex halo push
            o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
            wz = CudaVector.solve(op, CudaVector(x, dict(o)), zs).array
            res = np.linalg.norm(x - (zs * wz - H @ wz)) / np.linalg.norm(x)
            check(f"{name} complex-shift gcrotmk residual {res:.2e}", res < 1e-8)
            S = CudaVector.overlapMatrix([X, Y])
            check(f"{name} overlap", np.allclose(S, np.array([[x @ x, x @ y], [x @ y, y @ y]]), rtol=1e-12))

    if "zherm" in cases:
        # complex Hermitian H, row-sharded: second CSR value stream over the general halo plan
        import scipy.sparse as sp
        n = 3000
        zr = np.random.default_rng(11)                          # same matrix on every rank
        k = int(0.006 * n * n)
        Z = sp.coo_matrix((zr.random(k) + 1j * zr.random(k), (zr.integers(0, n, k), zr.integers(0, n, k))), shape=(n, n)).tocsr()
        d = np.arange(1, n + 1, dtype=np.float64)
        d[n // 2:] += 200.0
        Hz = ((Z + Z.conj().T) * 0.5 + sp.diags(d)).tocsr()
        zop = DeviceOperator.from_host(Hz)
        check("zherm format/dtype", zop.format == "csr" and zop.dtype == np.complex128 and zop.n_halo > 0)
        xr = rng.standard_normal(n)
        xz = xr + 1j * rng.standard_normal(n)
        for tag, xx in (("real", xr), ("complex", xz)):
            got = CudaVector(xx).applyOp(zop).array
            check(f"zherm spmv on a {tag} vector", np.allclose(got, Hz @ xx, rtol=1e-12, atol=1e-12))
        zs = n // 2 + 100.5              # mid-gap: a few hundred applications
        oz = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
        wz = CudaVector.solve(zop, CudaVector(xr, dict(oz)), zs).array
        res = np.linalg.norm(xr - (zs * wz - Hz @ wz)) / np.linalg.norm(xr)
        check(f"zherm gcrotmk residual {res:.2e}", res < 1e-8)
        vs = [CudaVector(rng.standard_normal(n) + 1j * rng.standard_normal(n)) for _ in range(3)]
        V = np.stack([v.array for v in vs], axis=1)
        M = CudaVector.matrixRepresentation(zop, vs)
        refM = V.conj().T @ (Hz @ V)
        check("zherm matrixRepresentation", np.allclose(M, refM, rtol=1e-11, atol=1e-11 * np.abs(refM).max()))

    if "onesided" in cases:
        # structurally one-sided coupling: only the FIRST rank's rows reference columns of other ranks, so
        # the other ranks have n_halo == 0 but non-empty send lists (such a rank used to skip the
        # exchange and its peers blocked).  General sparsity -> CSR / SELL with gather lists.
        import scipy.sparse as sp
        n = 4096
        r1 = n // world
        A = sp.random(n, n, density=2e-3, random_state=5, format="lil")
        A[r1:, :] = 0                                            # rows of ranks >= 1: diagonal only
        A = (A + sp.identity(n) * 2.0).tocsr()
        x = rng.standard_normal(n)
        for fmt in ("csr", "sell"):
            op = DeviceOperator.from_host(A, fmt=fmt)
            mine_has_halo = op.n_halo > 0
            check(f"onesided {fmt}: halo only on rank 0", mine_has_halo == (rank == 0))
            X = CudaVector(x)
            for rep in range(3):                                 # back-to-back exchanges, no reduction between
                check(f"onesided {fmt} spmv #{rep}", np.allclose(X.applyOp(op).array, A @ x, rtol=1e-12, atol=1e-12))
            oo = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 500, "linear_tol": 1e-10, "linear_atol": 0.0}}
            wv = CudaVector.solve(op, CudaVector(x, dict(oo)), 5.0).array
            res = np.linalg.norm(x - (5.0 * wv - A @ wv)) / np.linalg.norm(x)
            check(f"onesided {fmt} gcrotmk residual {res:.2e}", res < 1e-9)

    if "lanczos" in cases:
        # full driver run against the reference golden (tests/golden/osc_1.npz)
        g = np.load(os.path.join(ROOT, "tests", "golden", "osc_1.npz"))
        H, om = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
        o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-4, "linear_atol": 1e-4}}
        ev, vecs, st = run(H, CudaVector(g["y0"].copy(), o), float(g["sigma"]), 8, 20, 1e-10)
        check("lanczos converged", st["isConverged"])
        check("lanczos eigenvalue", abs(ev[0] - g["ev"][0]) <= 1e-10 * abs(g["ev"][0]))
        check("lanczos overlap", abs(np.vdot(vecs[0].array, g["vecs"][0])) >= 1 - 1e-8)

    if "lindep" in cases:
        # BASELINE config 4 at reduced N: near-parallel block start, loose solves, eConv out of reach; after the
        # first restart Gram-Schmidt returns None and the driver aborts with NaN eigenvalues.  The abort must
        # happen at the same (outer, inner, iBlock, cumIter) as in the CPU oracle on the same inputs.
        from eigensolvers_b200.workloads import build_workload, solver_options
        w = build_workload("c4small")
        o = solver_options(w)
        L = 20
        ev_o, Y_o, st_o = run(w["H"], [NV(gv.copy(), dict(o)) for gv in w["guesses"]], w["sigma"], L, 3, 1e-15)
        ev, Y, st = run(w["H"], [CudaVector(gv.copy(), dict(o)) for gv in w["guesses"]], w["sigma"], L, 3, 1e-15)
        check("lindep: oracle aborts with NaN", bool(np.all(np.isnan(ev_o))))
        check("lindep: NaN eigenvalues", bool(np.all(np.isnan(ev))) and len(ev) == len(ev_o))
        where = tuple(int(st[k]) for k in ("outerIter", "innerIter", "iBlock", "cumIter"))
        where_o = tuple(int(st_o[k]) for k in ("outerIter", "innerIter", "iBlock", "cumIter"))
        check(f"lindep: abort position {where} vs oracle {where_o}", where == where_o)
        check("lindep: vectors returned", len(Y) == len(Y_o))

    flag = torch.tensor([1.0 if ok else 0.0])
    if not share:
        flag = flag.to(rt.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    allok = flag.item() == 1.0
    print(f"[rank {rank}] {'PASS' if ok else 'FAIL'} (all ranks: {'PASS' if allok else 'FAIL'}; "
          f"{'shared cuda:0 + gloo' if share else 'one GPU per rank + nccl'}; transport {rt.transport})", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
