"""Row-sharded mode check on real GPUs (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py

Every rank builds the same small Hamiltonians on the host, shards them through the product path
and compares with the CPU oracle / scipy on the full problem: sharded fused SpMV (halo exchange
over NCCL), reductions (batched all-reduce), GCROT / MINRES solves, Gram-Schmidt, and a complete
inexact-Lanczos run against the reference golden result.  Prints PASS/FAIL per rank.
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
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, hamiltonians as hm
    from eigensolvers_b200.lanczos import inexactLanczosDiagonalization
    from oracle.numpy_vector import NumpyVectorOracle as NV
    rt = Runtime.get().init_distributed()
    assert rt.world == world and rt.rank == rank
    ok = True

    def check(name, cond):
        nonlocal ok
        if not cond:
            ok = False
            print(f"[rank {rank}] FAIL {name}", flush=True)

    rng = np.random.default_rng(0)
    for name, H in (("lap", hm.laplacian3d(21)), ("osc", hm.coupled_oscillators((8, 6, 5, 5, 4))[0])):
        n = H.shape[0]
        x = rng.standard_normal(n)
        y = rng.standard_normal(n)
        X, Y = CudaVector(x), CudaVector(y)
        check(f"{name} roundtrip", np.array_equal(X.array, x))
        check(f"{name} dot", abs(X.vdot(Y) - x @ y) <= 1e-11 * np.sqrt(n))
        check(f"{name} norm", abs(X.norm() - np.linalg.norm(x)) <= 1e-12 * np.linalg.norm(x))
        for fmt in ("csr", "sell", "dia"):
            op = DeviceOperator.from_host(H, fmt=fmt)
            check(f"{name} {fmt} halo>0", op.n_halo > 0)
            check(f"{name} {fmt} format", op.format == fmt)
            hx = X.applyOp(op).array
            check(f"{name} {fmt} spmv", np.allclose(hx, H @ x, rtol=1e-12, atol=1e-12))
            # fused Arnoldi step with this format's halo push (gather lists for csr/sell, ranges for dia)
            sg = 0.9 if name == "lap" else 4.6
            oo = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
            wf = CudaVector.solve(op, CudaVector(x, dict(oo)), sg).array
            resf = np.linalg.norm(x - (sg * wf - H @ wf)) / np.linalg.norm(x)
            check(f"{name} {fmt} gcrotmk residual {resf:.2e}", resf < 1e-8)
        # row-block construction equals slicing the full matrix
        r0, r1 = rt.local_range(n)
        op2 = DeviceOperator.from_local_rows(H[r0:r1], n)
        check(f"{name} local rows", np.allclose(X.applyOp(op2).array, H @ x, rtol=1e-12, atol=1e-12))
        # complex vectors (FEAST)
        z = x + 1j * y
        Z = CudaVector(z)
        check(f"{name} complex spmv", np.allclose(Z.applyOp(op).array, H @ z, rtol=1e-12, atol=1e-12))
        # Gram-Schmidt
        qs = [rng.standard_normal(n) for _ in range(4)]
        g = CudaVector.orthogonalize_against_set(X, [CudaVector(q) for q in qs]).array
        gr = NV.orthogonalize_against_set(NV(x.copy()), [NV(q.copy()) for q in qs]).array
        check(f"{name} gs", np.allclose(g, gr, rtol=1e-10, atol=1e-12))
        # solves
        sigma = 0.9 if name == "lap" else 4.6
        for solver in ("gcrotmk", "minres"):
            o = {"linearSystemArgs": {"linearSolver": solver, "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
            w = CudaVector.solve(op, CudaVector(x, dict(o)), sigma).array
            res = np.linalg.norm(x - (sigma * w - H @ w)) / np.linalg.norm(x)
            check(f"{name} {solver} residual {res:.2e}", res < (1e-8 if solver == "gcrotmk" else 1e-5))
        # complex shift (FEAST's solves) on the sharded operator: complex Arnoldi step + complex halo push
        zs = sigma + 0.05j
        o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 3000, "linear_tol": 1e-9, "linear_atol": 0.0}}
        wz = CudaVector.solve(op, CudaVector(x, dict(o)), zs).array
        res = np.linalg.norm(x - (zs * wz - H @ wz)) / np.linalg.norm(x)
        check(f"{name} complex-shift gcrotmk residual {res:.2e}", res < 1e-8)
        S = CudaVector.overlapMatrix([X, Y])
        check(f"{name} overlap", np.allclose(S, np.array([[x @ x, x @ y], [x @ y, y @ y]]), rtol=1e-12))

    # full driver run against the reference golden (tests/golden/osc_1.npz)
    g = np.load(os.path.join(ROOT, "tests", "golden", "osc_1.npz"))
    H, om = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-4, "linear_atol": 1e-4}}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev, vecs, st = inexactLanczosDiagonalization(H, CudaVector(g["y0"].copy(), o), float(g["sigma"]), 8, 20, 1e-10,
                                                     writeOut=False)
    check("lanczos converged", st["isConverged"])
    check("lanczos eigenvalue", abs(ev[0] - g["ev"][0]) <= 1e-10 * abs(g["ev"][0]))
    v = vecs[0].array
    check("lanczos overlap", abs(np.vdot(v, g["vecs"][0])) >= 1 - 1e-8)

    flag = torch.tensor([1.0 if ok else 0.0], device=rt.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"[rank {rank}] {'PASS' if ok else 'FAIL'} (all ranks: {'PASS' if flag.item() == 1.0 else 'FAIL'})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
