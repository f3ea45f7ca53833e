"""FEAST with the quadrature nodes distributed over GPUs (BASELINE config 5 at reduced N):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 tools/feast_nodes_check.py [dims...]

H is replicated (every rank holds the full operator, the runtime is NOT row-sharded), node k of
the nc/2 retained Legendre nodes is solved on rank k % world, and the m0 accumulated vectors are
summed over ranks once per FEAST iteration.  Checks the eigenvalues inside the window against
the analytic levels of the oscillator family and that all ranks hold identical results."""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime, hamiltonians as hm
    from eigensolvers_b200.contour import feastDiagonalization
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    distribute = "tasks" if "--tasks" in sys.argv else "nodes"
    dims = tuple(int(a) for a in args) or (10, 10, 10, 10, 10)
    rt = Runtime.get()                      # unsharded: do NOT call init_distributed()
    H, om = hm.coupled_oscillators(dims, coupling=0.1, seed=1)
    levels = hm.oscillator_levels(om, 0.1, 12, max_quanta=6)
    eMin, eMax = 0.5 * (levels[0] + levels[1]), 0.5 * (levels[2] + levels[3])
    inside = levels[(levels > eMin) & (levels < eMax)]
    Q = np.linalg.qr(np.random.default_rng(7).standard_normal((H.shape[0], 4)))[0]
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 2000, "linear_tol": 1e-2}}
    op = DeviceOperator.from_host(H)
    guess = [CudaVector(np.ascontiguousarray(Q[:, i]), dict(o)) for i in range(4)]
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev, vecs, st = feastDiagonalization(op, guess, 16, "legendre", eMin, eMax, 1e-8, 12, writeOut=False,
                                            distribute=distribute)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    got = np.sort([e for e in ev if eMin < e < eMax])
    ok = len(got) == len(inside) and bool(np.allclose(got, inside, rtol=0, atol=5e-6))
    box = [None] * world
    dist.all_gather_object(box, [float(x) for x in ev])
    ok = ok and all(b == box[0] for b in box)
    print(f"[rank {rank}] {'PASS' if ok else 'FAIL'} distribute={distribute} N={H.shape[0]} world={world} iterations={st['outerIter'] + 1} "
          f"seconds={dt:.2f} solves={rt.stats['solves']} matvecs={rt.stats['matvecs']} inside={got.tolist()} "
          f"analytic={inside.tolist()}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
