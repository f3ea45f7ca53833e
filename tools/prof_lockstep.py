"""Two GCROT outer cycles of C2's four block solves, one by one and in lock step, for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_lockstep.py
"""
import sys
import warnings

sys.path.insert(0, ".")
from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime  # noqa: E402
from eigensolvers_b200.workloads import build_workload, solver_options  # noqa: E402

w = build_workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
rt = Runtime.get()
op = DeviceOperator.from_host(w["H"])
o = solver_options(w)
o["linearSystemArgs"]["linearIter"] = 2
B = [CudaVector(g.copy(), dict(o)) for g in w["guesses"]]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for b in B:
        try:
            CudaVector.solve(op, b, w["sigma"])
        except Exception:
            pass
    try:
        CudaVector.solveBlock(op, B, w["sigma"])
    except Exception:
        pass
rt.torch.cuda.synchronize()
print("done", rt.stats)
