"""Launch the two dominant kernels of the C3 workload at full size for `ncu --set full`
(the fused shifted SpMV in DIA storage and the fused Arnoldi-step kernel, N = 2e7):

    ncu --set full --clock-control none --import-source on -k regex:'k_spmv_dia|k_orth_step' \
        --launch-skip 60 --launch-count 4 -o gpurun_out/prof_c3 python tools/prof_c3.py

One GCROT outer cycle (40 Arnoldi steps) of the first shifted solve; the skipped launches are the
short-basis steps, the captured ones have ~30 basis vectors like the average step of a full run."""
import sys
import warnings

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from eigensolvers_b200 import CudaVector, DeviceOperator, Runtime  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    w = bench.build_workload(name)
    rt = Runtime.get()
    op = DeviceOperator.from_host(w["H"])
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1, "linear_tol": 1e-12, "linear_atol": 0.0}}
    b = CudaVector(w["guesses"][0], o)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            CudaVector.solve(op, b, w["sigma"])
        except Exception as e:  # one outer cycle does not converge: the reference raises, so do we
            print("solve:", type(e).__name__)
    warnings.resetwarnings()
    rt.torch.cuda.synchronize()
    print("format", op.format, "launches", rt.launch_count(), "matvecs", rt.last_solve.n_matvec)


if __name__ == "__main__":
    main()
