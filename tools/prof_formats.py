"""A few launches of the fused shifted SpMV in SELL-32x2 and matrix-free Kronecker form on the C3
Hamiltonian (N = 2e7) for `ncu --set full`:

    ncu --set full --clock-control none --import-source on -k regex:'k_spmv_sell|k_spmv_kron' \
        -s 4 -c 4 -o gpurun_out/prof_formats python tools/prof_formats.py
"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from eigensolvers_b200 import CudaVector, DeviceOperator, KroneckerSumOperator, Runtime, _lib  # noqa: E402
from eigensolvers_b200.workloads import build_workload  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    w = build_workload(name)
    rt = Runtime.get()
    n = w["N"]
    X = CudaVector(np.random.default_rng(0).standard_normal(n))
    y = rt.empty(n, 0)
    out3 = (C.c_double * 3)()
    for op in (DeviceOperator.from_host(w["H"], fmt="sell"),
               KroneckerSumOperator.coupled_oscillators(w["dims"], coupling=0.1, seed=1)):
        for _ in range(4):
            _lib.check(rt.lib.cv_spmv_dots(rt.ctx, op.handle, 0, 1, w["sigma"], 0.0, X._ptr, y.data_ptr(), out3, rt.stream))
        print(op.format, [out3[i] for i in range(3)])
        del op


if __name__ == "__main__":
    main()
