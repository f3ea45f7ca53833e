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


def test_kronecker_terms_reproduce_the_assembled_generator():
    """The 1-D factors handed to the matrix-free operator assemble (scipy.sparse.kron) to exactly the
    CSR matrix of BASELINE config 3's generator — host-only check of eigensolvers_b200/kronecker.py."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from eigensolvers_b200 import hamiltonians as hm
    from eigensolvers_b200.kronecker import assemble_csr, oscillator_terms
    for dims in ((6, 5, 5, 4), (4, 3, 5), (7, 2)):
        terms, omega = oscillator_terms(dims, coupling=0.1, seed=1)
        A = assemble_csr(dims, terms)
        B, omega2 = hm.coupled_oscillators(dims, coupling=0.1, seed=1)
        assert np.array_equal(omega, omega2) and A.nnz == B.nnz
        assert abs(A - B).max() < 1e-15


def test_c4_and_c5_workloads():
    """BASELINE configs 4 and 5 at reduced size: two ORTHOGONAL near-parallel guesses built from the two
    product states next to the target level; FEAST window holding three analytic levels, H replicated."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    w = bench.build_workload("c4small")
    g1, g2 = w["guesses"]
    assert w["nBlock"] == 2 and abs(g1 @ g2) < 1e-15 and abs(g1 @ g1 - 1) < 1e-15
    assert np.count_nonzero(g1) == 2 and np.array_equal(np.nonzero(g1)[0], np.nonzero(g2)[0])
    assert w["L"] == 100 and w["tol"] == 1e-1                      # unittests/test_lanczosLINDEP.py:18,28-30
    w5 = bench.build_workload("c5small", rank=1, world=4)
    assert w5["H"].shape == (w5["N"], w5["N"])                     # replicated, not a row block
    inside = [x for x in w5["analytic"] if w5["eMin"] < x < w5["eMax"]]
    assert len(inside) == 3 and len(w5["guesses"]) == w5["m0"] == 6 and w5["nc"] == 16
    Q = np.stack(w5["guesses"], axis=1)
    assert np.allclose(Q.T @ Q, np.eye(6), atol=1e-12)
