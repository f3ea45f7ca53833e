"""The drop-in claim: the reference's OWN, UNCHANGED `inexactLanczosDiagonalization`
(inexact_Lanczos.py:229-443) and `feastDiagonalization` (feast.py:126-244), imported from the
installation under baseline/_ref (baseline/install_reference.py; sha256 manifest next to it), run on
`CudaVector` and reproduce the results the same drivers gave with `NumpyVector` (tests/golden,
oracle/ref_harness/make_golden.py): same number of Krylov steps, eigenvalues within
max(eConv, 1e-10 relative), eigenvector overlaps >= 1 - 1e-8.

Nothing of eigensolvers_b200.lanczos / .contour / .hostmath is used here.
"""
import hashlib
import json
import os
import warnings

import numpy as np
import pytest
import scipy.linalg as la

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def summary():
    with open(os.path.join(GOLD, "summary.json")) as fh:
        return json.load(fh)


def opts(solver="gcrotmk", tol=1e-4, it=1000):
    return {"linearSystemArgs": {"linearSolver": solver, "linearIter": it, "linear_tol": tol}}


def _overlap(a, b):
    return abs(np.vdot(a, b)) / (np.linalg.norm(a) * np.linalg.norm(b))


@pytest.fixture(scope="module")
def ref(rt):
    from eigensolvers_b200 import refdrivers
    assert refdrivers.available(), ("the reference is not installed under baseline/_ref — run "
                                    "`python baseline/install_reference.py` in the build container")
    ns = refdrivers.load()
    manifest = json.load(open(os.path.join(ns.path, "MANIFEST.json")))
    for fn, digest in manifest["sha256"].items():   # the installed files are the ones that were hashed
        assert hashlib.sha256(open(os.path.join(ns.path, fn), "rb").read()).hexdigest() == digest, fn
    assert ns.inexact_Lanczos.__file__.startswith(ns.path)
    assert ns.feast.__file__.startswith(ns.path)
    return ns


@pytest.fixture(autouse=True)
def _scratch_cwd(tmp_path, monkeypatch):
    """the reference writes iterations_*.out / summary_*.out / saveTNSs/ into the CWD"""
    monkeypatch.chdir(tmp_path)
    yield
    warnings.resetwarnings()


def test_cudavector_is_an_abstractvector(ref):
    from eigensolvers_b200 import CudaVector
    assert issubclass(CudaVector, ref.abstractVector.AbstractVector)
    v = CudaVector(np.arange(4.0))
    assert isinstance(v, ref.abstractVector.AbstractVector)


def test_c1_reference_driver_with_default_arguments(ref, tmp_path):
    """examples/driver_numpyVector.py:27-43 (BASELINE config 1) with the driver's DEFAULT keyword
    arguments: writeOut=True (printUtils files) and saveTNSsEachIteration=True (the `.ttns` shim)."""
    from eigensolvers_b200 import CudaVector
    g = gold("lanczos_c1")
    with warnings.catch_warnings():
        warnings.simplefilter("default")
        ev, vecs, st = ref.inexactLanczosDiagonalization(g["A"], CudaVector(g["Y0"].copy(), opts()), 30, 6, 4, 1e-8)
    want = summary()["lanczos_c1"]
    assert st["isConverged"] and st["cumIter"] == want["cumIter"] and st["outerIter"] == want["outerIter"]
    assert isinstance(vecs[0], CudaVector) and len(ev) == len(g["ev"])
    i, j = np.argmin(abs(ev - 30)), np.argmin(abs(g["ev"] - 30))
    assert abs(ev[i] - g["ev"][j]) <= 1e-8 * abs(g["ev"][j])
    assert _overlap(vecs[i].array, g["vecs"][j]) >= 1 - 1e-8
    files = sorted(os.listdir(tmp_path))
    assert any(f.startswith("summary") for f in files) and any(f.startswith("iterations") for f in files), files
    text = open([f for f in files if f.startswith("summary")][0]).read()
    assert "endingPoint" in text and f"{want['cumIter']:>4d}" in text
    saved = sorted(os.listdir(tmp_path / "saveTNSs"))
    assert f"tns_{want['cumIter']}_0.h5.npz" in saved
    chk = np.load(tmp_path / "saveTNSs" / f"tns_{want['cumIter']}_0.h5.npz")
    assert chk["array"].shape == (100,) and "eigenvalues" in chk.files


def test_unit_test_lanczos_setup(ref):
    """unittests/test_lanczos.py:14-41 set-up and its assertions (:44-93) on CudaVector."""
    from eigensolvers_b200 import CudaVector
    uf = ref.util_funcs
    g = gold("lanczos_t1")
    A = g["A"]
    ev, vecs, st = ref.inexactLanczosDiagonalization(A, CudaVector(g["Y0"].copy(), opts()), 30, 6, 4, 1e-6,
                                                     pick=uf.get_pick_function_close_to_sigma(30), writeOut=False,
                                                     saveTNSsEachIteration=False)
    assert st["cumIter"] == summary()["lanczos_t1"]["cumIter"]
    np.testing.assert_allclose(np.sort(ev), np.sort(g["ev"]), rtol=1e-6)
    S = CudaVector.overlapMatrix(vecs)
    np.testing.assert_allclose(S, np.eye(S.shape[0]), atol=1e-5)
    Hm = CudaVector.matrixRepresentation(A, vecs)
    S1 = CudaVector.overlapMatrix(vecs[:-1])
    np.testing.assert_allclose(CudaVector.extendOverlapMatrix(vecs, S1), S, atol=1e-9)
    H1 = CudaVector.matrixRepresentation(A, vecs[:-1])
    np.testing.assert_allclose(CudaVector.extendMatrixRepresentation(A, vecs, H1), Hm, atol=1e-9)
    evE, uvE = np.linalg.eigh(A)
    assert abs(uf.find_nearest(ev, 30)[1] - uf.find_nearest(evE, 30)[1]) <= 1e-4
    iE, iT = uf.find_nearest(evE, 30)[0], uf.find_nearest(ev, 30)[0]
    ov = np.vdot(uvE[:, iE], vecs[iT].array)
    np.testing.assert_allclose(abs(ov), 1, rtol=1e-5)
    for j in range(len(g["ev"])):
        i = int(np.argmin(abs(ev - g["ev"][j])))
        assert _overlap(vecs[i].array, g["vecs"][j]) >= 1 - 1e-8


def test_unit_test_block_setup(ref):
    """unittests/test_lanczosBlock.py:13-46 (3-fold degenerate target, three guesses)."""
    from eigensolvers_b200 import CudaVector
    g = gold("lanczos_blk")
    sigma = float(g["sigma"])
    guess = [CudaVector(g["Ys"][:, i].copy(), opts()) for i in range(3)]
    ev, vecs, st = ref.inexactLanczosDiagonalization(g["A"], guess, sigma, 6, 4, 1e-6,
                                                     pick=ref.util_funcs.get_pick_function_close_to_sigma(sigma),
                                                     writeOut=False, saveTNSsEachIteration=False)
    want = summary()["lanczos_blk"]
    assert st["isConverged"] and st["cumIter"] == want["cumIter"] and len(vecs) == len(g["ev"])
    np.testing.assert_allclose(ev[:3], g["exact"][5:8], rtol=1e-6)
    mine = np.vstack([vecs[i].array for i in range(3)]).T
    assert abs(np.abs(la.eigvals(mine.T.conj() @ g["vecs"][:3].T)).sum() - 3) < 1e-6


def test_state_following_setup(ref):
    """unittests/test_stateFollowingHO.py:13-43 with the reference's max-overlap pick (util_funcs.py:308-327)."""
    from eigensolvers_b200 import CudaVector
    g = gold("lanczos_ho")
    o = opts("gcrotmk", 1e-4, 30000)
    pick = ref.util_funcs.get_pick_function_maxOvlp(CudaVector(g["ovlpRef"].copy(), o))
    ev, vecs, st = ref.inexactLanczosDiagonalization(g["H"], CudaVector(g["Y0"].copy(), o), float(g["sigma"]), 16, 200,
                                                     1e-10, pick=pick, writeOut=False, saveTNSsEachIteration=False)
    assert st["isConverged"] and st["cumIter"] == summary()["lanczos_ho"]["cumIter"]
    assert abs(ev[0] - g["ev"][0]) <= 1e-8 * abs(g["ev"][0])
    assert _overlap(vecs[0].array, g["vecs"][0]) >= 1 - 1e-8


def test_lindep_setup(ref):
    """unittests/test_lanczosLINDEP.py:8-34 (n = 1200, rtol 1e-1, L = 100)."""
    from eigensolvers_b200 import CudaVector
    g = gold("lanczos_lindep")
    n = 1200
    np.random.seed(10)
    Q = la.qr(np.random.rand(n, n))[0]
    A = Q.T @ np.diag(np.linspace(1, 400, n)) @ Q
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 500, "linear_tol": 1e-1}}
    ev, vecs, st = ref.inexactLanczosDiagonalization(A, CudaVector(g["Y0"].copy(), o), 390, 100, 1000, 1e-12,
                                                     writeOut=False, saveTNSsEachIteration=False)
    want = summary()["lanczos_lindep"]
    assert st["isConverged"] and abs(st["cumIter"] - want["cumIter"]) <= 3
    i, j = np.argmin(abs(ev - 390)), np.argmin(abs(g["ev"] - 390))
    assert abs(ev[i] - g["ev"][j]) <= 1e-10 * abs(g["ev"][j])
    assert _overlap(vecs[i].array, g["vecs"][j]) >= 1 - 1e-8


def test_sparse_oscillator_setup(ref):
    """C3's generator at N = 600 (golden osc_1): sparse H through the reference driver."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    g = gold("osc_1")
    H, _ = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 1000, "linear_tol": 1e-4, "linear_atol": 1e-4}}
    ev, vecs, st = ref.inexactLanczosDiagonalization(H, CudaVector(g["y0"].copy(), o), float(g["sigma"]), 8, 20, 1e-10,
                                                     writeOut=False, saveTNSsEachIteration=False)
    assert st["isConverged"] and abs(st["cumIter"] - summary()["osc_1"]["cumIter"]) <= 1
    assert abs(ev[0] - g["ev"][0]) <= 1e-10 * abs(g["ev"][0])
    assert _overlap(vecs[0].array, g["vecs"][0]) >= 1 - 1e-8


def test_feast_unit_test_setup(ref):
    """unittests/test_feast.py:14-50 through the reference's feastDiagonalization."""
    from eigensolvers_b200 import CudaVector
    g = gold("feast_t1")
    Y = [CudaVector(g["Y1"][:, i].copy(), opts("gcrotmk", 1e-2)) for i in range(6)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev, vecs, st = ref.feastDiagonalization(g["A"], Y, 8, "legendre", 160.0, 166.0, 1e-10, 20, writeOut=False)
    inside = [e for e in g["exact"] if 160.0 <= e <= 166.0]
    for e in inside:
        mine = ev[np.argmin(abs(ev - e))]
        assert abs(mine - e) <= 1e-4
        assert abs(mine - g["ev"][np.argmin(abs(g["ev"] - e))]) <= 1e-6
    S = CudaVector.overlapMatrix(vecs)
    np.testing.assert_allclose(S, np.eye(S.shape[0]), atol=1e-5)


def test_feast_sparse_oscillator_setup(ref):
    """C5's structure at N = 600 (golden feast_osc: nc = 16 -> 8 retained nodes, complex shifted solves)."""
    from eigensolvers_b200 import CudaVector, hamiltonians as hm
    g = gold("feast_osc")
    H, _ = hm.coupled_oscillators((6, 5, 5, 4), coupling=0.1, seed=1)
    o = {"linearSystemArgs": {"linearSolver": "gcrotmk", "linearIter": 2000, "linear_tol": 1e-2}}
    Y = [CudaVector(np.ascontiguousarray(g["Q"][:, i]), dict(o)) for i in range(4)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev, vecs, st = ref.feastDiagonalization(H, Y, 16, "legendre", float(g["eMin"]), float(g["eMax"]), 1e-8, 12,
                                                writeOut=False)
    inside_ref = np.sort([e for e in g["ev"] if g["eMin"] < e < g["eMax"]])
    inside = np.sort([e for e in ev if g["eMin"] < e < g["eMax"]])
    assert len(inside) == len(inside_ref)
    np.testing.assert_allclose(inside, inside_ref, rtol=0, atol=5e-6)
