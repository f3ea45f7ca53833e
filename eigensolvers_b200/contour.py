"""FEAST contour-integral subspace iteration (Polizzi, PRB 79, 115112 (2009)) — host control
flow over the vector plug-in, with the call signature, status dictionary, return values and
quirks of the reference's ``feastDiagonalization`` (feast.py:126-244):

  per FEAST iteration, for each retained quadrature node k and each of the m0 subspace vectors:
  solve (z_k I - A) Qe = Y, accumulate Re[-1/2 w_k r (f cos(theta) + i sin(theta)) Qe] into Q;
  then Rayleigh-Ritz on Q in the Loewdin basis and Y <- Q uSH; stop when the relative change of
  the eigenvalues inside [eMin, eMax] drops below eConv.

`positiveHalf=True` keeps the Legendre nodes with g_k > 0 only, i.e. nc/2 nodes in the upper
right quadrant of the contour — kept as is, it is the parity target.

All (node, vector) solves of one iteration are independent (H replicated on every rank, the
partial contour sums are added over ranks once per iteration):
  distribute="nodes"   node k is owned by rank k % world (BASELINE config 5: one node per GPU);
  distribute="dynamic" the (node, vector) solves are PULLED by the ranks from one shared counter
                       (torch.distributed's store), most expensive first (cost = measured time of the
                       same solve in the previous iteration, 1/Im(z) in the first): no rank idles while
                       work is left, whatever the cost estimate is worth.
  distribute="tasks"   the nc/2 * m0 (node, vector) solves are spread over the ranks by longest-
                       processing-time-first with MEASURED costs: nodes close to the real axis need
                       ~2x the matvecs of the far ones, so whole nodes per GPU leave
                       GPUs idle.  Iteration 0 uses 1/Im(z) as the cost guess, later iterations the
                       wall time each solve took in the previous one (all-gathered, so every rank
                       computes the same assignment).
"""
import math
import time
import warnings

import numpy as np

from .hostmath import (basisTransformation, diagonalizeHamiltonian, eigenvalueResidual,
                       lowdinOrthoMatrix, quadraturePointsWeights)
from .runlog import FeastRunLog


def _getStatus(status, guess):
    """feast.py:16-43."""
    out = {"flagAddition": guess[0].hasExactAddition, "outerIter": 0, "quadrature": 0,
           "isConverged": False, "phase": 1, "residual": None,
           "startTime": time.time(), "runTime": 0.0}
    if status is not None:
        out.update(status)
    return out


def calculateQuadrature(Amat, guess_b, z, radius, angle, weight, contourEllipseFactor):
    """Contribution of one contour point to the filtered vector (feast.py:45-103).

    Exact-addition vectors need one complex solve: Re[mult * (zI-A)^-1 b]; others need the
    solves at z and conj(z) and a fitted sum (Polizzi eq. 12)."""
    b = guess_b
    typeClass = b.__class__
    if abs(z.imag) < 1e-15:  # contour point on the real axis
        opType = "her"
        z = z.real
    else:
        opType = "gen"
    c, s = contourEllipseFactor * math.cos(angle), math.sin(angle)
    if b.hasExactAddition:
        Qe = typeClass.solve(Amat, b, z, opType=opType)
        mult = -0.50 * weight * radius * (c + s * 1j)
        return typeClass.real(mult * Qe)
    mult = -0.25 * weight * radius
    part1 = typeClass.solve(Amat, b, z, opType=opType)
    part2 = typeClass.solve(Amat, b, z.conj(), opType=opType)
    return typeClass.linearCombination([part1, part2], [mult * (c + s * 1j), mult * (c - s * 1j)])


def updateQ(Q, im0, Qquad_k, k):
    """Q[im0] (+)= Qquad_k (feast.py:105-121)."""
    typeClass = Qquad_k.__class__
    if k == 0:
        Q[im0] = Qquad_k
    else:
        Q[im0] = typeClass.linearCombination([Q[im0], Qquad_k], [1.0, 1.0])
    return Q


_CALLS = [0]


def _task_order(nodes, m0, cost):
    """All (node, vector) solves, most expensive first; identical on every rank."""
    tasks = [(k, i) for k in range(len(nodes)) for i in range(m0)]
    if cost is None or any(t not in cost for t in tasks):   # first iteration / subspace size changed
        cost = {(k, i): 1.0 / max(abs(nodes[k][1].imag), 1e-3 * abs(nodes[0][1].imag) + 1e-300) for (k, i) in tasks}
    return sorted(tasks, key=lambda t: (-cost[t], t))


def _assign_tasks(distribute, rank, world, nodes, m0, cost):
    """The (node, vector) solves this rank performs.  Deterministic: every rank computes the same map."""
    tasks = [(k, i) for k in range(len(nodes)) for i in range(m0)]
    if world == 1:
        return set(tasks)
    if distribute == "nodes":
        return {(k, i) for (k, i) in tasks if k % world == rank}
    if cost is None or any(t not in cost for t in tasks):   # first iteration / subspace size changed
        cost = {(k, i): 1.0 / max(abs(nodes[k][1].imag), 1e-3 * abs(nodes[0][1].imag) + 1e-300) for (k, i) in tasks}
    load = [0.0] * world
    mine = set()
    for t in sorted(tasks, key=lambda t: (-cost[t], t)):    # longest processing time first
        r = min(range(world), key=lambda q: (load[q], q))
        load[r] += cost[t]
        if r == rank:
            mine.add(t)
    return mine


def feastDiagonalization(A, Y, nc, quad, eMin, eMax, eConv, maxit, contourEllipseFactor=1.0,
                         writeOut=True, eShift=0.0, convertUnit="au", outFileName=None,
                         summaryFileName=None, distribute=None, lockstep=True):
    """Eigenpairs of the Hermitian A inside [eMin, eMax] from the m0 = len(Y) guess vectors.
    Returns (ev, Y, status) with ALL m0 (or fewer, after a rank drop) Ritz pairs."""
    typeClass = type(Y[0])
    N_SUBSPACE = len(Y)
    assert eMax > eMin
    eRadius = (eMax - eMin) * 0.5
    gk, wk = quadraturePointsWeights(nc, quad, positiveHalf=True)

    status = _getStatus(None, Y)  # the reference ignores a user status here (feast.py:178)
    printObj = FeastRunLog(Y, nc, quad, eMin, eMax, eConv, maxit, writeOut, eShift, convertUnit,
                           status, outFileName, summaryFileName)
    printObj.fileHeader()
    rank, world = 0, 1
    if distribute not in (None, "nodes", "tasks", "dynamic"):
        raise ValueError(f"distribute={distribute!r}: expected None, 'nodes', 'tasks' or 'dynamic'")
    _CALLS[0] += 1        # every rank makes the same sequence of calls: a common name for this call's counters
    if distribute in ("nodes", "tasks", "dynamic"):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
    task_cost = None  # (node, vector) -> cost of that solve, identical on all ranks
    profile = {"distribute": distribute, "world": world, "iterations": []}
    feastDiagonalization.last_profile = profile

    ev = None
    ref_ev = None
    for it in range(maxit):
        status["outerIter"] = it
        Q = [None] * N_SUBSPACE
        nodes = []
        for k in range(len(gk)):
            theta = -(np.pi * 0.5) * (gk[k] - 1)  # Polizzi (13)
            z = (eMin + eMax) * 0.5 + eRadius * (math.cos(theta) + contourEllipseFactor * 1.0j * math.sin(theta))
            nodes.append((theta, z))
        spent, work = {}, {}
        counter = getattr(typeClass, "matvecCount", None)

        def do_task(k, im0, Q):
            theta, z = nodes[k]
            status["quadrature"] = k
            t_task = time.perf_counter()
            c0 = counter() if counter else 0
            Qk = calculateQuadrature(A, Y[im0], z, eRadius, theta, wk[k], contourEllipseFactor)
            Q = updateQ(Q, im0, Qk, 0 if Q[im0] is None else 1)
            spent[(k, im0)] = time.perf_counter() - t_task
            work[(k, im0)] = (counter() - c0) if counter else 0
            return Q

        block_solve = getattr(typeClass, "solveBlock", None) if (lockstep and Y[0].hasExactAddition) else None

        def do_group(grp, Q):
            """Several (node, vector) solves advanced in lock step (each carries its own complex shift z_k)."""
            zs = [nodes[k][1].real if abs(nodes[k][1].imag) < 1e-15 else nodes[k][1] for k, _ in grp]
            t_task = time.perf_counter()
            c0 = counter() if counter else 0
            Qe = block_solve(A, [Y[im0] for _, im0 in grp], zs)
            dt, dc = time.perf_counter() - t_task, (counter() - c0) if counter else 0
            per = getattr(typeClass, "lastBlockMatvecs", lambda: [])()
            tot = float(sum(per)) if len(per) == len(grp) and sum(per) > 0 else 0.0
            for j, (k, im0) in enumerate(grp):
                theta = nodes[k][0]
                mult = -0.50 * wk[k] * eRadius * (contourEllipseFactor * math.cos(theta) + math.sin(theta) * 1j)
                Q = updateQ(Q, im0, typeClass.real(mult * Qe[j]), 0 if Q[im0] is None else 1)   # feast.py:90-92
                share = per[j] / tot if tot else 1.0 / len(grp)
                spent[(k, im0)] = dt * share
                work[(k, im0)] = per[j] if tot else dc // len(grp)
            return Q

        if distribute == "dynamic" and world > 1:
            import torch.distributed as dist
            order = _task_order(nodes, N_SUBSPACE, task_cost)
            store = dist.distributed_c10d._get_default_store()
            key = f"eigb200_feast_{_CALLS[0]}_{it}"
            chunk = 2 if block_solve is not None else 1   # neighbours in cost order: similar length, good lock-step pairs
            while True:
                idx = store.add(key, chunk) - chunk      # atomic fetch-and-add shared by all ranks
                if idx >= len(order):
                    break
                grp = order[idx:idx + chunk]
                if len(grp) > 1:
                    Q = do_group(grp, Q)
                else:
                    Q = do_task(grp[0][0], grp[0][1], Q)
        else:
            mine = _assign_tasks(distribute if distribute != "dynamic" else None, rank, world, nodes, N_SUBSPACE, task_cost)
            todo = [(k, im0) for k in range(len(nodes)) for im0 in range(N_SUBSPACE) if (k, im0) in mine]
            if block_solve is None or len(todo) < 2:
                for k, im0 in todo:
                    Q = do_task(k, im0, Q)
            else:
                for g0 in range(0, len(todo), 8):   # up to 8 at a time
                    Q = do_group(todo[g0:g0 + 8], Q)
        it_prof = {"solve_seconds_this_rank": sum(spent.values()), "matvecs_this_rank": sum(work.values())}
        if world > 1:
            import torch.distributed as dist
            t_red = time.perf_counter()
            Q = typeClass.sumOverRanks(Q, like=Y)          # ONE bucketed all-reduce of the m0 partial sums
            it_prof["reduction_seconds"] = time.perf_counter() - t_red
            box = [None] * world
            dist.all_gather_object(box, (spent, work))
            if distribute in ("tasks", "dynamic"):
                task_cost = {}
                for b in box:
                    task_cost.update(b[0])
            busy = [sum(b[0].values()) for b in box]
            it_prof["busy_seconds_per_rank"] = busy
            it_prof["idle_fraction"] = 1.0 - (sum(busy) / len(busy)) / max(max(busy), 1e-30)
            per_node = {}
            for b in box:
                for (k, _i), c in b[1].items():
                    per_node[k] = per_node.get(k, 0) + c
            it_prof["matvecs_per_node"] = [per_node.get(k, 0) for k in range(len(nodes))]
        else:
            per_node = {}
            for (k, _i), c in work.items():
                per_node[k] = per_node.get(k, 0) + c
            it_prof["matvecs_per_node"] = [per_node.get(k, 0) for k in range(len(nodes))]
        profile["iterations"].append(it_prof)

        Smat = typeClass.overlapMatrix(Q)
        Hmat = typeClass.matrixRepresentation(A, Q)
        printObj.writeFile("iteration", status)
        printObj.writeFile("overlap", Smat)
        status, uS = lowdinOrthoMatrix(Smat, status)
        ev, uv = diagonalizeHamiltonian(uS, Hmat, printObj)
        uSH = uS @ uv
        Y = basisTransformation(Q, uSH)
        del Q

        if it != 0:
            if len(ref_ev) > len(ev):
                nearest = np.argmin(np.abs(ref_ev[:, None] - ev[None, :]), axis=0)
                ref_ev = ref_ev[nearest]
            elif len(ref_ev) < len(ev):
                raise RuntimeError(f"{ref_ev=} but {ev=}. Enlarged space?")
            residual = eigenvalueResidual(ev, ref_ev, [eMin, eMax])
            status["runTime"] = time.time() - status["startTime"]
            status["residual"] = residual
            printObj.writeFile("summary", ev, residual, status)
            if residual < eConv:
                break
        if N_SUBSPACE != len(Y):
            warnings.warn(f"Alert! Got {N_SUBSPACE - len(Y)} dependent vectors")
        N_SUBSPACE = len(Y)
        ref_ev = ev

    printObj.writeFile("results", ev)
    printObj.fileFooter()
    return ev, Y, status
