"""ORACLE — test infrastructure, NOT product code.

CPU restatement of the reference's association integer programme ``ortools_solve``
(reference: solvers.py:9-138).  **Parity unpinned**: the reference solves it with OR-tools CBC
(``ortools``, version unpinned in requirements.txt:5), which is not installed and cannot be
installed offline, and the reference holds no test or golden vector for it.
What is restated here is the *model* — variables solvers.py:17-30, objective :31-49,
constraints :83-111, output layout :115-138 — handed to HiGHS (``scipy.optimize.milp``) instead
of CBC.  Both are exact MIP solvers, so on instances with a unique optimum they must agree;
``brute_force`` (exhaustive enumeration, tiny cases) pins the restatement itself.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file.
"""
import itertools

import numpy as np
import torch


def _build(det_score, link_score, new_score, end_score, det_split):
    """Variable layout [y_det(L) | y_new(L) | y_end(L) | y_link_0(N0*M0) | ...], objective
    vector c (maximise c.y) and equality rows A y = 0 exactly as solvers.py:83-111 adds them
    (A is returned as a scipy.sparse CSR matrix: 3 + M non-zeros per row)."""
    from scipy.sparse import csr_matrix
    split = [int(s) for s in det_split]
    L = int(det_score.shape[0])
    assert sum(split) == L
    det = np.asarray(det_score, dtype=np.float64).reshape(-1)
    new = np.asarray(new_score, dtype=np.float64).reshape(-1)
    end = np.asarray(end_score, dtype=np.float64).reshape(-1)
    links = [np.asarray(l, dtype=np.float64).reshape(l.shape[-2], l.shape[-1]) for l in link_score]
    off_link = [3 * L]
    for l in links:
        off_link.append(off_link[-1] + l.size)
    nvar = off_link[-1]
    c = np.concatenate([det, new, end] + [l.reshape(-1) for l in links])
    ri, ci, vals = [], [], []
    nrow = 0

    def add_row(cols, coef):
        nonlocal nrow
        ri.append(np.full(len(cols), nrow))
        ci.append(np.asarray(cols))
        vals.append(np.asarray(coef, dtype=np.float64))
        nrow += 1
    start = 0
    for i in range(len(split) - 1):
        n, m = split[i], split[i + 1]
        assert links[i].shape == (n, m)
        for j in range(n):                      # end + successors = det          (:88-98)
            idx = start + j
            succ = np.arange(off_link[i] + j * m, off_link[i] + (j + 1) * m)
            add_row(np.concatenate([[2 * L + idx, idx], succ]), np.concatenate([[1.0, -1.0], np.ones(m)]))
            if i == 0:                          # first frame: new = det           (:99-101)
                add_row([L + idx, idx], [1.0, -1.0])
        start += n
        for k in range(m):                      # new + predecessors = det        (:103-109)
            idx = start + k
            pred = np.arange(off_link[i] + k, off_link[i] + n * m, m)
            add_row(np.concatenate([[L + idx, idx], pred]), np.concatenate([[1.0, -1.0], np.ones(n)]))
            if i == len(split) - 2:             # last frame: end = det           (:110-111)
                add_row([2 * L + idx, idx], [1.0, -1.0])
    if nrow:
        A = csr_matrix((np.concatenate(vals), (np.concatenate(ri), np.concatenate(ci))), shape=(nrow, nvar))
    else:
        A = csr_matrix((0, nvar))
    return c, A, L, split, off_link


def _unpack(y, L, split, off_link, like):
    """Output layout of solvers.py:115-138: fp32 0/1 tensors, link as list of 1xNxM."""
    y = np.rint(y)
    t = lambda a: torch.as_tensor(a, dtype=like.dtype)
    links = [t(y[off_link[i]:off_link[i + 1]].reshape(1, split[i], split[i + 1]))
             for i in range(len(split) - 1)]
    return t(y[:L]), links, t(y[L:2 * L]), t(y[2 * L:3 * L])


def milp_solve(det_score, link_score, new_score, end_score, det_split, exclude=None):
    """Same signature and outputs as ``ortools_solve`` (gt branch omitted: never used on the
    predict path, SURVEY a-15).  ``exclude``: optional 0/1 solution vector to cut off (used to
    measure the optimality gap to the second-best solution)."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    c, A, L, split, off = _build(det_score, link_score, new_score, end_score, det_split)
    cons = [LinearConstraint(A, 0, 0)] if A.shape[0] else []
    if exclude is not None:
        e = np.asarray(exclude, dtype=np.float64)
        cons.append(LinearConstraint((2 * e - 1)[None, :], -np.inf, e.sum() - 1))
    res = milp(-c, constraints=cons, integrality=np.ones_like(c), bounds=Bounds(0, 1),
               options={"mip_rel_gap": 0.0, "presolve": True})
    assert res.status == 0, res.message
    y = np.rint(res.x)
    return _unpack(y, L, split, off, det_score), float(c @ y), y


def brute_force(det_score, link_score, new_score, end_score, det_split):
    """Exhaustive enumeration of every 0/1 vector satisfying the constraints (2-frame, tiny)."""
    c, A, L, split, off = _build(det_score, link_score, new_score, end_score, det_split)
    n, m = split
    best, best_y, second = -np.inf, None, -np.inf
    # enumerate partial matchings + activity flags; all other variables are then forced
    for match in itertools.product(range(-2, m), repeat=n):      # -2 inactive, -1 active unmatched
        used = [k for k in match if k >= 0]
        if len(used) != len(set(used)):
            continue
        free = [k for k in range(m) if k not in used]
        for act in itertools.product((0, 1), repeat=len(free)):
            y = np.zeros(len(c))
            for j, k in enumerate(match):
                if k == -2:
                    continue
                y[j] = 1
                y[L + j] = 1
                if k == -1:
                    y[2 * L + j] = 1
                else:
                    y[off[0] + j * m + k] = 1
                    y[n + k] = 1
                    y[2 * L + n + k] = 1
            for k, a in zip(free, act):
                if a:
                    y[n + k] = 1
                    y[L + n + k] = 1
                    y[2 * L + n + k] = 1
            assert not A.shape[0] or np.all(A @ y == 0)
            v = float(c @ y)
            if v > best:
                second, best, best_y = best, v, y
            elif v > second:
                second = v
    return _unpack(best_y, L, split, off, det_score), best, best - second


def objective(det_score, link_score, new_score, end_score, assign):
    """Value of a returned assignment under the reference objective (solvers.py:31-49)."""
    a_det, a_link, a_new, a_end = assign
    v = (det_score.double() * a_det.double()).sum() + (new_score.double() * a_new.double()).sum() \
        + (end_score.double() * a_end.double()).sum()
    for l, a in zip(link_score, a_link):
        v = v + (l.double() * a.double()).sum()
    return float(v)


def assignment_solve(det_score, link_score, new_score, end_score, det_split):
    """Second, independent oracle for 2-frame samples: the (N+M) x (M+N) assignment reduction of the
    programme (derivation from solvers.py:83-111) solved by ``scipy.optimize.
    linear_sum_assignment`` — rows = previous detections + one private "next det stays unmatched"
    row per next detection, columns = next detections + one private "track ends / inactive" column
    per previous detection.  Returns the same tuple as ``milp_solve`` (assignment, objective)."""
    from scipy.optimize import linear_sum_assignment
    n, m = (int(s) for s in det_split)
    d = np.asarray(det_score, dtype=np.float64).reshape(-1)
    nw = np.asarray(new_score, dtype=np.float64).reshape(-1)
    e = np.asarray(end_score, dtype=np.float64).reshape(-1)
    l = np.asarray(link_score[0], dtype=np.float64).reshape(n, m)
    u = d[:n] + nw[:n] + e[:n]                 # prev det active, track ends here
    v = d[n:] + e[n:] + nw[n:]                 # next det active, track starts here
    NEG = -1e9
    C = np.full((n + m, m + n), NEG)
    C[:n, :m] = (d[:n] + nw[:n])[:, None] + (d[n:] + e[n:])[None, :] + l
    C[np.arange(n), m + np.arange(n)] = np.maximum(u, 0.0)
    C[n + np.arange(m), np.arange(m)] = np.maximum(v, 0.0)
    C[n:, m:] = 0.0
    rr, cc = linear_sum_assignment(C, maximize=True)
    L = n + m
    a_det, a_new, a_end = np.zeros(L), np.zeros(L), np.zeros(L)
    a_link = np.zeros((n, m))
    for r, c in zip(rr, cc):
        if r < n and c < m:
            a_link[r, c] = 1
            a_det[r] = a_new[r] = 1
            a_det[n + c] = a_end[n + c] = 1
        elif r < n and c == m + r and u[r] > 0:
            a_det[r] = a_new[r] = a_end[r] = 1
        elif r >= n and c == r - n and v[c] > 0:
            a_det[n + c] = a_new[n + c] = a_end[n + c] = 1
    t = lambda a: torch.as_tensor(a, dtype=det_score.dtype)
    assign = (t(a_det), [t(a_link[None])], t(a_new), t(a_end))
    return assign, float(C[rr, cc].sum())
