#!/usr/bin/env python
"""CPU experiment (numpy, oracle operators) for SURVEY.md 8 f-4 / round-1 verdict item 10: does warm-starting the Schur
complement's inner CG from the previous outer iteration's inner solution save inner iterations?
Outer: CG on S = Q_xx - Q_xz Q_zz^-1 Q_zx (labelled rows); inner: CG on Q_zz to a relative tolerance.  Warm start: solve
Q_zz d = b - Q_zz y_prev to the tolerance rescaled by |b| / |r0| and return y_prev + d (same accuracy contract).
    python profiles/schur_warmstart_cpu.py [n] [labelled]"""
import json, sys
import numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 400
x = oracle.datasets.torus(n, seed=0)
idx, val = oracle.knn_graph(x, 12)
d2, _ = oracle.knn_search(x, x, 12)
eps = float(d2[:, 11].sqrt().median())
lap = oracle.LaplacianOracle(val.double(), idx, n, eps, "symmetric", True)
Q = oracle.dense_from_matmul(lambda t: oracle.precision_matmul(lap, 2, 0.7, t), n).numpy()
rng = np.random.default_rng(0)
lab = np.zeros(n, bool); lab[rng.choice(n, m, replace=False)] = True
Qxx, Qxz, Qzz = Q[np.ix_(lab, lab)], Q[np.ix_(lab, ~lab)], Q[np.ix_(~lab, ~lab)]


def cg(A, b, tol, x0=None, maxit=5000):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x if x0 is not None else b.copy()
    bn = np.linalg.norm(b)
    p = r.copy(); rr = r @ r; it = 0
    while np.sqrt(rr) > tol * bn and it < maxit:
        Ap = A @ p; a = rr / (p @ Ap); x += a * p; r -= a * Ap
        rn = r @ r; p = r + (rn / rr) * p; rr = rn; it += 1
    return x, it


def outer(warm, tol_in=1e-6, tol_out=1e-6):
    b = rng.standard_normal(m)
    inner_its, prev = [], None
    def S(v):
        nonlocal prev
        rhs = Qxz.T @ v
        y, it = cg(Qzz, rhs, tol_in, x0=prev if warm else None)
        prev = y
        inner_its.append(it)
        return Qxx @ v - Qxz @ y
    xs = np.zeros(m); r = b.copy(); p = r.copy(); rr = r @ r; k = 0
    while np.sqrt(rr) > tol_out * np.linalg.norm(b) and k < 500:
        Sp = S(p); a = rr / (p @ Sp); xs += a * p; r -= a * Sp
        rn = r @ r; p = r + (rn / rr) * p; rr = rn; k += 1
    return k, inner_its


rng = np.random.default_rng(1); k0, cold = outer(False)
rng = np.random.default_rng(1); k1, warm = outer(True)
print(json.dumps({"n": n, "labelled": m, "outer_iterations": [k0, k1], "inner_iterations_per_outer_matvec_cold": round(float(np.mean(cold)), 1),
                  "inner_iterations_per_outer_matvec_warm": round(float(np.mean(warm)), 1),
                  "note": "warm start = previous outer iteration's inner solution as the initial guess (same stopping rule relative to |rhs|)"}))
