"""CPU oracle for the iRotAvg IRLS rotation-averaging hot path (numpy / scipy, float64).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (irotavg_b200/, include/) may import,
call, link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs use it, and only as the checker / reported baseline.

PARITY PIN: the reference ships no expected output, known-answer test or golden vector for this path
(ral/test.cpp is a CLI, ral/data/ravg_input.txt is input only), and its dependencies (Eigen, SuiteSparseQR,
UMFPACK) are absent from this image.  Since round 2 the oracle is nevertheless pinned to the reference's OWN
CODE: oracle/build_ref.py compiles ral/l1_irls.cpp and ral/test.cpp unmodified from /root/reference against
stand-in headers for the three absent libraries (oracle/ref_shim/: eager-evaluation Eigen subset, dense
Householder-QR SuiteSparseQR, dense partial-pivoting-LU UMFPACK) into oracle/_ref/, and
tests/golden/make_golden_ref.py commits what that build produces (the CLI on the bundled fixture for four
costs; all 14 costs, make_A, residuals, exp_map, init_mst, l1ra, l1ra->irls flows on seeded graphs).
tests/test_ref_pin.py holds this oracle to those numbers (<= 1e-12 rad RMS) and, where oracle/_ref is present,
to the live reference on fresh random graphs.  What stays unpinned is the arithmetic INSIDE the three third-party
libraries (SPQR's / UMFPACK's elimination order, Eigen's vectorised reductions): rounding-level effects.
Further pins (tests/test_oracle.py):
  1. noise-free known-answer graphs (ground truth recovered to ~1e-15 rad),
  2. agreement of three independent formulations of the linear step
     (dense QR least squares on D*A  ==  sparse LU on A^T D^2 A  ==  Jacobi-PCG),
  3. agreement with the independent plain-C restatement oracle/irls_oracle.c.

Every function cites the reference lines it restates (paths relative to /root/reference).
The third-party arithmetic is restated from its definition:
  * SuiteSparseQR<double>(DA, DB)  (ral/l1_irls.cpp:550, system package, version unpinned):
    X = argmin ||DA X - DB||_F.  Restated as numpy lstsq (dense QR/SVD), sparse LU of the
    normal equations, or Jacobi-PCG on the normal equations.
  * Eigen::Quaterniond product (ral/l1_irls.cpp:99-105): the Hamilton product.
Quaternions are rows [x y z w] (ral/test.cpp:193,221).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

EPS = 2.2204e-16  # ral/l1_irls.hpp:40

# enum Cost, ral/l1_irls.hpp:56-57 (integer values 0..13 in this order)
L2, L1, L15, L05, GEMAN_MCCLURE, HUBER, PSEUDO_HUBER, ANDREWS, BISQUARE, CAUCHY, FAIR, \
    LOGISTIC, TALWAR, WELSCH = range(14)
COST_NAMES = ["L2", "L1", "L1.5", "L0.5", "Geman-McClure", "Huber", "Pseudo-Huber", "Andrews",
              "Bisquare", "Cauchy", "Fair", "Logistic", "Talwar", "Welsch"]


def parse_cost(name: str) -> int:
    """ral/test.cpp:35-72 (case-insensitive name -> enum)."""
    low = name.lower()
    for k, nm in enumerate(COST_NAMES):
        if nm.lower() == low:
            return k
    raise ValueError("Unknown string. " + name)


# --------------------------------------------------------------------------------------------
# quaternion arithmetic
# --------------------------------------------------------------------------------------------
def quat_mult(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Hamilton product a (x) b on rows [x y z w].  ral/l1_irls.cpp:99-105."""
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    out = np.empty(np.broadcast(a, b).shape, dtype=np.float64)
    out[..., 0] = aw * bx + ax * bw + ay * bz - az * by
    out[..., 1] = aw * by + ay * bw + az * bx - ax * bz
    out[..., 2] = aw * bz + az * bw + ax * by - ay * bx
    out[..., 3] = aw * bw - ax * bx - ay * by - az * bz
    return out


def delta_rel(I: np.ndarray, QQ: np.ndarray, Q: np.ndarray) -> np.ndarray:
    """resp_k = Q_inv[j] (x) (QQ[k] (x) Q[i]), Q_inv = Q with column w negated.
    ral/l1_irls.cpp:109-127 (note: -conj(q), not conj(q); the wrap in log_map absorbs it)."""
    Q_inv = Q.copy()
    Q_inv[:, 3] *= -1.0
    return quat_mult(Q_inv[I[:, 1]], quat_mult(QQ, Q[I[:, 0]]))


def log_map(w: np.ndarray) -> np.ndarray:
    """In place: rows -> [axis*theta, theta], theta wrapped to [-pi, pi).  ral/l1_irls.cpp:498-532."""
    s2 = np.sqrt(w[:, 0] ** 2 + w[:, 1] ** 2 + w[:, 2] ** 2)
    theta = 2.0 * np.arctan2(s2, w[:, 3])
    theta = np.where(theta < -np.pi, theta + 2 * np.pi,
                     np.where(theta >= np.pi, theta - 2 * np.pi, theta))
    with np.errstate(divide="ignore", invalid="ignore"):
        aux = theta / s2
        w[:, 0] *= aux
        w[:, 1] *= aux
        w[:, 2] *= aux
    w[:, 3] = theta
    w[s2 < EPS, 0:3] = 0.0
    return w


def exp_map(W: np.ndarray) -> np.ndarray:
    """In place: rows [v, *] -> [v sin(|v|/2)/|v|, cos(|v|/2)], non-finite -> 0.  ral/l1_irls.cpp:471-492."""
    theta = np.sqrt(W[:, 0] ** 2 + W[:, 1] ** 2 + W[:, 2] ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        ang = np.sin(theta / 2.0) / theta
        W[:, 3] = np.cos(theta / 2.0)
        W[:, 0] *= ang
        W[:, 1] *= ang
        W[:, 2] *= ang
    W[~np.isfinite(W)] = 0.0
    return W


def quat_normalised(Q: np.ndarray, f: int) -> np.ndarray:
    """Normalise rows f..n in place.  ral/l1_irls.cpp:982-991."""
    nrm = np.sqrt((Q[f:] ** 2).sum(axis=1))
    Q[f:] /= nrm[:, None]
    return Q


def rmat2quat(R: np.ndarray) -> np.ndarray:
    """3x3 rotation -> [x y z w], exact branch rule of src/ViewGraph.cpp:1175-1203."""
    q = np.zeros(4)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0.0:
        s = np.sqrt(tr + 1.0)
        q[3] = s * 0.5
        s = 0.5 / s
        q[0] = (R[2, 1] - R[1, 2]) * s
        q[1] = (R[0, 2] - R[2, 0]) * s
        q[2] = (R[1, 0] - R[0, 1]) * s
    else:
        i = (2 if R[1, 1] < R[2, 2] else 1) if R[0, 0] < R[1, 1] else (2 if R[0, 0] < R[2, 2] else 0)
        j = (i + 1) % 3
        k = (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = s * 0.5
        s = 0.5 / s
        q[3] = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
    return q


def quat2rmat(q: np.ndarray) -> np.ndarray:
    """Normalised [x y z w] -> 3x3 (Eigen toRotationMatrix; src/ViewGraph.cpp:1426-1433)."""
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


# --------------------------------------------------------------------------------------------
# incidence matrix and its mask
# --------------------------------------------------------------------------------------------
def edge_mask(I: np.ndarray, f: int):
    """(has_j, has_i) per edge under make_A's rule: +1 at j-f if j>=f; -1 at i-f only if ALSO
    i>=f (the `continue` at ral/l1_irls.cpp:771 drops the whole edge when j is fixed)."""
    has_j = I[:, 1] >= f
    has_i = has_j & (I[:, 0] >= f)
    return has_j, has_i


def make_A(n: int, f: int, I: np.ndarray):
    """m x (n-f) sparse incidence matrix.  ral/l1_irls.cpp:755-780.  `n` is the TOTAL node count
    (ral/test.cpp:288 passes n, src/ViewGraph.cpp:1400 passes num_of_vertices)."""
    import scipy.sparse as sp
    m = I.shape[0]
    has_j, has_i = edge_mask(I, f)
    rows = np.concatenate([np.nonzero(has_j)[0], np.nonzero(has_i)[0]])
    cols = np.concatenate([I[has_j, 1] - f, I[has_i, 0] - f])
    vals = np.concatenate([np.ones(has_j.sum()), -np.ones(has_i.sum())])
    return sp.csc_matrix((vals, (rows, cols)), shape=(m, n - f))


# --------------------------------------------------------------------------------------------
# robust weights
# --------------------------------------------------------------------------------------------
def update_weights(cost: int, sigma: float, E: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """The cost switch of ral/l1_irls.cpp:617-727.  `weights` are square-root weights (they scale
    rows of A); returns the new vector (Huber / L2 keep entries of the old one)."""
    e2 = E[:, 0] ** 2 + E[:, 1] ** 2 + E[:, 2] ** 2  # rowwise().squaredNorm()
    e = np.sqrt(e2)                                   # rowwise().norm()
    with np.errstate(divide="ignore", invalid="ignore"):
        if cost == L2:                                                     # :619-620
            return weights
        if cost == L05:                                                    # :621-625
            w = 1.0 / np.power(e2, 3.0 / 8.0)
            return np.where(w > 1e4, 1e4, w)
        if cost == L1:                                                     # :626-630
            w = 1.0 / np.sqrt(e)
            return np.where(w > 1e4, 1e4, w)
        if cost == L15:                                                    # :631-635
            w = 1.0 / np.sqrt(np.sqrt(e))
            return np.where(w > 1e4, 1e4, w)
        if cost == GEMAN_MCCLURE:                                          # :636-642
            return 1.0 / (e2 + sigma * sigma)
        if cost == HUBER:                                                  # :643-651 (sticky)
            tun = 1.345 * sigma
            ee = e / tun
            w = weights.copy()
            sel = ee >= 1
            w[sel] = np.sqrt(1.0 / ee[sel])
            return w
        if cost == PSEUDO_HUBER:                                           # :652-658
            tun = sigma
            return 1.0 / np.sqrt(np.sqrt(1.0 + e2 / (tun * tun)))
        if cost == ANDREWS:                                                # :659-677
            tun = 1.339 * sigma
            ee = e / tun
            w = np.sqrt(np.sin(ee) / ee)
            w = np.where(ee >= np.pi, 0.0, np.where(ee < 1e-4, 1.0, w))
            # `weights(i) < 0.0001` is false for NaN, so NaN would survive; sin(e)/e<0 only for e>=pi
            return np.where(w < 1e-4, 1e-4, w)
        if cost == BISQUARE:                                               # :678-684
            tun = 4.685 * sigma
            w = 1.0 - e2 / (tun * tun)
            return np.where(w < 1e-4, 1e-4, w)
        if cost == CAUCHY:                                                 # :685-691
            tun = 2.385 * sigma
            return 1.0 / np.sqrt(1.0 + e2 / (tun * tun))
        if cost == FAIR:                                                   # :692-698
            tun = 1.400 * sigma
            return 1.0 / np.sqrt(1.0 + e / tun)
        if cost == LOGISTIC:                                               # :699-707
            tun = 1.205 * sigma
            ee = e / tun
            w = np.sqrt(np.tanh(ee) / ee)
            return np.where(ee < 1e-4, 1.0, w)
        if cost == TALWAR:                                                 # :708-714
            tun = 2.795 * sigma
            return np.where(e2 < tun * tun, 1.0001, 0.0)
        if cost == WELSCH:                                                 # :715-722
            tun = 2.985 * sigma
            w = np.exp(-0.5 * e2 / (tun * tun))
            return np.where(w < 1e-4, 1e-4, w)
    raise ValueError("Unknown cost!!")                                     # :723-726


# --------------------------------------------------------------------------------------------
# linear step: X = argmin || D A X - D w ||_F      (ral/l1_irls.cpp:596-612, ls_solve :536-556)
# --------------------------------------------------------------------------------------------
def pcg_jacobi(L, d, B, rtol=1e-13, max_iters=20000):
    """Jacobi-preconditioned CG on L X = B (3 right-hand sides, independent alpha/beta per column),
    x0 = 0.  Returns (X, iterations)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        dinv = np.where(d > 0, 1.0 / d, 0.0)
    X = np.zeros_like(B)
    R = B.copy()
    Z = R * dinv[:, None]
    P = Z.copy()
    rz = (R * Z).sum(axis=0)
    bn = np.sqrt((B * B).sum(axis=0))
    it = 0
    while it < max_iters:
        rn = np.sqrt((R * R).sum(axis=0))
        if np.all(rn <= rtol * bn):
            break
        AP = L @ P
        pAp = (P * AP).sum(axis=0)
        alpha = np.where(pAp > 0, rz / np.where(pAp > 0, pAp, 1.0), 0.0)
        X += alpha * P
        R -= alpha * AP
        Z = R * dinv[:, None]
        rz_new = (R * Z).sum(axis=0)
        beta = np.where(rz > 0, rz_new / np.where(rz > 0, rz, 1.0), 0.0)
        P = Z + beta * P
        rz = rz_new
        it += 1
    return X, it


def ls_solve(A, weights: np.ndarray, w3: np.ndarray, solver: str = "direct", pcg_rtol=1e-13):
    """Restates ls_solve(W3, DA, DB) with DA = diag(weights) A, DB = weights .* w3.
    solver: 'lstsq' (dense QR/SVD on DA: nearest to SPQR semantics; small graphs only),
            'direct' (sparse LU of A^T D^2 A), 'pcg' (Jacobi-PCG on A^T D^2 A, x0=0)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    if solver == "lstsq":
        DA = sp.diags(weights) @ A
        DB = weights[:, None] * w3
        X, *_ = np.linalg.lstsq(DA.toarray(), DB, rcond=None)
        return X, 0
    w2 = weights * weights
    At = A.T.tocsr()
    Lm = (At @ sp.diags(w2) @ A).tocsc()
    B = At @ (w2[:, None] * w3)
    if solver == "direct":
        lu = spla.splu(Lm)
        X = lu.solve(B)
        # one step of iterative refinement keeps the normal-equation solve at QR-level accuracy
        X += lu.solve(B - Lm @ X)
        return X, 0
    if solver == "pcg":
        return pcg_jacobi(Lm.tocsr(), Lm.diagonal(), B, rtol=pcg_rtol)
    raise ValueError(solver)


# --------------------------------------------------------------------------------------------
# the IRLS loop
# --------------------------------------------------------------------------------------------
@dataclass
class IrlsResult:
    Q: np.ndarray
    weights: np.ndarray
    iters: int
    runtime: float
    scores: list = field(default_factory=list)
    cg_iters: list = field(default_factory=list)


def irls(QQ, I, A, cost, sigma, Q, f, max_iters, change_th, solver="direct", pcg_rtol=1e-13,
         verbose=False) -> IrlsResult:
    """irotavg::irls, ral/l1_irls.cpp:559-752.  Q is NOT modified; the updated copy is returned.
    `A` may be None (it is a pure function of (n, f, I), ral/l1_irls.cpp:755-780)."""
    tic = time.perf_counter()
    QQ = np.asarray(QQ, dtype=np.float64)
    I = np.asarray(I, dtype=np.int64).reshape(-1, 2)
    Q = np.array(Q, dtype=np.float64, copy=True)
    m = QQ.shape[0]
    n = Q.shape[0] - f                       # number of variables (:568)
    if A is None:
        A = make_A(Q.shape[0], f, I)
    A = A.tocsr()
    weights = np.ones(m)                     # :577
    score = np.finfo(np.float64).max         # :574
    iters = 0
    scores, cgs = [], []
    while score > change_th and iters < max_iters:       # :590
        w = log_map(delta_rel(I, QQ, Q))                 # :592-593
        W3, k = ls_solve(A, weights, w[:, :3], solver, pcg_rtol)   # :596-612
        E = A @ W3 - w[:, :3]                            # :614
        weights = update_weights(cost, sigma, E, weights)  # :617-727
        score = float(np.sqrt((W3 * W3).sum(axis=1)).mean())  # :729
        W = np.zeros((n, 4))
        W[:, :3] = W3
        exp_map(W)                                       # :731
        Q[f:] = quat_mult(Q[f:], W)                      # :734-737 (no renormalisation)
        iters += 1
        scores.append(score)
        cgs.append(k)
        if verbose:
            print(f"IRLS iteration: {iters:4d}; Change: {score:10.6g}; solver iters {k}")
    return IrlsResult(Q=Q, weights=weights, iters=iters, runtime=time.perf_counter() - tic,
                      scores=scores, cg_iters=cgs)


# --------------------------------------------------------------------------------------------
# spanning-tree initialisation (NEXT row #3 of SURVEY 8(f); needed for the CLI flow / config 1)
# --------------------------------------------------------------------------------------------
def init_mst(Q, QQ, I, f):
    """irotavg::init_mst, ral/l1_irls.cpp:915-979: repeated sweeps over the edge list in order,
    propagating from flagged to unflagged endpoints.  Order dependent - restated literally."""
    Q = np.array(Q, dtype=np.float64, copy=True)
    n = Q.shape[0]
    m = QQ.shape[0]
    flags = np.zeros(n, dtype=bool)
    flags[0] = True
    count = 1
    e1s = [int(v) for v in I[:, 0]]
    e2s = [int(v) for v in I[:, 1]]
    while count < n:
        span = False
        for k in range(m):
            e1, e2 = e1s[k], e2s[k]
            if flags[e1] and not flags[e2]:
                if e2 >= f:
                    Q[e2] = quat_mult(QQ[k], Q[e1])                       # :941
                count += 1
                flags[e2] = True
                span = True
            if (not flags[e1]) and flags[e2]:
                if e1 >= f:
                    qinv = QQ[k].copy()
                    qinv[3] *= -1.0                                       # :956-957
                    Q[e1] = quat_mult(qinv, Q[e2])
                count += 1
                flags[e1] = True
                span = True
        if not span and count < n:
            raise RuntimeError("Relative rotations DO NOT SPAN all the nodes in the VIEW GRAPH")
    return Q


# --------------------------------------------------------------------------------------------
# metrics
# --------------------------------------------------------------------------------------------
def geodesic_angles(Qa: np.ndarray, Qb: np.ndarray) -> np.ndarray:
    """ang(Qa_i^-1 (x) Qb_i) = 2 atan2(|v|, |w|) per row; inputs need not be unit."""
    Qa = Qa / np.linalg.norm(Qa, axis=1, keepdims=True)
    Qb = Qb / np.linalg.norm(Qb, axis=1, keepdims=True)
    conj = Qa * np.array([-1.0, -1.0, -1.0, 1.0])
    d = quat_mult(conj, Qb)
    return 2.0 * np.arctan2(np.sqrt((d[:, :3] ** 2).sum(axis=1)), np.abs(d[:, 3]))


def geodesic_rms(Qa, Qb, f=0) -> float:
    a = geodesic_angles(np.asarray(Qa)[f:], np.asarray(Qb)[f:])
    return float(np.sqrt((a * a).mean())) if a.size else 0.0


# --------------------------------------------------------------------------------------------
# L1RA initial stage (NEXT row #1 of SURVEY 8(f)):  l1ra + l1decode_pd + make_AtA
# --------------------------------------------------------------------------------------------
def make_A_noquirk(n: int, f: int, I: np.ndarray):
    """The incidence pattern behind make_AtA (ral/l1_irls.cpp:811-848): +1 at j-f if j>=f, -1 at i-f if
    i>=f, INDEPENDENTLY of each other - make_AtA has no `continue`, so an edge (free i, fixed j) still
    puts sigma on H(i,i) although make_A dropped it (App. A.6.1).  reshape(AtA*sigma) = A'^T diag(sigma) A'
    up to the sign convention (AtA stores +1 on both diagonals and -1 off-diagonal)."""
    import scipy.sparse as sp
    m = I.shape[0]
    hj = I[:, 1] >= f
    hi = I[:, 0] >= f
    rows = np.concatenate([np.nonzero(hj)[0], np.nonzero(hi)[0]])
    cols = np.concatenate([I[hj, 1] - f, I[hi, 0] - f])
    vals = np.concatenate([np.ones(hj.sum()), -np.ones(hi.sum())])
    return sp.csc_matrix((vals, (rows, cols)), shape=(m, n - f))


def l1decode_pd(x0, A, y, pdmaxiter, Ah, newton="direct", pcg_rtol=1e-13, trace=None):
    """Primal-dual interior-point L1 regression min ||A x - y||_1, restating ral/l1_irls.cpp:228-468
    line by line (l1-magic's l1decode_pd).  `Ah` is the pattern of make_AtA (make_A_noquirk): the
    Newton matrix is H = Ah^T diag(sigx) Ah (:308-319), solved exactly by UMFPACK in the reference
    (linsolve, :131-184; third-party, unpinned) and here by sparse LU ('direct') or Jacobi-PCG ('pcg')."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    PDTOL, alpha, beta, mu = 1e-3, 0.01, 0.5, 10.0                         # :231-238
    n, m = x0.size, y.size
    At = A.T.tocsr()
    x = x0.copy()
    Ax = A @ x
    yab = np.abs(y - Ax)
    u = 0.95 * yab + 0.10 * yab.max()                                      # :248-253
    fu1 = Ax - y - u
    fu2 = -Ax + y - u
    lamu1 = -1.0 / fu1
    lamu2 = -1.0 / fu2
    Atv = At @ (lamu1 - lamu2)                                             # :262
    sdg = -(fu1 @ lamu1 + fu2 @ lamu2)
    tau = mu * 2 * m / sdg                                                 # :265
    rcent = np.concatenate([-lamu1 * fu1, -lamu2 * fu2]) - 1.0 / tau
    rdual_n = Atv.copy()                                                   # gradf0 = [0; 1]
    rdual_m = 1.0 - lamu1 - lamu2
    resnorm = np.sqrt(rdual_n @ rdual_n + rdual_m @ rdual_m + rcent @ rcent)
    pditer = 0
    done = (sdg < PDTOL) or (pditer >= pdmaxiter)                          # :284
    xp = x.copy()       # the reference returns an unassigned Vec when the loop never runs (A.6.7)
    Aht = Ah.T.tocsr()
    while not done:
        pditer += 1
        w2 = -1.0 - (1.0 / tau) * (1.0 / fu1 + 1.0 / fu2)                  # :293
        sig1 = -lamu1 / fu1 - lamu2 / fu2
        sig2 = lamu1 / fu1 - lamu2 / fu2
        sigx = sig1 - sig2 ** 2 / sig1                                     # :296-298
        w1 = -(1.0 / tau) * (At @ (-1.0 / fu1 + 1.0 / fu2))                # :301-302
        w1p = w1 - At @ ((sig2 / sig1) * w2)                               # :305-306
        H = (Aht @ sp.diags(sigx) @ Ah).tocsc()                            # :308-317
        if newton == "direct":
            lu = spla.splu(H)
            dx = lu.solve(w1p)
            dx += lu.solve(w1p - H @ dx)
        else:
            dx = pcg_jacobi(H.tocsr(), H.diagonal(), w1p[:, None], rtol=pcg_rtol)[0][:, 0]
        Adx = A @ dx                                                       # :324
        du = (w2 - sig2 * Adx) / sig1                                      # :327
        dlamu1 = -(lamu1 / fu1) * (Adx - du) - lamu1 - (1.0 / tau) / fu1   # :330-333
        dlamu2 = (lamu2 / fu2) * (Adx + du) - lamu2 - (1.0 / tau) / fu2    # :336-339
        Atdv = At @ (dlamu1 - dlamu2)                                      # :342
        s = 1.0                                                            # :347-381
        neg = dlamu1 < 0
        if neg.any():
            s = min(s, (-lamu1[neg] / dlamu1[neg]).min())
        neg = dlamu2 < 0
        if neg.any():
            s = min(s, (-lamu2[neg] / dlamu2[neg]).min())
        d1 = Adx - du
        pos = d1 > 0
        if pos.any():
            s = min(s, (-fu1[pos] / d1[pos]).min())
        d2 = -Adx - du
        pos = d2 > 0
        if pos.any():
            s = min(s, (-fu2[pos] / d2[pos]).min())
        s *= 0.99
        suffdec = False
        backiter = 0
        while not suffdec:                                                 # :392-429
            xp = x + s * dx
            up = u + s * du
            Axp = Ax + s * Adx
            Atvp = Atv + s * Atdv
            lamu1p = lamu1 + s * dlamu1
            lamu2p = lamu2 + s * dlamu2
            fu1p = Axp - y - up
            fu2p = -Axp + y - up
            rdp_n = Atvp
            rdp_m = 1.0 - lamu1p - lamu2p
            rcp = np.concatenate([-lamu1p * fu1p, -lamu2p * fu2p]) - 1.0 / tau
            suffdec = np.sqrt(rdp_n @ rdp_n + rdp_m @ rdp_m + rcp @ rcp) <= (1 - alpha * s) * resnorm
            s *= beta
            backiter += 1
            if backiter > 32:                                              # :423-428
                return x.copy()
        x, u, Ax, Atv = xp, up, Axp, Atvp                                  # :432-442
        lamu1, lamu2, fu1, fu2 = lamu1p, lamu2p, fu1p, fu2p
        sdg = -(fu1 @ lamu1 + fu2 @ lamu2)                                 # :446
        tau = mu * 2 * m / sdg
        rcent = np.concatenate([-lamu1 * fu1, -lamu2 * fu2]) - 1.0 / tau
        resnorm = np.sqrt(rdp_n @ rdp_n + rdp_m @ rdp_m + rcent @ rcent)   # rdual = rdp (:455-458)
        if trace is not None:
            trace.append(dict(pditer=pditer, backiter=backiter, s=s / beta, sdg=sdg, tau=tau, resnorm=resnorm,
                              sigx_min=float(sigx.min()), sigx_max=float(sigx.max())))
        done = (sdg < PDTOL) or (pditer >= pdmaxiter)
    return xp


@dataclass
class L1raResult:
    Q: np.ndarray
    iters: int
    runtime: float
    scores: list = field(default_factory=list)
    trace: list = field(default_factory=list)


def l1ra(QQ, I, A, Q, f, max_iters, change_th, newton="direct", pcg_rtol=1e-13) -> L1raResult:
    """irotavg::l1ra, ral/l1_irls.cpp:851-912.  l1_step stays 2 for ever: the `if (score<change_th)` at
    :879 is unreachable because the loop condition already requires score >= change_th (A.6.7)."""
    tic = time.perf_counter()
    QQ = np.asarray(QQ, dtype=np.float64)
    I = np.asarray(I, dtype=np.int64).reshape(-1, 2)
    Q = np.array(Q, dtype=np.float64, copy=True)
    ntot = Q.shape[0]
    n = ntot - f
    if A is None:
        A = make_A(ntot, f, I)
    A = A.tocsr()
    Ah = make_A_noquirk(ntot, f, I).tocsr()
    score = np.finfo(np.float64).max
    it = 0
    l1_step = 2
    scores, trace = [], []
    while (score >= change_th or l1_step < 2) and it < max_iters:          # :877
        w = log_map(delta_rel(I, QQ, Q))                                   # :885-887
        W = np.zeros((n, 4))
        for c in range(3):                                                 # :890-892
            tr = []
            W[:, c] = l1decode_pd(np.zeros(n), A, w[:, c].copy(), l1_step, Ah, newton, pcg_rtol, tr)
            trace.append(tr)
        score = float(np.sqrt((W[:, :3] ** 2).sum(axis=1)).mean())         # :894
        exp_map(W)
        Q[f:] = quat_mult(Q[f:], W)                                        # :899-902
        it += 1
        scores.append(score)
    return L1raResult(Q=Q, iters=it, runtime=time.perf_counter() - tic, scores=scores, trace=trace)
