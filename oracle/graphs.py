"""Synthetic SO(3) view graphs and the RAL text format (ral/test.cpp:161-247,314-326).

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/irls_oracle.py header).  Everything is seeded
(numpy default_rng, seed 20190319 unless stated) so the oracle, the CPU baseline and the GPU path
read identical bytes.  Quaternions are rows [x y z w]; edge (i, j) means Q_j = QQ_ij (x) Q_i
(ral/l1_irls.cpp:941).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .irls_oracle import quat_mult

SEED = 20190319


@dataclass
class Graph:
    I: np.ndarray        # (m, 2) int32, rows (i, j)
    QQ: np.ndarray       # (m, 4) float64 relative rotations [x y z w]
    Q0: np.ndarray       # (n, 4) float64 initial absolute rotations (rows < f are the fixed ones)
    Qgt: np.ndarray      # (n, 4) float64 ground truth (or None)
    f: int
    name: str = ""

    @property
    def m(self):
        return self.I.shape[0]

    @property
    def n(self):
        return self.Q0.shape[0]


def _exp_quat(v: np.ndarray) -> np.ndarray:
    th = np.sqrt((v * v).sum(axis=1))
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(th > 0, np.sin(th / 2) / np.where(th > 0, th, 1.0), 0.5)
    return np.concatenate([v * k[:, None], np.cos(th / 2)[:, None]], axis=1)


def _conj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def _rand_quat(rng, k):
    q = rng.standard_normal((k, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 3] < 0] *= -1.0
    return q


def _measure(rng, Qgt, I, sigma_n, outlier_mask):
    """QQ_k = Q*_j (x) Exp(eta_k) (x) Q*_i^-1, outliers replaced by uniform random rotations."""
    m = I.shape[0]
    eta = rng.standard_normal((m, 3)) * sigma_n
    QQ = quat_mult(Qgt[I[:, 1]], quat_mult(_exp_quat(eta), _conj(Qgt[I[:, 0]])))
    k = int(outlier_mask.sum())
    if k:
        QQ[outlier_mask] = _rand_quat(rng, k)
    return QQ


def _init(rng, Qgt, f, sigma_init):
    Q0 = Qgt.copy()
    n = Qgt.shape[0]
    if sigma_init > 0:
        Q0[f:] = quat_mult(Qgt[f:], _exp_quat(rng.standard_normal((n - f, 3)) * sigma_init))
    return Q0


def _distinct_pairs(rng, n, k, exclude_adjacent=True):
    """k distinct uniform pairs i<j (not (t,t+1) when exclude_adjacent)."""
    seen = np.empty(0, dtype=np.int64)
    out = np.empty(0, dtype=np.int64)
    while out.size < k:
        need = int((k - out.size) * 1.2) + 16
        a = rng.integers(0, n, need)
        b = rng.integers(0, n, need)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        ok = lo != hi
        if exclude_adjacent:
            ok &= (hi - lo) != 1
        key = lo[ok] * n + hi[ok]
        _, first = np.unique(key, return_index=True)
        key = key[np.sort(first)]
        key = key[~np.isin(key, seen)]
        seen = np.concatenate([seen, key])
        out = np.concatenate([out, key])
    out = out[:k]
    return np.stack([out // n, out % n], axis=1)


def random_graph(n=100_000, m=1_000_000, sigma_n=0.05, outlier_frac=0.10, sigma_init=0.1,
                 seed=SEED, f=1, name=None) -> Graph:
    """Config 3 / 4 recipe: path (k,k+1) + (m-n+1) distinct uniform pairs; 10 % of the non-path
    edges are outliers; node 0 fixed at ground truth."""
    rng = np.random.default_rng(seed)
    Qgt = _rand_quat(rng, n)
    path = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
    extra = _distinct_pairs(rng, n, m - (n - 1))
    I = np.concatenate([path, extra]).astype(np.int32)
    out = np.zeros(m, dtype=bool)
    out[n - 1:] = rng.random(m - (n - 1)) < outlier_frac
    QQ = _measure(rng, Qgt, I, sigma_n, out)
    Q0 = _init(rng, Qgt, f, sigma_init)
    return Graph(I=I, QQ=QQ, Q0=Q0, Qgt=Qgt, f=f, name=name or f"random_n{n}_m{m}")


def kitti_like_graph(n=4541, m=50_000, band=10, sigma_n=0.005, outlier_frac=0.02, sigma_init=0.1,
                     seed=SEED, f=1, name=None) -> Graph:
    """Config 2 recipe ("KITTI-00 scale"): smooth trajectory, band edges (k-d,k) d=1..band, the
    rest loop closures between indices > 500 apart; 2 % of the loop edges are outliers."""
    rng = np.random.default_rng(seed)
    Qgt = np.empty((n, 4))
    Qgt[0] = _rand_quat(rng, 1)[0]
    steps = _exp_quat(rng.standard_normal((n - 1, 3)) * 0.02)
    for k in range(1, n):
        Qgt[k] = quat_mult(Qgt[k - 1], steps[k - 1])
    Qgt /= np.linalg.norm(Qgt, axis=1, keepdims=True)
    bands = [np.stack([np.arange(0, n - d), np.arange(d, n)], axis=1) for d in range(1, band + 1)]
    bandI = np.concatenate(bands)
    bandI = bandI[np.lexsort((bandI[:, 0], bandI[:, 1]))]       # ordered by (j, i) like a SLAM run
    n_loop = m - bandI.shape[0]
    loops = np.empty((0, 2), dtype=np.int64)
    while loops.shape[0] < n_loop:
        cand = _distinct_pairs(rng, n, n_loop * 2, exclude_adjacent=False)
        cand = cand[(cand[:, 1] - cand[:, 0]) > 500]
        loops = np.unique(np.concatenate([loops, cand]), axis=0)
        loops = loops[rng.permutation(loops.shape[0])]
    loops = loops[:n_loop]
    I = np.concatenate([bandI, loops]).astype(np.int32)
    out = np.zeros(m, dtype=bool)
    out[bandI.shape[0]:] = rng.random(n_loop) < outlier_frac
    QQ = _measure(rng, Qgt, I, sigma_n, out)
    Q0 = _init(rng, Qgt, f, sigma_init)
    return Graph(I=I, QQ=QQ, Q0=Q0, Qgt=Qgt, f=f, name=name or f"kitti_like_n{n}_m{m}")


def small_graph(n=200, extra=1500, sigma_n=0.0, outlier_frac=0.0, sigma_init=0.2, seed=7, f=1,
                fixed_anywhere=False, name=None) -> Graph:
    """Small test graph.  sigma_n=0, outlier_frac=0 gives a noise-free known-answer graph whose
    ground truth is independent of any implementation.  fixed_anywhere=True scatters edges so that
    (free i, fixed j) edges exist and make_A's dropped-edge quirk (App. A.6.1) is exercised."""
    rng = np.random.default_rng(seed)
    Qgt = _rand_quat(rng, n)
    path = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
    ex = _distinct_pairs(rng, n, extra)
    I = np.concatenate([path, ex])
    if fixed_anywhere:
        flip = rng.random(I.shape[0]) < 0.3
        I[flip] = I[flip][:, ::-1]          # some edges now have i > j, i.e. possibly j < f <= i
    I = I.astype(np.int32)
    out = np.zeros(I.shape[0], dtype=bool)
    out[n - 1:] = rng.random(extra) < outlier_frac
    QQ = _measure(rng, Qgt, I, sigma_n, out)
    Q0 = _init(rng, Qgt, f, sigma_init)
    return Graph(I=I, QQ=QQ, Q0=Q0, Qgt=Qgt, f=f, name=name or f"small_n{n}")


def banded_graph(n=400, band=6, loops=120, sigma_n=0.005, outlier_frac=0.05, sigma_init=0.1, seed=5, f=1,
                 name=None) -> Graph:
    """A small view-graph-like case for the dense reference build (oracle/_ref): smooth trajectory, band edges
    (k-d, k) d = 1..band ordered by (j, i) like a SLAM run, plus `loops` random loop closures (5 % outliers)."""
    rng = np.random.default_rng(seed)
    Qgt = np.empty((n, 4))
    Qgt[0] = _rand_quat(rng, 1)[0]
    steps = _exp_quat(rng.standard_normal((n - 1, 3)) * 0.02)
    for k in range(1, n):
        Qgt[k] = quat_mult(Qgt[k - 1], steps[k - 1])
    Qgt /= np.linalg.norm(Qgt, axis=1, keepdims=True)
    bandI = np.concatenate([np.stack([np.arange(0, n - d), np.arange(d, n)], axis=1) for d in range(1, band + 1)])
    bandI = bandI[np.lexsort((bandI[:, 0], bandI[:, 1]))]
    lp = _distinct_pairs(rng, n, loops)
    lp = lp[(lp[:, 1] - lp[:, 0]) > band]
    I = np.concatenate([bandI, lp]).astype(np.int32)
    out = np.zeros(I.shape[0], dtype=bool)
    out[bandI.shape[0]:] = rng.random(lp.shape[0]) < outlier_frac
    QQ = _measure(rng, Qgt, I, sigma_n, out)
    Q0 = _init(rng, Qgt, f, sigma_init)
    return Graph(I=I, QQ=QQ, Q0=Q0, Qgt=Qgt, f=f, name=name or f"banded_n{n}")


# --------------------------------------------------------------------------------------------
# RAL text format
# --------------------------------------------------------------------------------------------
def read_ral_text(path: str):
    """Parse `m n f` / m lines `i j w x y z` / up to n lines `w x y z` exactly as ral/test.cpp:161-247
    does: vertex ids compacted in sorted order, file quaternions [w x y z] stored as [x y z w].
    Returns (I, QQ, Q, f_file, n_given)."""
    with open(path) as fh:
        tok = fh.read().split()
    m, n, f = int(tok[0]), int(tok[1]), int(tok[2])
    p = 3
    rel = np.array(tok[p:p + 6 * m], dtype=np.float64).reshape(m, 6)
    if rel.shape[0] != m:
        raise ValueError("Corrupt input file: inconsistent number of connections.")
    p += 6 * m
    ids = rel[:, :2].astype(np.int64)
    verts = np.unique(ids)
    I = np.searchsorted(verts, ids).astype(np.int32)
    QQ = rel[:, [3, 4, 5, 2]].copy()
    rest = tok[p:]
    n_given = min(n, len(rest) // 4)
    Q = np.zeros((n, 4))
    if n_given:
        a = np.array(rest[:4 * n_given], dtype=np.float64).reshape(n_given, 4)
        Q[:n_given] = a[:, [1, 2, 3, 0]]
    if n_given < f:
        raise ValueError(f"Insuficient number of absolute rotations. At least {f} must be given.")
    if n != int(I[:, 1].max()) + 1:
        raise ValueError("Corrupt input file: check abs rotations")
    return I, QQ, Q, f, n_given


def write_ral_text(path: str, I, QQ, Q_given, f):
    with open(path, "w") as fh:
        n = int(max(np.unique(np.asarray(I)).size, len(Q_given)))
        fh.write(f"{len(I)} {n} {f}\n")
        for (i, j), q in zip(I, QQ):
            fh.write(f"{i} {j} {q[3]:.17g} {q[0]:.17g} {q[1]:.17g} {q[2]:.17g}\n")
        for q in Q_given:
            fh.write(f"{q[3]:.17g} {q[0]:.17g} {q[1]:.17g} {q[2]:.17g}\n")


def write_ral_output(path: str, Q, weights):
    """n rows `w x y z` then m weights, full precision (ral/test.cpp:314-326)."""
    with open(path, "w") as fh:
        for q in Q:
            fh.write(f"{q[3]:.17g} {q[0]:.17g} {q[1]:.17g} {q[2]:.17g}\n")
        for w in weights:
            fh.write(f"{w:.17g}\n")
