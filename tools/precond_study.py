"""CPU study of preconditioners for the IRLS normal equations A^T D^2 A X = A^T D^2 w (numpy/scipy; no GPU).

Runs L1-IRLS on the configs[2] recipe (n, m scalable) with a strong block preconditioner so that it is quick,
snapshots the weights at chosen IRLS iterations, then counts PCG iterations (rtol 1e-10, 3 RHS in lock step,
stop when all converged - like the device kernels) for candidate preconditioners on those snapshots:

  jacobi            diagonal
  agg(theta, cap)   block-Jacobi over aggregates grown greedily along edges of normalised strength
                    w2_uv / sqrt(d_u d_v) >= theta, strongest edges first, at most `cap` nodes per aggregate,
                    every block inverted exactly (cap 2 ~ the mutual pairs, cap 3 ~ the 3x3 blocks of
                    irotavg_b200/csrc/ira_pcg.cuh)
  agg + coarse      the same plus an additive piecewise-constant coarse correction over unlimited stiff components

    python tools/precond_study.py --n 100000 --m 1000000 --snap 5,10,20,29
    python tools/precond_study.py --coarse          # chain-like graphs: Jacobi vs Jacobi + coarse space (ira_coarse.cuh)

Findings (n = 100 000, m = 1 000 000, L1): PCG iterations of the solve at IRLS iteration 10 / 20 / 29 - mutual pairs
67 / 221 / 436, <= 3-node aggregates 45 / 89 / 132, <= 8 nodes 43 / 80 / 120, <= 32 nodes 43 / 80 / 121: exact blocks
larger than 3 buy nothing (the stiff components are tiny, the rest of the spectrum is the spread of the robust weights).
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import graphs as G  # noqa: E402
from oracle import irls_oracle as O  # noqa: E402


def laplacian(n, f, I, w2):
    """Grounded weighted Laplacian over the free nodes with make_A's mask; returns (L csr, rows->node offset f)."""
    i, j = I[:, 0].astype(np.int64), I[:, 1].astype(np.int64)
    keep = j >= f
    both = keep & (i >= f)
    nf = n - f
    d = np.zeros(nf)
    np.add.at(d, j[keep] - f, w2[keep])
    np.add.at(d, i[both] - f, w2[both])
    r = np.concatenate([i[both] - f, j[both] - f])
    c = np.concatenate([j[both] - f, i[both] - f])
    v = np.concatenate([-w2[both], -w2[both]])
    L = sp.csr_matrix((v, (r, c)), shape=(nf, nf)) + sp.diags(d)
    return L.tocsr(), d


def rhs(n, f, I, w2, w3):
    i, j = I[:, 0].astype(np.int64), I[:, 1].astype(np.int64)
    keep = j >= f
    both = keep & (i >= f)
    B = np.zeros((n - f, 3))
    np.add.at(B, j[keep] - f, w2[keep, None] * w3[keep])
    np.add.at(B, i[both] - f, -w2[both, None] * w3[both])
    return B


def pcg(L, B, apply_M, rtol=1e-10, max_iters=50000):
    X = np.zeros_like(B)
    R = B.copy()
    Z = apply_M(R)
    P = Z.copy()
    rz = (R * Z).sum(0)
    bb = (B * B).sum(0)
    it = 0
    while it < max_iters:
        rr = (R * R).sum(0)
        if np.all(rr <= rtol * rtol * bb):
            break
        AP = L @ P
        pap = (P * AP).sum(0)
        al = np.where(pap > 0, rz / np.where(pap > 0, pap, 1), 0.0)
        X += al * P
        R -= al * AP
        Z = apply_M(R)
        rzn = (R * Z).sum(0)
        be = np.where(rz > 0, rzn / np.where(rz > 0, rz, 1), 0.0)
        rz = rzn
        P = Z + be * P
        it += 1
    return X, it


def greedy_aggregates(L, d, theta, cap):
    """Union-find over edges sorted by normalised strength (desc), merging while size <= cap."""
    Lc = sp.triu(L, 1).tocoo()
    st = -Lc.data / np.sqrt(d[Lc.row] * d[Lc.col])
    sel = st >= theta
    r, c, s = Lc.row[sel], Lc.col[sel], st[sel]
    order = np.argsort(-s, kind="stable")
    nf = L.shape[0]
    parent = np.arange(nf)
    size = np.ones(nf, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for k in order:
        a, b = find(r[k]), find(c[k])
        if a != b and size[a] + size[b] <= cap:
            if size[a] < size[b]:
                a, b = b, a
            parent[b] = a
            size[a] += size[b]
    lab = np.array([find(a) for a in range(nf)])
    return lab


def block_jacobi(L, lab):
    """Exact inverse of the diagonal blocks given by labels (block-diagonal sparse LU)."""
    Lc = L.tocoo()
    same = lab[Lc.row] == lab[Lc.col]
    Bd = sp.csc_matrix((Lc.data[same], (Lc.row[same], Lc.col[same])), shape=L.shape)
    perm = np.argsort(lab, kind="stable")
    Bp = Bd[perm][:, perm].tocsc()
    lu = spla.splu(Bp, permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)

    def apply(R):
        return lu.solve(R[perm])[inv]
    return apply


def coarse_additive(L, d, lab, base_apply, exact=False):
    """M^-1 = base + P (P^T L P)^-1 P^T with P piecewise constant over labels (diagonal approx unless exact)."""
    nf = L.shape[0]
    uniq, idx = np.unique(lab, return_inverse=True)
    nc = uniq.size
    P = sp.csr_matrix((np.ones(nf), (np.arange(nf), idx)), shape=(nf, nc))
    Lc = (P.T @ L @ P).tocsc()
    if exact:
        lu = spla.splu(Lc)
        cs = lu.solve
    else:
        dc = Lc.diagonal()
        cs = lambda y: y / dc[:, None]

    def apply(R):
        return base_apply(R) + P @ cs(P.T @ R)
    return apply, nc


def coarse_study():
    """PCG iterations (rtol 1e-10) with Jacobi and with Jacobi + piecewise-constant coarse space over contiguous index
    blocks (exact coarse solve), on the chain-like graphs: a config-5 stream graph, the reference's bundled fixture,
    config 2.  The numbers quoted in irotavg_b200/csrc/ira_coarse.cuh."""
    from oracle import rotavg_stream as RS

    def test(name, n, f, I, w2, ncs=(32, 64, 128)):
        L, d = laplacian(n, f, I, w2)
        nf = n - f
        B = np.random.default_rng(0).standard_normal((nf, 3))
        dj = lambda R: R / d[:, None]
        out = {"jacobi": pcg(L, B, dj, max_iters=20000)[1]}
        for nc in ncs:
            bsz = int(np.ceil(nf / nc))
            M, ncc = coarse_additive(L, d, np.arange(nf) // bsz, dj, exact=True)
            out[f"coarse{ncc}"] = pcg(L, B, M, max_iters=20000)[1]
        print(name, "n", n, "m", len(I), out, flush=True)
    ops, _ = RS.make_stream(n_frames=3000, loop_every=500, min_loop_gap=500)
    I = np.array([(op[1], op[2]) for op in ops if op[0] == "E"])
    test("stream, 3000 views, unit weights", 3000, 1, I, np.ones(len(I)))
    test("stream, 3000 views, weights 1e-2..1e2", 3000, 1, I, 10 ** np.random.default_rng(1).uniform(-2, 2, len(I)))
    b = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bundled_graph.npz"))
    test("bundled fixture, unit weights", 1832, 1, b["I"], np.ones(len(b["I"])))
    g = G.kitti_like_graph()
    test("config 2, unit weights", g.n, 1, g.I, np.ones(g.m))


def coarse_tri_study(n=9500):
    """The numbers behind the tridiagonal coarse operator (ira_coarse.cuh TRI) on a config-5 stream graph:
    PCG iterations (rtol 1e-10, random right-hand sides) with Jacobi + piecewise-constant coarse space over nc
    contiguous index blocks, the coarse operator P^T L P (a) solved exactly, (b) cut to its tridiagonal part with the far
    couplings kept on the diagonal ("lump", what the kernel does) or removed from it ("drop"), (c) (b) plus the dense
    64-block space, (d) a BPX-style multilevel sum of diagonally scaled levels above an exact 64-block level."""
    from oracle import rotavg_stream as RS
    ops, _ = RS.make_stream(n_frames=n, loop_every=500, min_loop_gap=500)
    I = np.array([(op[1], op[2]) for op in ops if op[0] == "E"])
    nf = n - 1

    def P_of(bsz):
        lab = np.arange(nf) // bsz
        return sp.csr_matrix((np.ones(nf), (np.arange(nf), lab)), shape=(nf, lab.max() + 1))

    def tri(L, d, nc, mode):
        P = P_of(int(np.ceil(nf / nc)))
        Ac = (P.T @ L @ P).tocoo()
        near = np.abs(Ac.row - Ac.col) <= 1
        T = sp.csc_matrix((Ac.data[near], (Ac.row[near], Ac.col[near])), shape=Ac.shape)
        if mode == "drop":
            far = ~near
            T = T + sp.csc_matrix((Ac.data[far], (Ac.row[far], Ac.row[far])), shape=Ac.shape)
        lu = spla.splu(T.tocsc())
        return (lambda R: R / d[:, None] + P @ lu.solve(P.T @ R)), P.shape[1]

    def multilevel(L, d, factor):
        B = int(np.ceil(nf / 64))
        Ps, Ds, b = [], [], factor
        while b < B:
            P = P_of(b)
            Ps.append(P)
            Ds.append((P.T @ L @ P).diagonal())
            b *= factor
        PL = P_of(B)
        lu = spla.splu((PL.T @ L @ PL).tocsc())

        def apply(R):
            U = R / d[:, None]
            for P, D in zip(Ps, Ds):
                U = U + P @ ((P.T @ R) / D[:, None])
            return U + PL @ lu.solve(PL.T @ R)
        return apply

    for wname, w2 in (("unit weights", np.ones(len(I))), ("weights 1e-2..1e2", 10 ** np.random.default_rng(1).uniform(-2, 2, len(I)))):
        L, d = laplacian(n, 1, I, w2)
        B = np.random.default_rng(0).standard_normal((nf, 3))
        dj = lambda R: R / d[:, None]
        out = {}
        for nc in (64, 128, 256, 512, 1024):
            M, ncc = coarse_additive(L, d, np.arange(nf) // int(np.ceil(nf / nc)), dj, exact=True)
            out[f"exact{ncc}"] = pcg(L, B, M, max_iters=20000)[1]
        for nc in (64, 256, 512, 1024, 2048):
            for mode in ("lump", "drop"):
                M, ncc = tri(L, d, nc, mode)
                out[f"tri{ncc}{mode}"] = pcg(L, B, M, max_iters=20000)[1]
        for nc in (512, 1024):
            M1, ncc = tri(L, d, nc, "lump")
            M2, _ = coarse_additive(L, d, np.arange(nf) // int(np.ceil(nf / 64)), lambda R: 0 * R, exact=True)
            out[f"tri{ncc}lump+dense64"] = pcg(L, B, lambda R: M1(R) + M2(R), max_iters=20000)[1]
        for factor in (2, 4, 8):
            out[f"multilevel x{factor}"] = pcg(L, B, multilevel(L, d, factor), max_iters=20000)[1]
        print(f"stream, {n} views, {wname}:", out, flush=True)


def main():
    if "--coarse-tri" in sys.argv:
        return coarse_tri_study()
    if "--coarse" in sys.argv:
        return coarse_study()
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--m", type=int, default=1000000)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--snap", default="5,10,20,29")
    ap.add_argument("--cost", default="L1")
    ap.add_argument("--save", default="")
    a = ap.parse_args()
    snaps = [int(s) for s in a.snap.split(",")]
    g = G.random_graph(n=a.n, m=a.m)
    cost = O.parse_cost(a.cost)
    sigma = 5 * np.pi / 180
    n, f, I = g.n, g.f, g.I
    Q = g.Q0.copy()
    weights = np.ones(g.m)
    i64, j64 = I[:, 0].astype(np.int64), I[:, 1].astype(np.int64)
    for it in range(a.iters):
        w = O.log_map(O.delta_rel(I, g.QQ, Q))
        w3 = w[:, :3]
        w2 = weights * weights
        L, d = laplacian(n, f, I, w2)
        B = rhs(n, f, I, w2, w3)
        t0 = time.time()
        lab = greedy_aggregates(L, d, 0.05, 8) if it > 0 else np.arange(n - f)
        M = block_jacobi(L, lab)
        X, k = pcg(L, B, M)
        print(f"irls {it}: solve pcg(agg8) {k} its, {time.time() - t0:.1f}s, |X| mean {np.linalg.norm(X, axis=1).mean():.3e}", flush=True)
        if it in snaps:
            print(f"  -- snapshot at IRLS iteration {it}: weights^2 quantiles "
                  f"{np.quantile(w2, [0, .5, .9, .99, .999, 1])}")
            dj = lambda R: R / d[:, None]
            res = {}
            _, res["jacobi"] = pcg(L, B, dj, max_iters=300)
            for theta, cap in ((0.2, 2), (0.05, 3), (0.05, 8), (0.05, 32), (0.02, 32), (0.3, 256)):
                t1 = time.time()
                lab2 = greedy_aggregates(L, d, theta, cap)
                sizes = np.bincount(np.unique(lab2, return_inverse=True)[1])
                M2 = block_jacobi(L, lab2)
                _, kk = pcg(L, B, M2)
                res[f"agg(th={theta},cap={cap})"] = kk
                print(f"     agg theta={theta} cap={cap}: {kk} its; aggregates>1: {(sizes > 1).sum()}, max size {sizes.max()}, "
                      f"nodes in aggregates {(sizes[sizes > 1]).sum()}  ({time.time() - t1:.1f}s)", flush=True)
                if cap == 3:
                    lab_inf = greedy_aggregates(L, d, 0.3, 10**9)
                    for exact in (False,):
                        M3, nc = coarse_additive(L, d, lab_inf, M2, exact)
                        _, kk = pcg(L, B, M3)
                        res[f"agg3+coarse(exact={exact})"] = kk
                        print(f"     agg3 + additive coarse over unlimited th=0.3 components (nc={nc}, exact={exact}): {kk} its", flush=True)
            print("  ", res, flush=True)
        # finish the IRLS iteration
        Xf = np.zeros((n, 3))
        Xf[f:] = X
        E = -w3.copy()
        keep = j64 >= f
        E[keep] += Xf[j64[keep]]
        both = keep & (i64 >= f)
        E[both] -= Xf[i64[both]]
        weights = O.update_weights(cost, sigma, E, weights)
        W = np.zeros((n - f, 4))
        W[:, :3] = X
        W = O.exp_map(W)
        Q[f:] = O.quat_mult(Q[f:], W)
    if a.save:
        np.savez(a.save, Q=Q, weights=weights)


if __name__ == "__main__":
    main()
