"""CPU checks of the two derivations the device code rests on (restated here in numpy, no device):

1. ira_mst.cuh: the time stamp label(v) = (sweep << 32) | (edge + 1) defined by the monotone relaxation
   "edge k can flag v from u at the first time (s, k) after label(u)" has a unique fixed point, and it is the order
   in which irotavg::init_mst's sequential sweeps (ral/l1_irls.cpp:915-979) flag the nodes - so the parent edge of
   every node is the same as in the reference.
2. ira_pcg.cuh (k_attach_block): the cancellation-free inverse of a 3x3 diagonal block of A^T D^2 A written with the
   'excess' e = d - (block couplings) equals the matrix inverse, also when the couplings are 1e8 and the excess 10.
"""
import numpy as np
import pytest

from oracle import graphs as G

INF = np.uint64(2**64 - 1)


def sequential_flag_order(I, n):
    """The literal sweeps of ral/l1_irls.cpp:925-968: returns per node (sweep, edge) of the moment it is flagged."""
    flags = np.zeros(n, dtype=bool)
    flags[0] = True
    when = {0: (0, -1)}
    count, sweep = 1, 0
    while count < n:
        progressed = False
        for k, (a, b) in enumerate(I):
            if flags[a] and not flags[b]:
                flags[b] = True; when[int(b)] = (sweep, k); count += 1; progressed = True
            if not flags[a] and flags[b]:
                flags[a] = True; when[int(a)] = (sweep, k); count += 1; progressed = True
        if not progressed:
            break
        sweep += 1
    return when


def relaxed_labels(I, n):
    """Jacobi-style relaxation to the fixed point (what k_mst_labels computes with atomicMin)."""
    lab = np.full(n, INF, dtype=np.uint64)
    lab[0] = 0
    k1 = np.arange(1, len(I) + 1, dtype=np.uint64)
    a, b = I[:, 0], I[:, 1]

    def nxt(t):
        sweep = (t >> np.uint64(32)) + ((t & np.uint64(0xffffffff)) >= k1).astype(np.uint64)
        return (sweep << np.uint64(32)) | k1

    for _ in range(10 * n):
        new = lab.copy()
        for src, dst in ((a, b), (b, a)):
            ok = (lab[src] != INF) & (src != dst)
            cand = np.where(ok, nxt(np.where(ok, lab[src], np.uint64(0))), INF)
            np.minimum.at(new, dst, cand)
        if np.array_equal(new, lab):
            return lab
        lab = new
    raise AssertionError("no fixed point")


@pytest.mark.parametrize("seed", range(6))
def test_time_stamp_fixed_point_is_the_sequential_order(seed):
    rng = np.random.default_rng(seed)
    g = G.small_graph(n=120, extra=int(rng.integers(0, 400)), seed=100 + seed)
    I = g.I[rng.permutation(g.m)].copy()
    flip = rng.random(len(I)) < 0.5
    I[flip] = I[flip][:, ::-1]
    when = sequential_flag_order(I, g.n)
    lab = relaxed_labels(I, g.n)
    assert len(when) == g.n
    for v in range(1, g.n):
        s, k = when[v]
        assert int(lab[v]) == (s << 32) | (k + 1), (v, when[v], int(lab[v]))


def test_time_stamp_reversed_path_needs_one_sweep_per_node():
    n = 40
    I = np.array([(k, k + 1) for k in range(n - 1)][::-1])
    lab = relaxed_labels(I, n)
    assert [int(l >> np.uint64(32)) for l in lab[1:]] == list(range(n - 1))


def block3_inverse(da, db, dt, wab, wat, wbt):
    """k_attach_block's arithmetic: eliminate a, Schur complement of (b, t) written with the excesses."""
    ea, eb, et = max(da - wab - wat, 0.0), max(db - wab - wbt, 0.0), max(dt - wat - wbt, 0.0)
    lb, lt = wab / da, wat / da
    xb, xt = eb + lb * ea, et + lt * ea
    wp = wbt + lb * wat
    det = wp * (xb + xt) + xb * xt
    sbb, sbt, stt = (wp + xt) / det, wp / det, (wp + xb) / det
    iab, iat = lb * sbb + lt * sbt, lb * sbt + lt * stt
    iaa = 1.0 / da + lb * iab + lt * iat
    return np.array([[iaa, iab, iat], [iab, sbb, sbt], [iat, sbt, stt]])


@pytest.mark.parametrize("stiff", [1.0, 1e3, 1e8])
def test_three_by_three_block_inverse(stiff):
    rng = np.random.default_rng(int(np.log10(stiff)) + 1)
    for _ in range(200):
        wab, wat, wbt = stiff * rng.uniform(0.5, 2), stiff * rng.uniform(0, 2) * (rng.random() < 0.8), rng.uniform(0, 30)
        if wat == 0.0 and wbt == 0.0:
            wbt = 1.0
        ea, eb, et = rng.uniform(1, 40, 3)
        da, db, dt = wab + wat + ea, wab + wbt + eb, wat + wbt + et
        M = np.array([[da, -wab, -wat], [-wab, db, -wbt], [-wat, -wbt, dt]])
        inv = block3_inverse(da, db, dt, wab, wat, wbt)
        assert np.all(inv > 0) and np.allclose(inv, inv.T)
        # the excesses are recovered from d by subtraction (as on the device): relative accuracy eps * stiff / e
        tol = 1e-13 * max(1.0, stiff / 1.0)
        assert np.max(np.abs(inv @ M - np.eye(3))) <= max(tol, 1e-12)
        ref = np.linalg.inv(M)
        assert np.allclose(inv, ref, rtol=1e-6, atol=0)
