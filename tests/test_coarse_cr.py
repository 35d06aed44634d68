"""The cyclic-reduction schedule of the tridiagonal coarse solve (irotavg_b200/csrc/ira_coarse.cuh: tri_factor /
tri_solve), restated index for index in numpy and checked against a dense solve: level strides, the slot of every
stored multiplier, the back-substitution order, padding to a power of two and the dead-pivot rule.  The CUDA kernel is
held to the oracle by tests/test_gpu_parity.py::test_two_level_*; this file pins the algorithm it implements."""
import numpy as np
import pytest


def _guard(b, s):
    return 1.0 / b if (s > 0.0 and b > 1e-12 * s) else 0.0


def tri_factor(d, lo):
    N = d.size
    BI, D0, L1 = d.copy(), d.copy(), lo.copy()
    L2, M1, M2 = np.zeros(N), np.zeros(N), np.zeros(N)
    s = 1
    while s < N:
        cnt, off = N // (2 * s), N - N // s
        upd = []
        for kk in range(cnt):                                   # one thread per kept unknown; reads precede writes
            i = 2 * s * (kk + 1) - 1
            jl, jr = i - s, i + s
            ci = L1[i]
            k1 = ci * _guard(BI[jl], D0[jl])
            cr = L1[jr] if jr < N else 0.0
            k2 = cr * _guard(BI[jr], D0[jr]) if jr < N else 0.0
            M1[off + kk], M2[off + kk], L2[jl] = k1, k2, ci
            upd.append((i, BI[i] - (ci * k1 + cr * k2), -L1[jl] * k1))
        for i, b, l in upd:
            BI[i], L1[i] = b, l
        for kk in range(cnt):                                   # the eliminated unknowns keep 1 / pivot
            j = s - 1 + 2 * s * kk
            BI[j] = _guard(BI[j], D0[j])
        s *= 2
    BI[N - 1] = _guard(BI[N - 1], D0[N - 1])
    return BI, L1, L2, M1, M2


def tri_solve(F, r):
    BI, L1, L2, M1, M2 = F
    N = r.size
    Y = r.copy()
    s = 1
    while s < N:
        cnt, off = N // (2 * s), N - N // s
        for kk in range(cnt):
            i = 2 * s * (kk + 1) - 1
            Y[i] -= M1[off + kk] * Y[i - s] + (M2[off + kk] * Y[i + s] if i + s < N else 0.0)
        s *= 2
    Y[N - 1] *= BI[N - 1]
    s = N // 2
    while s >= 1:
        for kk in range(N // (2 * s)):
            j = s - 1 + 2 * s * kk
            y = Y[j] - L2[j] * Y[j + s]
            if j >= s:
                y -= L1[j] * Y[j - s]
            Y[j] = y * BI[j]
        s //= 2
    return Y


def _system(nc, N, rng, grounded=0.3):
    lo = np.zeros(N)
    lo[1:nc] = -rng.uniform(0.1, 10.0, nc - 1)                  # coupling of block a to block a - 1
    d = np.ones(N)                                              # padding rows: identity
    for a in range(nc):
        d[a] = -lo[a] + (-lo[a + 1] if a + 1 < nc else 0.0) + (rng.uniform(0, 1) if rng.random() < grounded else 0.0)
    d[0] += 1.0                                                 # the fixed node grounds the first block
    return d, lo


@pytest.mark.parametrize("nc,N", [(2, 2), (5, 8), (65, 128), (229, 256), (594, 1024), (1024, 1024)])
def test_cyclic_reduction_matches_dense_solve(nc, N):
    rng = np.random.default_rng(nc)
    d, lo = _system(nc, N, rng)
    T = np.diag(d) + np.diag(lo[1:], -1) + np.diag(lo[1:], 1)
    r = np.zeros(N)
    r[:nc] = rng.standard_normal(nc)
    y = tri_solve(tri_factor(d, lo), r)
    ref = np.linalg.solve(T, r)
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.all(y[nc:] == 0.0)


def test_dead_pivot_switches_the_unknown_off():
    """A floating component (blocks 3..5 tied to each other only) has a singular diagonal block: its last pivot is zero,
    that unknown is switched off (y = 0) and the others solve T restricted to them - a symmetric PSD operator."""
    N = 8
    d = np.array([2.0, 2.0, 1.0, 1.0, 2.0, 1.0, 1.0, 1.0])
    lo = np.array([0.0, -1.0, -1.0, 0.0, -1.0, -1.0, 0.0, 0.0])
    F = tri_factor(d, lo)
    M = np.column_stack([tri_solve(F, e) for e in np.eye(N)])
    assert np.allclose(M, M.T, atol=1e-14)
    assert np.linalg.eigvalsh(0.5 * (M + M.T)).min() >= -1e-14
    dead = [j for j in range(N) if np.all(M[j] == 0.0)]
    assert dead == [3]
    keep = [j for j in range(N) if j != 3]
    T = np.diag(d) + np.diag(lo[1:], -1) + np.diag(lo[1:], 1)
    assert np.allclose(M[np.ix_(keep, keep)], np.linalg.inv(T[np.ix_(keep, keep)]), atol=1e-12)


def test_empty_blocks_of_the_partition():
    """Blocks whose rows are all fixed assemble to zero rows (diag 0, couplings 0): switched off, the rest unaffected."""
    rng = np.random.default_rng(3)
    nc, N = 40, 64
    d, lo = _system(nc, N, rng)
    d[:5] = 0.0
    lo[:6] = 0.0
    d[5] += 2.0                                                 # grounded through the fixed prefix
    F = tri_factor(d, lo)
    r = np.zeros(N)
    r[5:nc] = rng.standard_normal(nc - 5)
    y = tri_solve(F, r)
    T = np.diag(d) + np.diag(lo[1:], -1) + np.diag(lo[1:], 1)
    ref = np.linalg.solve(T[5:nc, 5:nc], r[5:nc])
    assert np.all(y[:5] == 0.0)
    assert np.abs(y[5:nc] - ref).max() <= 1e-12 * np.abs(ref).max()
