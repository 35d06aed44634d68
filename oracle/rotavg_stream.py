"""TEST INFRASTRUCTURE (CPU oracle) - ViewGraph::rotAvg and the driver's call pattern, restated.

  rot_avg()         src/ViewGraph.cpp:1263-1435 on plain Python containers (views = list of dicts with the
                    absolute rotation `R` and the adjacency `conn` {other id: relative R}; fixed mask = list)
  make_stream()     synthetic config-5 op list: per frame a new view (identity pose, src/Pose.hpp:43), edges to
                    the previous <= 4 views (src/IRotAvg.cpp:159,278), optionally a loop-closure edge / a ground-
                    truth fix, then rotAvg(10) or, after a loop edge or a fix, rotAvg(5000000) (src/IRotAvg.cpp:371-378)
  write_ops()       the op list as text for tests/cpp/rotavg_main.cpp and tools/rotavg_stream (same bytes for both arms)
  replay()          runs an op list through rot_avg() with the oracle's l1ra / irls

Only tests/, bench.py's CPU leg and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import numpy as np

from . import irls_oracle as O
from .graphs import _exp_quat

SIGMA = 5 * np.pi / 180.0


def rot_avg(views, fixed_mask, win, l1_iters=100, irls_iters=100, change_th=1e-3, solver="direct"):
    """Returns a report dict; updates views[k]['R'] of the free window views in place."""
    m_views = len(views)
    win = min(m_views, win)                                                  # :1269
    rep = dict(solved=False, vertices=0, edges=0, fixed=0, l1_iters=0, irls_iters=0)
    if win < 2:
        return rep
    first = m_views - win
    I, qq, vertices = [], [], set()
    for t in range(first, m_views):                                          # :1282-1307
        j = t
        for i in sorted(views[t]["conn"]):        # std::map<View*,...> order is by address; any order is the same problem
            if i < j:
                I.append((i, j))
                vertices.update((i, j))
                qq.append(O.rmat2quat(views[t]["conn"][i]))
    rep["edges"], rep["vertices"] = len(qq), len(vertices)
    if len(qq) < win or len(vertices) < win:                                 # :1313-1321
        return rep
    vs = sorted(vertices)
    f = len(vs) - win + sum(1 for x in vs if x >= first and fixed_mask[x])   # :1329-1338
    to_idx, to_vertex = {}, {}
    t, k = 0, f
    for x in vs:                                                             # :1343-1363
        if x >= first and not fixed_mask[x]:
            to_idx[x], to_vertex[k] = k, x
            k += 1
        else:
            to_idx[x], to_vertex[t] = t, x
            t += 1
    Ia = np.array([(to_idx[a], to_idx[b]) for a, b in I], dtype=np.int32).reshape(-1, 2)
    Q = np.zeros((len(vs), 4))
    for x in vs:
        Q[to_idx[x]] = O.rmat2quat(views[x]["R"])
    if f == 0:                                                               # :1382-1386
        Q[0] = [0, 0, 0, 1]
        f = 1
    rep["fixed"] = f
    QQ = np.array(qq).reshape(-1, 4)
    la = O.l1ra(QQ, Ia, None, Q, f, l1_iters, change_th)                     # :1402-1407
    r = O.irls(QQ, Ia, None, O.GEMAN_MCCLURE, SIGMA, la.Q, f, irls_iters, change_th, solver=solver)   # :1409-1417
    for k in range(f, len(vs)):                                              # :1420-1434
        views[to_vertex[k]]["R"] = O.quat2rmat(r.Q[k])
    rep.update(solved=True, l1_iters=la.iters, irls_iters=r.iters)
    return rep


def make_stream(n_frames=300, loop_every=100, fix_every=0, sigma_n=0.005, step=0.02, seed=20190319, local_win=10,
                min_loop_gap=50):
    """Op list [('V',), ('E', i, j, R), ('F', idx, R), ('A', win)] and the ground-truth rotations."""
    rng = np.random.default_rng(seed)
    Qgt = np.zeros((n_frames, 4))
    Qgt[0] = [0, 0, 0, 1]
    for k in range(1, n_frames):
        Qgt[k] = O.quat_mult(Qgt[k - 1][None, :], _exp_quat(rng.normal(0, step, (1, 3))))[0]
    conj = lambda q: np.array([-q[0], -q[1], -q[2], q[3]])

    def rel(i, j, outlier=False):
        noise = _exp_quat(rng.normal(0, sigma_n, (1, 3)))[0]
        q = O.quat_mult(O.quat_mult(Qgt[j][None, :], noise[None, :]), conj(Qgt[i])[None, :])[0]
        return O.quat2rmat(q)

    ops = []
    for t in range(n_frames):
        ops.append(("V",))
        for d in range(1, 5):
            if t - d >= 0:
                ops.append(("E", t - d, t, rel(t - d, t)))
        glob = False
        if loop_every and t > 0 and t % loop_every == 0 and t > min_loop_gap:
            i = int(rng.integers(0, t - min_loop_gap))
            ops.append(("E", i, t, rel(i, t)))
            glob = True
        if fix_every and t % fix_every == 0:
            ops.append(("F", t, O.quat2rmat(Qgt[t])))
            glob = True
        ops.append(("A", 5000000 if glob else local_win))
    return ops, Qgt


def write_ops(path, ops):
    with open(path, "w") as fh:
        fh.write(f"{len(ops)}\n")
        for op in ops:
            if op[0] == "V":
                fh.write("V\n")
            elif op[0] == "E":
                fh.write(f"E {op[1]} {op[2]} " + " ".join(f"{v:.17g}" for v in np.asarray(op[3]).ravel()) + "\n")
            elif op[0] == "F":
                fh.write(f"F {op[1]} " + " ".join(f"{v:.17g}" for v in np.asarray(op[2]).ravel()) + "\n")
            else:
                fh.write(f"A {op[1]}\n")


def replay(ops, solver="direct"):
    views, mask, reports = [], [], []
    for op in ops:
        if op[0] == "V":
            views.append(dict(R=np.eye(3), conn={}))
            mask.append(False)
        elif op[0] == "E":
            _, i, j, R = op
            if j not in views[i]["conn"]:                                     # View::connect dedups (:1441-1445)
                views[i]["conn"][j] = np.asarray(R)
                views[j]["conn"][i] = np.asarray(R)
        elif op[0] == "F":
            mask[op[1]] = True                                                # ViewGraph::fixPose (:1234-1246)
            views[op[1]]["R"] = np.asarray(op[2])
        else:
            reports.append(rot_avg(views, mask, op[1], solver=solver))
    return np.array([v["R"] for v in views]), reports
