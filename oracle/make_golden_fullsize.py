"""Full-size oracle results for BASELINE configs[2] (C3: 65,536 candidates) and configs[4] (C5: 1,048,576 candidates).

TEST INFRASTRUCTURE.  Runs the vectorised float64 oracle (oracle/jssp_oracle.py::plan_cycle) over the complete synthetic
cycle in seed chunks on a process pool (candidates are independent until the argmin, jerk_space_sampling_planner.py:288-294,
326-327) and stores, per workload: the packed feasible / collision masks of EVERY candidate, the oracle's argmin
(lowest index among equal costs) and its cost, and the float64 costs of every 64th candidate.  The GPU tests
(tests/test_gpu_parity.py) compare the CUDA path at full size against these - masks and selected index bit-exact.

    python oracle/make_golden_fullsize.py [C3] [C5]     -> tests/golden/fullsize_golden.npz   (~280 KB)

The workloads are the seeded generators of dlii_jssp_b200/scenario.py (dense_cycle), the same ones bench.py times.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SPECS = {"C3": dict(n_seeds=16384, n_widths=4, O=32, M=3, plan_horizon=5.0, seed=20251019),
         "C5": dict(n_seeds=131072, n_widths=8, O=64, M=3, plan_horizon=8.0, seed=20251020)}
COST_STRIDE = 64
_CYC = {}


def _cycle(name):
    from dlii_jssp_b200 import scenario as S
    if name not in _CYC:
        _CYC[name] = S.dense_cycle(name=name, **SPECS[name])
    return _CYC[name]


def _chunk(args):
    name, lo, hi = args
    from oracle import jssp_oracle as O
    cyc = _cycle(name)
    cfg = O.OracleConfig(ego_length=cyc.ego_length, ego_width=cyc.ego_width, plan_horizon=cyc.plan_horizon)
    spline = O.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1])
    o = cyc.obstacles
    with np.errstate(all="ignore"):
        ref = O.plan_cycle(cfg, spline, cyc.frenet0, (o.state, o.valid.astype(bool), o.size, o.prob), cyc.time_step_now, cyc.final_time_step,
                           seeds=cyc.seeds[lo:hi], targets=cyc.targets)
    W, n = len(cyc.targets), hi - lo
    # chunk-local candidate w * n + k  ->  [W, n]
    return lo, hi, ref["feasible"].reshape(W, n), ref["collision"].reshape(W, n), ref["cost"].reshape(W, n)


def full_cycle(name: str, chunk: int = 512, procs=None):
    import multiprocessing as mp
    cyc = _cycle(name)
    Ns, W = len(cyc.seeds), len(cyc.targets)
    feas = np.zeros((W, Ns), dtype=bool); coll = np.zeros((W, Ns), dtype=bool); cost = np.zeros((W, Ns))
    jobs = [(name, lo, min(lo + chunk, Ns)) for lo in range(0, Ns, chunk)]
    with mp.get_context("fork").Pool(procs or os.cpu_count()) as pool:
        for lo, hi, f, c, k in pool.imap_unordered(_chunk, jobs):
            feas[:, lo:hi], coll[:, lo:hi], cost[:, lo:hi] = f, c, k
    feas, coll, cost = feas.reshape(-1), coll.reshape(-1), cost.reshape(-1)      # global candidate n = w * Ns + k
    ok = feas & ~coll & np.isfinite(cost)
    best = -1
    if ok.any():
        cmin = cost[ok].min()
        best = int(np.nonzero(ok & (cost == cmin))[0][0])                        # lowest index among equal costs
    return feas, coll, cost, best


def main():
    names = [a for a in sys.argv[1:] if a in SPECS] or list(SPECS)
    path = os.path.join(ROOT, "tests", "golden", "fullsize_golden.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for name in names:
        t0 = time.time()
        feas, coll, cost, best = full_cycle(name)
        out[f"{name}_feasible_bits"] = np.packbits(feas)
        out[f"{name}_collision_bits"] = np.packbits(coll)
        out[f"{name}_cost_stride{COST_STRIDE}"] = cost[::COST_STRIDE].copy()
        out[f"{name}_best"] = np.array([best], dtype=np.int64)
        out[f"{name}_best_cost"] = np.array([cost[best] if best >= 0 else np.nan])
        out[f"{name}_n"] = np.array([len(feas)], dtype=np.int64)
        # runner-up margin: how far the second-best survivor is from the best (says how robust the argmin is to rounding)
        ok = feas & ~coll & np.isfinite(cost)
        rest = np.delete(cost[ok], np.nonzero(np.nonzero(ok)[0] == best)[0]) if best >= 0 else np.array([])
        out[f"{name}_margin"] = np.array([rest.min() - cost[best] if rest.size else np.inf])
        print(f"{name}: {len(feas)} candidates, feasible {int(feas.sum())}, collision {int(coll.sum())}, survivors {int(ok.sum())}, best {best} "
              f"cost {cost[best] if best >= 0 else None}, margin {out[f'{name}_margin'][0]:.3e}, {time.time() - t0:.0f} s")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
