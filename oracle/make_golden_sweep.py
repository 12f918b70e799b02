"""Oracle replays of a seeded sample of the 800-scenario sweep (BASELINE configs[3], "C4").

TEST INFRASTRUCTURE.  bench.py times the batched device-resident replay on ``scenario.synthetic_intersection(i)`` for
i in range(800), <= 100 plan cycles each.  This script replays a seeded sample of those ids with the float64 oracle planner in
the reference driver's loop (oracle/replay_oracle.py; code/__main__.py:124-167, env.py:21-33) and stores per scenario the end
status, the number of cycles, the selected index and candidate count of every cycle and the ego row after every cycle.
tests/test_gpu_replay.py checks the batched CUDA replay of the same ids against it, and - across ranks - that a sharded sweep
returns the same per-scenario rows as the 1-GPU run.

    python oracle/make_golden_sweep.py [n_sample=40]   -> tests/golden/sweep_golden.npz
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N_SWEEP, MAX_CYCLES, PAD = 800, 100, -9


def _one(i):
    from dlii_jssp_b200 import batch as B, replay as R, scenario as S
    from oracle import replay_oracle
    sc = S.synthetic_intersection(int(i))
    p = B.pack_scenario(sc, max_cycles=MAX_CYCLES)
    if p.initial_end != -1:
        return int(i), dict(best_idx=[], n=[], rows=[], end=p.initial_end)
    return int(i), replay_oracle.replay(sc, tuple(p.frenet0), S.pack_obstacle_dict, R.end_status, R.goal_pose_of(sc), max_cycles=MAX_CYCLES)


def main():
    import multiprocessing as mp
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    ids = np.sort(np.random.default_rng(20261020).choice(N_SWEEP, size=n, replace=False))
    t0 = time.time()
    with mp.get_context("spawn").Pool(min(os.cpu_count(), n)) as pool:
        res = dict(pool.imap_unordered(_one, ids.tolist()))
    best = np.full((n, MAX_CYCLES), PAD, dtype=np.int32); ncand = np.zeros((n, MAX_CYCLES), dtype=np.int32)
    rows = np.full((n, MAX_CYCLES, 6), np.nan); end = np.zeros(n); cycles = np.zeros(n, dtype=np.int32); nrows = np.zeros(n, dtype=np.int32)
    for k, i in enumerate(ids.tolist()):
        r = res[i]
        c = len(r["best_idx"])
        best[k, :c], ncand[k, :c], end[k], cycles[k], nrows[k] = r["best_idx"], r["n"], r["end"], c, len(r["rows"])
        if r["rows"]:
            rows[k, :len(r["rows"])] = np.array(r["rows"])
    path = os.path.join(ROOT, "tests", "golden", "sweep_golden.npz")
    np.savez_compressed(path, ids=ids, end=end, cycles=cycles, best_idx=best, n_candidates=ncand, rows=rows, n_rows=nrows,
                        max_cycles=np.array([MAX_CYCLES]))
    print(f"{n} scenarios, {int(cycles.sum())} oracle plan cycles in {time.time() - t0:.0f} s; end status counts:",
          {float(e): int((end == e).sum()) for e in np.unique(end)}, "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
