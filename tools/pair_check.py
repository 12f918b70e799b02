"""Paired walk against the tiled walk at full size (tools/, GPU): masks, fp32 costs, selection and exported trajectory must be
identical on C3 and C5, full and select-only.   python tools/pair_check.py [C3 C5]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import dlii_jssp_b200 as J

for name in (sys.argv[1:] or ["C3", "C5"]):
    cyc = bench.workload(name)
    eng = bench.make_engine(J, cyc, len(cyc.seeds), 0)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    out = {}
    for kern in ("tiled", "pair"):
        r = eng.evaluate(cyc.frenet0, cyc.time_step_now, cyc.final_time_step, seeds=seeds, targets=cyc.targets, kernel=kern)
        s = eng.evaluate(cyc.frenet0, cyc.time_step_now, cyc.final_time_step, seeds=seeds, targets=cyc.targets, kernel=kern, select_only=True)
        out[kern] = (r.feasible.cpu().numpy().copy(), r.collision.cpu().numpy().copy(), r.cost.cpu().numpy().copy(), r.best_idx, r.best_cost,
                     np.nan_to_num(r.best_traj).copy(), s.best_idx, s.best_cost)
    a, b = out["tiled"], out["pair"]
    print(name, "feasible equal", np.array_equal(a[0], b[0]), int((a[0] != b[0]).sum()), "collision equal", np.array_equal(a[1], b[1]), int((a[1] != b[1]).sum()),
          "cost equal", np.array_equal(a[2], b[2], equal_nan=True), "best", a[3], b[3], a[4] == b[4], "traj", np.array_equal(a[5], b[5]),
          "select-only", a[6], b[6], a[7] == b[7], flush=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2], equal_nan=True)
    assert a[3] == b[3] and a[4] == b[4] and np.array_equal(a[5], b[5]) and a[6] == b[6] == a[3] and a[7] == b[7]
    eng.close()
print("pair_check ok")
