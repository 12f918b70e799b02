"""A/B timing of the evaluation kernels on a dense cycle: python tools/kernel_ab.py [C3|C5|W3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J
which = sys.argv[1] if len(sys.argv) > 1 else "C3"
S = J.scenario
cyc = {"C3": lambda: S.dense_cycle(16384, 4, 32, 3, 5.0, 20251019), "C5": lambda: S.dense_cycle(131072, 8, 64, 3, 8.0, 20251020),
       "W3": lambda: S.dense_cycle(21845, 3, 6, 1, 5.0, 7)}[which]()
O, M, Tt = cyc.obstacles.shape
cfg = J.JsspConfig(plan_horizon=cyc.plan_horizon, max_seeds=len(cyc.seeds), max_obstacle_modes=O * M, max_total_steps=Tt, max_knots=1024)
eng = J.JsspEngine(cfg)
sp = J.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1]); eng.set_refline(sp.s_list, sp.coefficient_table())
o = cyc.obstacles; eng.set_obstacles(o.state, o.valid, o.size, o.prob, o.n_dict)
seeds = torch.from_numpy(cyc.seeds).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for kern in ("thread", "group"):
    for cull in (True, False):
        for _ in range(5):
            eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, sync=False, kernel=kern, cull=cull)
        ms = []
        for _ in range(20):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, sync=False, kernel=kern, cull=cull); e1.record()
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        r = eng.fetch()
        sig = (r.best_idx, r.best_cost, int(r.feasible.sum()), int(r.collision.sum()))
        ref = ref or sig
        print(f"{which} kernel={kern} cull={cull}: median {np.median(ms)*1e3:.1f} us min {min(ms)*1e3:.1f} us  {sig} same={sig == ref}")
