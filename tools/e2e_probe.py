"""Host-side breakdown of the end-to-end C3 cycle (what bench.py's e2e leg times)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J
S = J.scenario
cyc = S.dense_cycle(16384, 4, 32, 3, 5.0, 20251019)
O, M, Tt = cyc.obstacles.shape
cfg = J.JsspConfig(plan_horizon=cyc.plan_horizon, max_seeds=len(cyc.seeds), max_obstacle_modes=O * M, max_total_steps=Tt, max_knots=1024)
eng = J.JsspEngine(cfg)
sp = J.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1]); eng.set_refline(sp.s_list, sp.coefficient_table())
o = cyc.obstacles
pinned = torch.from_numpy(cyc.seeds).pin_memory()
dev = torch.device("cuda", 0)
def step(detail=None):
    t0 = time.perf_counter()
    sd = pinned.to(dev, non_blocking=True)
    t1 = time.perf_counter()
    eng.set_obstacles(o.state, o.valid, o.size, o.prob, o.n_dict)
    t2 = time.perf_counter()
    r = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=sd, targets=cyc.targets, sync=True)
    t3 = time.perf_counter()
    if detail is not None:
        detail.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
for _ in range(20): step()
d = []
for _ in range(200): step(d)
d = np.array(d) * 1e6
print("median us: seeds.to() enqueue %.1f | set_obstacles %.1f | evaluate+sync %.1f | total %.1f" % tuple(np.median(d, 0)))
# pieces alone, synchronised
def timed(fn, n=200):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e6
sd = pinned.to(dev)
print("alone (enqueue + full sync): seeds H2D %.1f us | set_obstacles %.1f us | evaluate(sync) %.1f us | evaluate(no export) %.1f us" % (
    timed(lambda: pinned.to(dev, non_blocking=True)), timed(lambda: eng.set_obstacles(o.state, o.valid, o.size, o.prob, o.n_dict)),
    timed(lambda: eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=sd, targets=cyc.targets, sync=True)),
    timed(lambda: eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=sd, targets=cyc.targets, sync=False, export=False))))
