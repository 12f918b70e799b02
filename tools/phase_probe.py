"""Where the time of the group evaluation kernel goes, per block (needs a profiling build):
    JSSP_NVCC_EXTRA=-DJSSP_PROFILE_PHASES python -m dlii_jssp_b200.build && python tools/phase_probe.py [C3|C5|W3]
Stamps (globaltimer, ns): 0 entry, 1 tables staged, 2 lists built, 3 walk done, 4 select/export done."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J
which = sys.argv[1] if len(sys.argv) > 1 else "C3"
S = J.scenario
cyc = {"C3": lambda: S.dense_cycle(16384, 4, 32, 3, 5.0, 20251019), "C5": lambda: S.dense_cycle(131072, 8, 64, 3, 8.0, 20251020),
       "W3": lambda: S.dense_cycle(21845, 3, 6, 1, 5.0, 7), "C1": lambda: S.dense_cycle(150, 3, 6, 1, 5.0, 7),
       "C1x": lambda: S.dense_cycle(1500, 3, 6, 1, 5.0, 7)}[which]()
kern = sys.argv[2] if len(sys.argv) > 2 else "group"
O, M, Tt = cyc.obstacles.shape
cfg = J.JsspConfig(plan_horizon=cyc.plan_horizon, max_seeds=len(cyc.seeds), max_obstacle_modes=O * M, max_total_steps=Tt, max_knots=1024)
eng = J.JsspEngine(cfg)
sp = J.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1]); eng.set_refline(sp.s_list, sp.coefficient_table())
o = cyc.obstacles; eng.set_obstacles(o.state, o.valid, o.size, o.prob, o.n_dict)
seeds = torch.from_numpy(cyc.seeds).cuda()
nb = 1 << 16
ts = torch.zeros((nb, 16), dtype=torch.int64, device="cuda")
fn = eng.lib.jssp_debug_phase_buffer; fn.restype = C.c_int32; fn.argtypes = [C.c_void_p, C.c_void_p]
assert fn(eng._h, ts.data_ptr()) == 0
for _ in range(3):
    eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, sync=False, kernel=kern)
torch.cuda.synchronize(); ts.zero_()
eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, sync=True, kernel=kern)
t = ts.cpu().numpy()
used = t[:, 0] > 0
t = t[used].astype(np.float64)
t0 = t[:, 0].min()
print(f"{which}: {used.sum()} blocks; kernel span {(t[:, 4].max() - t0) / 1e3:.1f} us")
names = ["entry", "staged", "lists", "walk", "select"]
last = t[t[:, 5] > 0]
if len(last):
    print(f"  last block: argmin +{(last[0, 7] - last[0, 3]) / 1e3:.1f} us after its walk, reduce done +{(last[0, 5] - last[0, 3]) / 1e3:.1f} us, export {(last[0, 6] - last[0, 5]) / 1e3:.1f} us, "
          f"tail {(last[0, 4] - last[0, 6]) / 1e3:.1f} us")
for k in range(5):
    col = (t[:, k] - t0) / 1e3
    print(f"  {names[k]:7s} stamp: min {col.min():7.1f}  median {np.median(col):7.1f}  max {col.max():7.1f} us")
d = (t[:, 7] - t[:, 3]) / 1e3
print(f"  walk->block argmin done (before the fence): median {np.median(d):6.2f}  p90 {np.quantile(d, 0.9):6.2f}  max {d.max():6.2f} us")
d = (t[:, 4] - t[:, 7]) / 1e3
print(f"  fence + ticket + barrier:                   median {np.median(d):6.2f}  p90 {np.quantile(d, 0.9):6.2f}  max {d.max():6.2f} us")
for a, b in ((0, 1), (1, 2), (2, 3), (3, 4)):
    d = (t[:, b] - t[:, a]) / 1e3
    print(f"  {names[a]}->{names[b]}: median {np.median(d):6.1f}  p90 {np.quantile(d, 0.9):6.1f}  max {d.max():6.1f} us")
