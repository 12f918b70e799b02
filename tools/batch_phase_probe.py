"""Where a lock-step cycle of the batched replay spends its time (needs a profiling build, see phase_probe.py):
    python tools/batch_phase_probe.py [n_scenarios] [warm cycles] [thread|group|warp]
Stamps per block (globaltimer, ns): 0 entry, 1 tables staged, 2 lists built, 3 walk done, 4 select done; in the block that
reduces a scenario: 5 winner known, 6 exported, 4 hand-off done."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, scenario as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kern = sys.argv[3] if len(sys.argv) > 3 else None
packed = [B.pack_scenario(S.synthetic_intersection(i), max_cycles=warm + 2) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
cfg = J.JsspConfig(max_seeds=1024, max_knots=max(len(p.knot_s) for p in packed), max_obstacle_modes=max(1, max(p.state.shape[0] for p in packed)),
                   max_total_steps=max(p.state.shape[2] for p in packed))
bt = B.BatchReplay(cfg, max_scenarios=len(packed), max_cycles=warm + 2)
ts = torch.zeros((len(packed) * 4096, 16), dtype=torch.int64, device="cuda")
fn = bt.lib.jssp_batch_debug_phase_buffer; fn.restype = C.c_int32; fn.argtypes = [C.c_void_p, C.c_void_p]
assert fn(bt._h, ts.data_ptr()) == 0
bt.load(packed)
bt.run(warm, kernel=kern); torch.cuda.synchronize(); ts.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); bt.run(1, kernel=kern); e1.record(); torch.cuda.synchronize()
t = ts.cpu().numpy().astype(np.float64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
work = t[t[:, 1] > 0]
print(f"{len(packed)} scenarios, kernel={kern}: cycle (seed + evaluate) {1e3 * e0.elapsed_time(e1):.1f} us; {len(t)} blocks, {len(work)} with candidates; "
      f"evaluate span {(t[:, 4].max() - t0) / 1e3:.1f} us")
names = ["entry", "staged", "lists", "walk", "select"]
for k in range(5):
    col = (work[:, k] - t0) / 1e3
    print(f"  {names[k]:7s} stamp: min {col.min():7.1f}  median {np.median(col):7.1f}  p90 {np.quantile(col, 0.9):7.1f}  max {col.max():7.1f} us")
for a, b in ((0, 1), (1, 2), (2, 3), (3, 4)):
    d = (work[:, b] - work[:, a]) / 1e3
    print(f"  {names[a]}->{names[b]}: median {np.median(d):6.1f}  p90 {np.quantile(d, 0.9):6.1f}  max {d.max():6.1f} us")
last = t[t[:, 5] > 0]
if len(last):
    for nm, d in (("walk->winner known", last[:, 5] - last[:, 3]), ("export", last[:, 6] - last[:, 5]), ("hand-off", last[:, 4] - last[:, 6])):
        print(f"  reducing blocks ({len(last)}): {nm}: median {np.median(d) / 1e3:5.1f}  max {d.max() / 1e3:5.1f} us")
    print(f"  reducing blocks: winner known at median {np.median(last[:, 5] - t0) / 1e3:.1f}, max {(last[:, 5].max() - t0) / 1e3:.1f} us")
if t.shape[1] >= 16 and (work[:, 8] > 0).any():      # tiled kernels: progress of the block's first warp (thread 0)
    w = work[work[:, 8] > 0]
    for nm, a, b in (("lists -> tile 0 done", 2, 8), ("tile 0 -> tile 1", 8, 9), ("  tile 1: phase 1", 8, 14), ("  tile 1: cost sums", 14, 15), ("  tile 1: phase 2", 15, 9), ("tile 1 -> middle tile", 9, 10), ("middle -> last tile", 10, 11),
                     ("last tile -> walk returned", 11, 12), ("walk returned -> cost done", 12, 13), ("cost done -> block barrier passed", 13, 3)):
        d = (w[:, b] - w[:, a]) / 1e3
        print(f"  first warp: {nm}: median {np.median(d):6.1f}  p90 {np.quantile(d, 0.9):6.1f}  max {d.max():6.1f} us")
