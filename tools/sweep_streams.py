"""The batched sweep split over k sub-batches on k CUDA streams (seed enumeration of one overlaps the evaluation of another):
    python tools/sweep_streams.py [n_scenarios] [cycles] [k ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, scenario as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
cyc = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ks = [int(v) for v in sys.argv[3:]] or [1, 2, 4]
packed = [B.pack_scenario(S.synthetic_intersection(i), max_cycles=cyc) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
cfg = J.JsspConfig(max_seeds=1024, max_knots=max(len(p.knot_s) for p in packed), max_obstacle_modes=max(1, max(p.state.shape[0] for p in packed)),
                   max_total_steps=max(p.state.shape[2] for p in packed))
for k in ks:
    parts = [packed[i::k] for i in range(k)]
    streams = [torch.cuda.Stream() for _ in range(k)]
    bts = []
    for part, st in zip(parts, streams):
        with torch.cuda.stream(st):
            bt = B.BatchReplay(cfg, max_scenarios=len(part), max_cycles=cyc)
            bt.load(part)
            bts.append(bt)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        for bt, st in zip(bts, streams):
            with torch.cuda.stream(st):
                bt.reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        for bt, st in zip(bts, streams):
            with torch.cuda.stream(st):
                bt.run(cyc)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    sig = 0
    for bt, st in zip(bts, streams):
        with torch.cuda.stream(st):
            sig ^= hash(tuple(tuple(r.best_idx) for r in bt.results())) & 0xffffffff
    print(f"n={len(packed)} sub-batches={k}: {best:.2f} ms ({1e3 * best / cyc:.1f} us per lock-step cycle of the whole batch) sig {sig:08x}", flush=True)
    for bt in bts:
        bt.close()
