"""Per-stage time of the persistent replay kernel (tools/, GPU, profiling build):
    JSSP_NVCC_EXTRA=-DJSSP_PROFILE_PHASES python -m dlii_jssp_b200.build && JSSP_NVCC_EXTRA=-DJSSP_PROFILE_PHASES python tools/replay_phase_probe.py [n] [kernel] [first]
(or JSSP_B200_LIB=<profiling .so>).  The CTA of slot 0 stamps %globaltimer between the stages of every plan cycle; `first` = the scenario
id loaded into slot 0 (default 0), `kernel` = persistent | bestfirst."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
packed = [B.pack_scenario(J.scenario.synthetic_intersection(i), max_cycles=100) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "persistent"
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
packed = [packed[first]] + packed[:first] + packed[first + 1:]
e = packed[0].ego
cfg = J.JsspConfig(ego_length=float(e[5]), ego_width=float(e[6]), max_seeds=1024, max_knots=max(64, max(len(p.knot_s) for p in packed)),
                   max_obstacle_modes=max(1, max(p.state.shape[0] * p.state.shape[1] for p in packed)), max_total_steps=max(p.state.shape[2] for p in packed))
bt = B.BatchReplay(cfg, max_scenarios=len(packed), max_cycles=100)
bt.load(packed)
for _ in range(2):
    bt.reset(); bt.run(100, kernel=kernel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); bt.reset(); bt.run(100, kernel=kernel); e1.record(); torch.cuda.synchronize()
ns = (C.c_ulonglong * 8)()
bt.lib.jssp_debug_replay_phases.argtypes = [C.c_void_p]
assert bt.lib.jssp_debug_replay_phases(ns) == 0
cyc = max(int(ns[5]), 1)
names = ["slot load", "seed enumeration", "staging", "tiled evaluation", "select + export + hand-off"]
print(f"kernel {kernel}, slot 0 = scenario {first}")
print(f"{len(packed)} scenarios, sweep {e0.elapsed_time(e1):.3f} ms; scenario 0: {cyc} cycles, {ns[6] / cyc:.2f} passes and {ns[7] / cyc:.0f} seeds per cycle")
for k, nm in enumerate(names):
    print(f"  {nm:28s} {ns[k] / cyc / 1e3:8.2f} us per cycle")
print(f"  {'total':28s} {sum(ns[:5]) / cyc / 1e3:8.2f} us per cycle   ({sum(ns[:5]) / 1e6:.2f} ms over the scenario)")
if kernel == "bestfirst":
    bf = (C.c_ulonglong * 8)()
    bt.lib.jssp_debug_bestfirst.argtypes = [C.c_void_p]
    assert bt.lib.jssp_debug_bestfirst(bf) == 0
    print(f"  best-first: keys {bf[0] / cyc / 1e3:.2f} us, {bf[1] / cyc:.2f} rounds and {bf[2] / cyc:.2f} checks of warp 0 per cycle")
    nc = max(int(bf[2]), 1)
    print(f"  per check of warp 0: profile {bf[4] / nc / 1e3:.2f} us, samples {bf[5] / nc / 1e3:.2f} us, segments {bf[6] / nc / 1e3:.2f} us, exact cost {bf[7] / nc / 1e3:.2f} us")

# stage stamps of the LAST seed enumeration scenario 0 ran (SEED_STAMP in jssp_seed.cuh: 0 entry, 1 lists cleared, 2 level 1, 3 level 2, 4 level 3, 5 sorted, 6 barrier, 7 decoded)
st = (C.c_ulonglong * 16)()
bt.lib.jssp_debug_seed_stamps.argtypes = [C.c_void_p]
if bt.lib.jssp_debug_seed_stamps(st) == 0:
    t = [int(v) for v in st[:8]]
    names = ["clear", "level 1", "level 2", "level 3", "sort", "barrier", "decode"]
    print("  last seed enumeration of scenario 0: " + ", ".join(f"{nm} {(t[k + 1] - t[k]) / 1e3:.2f} us" for k, nm in enumerate(names)))
