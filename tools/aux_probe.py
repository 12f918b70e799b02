"""CUDA-event timing of the secondary kernels: IMM update, per-mode roll-out, batched Cartesian->Frenet projection.
    python tools/aux_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J

def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n

rng = np.random.default_rng(0)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
eng = J.JsspEngine(J.JsspConfig(max_knots=256))
for Ob in (32, 512, 8192, 131072):
    M = 3
    mean, cov, mu = d(rng.normal(0, 0.5, (Ob, M, 2))), d(np.tile(np.array([0.5, 0.0, 0.5]), (Ob, M, 1))), d(np.full((Ob, M), 1.0 / M))
    z = d(rng.normal(0, 1, (Ob, M)))
    us = timed(lambda: eng.imm_update(mean, cov, mu, z))
    byts = Ob * M * (2 + 3 + 1) * 8 * 2 + Ob * M * 8
    print(f"imm_update O={Ob} M={M}: {us:.1f} us per call ({byts / us / 1e3:.1f} GB/s of state traffic)")
lanes = np.stack([np.stack([np.linspace(-40, 40, 81), np.full(81, yv)], 1) for yv in (-16.5, -13.0, -9.5)] +
                 [np.stack([np.full(81, xv), np.linspace(-40, 40, 81)], 1) for xv in (11.5, 18.5)])
for Ob in (32, 512, 8192):
    M, Tt = 3, 152
    lane_of = d(np.stack([rng.permutation(5)[:M] for _ in range(Ob)]).astype(np.int32))
    s_start, speed = d(rng.uniform(0, 40, (Ob, M))), d(rng.uniform(1, 9, Ob))
    ld = d(lanes)
    us = timed(lambda: eng.predict_modes(ld, lane_of, s_start, speed, 0.1, Tt), 20)
    print(f"predict_modes O={Ob} M={M} Tt={Tt}: {us:.1f} us per call ({Ob * M * Tt * 25 / us / 1e3:.1f} GB/s written)")
sc = J.scenario.scenario_from_fixture(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenario_8_3_4_565.npz"))
sp = J.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1]); eng.set_refline(sp.s_list, sp.coefficient_table())
for n in (1, 100, 800, 8192):
    st = d(np.column_stack([rng.uniform(10, 55, n), rng.uniform(15, 40, n), rng.uniform(-3, 3, n), rng.uniform(0, 10, n), rng.normal(0, 1, n)]))
    us = timed(lambda: eng.project_states(st), 20)
    print(f"project_states n={n} (reference line {sp.s_list[-1]:.0f} m at 0.1 m): {us:.1f} us per call")
