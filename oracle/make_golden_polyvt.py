#!/usr/bin/env python
"""Generate tests/golden/polyvt_golden.npz by RUNNING the reference's own sampler.

TEST INFRASTRUCTURE - runs only in the build container, where /root/reference is
mounted; nothing on the GPU box executes this file.  It imports the *unmodified*
reference module

    /root/reference/code/intersection_local_planner/polyvt_sampling.py

with ``sys.modules`` stubs for the modules that are absent from the published
tree (``matplotlib*`` - import only; ``common.scenario.frenet`` - a trajectory
container with list fields) and ``tqdm`` replaced by
the identity.  It then records, for the start states and jerk limits listed in
SURVEY.md section 8c:

  * every main seed    (PolyVTSampling.sampling,                        :167-263)
  * every stop seed    (PolyVTSampling.sampling_stop_seeds_supplement,  :265-305)
  * the unrolled s, s_d, s_dd, s_ddd of a deterministic subsample of seeds
    (PolyVTSampling.get_one_frenet_trajectory,                        :311-385)

These arrays are the golden vectors that pin oracle/jssp_oracle.py (and through
it the CUDA seed-enumeration and unroll kernels) to executable reference code.

Usage:  python oracle/make_golden_polyvt.py  [out.npz]
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/code"
CASES = [(5.686, 0.0), (0.0, 0.0), (10.0, 1.0), (17.0, 0.0), (3.0, -2.0)]
JERKS = [10.0, 13.4112]
# horizon 8 s exercises the total_time-dependent bounds used by BASELINE config 5
EXTRA = [dict(v0=6.0, a0=0.5, J=10.0, total_time=8.0, num_samples=13)]


def _install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.collections"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.gridspec"].GridSpec = object
    sys.modules["matplotlib.collections"].LineCollection = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    tq = types.ModuleType("tqdm")
    tq.tqdm = lambda it, *a, **k: it
    sys.modules["tqdm"] = tq

    class JerkSpaceSamplingFrenetTrajectory:  # list-field container
        def __init__(self):
            for f in ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c", "c_d", "c_dd"):
                setattr(self, f, [])
            self.path_id = 0

    for name in ("common", "common.scenario", "common.scenario.frenet"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["common.scenario.frenet"].JerkSpaceSamplingFrenetTrajectory = JerkSpaceSamplingFrenetTrajectory
    return JerkSpaceSamplingFrenetTrajectory


def main(out):
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not mounted; golden vectors can only be regenerated in the build container")
    Traj = _install_stubs()
    sys.path.insert(0, os.path.join(REF, "intersection_local_planner"))
    sys.path.insert(0, REF)
    import polyvt_sampling as ref  # the unmodified reference module

    gold = {}
    runs = [dict(v0=v0, a0=a0, J=J, total_time=5.0, num_samples=25) for J in JERKS for (v0, a0) in CASES] + EXTRA
    meta = []
    for k, r in enumerate(runs):
        s0 = 3.25 if k % 2 else 0.0  # non-zero s0 must not change the seeds
        smp = ref.PolyVTSampling(total_time=r["total_time"], delta_t=0.1, jerk_min=-r["J"], jerk_max=r["J"],
                                 a_min=-8.0, a_max=8.0, v_max=18.0, num_samples=r["num_samples"])
        main_seeds = np.array(smp.sampling(s0, r["v0"], r["a0"]), dtype=np.float64).reshape(-1, 10)
        stop_seeds = np.array(smp.sampling_stop_seeds_supplement(s0, r["v0"], r["a0"]), dtype=np.float64).reshape(-1, 10)
        seeds = np.concatenate([main_seeds, stop_seeds], 0)
        rng = np.random.default_rng(1000 + k)
        pick = np.sort(rng.choice(len(seeds), size=min(24, len(seeds)), replace=False))
        unr = np.zeros((len(pick), 4, len(smp.time_array)))
        for q, idx in enumerate(pick):
            ft = smp.get_one_frenet_trajectory(s0=s0, v0=r["v0"], a0=r["a0"], seed=tuple(seeds[idx]), jssft=Traj())
            unr[q, 0], unr[q, 1], unr[q, 2], unr[q, 3] = ft.s, ft.s_d, ft.s_dd, ft.s_ddd
            assert ft.t == smp.time_array
        gold[f"c{k}_main"] = main_seeds
        gold[f"c{k}_stop"] = stop_seeds
        gold[f"c{k}_pick"] = pick
        gold[f"c{k}_unroll"] = unr
        gold[f"c{k}_time"] = np.array(smp.time_array)
        meta.append([s0, r["v0"], r["a0"], r["J"], r["total_time"], r["num_samples"]])
        print(f"case {k}: v0={r['v0']} a0={r['a0']} J={r['J']} T={r['total_time']} -> main {len(main_seeds)} stop {len(stop_seeds)}")
    gold["meta"] = np.array(meta)
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "tests", "golden", "polyvt_golden.npz"))
