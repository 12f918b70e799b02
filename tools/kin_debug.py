import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, replay as R, scenario as S
sc = S.scenario_from_fixture(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenario_8_3_4_565.npz"))
got = B.run_batch([sc], ego_update_mode="kinematics")[0]
solo = R.run_replay(sc, ego_update_mode="kinematics")
a, b = np.array(got.ego_rows), np.array(solo.ego_rows)
n = min(len(a), len(b))
print("cycles", got.cycles, solo.cycles, "ends", got.end, solo.end)
for i in range(n):
    d = np.abs(a[i] - b[i]).max()
    flag = "" if got.best_idx[i] == solo.best_idx[i] else f"  <-- best {got.best_idx[i]} vs {solo.best_idx[i]}"
    if d > 0 or flag:
        print(i, f"max|diff| {d:.3e}", np.array2string(a[i] - b[i], precision=2), got.n_candidates[i], solo.n_candidates[i], flag)
    if flag:
        break
