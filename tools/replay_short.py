import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlii_jssp_b200 as J
sc = J.scenario.scenario_from_fixture(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenario_8_3_4_565.npz"))
r = J.replay.run_replay(sc, max_cycles=int(sys.argv[1]) if len(sys.argv) > 1 else 12)
print(r.cycles, r.best_idx, r.p50_ms)
