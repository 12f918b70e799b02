import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import dlii_jssp_b200 as J
scs=[J.scenario.synthetic_intersection(i) for i in range(40)]
J.replay.run_sweep(scs[:2], max_cycles=5)
t0=time.perf_counter(); res=J.replay.run_sweep(scs, max_cycles=100); dt=time.perf_counter()-t0
lat=np.array([x for r in res for x in r.plan_seconds]); cyc=sum(r.cycles for r in res)
print('total',dt,'cycles',cyc,'sum plan',lat.sum(),'p50',np.median(lat)*1e3,'mean',lat.mean()*1e3,'p99',np.quantile(lat,0.99)*1e3,'max',lat.max()*1e3)
firsts=np.array([r.plan_seconds[0] for r in res]); print('first-cycle mean ms', firsts.mean()*1e3)
import cProfile, pstats
pr=cProfile.Profile(); pr.enable(); J.replay.run_sweep(scs[:10], max_cycles=100); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
