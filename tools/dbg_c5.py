import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import dlii_jssp_b200 as J
import bench
cyc = bench.workload("C5")
O_, M, Tt = cyc.obstacles.shape
cfg = J.JsspConfig(plan_horizon=cyc.plan_horizon, ego_length=cyc.ego_length, ego_width=cyc.ego_width, max_seeds=len(cyc.seeds), max_obstacle_modes=O_*M, max_total_steps=Tt)
eng = J.JsspEngine(cfg)
sp = J.CubicSpline2D(cyc.route_xy[:,0], cyc.route_xy[:,1]); eng.set_refline(sp.s_list, sp.coefficient_table())
o=cyc.obstacles; eng.set_obstacles(o.state,o.valid,o.size,o.prob,o.n_dict)
seeds=torch.from_numpy(cyc.seeds).cuda()
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
for mode in ['noflush','flush','noflush','flush']:
    ts=[]
    for i in range(6):
        if mode=='flush': flush.fill_(1)
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, sync=False); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(mode, [round(t,3) for t in ts])
r=eng.fetch(); print(r.n_candidates, r.best_idx, r.best_cost, int(r.collision.sum()), int(r.feasible.sum()))
