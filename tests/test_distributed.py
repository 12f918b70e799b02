"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: scenario sharding and the candidate-split reduction.
The per-rank "evaluation" is done by the oracle here - only the partitioning / reduction code of the product is under test."""
import os
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dlii_jssp_b200 import distributed as D

warnings.filterwarnings("ignore", category=RuntimeWarning)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import jssp_oracle as O
    from dlii_jssp_b200 import scenario as S, distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cyc = S.dense_cycle(n_seeds=96, n_widths=3, O=8, M=2, seed=17)
        cfg = O.OracleConfig(ego_length=cyc.ego_length, ego_width=cyc.ego_width)
        spline = O.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1])
        o = cyc.obstacles
        obs = (o.state, o.valid.astype(bool), o.size, o.prob)
        lo, hi = D.shard_range(len(cyc.seeds), world, rank)
        loc = O.plan_cycle(cfg, spline, cyc.frenet0, obs, 0, cyc.final_time_step, seeds=cyc.seeds[lo:hi], targets=cyc.targets)
        cost = float(loc["cost"][loc["best"]]) if loc["best"] >= 0 else float("inf")
        rec = torch.tensor([D.float_to_orderable_bits(cost), loc["best"]], dtype=torch.int64)
        rec = D.localize_record(rec, lo, hi - lo, len(cyc.seeds))
        bits, gidx, winner = D.allgather_argmin(rec)
        c32, gidx32 = D.min_reduce_packed(cost, D.global_candidate_index(loc["best"], lo, hi - lo, len(cyc.seeds)))
        rows = D.gather_results([(rank, s) for s in D.shard_scenarios(range(7), world, rank)])
        # the plumbing of the device-side exchange: per-rank byte strings (CUDA IPC handles) and the NCCL unique id
        handles = D.gather_bytes(bytes([rank]) * 64)
        uid = D.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128, 0)
        q.put((rank, D.orderable_bits_to_float(bits), gidx, winner, c32, gidx32, rows, handles, uid))
    finally:
        dist.destroy_process_group()


def test_candidate_split_and_scenario_shards_world2():
    from oracle import jssp_oracle as O
    from dlii_jssp_b200 import scenario as S
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    out = [q.get(timeout=180) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    cyc = S.dense_cycle(n_seeds=96, n_widths=3, O=8, M=2, seed=17)
    cfg = O.OracleConfig(ego_length=cyc.ego_length, ego_width=cyc.ego_width)
    spline = O.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1])
    o = cyc.obstacles
    full = O.plan_cycle(cfg, spline, cyc.frenet0, (o.state, o.valid.astype(bool), o.size, o.prob), 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
    assert full["best"] >= 0
    for rank, cost, gidx, winner, c32, gidx32, rows, handles, uid in out:
        assert handles == [bytes([r]) * 64 for r in range(world)] and uid == bytes(range(128))
        assert gidx == full["best"], (gidx, full["best"])                  # candidate-split argmin == single-GPU argmin
        assert cost == full["cost"][full["best"]]                            # float64 cost survives the record exactly
        lo, hi = D.shard_range(len(cyc.seeds), world, winner)
        assert lo <= full["best"] % len(cyc.seeds) < hi
        assert gidx32 == full["best"] and c32 == np.float32(full["cost"][full["best"]])
        flat = sorted(s for part in rows for (_, s) in part)
        assert flat == list(range(7))                                      # every scenario replayed exactly once


def test_shard_helpers_single_process():
    assert [D.shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sum((b - a) for a, b in (D.shard_range(1048576, 8, r) for r in range(8))) == 1048576
    assert D.shard_scenarios(range(10), 4, 1) == [1, 5, 9]
    assert D.global_candidate_index(5 * 100 + 7, 300, 100, 800) == 5 * 800 + 300 + 7 and D.global_candidate_index(-1, 0, 1, 1) == -1
    vals = [-3.5, -0.0, 0.0, 1e-300, 2.25, 7.0, float("inf")]
    bits = [D.float_to_orderable_bits(v) & 0xFFFFFFFFFFFFFFFF for v in vals]
    assert bits == sorted(bits) and [D.orderable_bits_to_float(b) for b in bits][2:] == vals[2:]
    # unsigned lexicographic order of (cost bits, index); -1 (= UINT64_MAX) loses against any index
    assert D._u64_less(D.float_to_orderable_bits(1.0), 9, D.float_to_orderable_bits(1.0), 10)
    assert D._u64_less(D.float_to_orderable_bits(-1.0), 99, D.float_to_orderable_bits(1.0), 0)
    assert D.pack_key32(1.0, 3) < D.pack_key32(1.0, 4) < D.pack_key32(1.0000002, 0) < D.pack_key32(2.0, 0) < D.pack_key32(0.0, -1)
    assert D.min_reduce_packed(4.25, 12) == (4.25, 12) and D.min_reduce_packed(0.0, -1) == (None, -1)
    assert D.allgather_argmin(torch.tensor([D.float_to_orderable_bits(2.5), 41], dtype=torch.int64)) == (D.float_to_orderable_bits(2.5), 41, 0)
    # sweep digest: independent of the order / sharding of the rows, sensitive to any selected index
    rows = [("a", 1.5, 3, [1, 2, 3]), ("b", -1.0, 2, [0, 1]), ("c", 2.0, 0, [])]
    assert D.sweep_digest(rows) == D.sweep_digest(rows[::-1]) == D.sweep_digest(rows[1:] + rows[:1])
    assert D.sweep_digest(rows) != D.sweep_digest([("a", 1.5, 3, [1, 2, 4])] + rows[1:])
    assert D.gather_bytes(b"xy") == [b"xy"] and D.broadcast_bytes(b"id", 2) == b"id"
