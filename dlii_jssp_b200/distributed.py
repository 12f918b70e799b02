"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the plumbing).

Two natural shardings exist and nothing else:

* independent scenarios (BASELINE config 4): scenario ids are dealt to the ranks, every rank replays its own shard, no
  data-path collective - only a host-side gather of the per-scenario result rows;
* one huge cycle (BASELINE config 5): contiguous candidate ranges per rank, obstacles / reference line replicated, and
  ONE exchange after the evaluation: the per-rank winners ``(orderable float64 cost bits, global candidate index)``.
  The exact reduction is an all-gather of 16 bytes per rank followed by a local lexicographic minimum, because a float64
  cost and a 32-bit index do not fit one 64-bit key; ``min_reduce_packed`` is the single u64 min-reduce the north star
  names, exact whenever the fp32-rounded costs of the rank winners differ (it falls back to the lowest index otherwise).

The candidate-split exchange itself runs on the device (``CandidateSplit``): either the evaluation kernel's last block pushes
the rank's record into every peer's mailbox over NVLink and reduces the world's records in the same launch, or one
``ncclAllGather`` of the records is enqueued behind the kernel (``jssp_allreduce_argmin``) - no ``.cpu()``, no Python loop inside
a timed step.  ``allgather_argmin`` / ``min_reduce_packed`` are the host-side forms of the same reduction; they work on CPU
tensors with the gloo backend, which is how the tests cover world_size 2 without GPUs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

_SIGN = -(1 << 63)   # int64 bit pattern 0x8000...: x ^ _SIGN maps unsigned order onto signed order
NO_CANDIDATE = -1    # int64 view of UINT64_MAX written by jssp_best_record_dev when nothing survived


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of ``n_items`` for ``rank``; the first ``n_items % world`` ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_scenarios(ids: Sequence, world: int, rank: int) -> list:
    """Round-robin deal of scenario ids (balances long and short scenarios better than blocks)."""
    return list(ids)[rank::world]


def global_candidate_index(local_idx: int, seed_lo: int, n_seeds_local: int, n_seeds_global: int) -> int:
    """Width-major candidate index of the full cycle from the index inside a rank's seed range
    (candidate n = w * n_seeds + k, jerk_space_sampling_planner.py:106,124-125)."""
    if local_idx < 0:
        return -1
    w, k = divmod(int(local_idx), n_seeds_local)
    return w * n_seeds_global + seed_lo + k


def localize_record(rec: torch.Tensor, seed_lo: int, n_seeds_local: int, n_seeds_global: int) -> torch.Tensor:
    """Rewrite word 1 of a best record from the local to the global candidate index (on the record's device)."""
    out = rec.clone()
    idx = rec[1]
    w = torch.div(idx, n_seeds_local, rounding_mode="floor")
    g = w * n_seeds_global + seed_lo + (idx - w * n_seeds_local)
    out[1] = torch.where(idx < 0, idx, g)
    return out


def gather_bytes(payload: bytes, device=None, group=None) -> List[bytes]:
    """All-gather one fixed-size byte string per rank (CUDA IPC handles, ...) through torch.distributed: a uint8 tensor on
    ``device`` (cuda for the nccl backend, cpu for gloo)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [payload]
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [bytes(t.cpu().tolist()) for t in out]


def broadcast_bytes(payload: Optional[bytes], n: int, src: int = 0, device=None, group=None) -> bytes:
    """Broadcast ``n`` bytes from rank ``src`` (the NCCL unique id)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return payload
    t = torch.tensor(list(payload) if dist.get_rank(group) == src else [0] * n, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def sweep_digest(rows) -> str:
    """Order-independent digest of a scenario sweep: rows = (name, end, cycles, best_idx list) per scenario.  Equal for every
    number of ranks iff every scenario was replayed identically (bench.py prints it for every --gpus N)."""
    import hashlib
    h = hashlib.sha256()
    for name, end, cycles, best in sorted((str(r[0]), float(r[1]), int(r[2]), tuple(int(b) for b in r[3])) for r in rows):
        h.update(repr((name, end, cycles, best)).encode())
    return h.hexdigest()[:16]


def _loaded_nccl_path() -> Optional[str]:
    """Path of the libnccl the process has already mapped (torch's bundled one), for the library's dlopen."""
    try:
        for line in open("/proc/self/maps"):
            if "libnccl.so" in line:
                return line.split()[-1]
    except OSError:
        pass
    return None


class CandidateSplit:
    """Device-side exchange of a candidate-split cycle (BASELINE configs[4]) for one JsspEngine: rank r evaluates seeds
    ``shard_range(n_seeds_global, world, r)`` for every lateral target, and the selection over all ranks is taken on the device.

    transport "peer": CUDA-IPC mailboxes over NVLink, exchange fused into the evaluation launch (default when the peers can
    map each other's memory); "nccl": one ncclAllGather of 16 bytes per rank behind the kernel (the library's own communicator
    from a unique id broadcast here).  ``exchange()`` enqueues what the transport needs after an ``evaluate(..., sync=False)``;
    ``fetch()`` waits and returns (global best index, cost, winner rank) - the same triple on every rank."""

    def __init__(self, engine, n_seeds_global: int, transport: str = "auto", group=None):
        import os
        self.eng, self.group = engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_seeds_global = int(n_seeds_global)
        self.lo, self.hi = shard_range(self.n_seeds_global, self.world, self.rank)
        engine.set_split(self.lo, self.n_seeds_global)
        self.comm = None
        self.transport = "local"
        if self.world == 1:
            return
        dev = engine.device
        if transport in ("auto", "peer"):
            ok = 1
            try:
                handles = gather_bytes(engine.peer_export(self.world), device=dev, group=group)
                engine.peer_connect(b"".join(handles), self.world, self.rank)
            except Exception as e:           # a box whose GPUs cannot map each other: fall back to NCCL on EVERY rank
                ok, self.peer_error = 0, repr(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.transport = "peer"
                return
            engine.peer_disconnect()
            if transport == "peer":
                raise RuntimeError("peer mailboxes unavailable: " + getattr(self, "peer_error", "a peer failed"))
        path = _loaded_nccl_path()
        if path and not os.environ.get("JSSP_NCCL_LIB"):
            os.environ["JSSP_NCCL_LIB"] = path
        import ctypes as C
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            engine._check(engine.lib.jssp_nccl_unique_id(uid), "jssp_nccl_unique_id")
        raw = broadcast_bytes(bytes(uid), 128, 0, device=dev, group=group)
        comm = C.c_void_p()
        engine._check(engine.lib.jssp_nccl_comm_create((C.c_uint8 * 128).from_buffer_copy(raw), self.world, self.rank, dev.index, C.byref(comm)),
                      "jssp_nccl_comm_create")
        self.comm, self.transport = comm, "nccl"

    def local_seeds(self, seeds):
        """This rank's rows of the global seed table."""
        return seeds[self.lo:self.hi]

    def exchange(self):
        """Enqueue the exchange behind the evaluation just enqueued on the engine's stream (nothing to do for the peer
        transport: the evaluation launch already contains it)."""
        if self.transport == "peer":
            return
        self.eng.allreduce_argmin(self.comm, self.world)

    def fetch(self):
        return self.eng.fetch_global_best()

    def close(self):
        if self.transport == "peer":
            self.eng.peer_disconnect()
        if self.comm is not None:
            self.eng.lib.jssp_nccl_comm_destroy(self.comm)
            self.comm = None
        self.eng.set_split(0, 0)


def _u64_less(a_cost: int, a_idx: int, b_cost: int, b_idx: int) -> bool:
    """(cost bits, index) lexicographic '<' with both words compared as unsigned 64-bit (index -1 == UINT64_MAX = none)."""
    ua, ub = a_cost & 0xFFFFFFFFFFFFFFFF, b_cost & 0xFFFFFFFFFFFFFFFF
    if ua != ub:
        return ua < ub
    return (a_idx & 0xFFFFFFFFFFFFFFFF) < (b_idx & 0xFFFFFFFFFFFFFFFF)


def allgather_argmin(rec: torch.Tensor, group=None, index_map=None) -> Tuple[int, int, int]:
    """rec = int64[2] (orderable cost bits, candidate index or -1).  One all-gather (16 B per rank), one device-to-host
    copy of the gathered table, then the lexicographic minimum on the host.  ``index_map(rank, idx)`` turns a rank-local
    index into the global one (default: the indices are already global).  Returns (cost_bits, global_index, winner_rank);
    global_index = -1 when no rank has a survivor."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        table = [rec.cpu().tolist()]
    else:
        out = torch.empty((world, 2), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(out, rec, group=group) if rec.is_cuda else dist.all_gather(list(out.unbind(0)), rec, group=group)
        table = out.cpu().tolist()
    best = None
    for r, (c, i) in enumerate(table):
        if i == NO_CANDIDATE:
            continue
        if index_map is not None:
            i = index_map(r, i)
        if best is None or _u64_less(c, i, best[0], best[1]):
            best = (c, i, r)
    return best if best is not None else (table[0][0], -1, 0)


def orderable_bits_to_float(bits: int) -> float:
    """Inverse of the kernels' orderable_bits()."""
    b = bits & 0xFFFFFFFFFFFFFFFF
    b = (b & 0x7FFFFFFFFFFFFFFF) if (b & 0x8000000000000000) else (~b & 0xFFFFFFFFFFFFFFFF)
    return float(np.array([b], dtype=np.uint64).view(np.float64)[0])


def float_to_orderable_bits(v: float) -> int:
    b = int(np.array([v], dtype=np.float64).view(np.uint64)[0])
    u = (~b & 0xFFFFFFFFFFFFFFFF) if (b & 0x8000000000000000) else (b | 0x8000000000000000)
    return u - (1 << 64) if u >= (1 << 63) else u     # as the int64 view torch carries


def pack_key32(cost: float, global_idx: int) -> int:
    """u64 key = orderable_u32(fp32(cost)) << 32 | index, returned as a python int in [0, 2^64)."""
    if global_idx < 0:
        return 0xFFFFFFFFFFFFFFFF
    b = int(np.array([cost], dtype=np.float32).view(np.uint32)[0])
    o = (~b & 0xFFFFFFFF) if (b & 0x80000000) else (b | 0x80000000)
    return (o << 32) | (global_idx & 0xFFFFFFFF)


def min_reduce_packed(cost: float, global_idx: int, device=None, group=None) -> Tuple[Optional[float], int]:
    """The single 8-byte min-reduce: all_reduce(MIN) of one u64 key per rank (carried as sign-flipped int64).
    Returns (fp32-rounded winning cost, global index) - index -1 when nothing survived anywhere."""
    key = pack_key32(cost, global_idx)
    signed = (key ^ 0x8000000000000000)
    signed = signed - (1 << 64) if signed >= (1 << 63) else signed
    t = torch.tensor([signed], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    k = (int(t.item()) & 0xFFFFFFFFFFFFFFFF) ^ 0x8000000000000000
    if k == 0xFFFFFFFFFFFFFFFF:
        return None, -1
    o = k >> 32
    b = (o & 0x7FFFFFFF) if (o & 0x80000000) else (~o & 0xFFFFFFFF)
    return float(np.array([b], dtype=np.uint32).view(np.float32)[0]), int(k & 0xFFFFFFFF)


def gather_results(rows: list, group=None) -> List[list]:
    """Host-side gather of per-scenario result rows (scenario sweep); returns the per-rank lists on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [rows]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, rows, group=group)
    return out
