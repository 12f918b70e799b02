// jssp_split.cuh - candidate-split exchange: the argmin over the candidates of ALL ranks (jerk_space_sampling_planner.py:288-294,
// 326-327 select over every candidate).  Each rank evaluates a contiguous seed range; its selection
// record is (orderable bits of the float64 cost, GLOBAL candidate index) - 16 bytes, lexicographic minimum = lowest cost, then
// lowest global index, i.e. exactly the 1-GPU selection.  Two transports:
//   * NCCL:  ncclAllGather of the 16-byte records on the caller's stream + split_reduce_kernel (one warp)
//   * peer mailboxes over NVLink (jssp_peer_*): the block that finishes the evaluation pushes the record into every peer's
//     mailbox (plain stores through peer-mapped memory, release fence, epoch flag) and one warp of the SAME kernel launch waits
//     for the world's records and reduces them: compute and exchange are one launch, no collective kernel, no host step.
#pragma once
#include "jssp_common.cuh"

namespace jssp {

struct SplitRecord {            // what travels between ranks
  unsigned long long key;       // orderable_bits(cost); UINT64_MAX = the rank has no survivor
  unsigned long long idx;       // global candidate index; UINT64_MAX = none
};

struct GlobalBest {             // result of the exchange, in mapped pinned host memory (and identical on every rank)
  unsigned long long key;
  long long idx;                // global candidate index, -1 = no candidate survived on any rank
  double cost;
  int rank;                     // rank that holds the winner (its exported trajectory is the global selection)
  int world;
  unsigned long long epoch;     // exchange counter (peer transport) / 0
  int timed_out;                // peer transport: a peer's record did not arrive within the spin budget
  int pad;
};

constexpr int SPLIT_MAX_WORLD = 32;

// Mailbox of one rank (device memory, opened by every peer through CUDA IPC): slot [parity][source rank]
struct PeerSlot { unsigned long long key, idx, epoch, pad; };
struct PeerView {
  PeerSlot* box[SPLIT_MAX_WORLD];   // box[r] = rank r's mailbox as mapped into THIS process ([2][world] slots); box[rank] = own
  int world, rank;
  GlobalBest* out;                  // mapped pinned host memory
  long long spin_budget;            // clock64 ticks before giving up on a peer
};

__device__ __forceinline__ bool split_less(unsigned long long ka, unsigned long long ia, unsigned long long kb, unsigned long long ib) {
  return ka < kb || (ka == kb && ia < ib);
}

__device__ __forceinline__ double orderable_to_double(unsigned long long u) {
  const unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)b);
}

// one warp: lexicographic minimum of `world` records held one per lane -> GlobalBest
__device__ __forceinline__ void split_reduce_warp(unsigned long long key, unsigned long long idx, int src_rank, int world, GlobalBest* out,
                                                  unsigned long long epoch, int timed_out) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long k2 = __shfl_down_sync(0xffffffffu, key, o), i2 = __shfl_down_sync(0xffffffffu, idx, o);
    const int r2 = __shfl_down_sync(0xffffffffu, src_rank, o);
    if (split_less(k2, i2, key, idx)) { key = k2; idx = i2; src_rank = r2; }
  }
  timed_out = __any_sync(0xffffffffu, timed_out) ? 1 : 0;
  if ((threadIdx.x & 31) == 0) {
    GlobalBest g;
    const bool none = idx == 0xffffffffffffffffull;
    g.key = key; g.idx = none ? -1 : (long long)idx; g.cost = none ? INFINITY : orderable_to_double(key);
    g.rank = none ? 0 : src_rank; g.world = world; g.epoch = epoch; g.timed_out = timed_out; g.pad = 0;
    *out = g;
  }
}

// NCCL transport, after ncclAllGather: table[world] records -> GlobalBest (one warp)
__global__ void split_reduce_kernel(const SplitRecord* __restrict__ table, int world, GlobalBest* __restrict__ out) {
  const int lane = threadIdx.x;
  unsigned long long key = 0xffffffffffffffffull, idx = 0xffffffffffffffffull;
  int src = lane;
  for (int r = lane; r < world; r += 32) {
    const SplitRecord t = table[r];
    if (split_less(t.key, t.idx, key, idx)) { key = t.key; idx = t.idx; src = r; }
  }
  split_reduce_warp(key, idx, src, world, out, 0ull, 0);
}

// Peer transport, called by ONE warp (threads 0..31) of the block that finished the rank's evaluation, once the rank's own
// record is final.  split_push_peers: lane r < world stores the record into rank r's mailbox (slot [epoch parity][this rank]),
// fences, then publishes the epoch.  split_wait_peers (later in the same launch - the block exports the rank's trajectory in
// between): lane r waits for rank r's record in the own mailbox, then the warp reduces the world's records into GlobalBest.
// Two slots per source (epoch parity): a rank can only write epoch e+2 after it has completed e+1, which needed this rank's
// e+1 record, which this rank writes only after it has read every epoch-e record.
__device__ __forceinline__ void split_push_peers(const PeerView& pv, unsigned long long epoch, unsigned long long key, unsigned long long idx) {
  const int lane = threadIdx.x & 31;
  if (lane < pv.world) {
    PeerSlot* dst = pv.box[lane] + (size_t)(epoch & 1ull) * pv.world + pv.rank;
    *reinterpret_cast<volatile unsigned long long*>(&dst->key) = key;
    *reinterpret_cast<volatile unsigned long long*>(&dst->idx) = idx;
    __threadfence_system();                                            // the record is visible system-wide before its flag
    *reinterpret_cast<volatile unsigned long long*>(&dst->epoch) = epoch;
  }
}
__device__ __forceinline__ void split_wait_peers(const PeerView& pv, unsigned long long epoch) {
  const int lane = threadIdx.x & 31;
  unsigned long long rk = 0xffffffffffffffffull, ri = 0xffffffffffffffffull;
  int to = 0;
  if (lane < pv.world) {
    const PeerSlot* src = pv.box[pv.rank] + (size_t)(epoch & 1ull) * pv.world + lane;
    const long long t0 = clock64();
    while (*reinterpret_cast<const volatile unsigned long long*>(&src->epoch) != epoch) {
      if (clock64() - t0 > pv.spin_budget) { to = 1; break; }
      __nanosleep(32);
    }
    __threadfence_system();
    if (!to) {
      rk = *reinterpret_cast<const volatile unsigned long long*>(&src->key);
      ri = *reinterpret_cast<const volatile unsigned long long*>(&src->idx);
    }
  }
  split_reduce_warp(rk, ri, lane, pv.world, pv.out, epoch, to);
}

}  // namespace jssp
