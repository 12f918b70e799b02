// jssp_common.cuh - launch constants, selection record, phase stamps, mbarrier / TMA bulk-copy helpers (sm_100a)
#pragma once
#include "jssp_device.cuh"
#include "../../include/jssp_b200.h"
#include <cooperative_groups.h>

namespace jssp {
namespace cg = cooperative_groups;

#ifndef JSSP_EVAL_THREADS
#define JSSP_EVAL_THREADS 256
#endif
#ifndef JSSP_EVAL_MIN_BLOCKS
#define JSSP_EVAL_MIN_BLOCKS 2
#endif
#ifndef JSSP_BATCH_THREADS
#define JSSP_BATCH_THREADS 128
#endif
#ifndef JSSP_BATCH_MIN_BLOCKS
#define JSSP_BATCH_MIN_BLOCKS 4
#endif
#ifndef JSSP_CHUNK_MIN_BLOCKS
#define JSSP_CHUNK_MIN_BLOCKS 5
#endif
constexpr int EVAL_THREADS = JSSP_EVAL_THREADS;
constexpr int EVAL_MIN_BLOCKS = JSSP_EVAL_MIN_BLOCKS;
// block size / residency of the thread and group evaluation kernels in the batched replay (a scenario has a few hundred
// candidates: small blocks leave fewer idle lanes in its last block and spread its work over more SMs)
constexpr int BATCH_EVAL_THREADS = JSSP_BATCH_THREADS;
constexpr int BATCH_MIN_BLOCKS = JSSP_BATCH_MIN_BLOCKS;
// the chunked tiled kernel doubles (or triples) the lanes of a plan cycle: 96 registers keep 5 of its blocks on an SM
constexpr int CHUNK_MIN_BLOCKS = JSSP_CHUNK_MIN_BLOCKS;
constexpr float SAT_EPS = 2e-3f;        // fp32 guard band [m]; inside it the pair is re-decided in float64
constexpr float FAR_AWAY = 1e30f;       // x of an absent obstacle state: squared distance overflows to +inf -> culled

#ifdef JSSP_PROFILE_PHASES
__device__ __forceinline__ void phase_stamp(unsigned long long* ts, int slot) {
  if (ts && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ts[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + slot] = t;
  }
}
#define PHASE(p, k) phase_stamp((p).phase_ts, k)
#else
#define PHASE(p, k)
#endif

// Programmatic dependent launch (the lock-step kernels of the batched replay are chained with it): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may be scheduled while its predecessor in the stream drains;
// pdl_wait() returns once the predecessor has completed and its writes are visible (a no-op for ordinary launches),
// pdl_launch_dependents() tells the scheduler that the successor may start being placed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct BestRecord {
  unsigned long long key;   // orderable_bits(cost)
  long long idx;            // candidate index (+ offset for candidate-split runs), -1 = none
  double cost;
  int n_pts;
  int n_candidates;
  int n_seeds;
  int pad;
};

// ------------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

}  // namespace jssp
