// jssp_kernels.cuh - sm_100a kernels of the JSSP candidate evaluation (included by jssp_b200.cu only).
//
//   K1  seed_flag_main / seed_flag_stop / seed_sort_emit   polyvt_sampling.py:167-305
//   K2  evaluate_kernel (unroll -> Frenet->Cartesian -> heading/curvature -> feasibility -> rectangle collision
//       against every obstacle mode -> cost -> block argmin -> last-block global argmin)
//                                                           polyvt_sampling.py:311-385, jerk_space_sampling_planner.py:131-294,326-327
//   K3  export_best_kernel / dump_states_kernel            consumers listed in SURVEY.md section 8 a12
//   K0  pack_obstacles_kernel                               obstacle dict -> device tables
//   K4  imm_update_kernel / predict_modes_kernel            builder-specified (no reference code)
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
constexpr int EVAL_THREADS = JSSP_EVAL_THREADS;
constexpr int EVAL_MIN_BLOCKS = JSSP_EVAL_MIN_BLOCKS;
constexpr float SAT_EPS = 2e-3f;        // fp32 guard band [m]; inside it the pair is re-decided in float64
constexpr float FAR_AWAY = 1e30f;       // x of an absent obstacle state: squared distance overflows to +inf -> culled

#ifdef JSSP_PROFILE_PHASES
__device__ __forceinline__ void phase_stamp(unsigned long long* ts, int slot) {
  if (ts && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); ts[(size_t)blockIdx.x * 8 + slot] = t; }
}
#define PHASE(p, k) phase_stamp((p).phase_ts, k)
#else
#define PHASE(p, k)
#endif

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

// ------------------------------------------------------------------------------------------------------------------
// K1: seed enumeration.  One thread per point of the flat grid; survivors append their flat index, a single block
// sorts the (few) indices so that the emitted order is the reference's nested-loop order, then decodes them.
// ------------------------------------------------------------------------------------------------------------------
struct SeedGrids {
  const double *j1, *t1, *t2, *j3, *t3, *t4;  // main grid axes (loop order j1, t1, t2, j3, t3, t4)
  int n_j1, n_t1, n_t2, n_j3, n_t3, n_t4;
  const double *jneg, *jpos, *ts;             // stop grid axes (j1 <= 0, t1, j2 >= 0, t2)
  int n_jneg, n_jpos, n_ts;
  double total_time, acc_max, vel_max, v_max, jmin5, jmax5;
  double lim2, lim3, lim4;                    // total - 0.8*3, total - 1.0*2, total - 1.5 (host-computed, same ops)
};

struct SeedState { double s0, v0, a0; };

__device__ __forceinline__ bool main_seed_ok(const SeedGrids& g, const SeedState& st, double j1, double t1, double t2,
                                             double j3, double t3, double t4, double& j5, double& t5) {
  const double T = g.total_time;
  if (t1 < 0.55) return false;                                                           // polyvt_sampling.py:183
  const double s1 = st_by_jt(st.s0, st.v0, st.a0, j1, t1), v1 = vt_by_jt(st.v0, st.a0, j1, t1), a1 = at_by_jt(st.a0, j1, t1);
  if (!check_state(s1, v1, a1, st.s0, g.acc_max, g.vel_max)) return false;               // :190
  if (t1 + t2 >= g.lim2) return false;                                                   // :195
  const double s2 = st_by_jt(s1, v1, a1, 0.0, t2), v2 = vt_by_jt(v1, a1, 0.0, t2), a2 = at_by_jt(a1, 0.0, t2);
  if (!check_state(s2, v2, a2, s1, g.acc_max, g.vel_max)) return false;                  // :202
  if (t1 + t2 > T + 1e-9) return false;                                                  // :205
  if (j3 * j1 > 0) return false;                                                         // :209
  if (t1 + t2 + t3 > g.lim3) return false;                                               // :213
  const double s3 = st_by_jt(s2, v2, a2, j3, t3), v3 = vt_by_jt(v2, a2, j3, t3), a3 = at_by_jt(a2, j3, t3);
  if (!check_state(s3, v3, a3, s2, g.acc_max, g.vel_max)) return false;                  // :220
  if (a3 * a1 > 0) return false;                                                         // :223
  if (a3 * a2 < 0) {                                                                     // :226-230
    const double t25 = (0.0 - a2) / j3;
    const double v25 = vt_by_jt(v2, a2, j3, t25);
    if (v25 > g.v_max || v25 < 0.0 + 1e-9) return false;
  }
  if (t1 + t2 + t3 + t4 > g.lim4) return false;                                          // :234
  const double s4 = st_by_jt(s3, v3, a3, 0.0, t4), v4 = vt_by_jt(v3, a3, 0.0, t4), a4 = at_by_jt(a3, 0.0, t4);
  if (!check_state(s4, v4, a4, s3, g.acc_max, g.vel_max)) return false;                  // :241
  t5 = T - t1 - t2 - t3 - t4;
  j5 = (0.0 - a4) / t5;
  if (j5 > g.jmax5 || j5 < g.jmin5) return false;                                        // :246
  const double s5 = st_by_jt(s4, v4, a4, j5, t5), v5 = vt_by_jt(v4, a4, j5, t5), a5 = at_by_jt(a4, j5, t5);
  return check_state(s5, v5, a5, s4, g.acc_max, g.vel_max);                              // :252
}

__device__ __forceinline__ void decode_main(const SeedGrids& g, uint32_t f, double& j1, double& t1, double& t2, double& j3,
                                            double& t3, double& t4) {
  uint32_t r = f;
  const uint32_t i4 = r % g.n_t4; r /= g.n_t4;
  const uint32_t i3 = r % g.n_t3; r /= g.n_t3;
  const uint32_t ij3 = r % g.n_j3; r /= g.n_j3;
  const uint32_t i2 = r % g.n_t2; r /= g.n_t2;
  const uint32_t i1 = r % g.n_t1; r /= g.n_t1;
  j1 = g.j1[r]; t1 = g.t1[i1]; t2 = g.t2[i2]; j3 = g.j3[ij3]; t3 = g.t3[i3]; t4 = g.t4[i4];
}

__device__ __forceinline__ void decode_stop(const SeedGrids& g, uint32_t f, double& j1, double& t1, double& j2, double& t2) {
  uint32_t r = f;
  const uint32_t i2 = r % g.n_ts; r /= g.n_ts;
  const uint32_t ij2 = r % g.n_jpos; r /= g.n_jpos;
  const uint32_t i1 = r % g.n_ts; r /= g.n_ts;
  j1 = g.jneg[r]; t1 = g.ts[i1]; j2 = g.jpos[ij2]; t2 = g.ts[i2];
}

// bitonic sort of npow2 <= blockDim.x keys, one per thread: compare-exchange distances below 32 are warp shuffles, only the
// longer ones go through shared memory
__device__ void bitonic_sort_regs(uint32_t* key, int npow2) {
  const int i = threadIdx.x;
  uint32_t v = i < npow2 ? key[i] : 0xffffffffu;
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      uint32_t o;
      if (j >= 32) {
        __syncthreads();
        if (i < npow2) key[i] = v;
        __syncthreads();
        o = i < npow2 ? key[i ^ j] : v;
      } else {
        o = __shfl_xor_sync(0xffffffffu, v, j);
      }
      const bool keep_min = ((i & k) == 0) == ((i & j) == 0);
      v = keep_min ? min(v, o) : max(v, o);
    }
  }
  __syncthreads();
  if (i < npow2) key[i] = v;
  __syncthreads();
}

// bitonic sort of n <= npow2 keys held in shared memory (padding = 0xffffffff)
__device__ void bitonic_sort_smem(uint32_t* key, int npow2) {
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = key[i], b = key[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { key[i] = b; key[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// K1: seed enumeration, ONE thread block per scenario and plan cycle.  The reference's nested loops
// (polyvt_sampling.py:167-305) are evaluated level by level: every level tests its own predicates once per surviving
// prefix, keeps the survivors' segment end states in shared memory and hands them to the next level - the pruning of the
// Python `continue` statements without their serial order.  Survivors append the flat position they have in the
// reference's loop nest; one bitonic sort restores the reference order (main seeds, then stop seeds) and the seeds are
// decoded from the sorted positions.  Same expressions and operand order as the reference at every level, so the seed
// set and the seeds are bit-identical to it.
//   counts [S][4]  results: 0 main found, 1 stop found, 2 seeds emitted, 3 overflow flag
struct SeedWork {
  int* counts;
  double* seeds;            // [S][cap_seeds][10]
  int cap_seeds;
  int cap_a_main, cap_a_stop, cap_b, cap_keys;   // shared-memory list capacities (seed_smem_bytes)
};
struct SeedPrefixA { double s, v, a; int i0, i1; };                            // after segment 1: (j1, t1) or (jneg, t1)
struct SeedPrefixB { double s, v, a, a1, t12; int ij1, i1, i2, pad; };         // after segment 2
constexpr uint32_t STOP_KEY = 0x80000000u;                                     // stop seeds sort behind the main seeds

__host__ __device__ inline size_t seed_smem_bytes(const SeedWork& w) {
  return 32 + (size_t)w.cap_keys * 4 + (size_t)(w.cap_a_main + w.cap_a_stop) * sizeof(SeedPrefixA) + (size_t)w.cap_b * sizeof(SeedPrefixB);
}

struct EvalParams;
__device__ __forceinline__ SeedState seed_state_of(const EvalParams* slots, int sc);
__device__ __forceinline__ bool slot_live(const EvalParams* slots, int sc);

// The block(s) of one scenario form a thread-block CLUSTER (1 CTA when many scenarios keep the GPU busy, up to 8 CTAs when a
// single scenario has to be answered with the lowest latency): the survivor lists live in the shared memory of the
// cluster's CTA 0 and the other CTAs append to and read from them through distributed shared memory; the levels are
// separated by cluster barriers.
constexpr int SEED_THREADS = 1024;
constexpr int SEED_MAX_CLUSTER = 8;
#ifdef JSSP_PROFILE_PHASES
__device__ unsigned long long g_seed_ts[16];
#define SEED_STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_seed_ts[k] = t_; } } while (0)
#else
#define SEED_STAMP(k)
#endif
template <bool kBatch>
__global__ void __launch_bounds__(SEED_THREADS) seed_enum_kernel(SeedGrids g, SeedState st1, const EvalParams* __restrict__ slots,
                                                                 SeedWork wk) {
  extern __shared__ __align__(16) unsigned char seed_smem_local[];
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  const int sc = blockIdx.x / csize;
  if (kBatch && !slot_live(slots, sc)) return;                  // cluster-uniform
  const SeedState st = kBatch ? seed_state_of(slots, sc) : st1;
  unsigned char* seed_smem = csize > 1 ? cluster.map_shared_rank(seed_smem_local, 0) : seed_smem_local;
  int* cnt = reinterpret_cast<int*>(seed_smem);                 // 0 A main, 1 A stop, 2 B, 3 keys, 4 main keys, 5 overflow
  SeedPrefixA* la_main = reinterpret_cast<SeedPrefixA*>(seed_smem + 32);
  SeedPrefixA* la_stop = la_main + wk.cap_a_main;
  SeedPrefixB* lb = reinterpret_cast<SeedPrefixB*>(la_stop + wk.cap_a_stop);
  uint32_t* key = reinterpret_cast<uint32_t*>(lb + wk.cap_b);
  const int tid = crank * blockDim.x + threadIdx.x, nt = csize * blockDim.x;     // position in the cluster
  SEED_STAMP(0);
  if (tid < 8) cnt[tid] = 0;
  cluster.sync();
  SEED_STAMP(1);
  const double T = g.total_time;
  // level 1: (j1, t1) of the main grid and (j1 <= 0, t1) of the stop grid
  const int n1m = g.n_j1 * g.n_t1, n1s = g.n_jneg * g.n_ts;
  for (int e = tid; e < n1m + n1s; e += nt) {
    if (e < n1m) {
      const int ij1 = e / g.n_t1, i1 = e - ij1 * g.n_t1;
      const double j1 = g.j1[ij1], t1 = g.t1[i1];
      if (t1 < 0.55) continue;                                                             // polyvt_sampling.py:183
      const double s1 = st_by_jt(st.s0, st.v0, st.a0, j1, t1), v1 = vt_by_jt(st.v0, st.a0, j1, t1), a1 = at_by_jt(st.a0, j1, t1);
      if (!check_state(s1, v1, a1, st.s0, g.acc_max, g.vel_max)) continue;                 // :190
      const int pos = atomicAdd(&cnt[0], 1);
      if (pos < wk.cap_a_main) la_main[pos] = SeedPrefixA{s1, v1, a1, ij1, i1}; else cnt[5] = 1;
    } else {
      const int f = e - n1m;
      const int ij1 = f / g.n_ts, i1 = f - ij1 * g.n_ts;
      const double j1 = g.jneg[ij1], t1 = g.ts[i1];
      const double s1 = st_by_jt(st.s0, st.v0, st.a0, j1, t1), v1 = vt_by_jt(st.v0, st.a0, j1, t1), a1 = at_by_jt(st.a0, j1, t1);
      if (!check_state(s1, v1, a1, st.s0, g.acc_max, g.vel_max)) continue;                 // :283
      const int pos = atomicAdd(&cnt[1], 1);
      if (pos < wk.cap_a_stop) la_stop[pos] = SeedPrefixA{s1, v1, a1, ij1, i1}; else cnt[5] = 1;
    }
  }
  cluster.sync();
  SEED_STAMP(2);
  // level 2: main prefixes x t2 -> list B; stop prefixes x (j2 >= 0) -> stop seeds.  A warp takes whole prefixes (the
  // entry is one broadcast read, also when it lives in another CTA's shared memory) and spreads t2 / j2 over its lanes.
  const int nam = min(cnt[0], wk.cap_a_main), nas = min(cnt[1], wk.cap_a_stop);
  const int lane = threadIdx.x & 31, warp = tid >> 5, nwarps = nt >> 5;
  {
    const bool fits = g.n_t2 <= 32;
    const int per = fits ? 32 / g.n_t2 : 1;                   // main prefixes per warp and round
    const int sub = fits ? lane / g.n_t2 : 0;
    const int i2_first = fits ? lane - sub * g.n_t2 : lane, i2_step = fits ? g.n_t2 : 32;
    for (int base = warp * per; base < nam; base += nwarps * per) {
      const int ia = base + sub;
      if (sub >= per || ia >= nam) continue;
      const SeedPrefixA A = la_main[ia];
      const double t1 = g.t1[A.i1];
      for (int i2 = i2_first; i2 < g.n_t2; i2 += i2_step) {
        const double t2 = g.t2[i2];
        if (t1 + t2 >= g.lim2) continue;                                                   // :195
        const double s2 = st_by_jt(A.s, A.v, A.a, 0.0, t2), v2 = vt_by_jt(A.v, A.a, 0.0, t2), a2 = at_by_jt(A.a, 0.0, t2);
        if (!check_state(s2, v2, a2, A.s, g.acc_max, g.vel_max)) continue;                 // :202
        if (t1 + t2 > T + 1e-9) continue;                                                  // :205
        const int pos = atomicAdd(&cnt[2], 1);
        if (pos < wk.cap_b) lb[pos] = SeedPrefixB{s2, v2, a2, A.a, t1 + t2, A.i0, A.i1, i2, 0}; else cnt[5] = 1;
      }
    }
  }
  {
    // stop prefixes x (j2 >= 0): a2 = a1 + j2*t2 is monotone in t2, so only a short run of the (uniform, np.arange) t2
    // grid can pass -0.3 <= a2 <= 0.  It is bracketed in closed form (fp32 is plenty for a one-sample margin) and the exact
    // float64 predicates decide inside the bracket.  The run is ~0.3 / (j2 * dt) samples long, so the work is balanced by
    // giving the lanes of a warp the SAME j2: with few threads per scenario (batch) lanes take prefixes and loop over j2,
    // with many (a cluster answering one scenario) a warp takes one prefix and its lanes the j2 values.
    const float dt_s = (float)(g.ts[g.n_ts > 1 ? 1 : 0] - g.ts[0]), t0_s = (float)g.ts[0];
    const bool lanes_over_prefixes = nas * 2 >= nt;
    const int ia_first = lanes_over_prefixes ? tid : warp, ia_step = lanes_over_prefixes ? nt : nwarps;
    const int j_first = lanes_over_prefixes ? 0 : lane, j_step = lanes_over_prefixes ? 1 : 32;
    for (int ia = ia_first; ia < nas; ia += ia_step) {
      const SeedPrefixA A = la_stop[ia];
      const double t1 = g.ts[A.i1];
      const double s1 = A.s, v1 = A.v, a1 = A.a;
      const float a1f = (float)a1, v1f = (float)v1;
      for (int ij2 = j_first; ij2 < g.n_jpos; ij2 += j_step) {
        const double j2 = g.jpos[ij2];
        int lo = 0, hi = g.n_ts - 1;
        if (j2 > 1e-6 && dt_s > 0.f) {
          // inside the bracket v2 lies in [v*, v* + 0.045 / j2] with v* = v1 - a1^2 / (2 j2), the minimum of v2 (reached
          // where a2 = 0): unless that band meets -0.3 <= v2 <= 0 (2 cm/s of slack for fp32) no t2 can pass - most pairs
          const float rj = __fdividef(1.0f, (float)j2);
          const float vstar = v1f - 0.5f * a1f * a1f * rj;
          if (vstar > 0.02f || vstar < -0.32f - 0.045f * rj) continue;
          const float inv = rj / dt_s;
          const float flo = (-0.3f - a1f) * inv - t0_s / dt_s, fhi = (0.0f - a1f) * inv - t0_s / dt_s;
          lo = (int)fmaxf(0.f, fminf(floorf(flo) - 1.f, (float)g.n_ts));
          hi = (int)fminf((float)(g.n_ts - 1), fmaxf(ceilf(fhi) + 1.f, -1.f));
        }
        for (int i2 = lo; i2 <= hi; ++i2) {
          const double t2 = g.ts[i2];
          if (t1 + t2 >= T) continue;                                                      // :291
          const double a2 = at_by_jt(a1, j2, t2);
          if (a2 > 0.0 || a2 < -0.3) continue;                                             // :293
          const double v2 = vt_by_jt(v1, a1, j2, t2);
          if (v2 > 0.0 || v2 < -0.3) continue;                                             // :296
          const double s2 = st_by_jt(s1, v1, a1, j2, t2);
          if (s2 < s1 - 0.1) continue;                                                     // :299
          const int pos = atomicAdd(&cnt[3], 1);
          const uint32_t flat = (((uint32_t)A.i0 * g.n_ts + A.i1) * g.n_jpos + ij2) * g.n_ts + i2;
          if (pos < wk.cap_keys) key[pos] = STOP_KEY | flat; else cnt[5] = 1;
        }
      }
    }
  }
  cluster.sync();
  SEED_STAMP(3);
  // level 3: list B x j3 x t3, then the t4 loop and the closing segment
  const int nb = min(cnt[2], wk.cap_b);
  const int n34 = g.n_j3 * g.n_t3;
  for (int e = tid; e < nb * n34; e += nt) {
    const int ib = e / n34, r = e - ib * n34;
    const int ij3 = r / g.n_t3, i3 = r - ij3 * g.n_t3;
    const SeedPrefixB B = lb[ib];
    const double j1 = g.j1[B.ij1], j3 = g.j3[ij3], t3 = g.t3[i3];
    if (j3 * j1 > 0) continue;                                                             // :209
    if (B.t12 + t3 > g.lim3) continue;                                                     // :213
    const double s3 = st_by_jt(B.s, B.v, B.a, j3, t3), v3 = vt_by_jt(B.v, B.a, j3, t3), a3 = at_by_jt(B.a, j3, t3);
    if (!check_state(s3, v3, a3, B.s, g.acc_max, g.vel_max)) continue;                     // :220
    if (a3 * B.a1 > 0) continue;                                                           // :223
    if (a3 * B.a < 0) {                                                                    // :226-230
      const double t25 = (0.0 - B.a) / j3;
      const double v25 = vt_by_jt(B.v, B.a, j3, t25);
      if (v25 > g.v_max || v25 < 0.0 + 1e-9) continue;
    }
    const double t1 = g.t1[B.i1], t2 = g.t2[B.i2];
    for (int i4 = 0; i4 < g.n_t4; ++i4) {
      const double t4 = g.t4[i4];
      if (B.t12 + t3 + t4 > g.lim4) continue;                                              // :234
      const double s4 = st_by_jt(s3, v3, a3, 0.0, t4), v4 = vt_by_jt(v3, a3, 0.0, t4), a4 = at_by_jt(a3, 0.0, t4);
      if (!check_state(s4, v4, a4, s3, g.acc_max, g.vel_max)) continue;                    // :241
      const double t5 = T - t1 - t2 - t3 - t4;
      const double j5 = (0.0 - a4) / t5;
      if (j5 > g.jmax5 || j5 < g.jmin5) continue;                                          // :246
      const double s5 = st_by_jt(s4, v4, a4, j5, t5), v5 = vt_by_jt(v4, a4, j5, t5), a5 = at_by_jt(a4, j5, t5);
      if (!check_state(s5, v5, a5, s4, g.acc_max, g.vel_max)) continue;                    // :252
      const int pos = atomicAdd(&cnt[3], 1);
      atomicAdd(&cnt[4], 1);
      const uint32_t flat = (((((uint32_t)B.ij1 * g.n_t1 + B.i1) * g.n_t2 + B.i2) * g.n_j3 + ij3) * g.n_t3 + i3) * g.n_t4 + i4;
      if (pos < wk.cap_keys) key[pos] = flat; else cnt[5] = 1;
    }
  }
  cluster.sync();
  SEED_STAMP(4);
  // reference order: one sort of the loop positions (stop seeds carry the top bit) by CTA 0, then every CTA decodes
  const int n_all = min(cnt[3], wk.cap_keys), n_main = cnt[4];
  if (crank == 0) {
    int np2 = 1;
    while (np2 < n_all) np2 <<= 1;
    for (int i = n_all + (int)threadIdx.x; i < np2; i += blockDim.x) key[i] = 0xffffffffu;
    __syncthreads();
    if (np2 <= (int)blockDim.x) bitonic_sort_regs(key, np2); else bitonic_sort_smem(key, np2);
  }
  SEED_STAMP(5);
  cluster.sync();
  SEED_STAMP(6);
  double* seeds = wk.seeds + (size_t)sc * wk.cap_seeds * 10;
  for (int i = tid; i < min(n_all, wk.cap_seeds); i += nt) {
    double* o = seeds + (size_t)i * 10;
    const uint32_t k = key[i];
    if (!(k & STOP_KEY)) {
      double j1, t1, t2, j3, t3, t4, j5 = 0, t5 = 0;
      decode_main(g, k, j1, t1, t2, j3, t3, t4);
      main_seed_ok(g, st, j1, t1, t2, j3, t3, t4, j5, t5);   // recompute j5, t5 with identical arithmetic
      o[0] = j1; o[1] = t1; o[2] = 0.0; o[3] = t2; o[4] = j3; o[5] = t3; o[6] = 0.0; o[7] = t4; o[8] = j5; o[9] = t5;
    } else {
      double j1, t1, j2, t2;
      decode_stop(g, k & ~STOP_KEY, j1, t1, j2, t2);
      o[0] = j1; o[1] = t1; o[2] = j2; o[3] = t2; o[4] = 0.0; o[5] = T - t1 - t2;          // :302
      o[6] = 0.0; o[7] = 0.0; o[8] = 0.0; o[9] = 0.0;
    }
  }
  if (tid == 0) {
    int* counts = wk.counts + (size_t)sc * 4;
    int overflow = cnt[5];
    int emitted = n_all;
    if (cnt[3] > wk.cap_keys || emitted > wk.cap_seeds) { overflow = 1; emitted = min(emitted, wk.cap_seeds); }
    counts[0] = n_main; counts[1] = cnt[3] - n_main; counts[2] = emitted; counts[3] = overflow;
  }
  SEED_STAMP(7);
  cluster.sync();                                               // CTA 0's shared memory stays alive until every CTA has read it
}

// ------------------------------------------------------------------------------------------------------------------
// K0: obstacle tables.  raw state [OM][Tt][3] (x, y, yaw), valid [OM][Tt] ->
//   broad  [Tt][OMpad] float4 = (x', y', |o'|^2 - R^2, prob) with o' = o - origin, R = r_ego + r_obs + pad; the
//          squared-distance test |o' - e'|^2 <= R^2 becomes  x'*(-2ex) + y'*(-2ey) + (|o'|^2 - R^2) <= -|e'|^2
//          (2 FFMA + 1 compare per pair).  Absent states and the padding up to a multiple of 32 hold +FAR.
//   narrow [Tt][OM] float4 = (cos yaw, sin yaw, half length + infl/2, half width + infl/2)
//   obs64  [Tt][OM][4]     = x, y, cos yaw, sin yaw in float64 for the guard-band re-decision
// ------------------------------------------------------------------------------------------------------------------
constexpr float BROAD_PAD = 0.05f;        // [m] added to the bounding-circle radius sum
constexpr int PACK_LANES = 8;             // lanes sharing one obstacle state in pack_obstacles_kernel
constexpr float BROAD_CANCEL = 0.5f;      // [m^2] covers fp32 cancellation of the expanded form for |coord| <= 1000 m
__global__ void pack_obstacles_kernel(const double* __restrict__ state, const uint8_t* __restrict__ valid,
                                      const double* __restrict__ size, const double* __restrict__ prob, int OM, int OMpad, int M,
                                      int Tt, double ox, double oy, double inflation, double r_ego, float4* __restrict__ broad,
                                      float4* __restrict__ narrow, double* __restrict__ obs64, double* __restrict__ ext64,
                                      float* __restrict__ probf, float2* __restrict__ sint, const double* __restrict__ knot_s,
                                      const double* __restrict__ knot_xy, int K, double cull_reach, double cull_ds) {
  // PACK_LANES consecutive lanes own one table entry idx = t * OMpad + om: lane 0 of the group writes the rows, all of them
  // share the scan over the reference-line knots
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gtid / PACK_LANES, sub = gtid % PACK_LANES;
  if (idx >= OMpad * Tt) return;               // whole groups leave together (blockDim is a multiple of PACK_LANES)
  const int t = idx / OMpad, om = idx - t * OMpad;
  if (om >= OM) {
    if (sub == 0) { broad[idx] = make_float4(0.f, 0.f, FAR_AWAY, 0.f); if (sint) sint[idx] = make_float2(FAR_AWAY, -FAR_AWAY); }
    return;
  }
  const int o = om / M;
  const double a2 = (size[2 * o] + inflation) / 2.0, b2 = (size[2 * o + 1] + inflation) / 2.0;
  const double* s = state + ((size_t)om * Tt + t) * 3;
  const bool ok = valid[(size_t)om * Tt + t] != 0;
  const float r = (float)(r_ego + sqrt(a2 * a2 + b2 * b2)) + BROAD_PAD;
  if (sub == 0) {
    if (t == 0) { ext64[2 * om] = a2; ext64[2 * om + 1] = b2; probf[om] = (float)prob[om]; }
    double sn, cs;
    sincos(s[2], &sn, &cs);
    const float fx = (float)(s[0] - ox), fy = (float)(s[1] - oy);
    const float q = fmaf(fx, fx, fy * fy) - (r * r * 1.0001f + BROAD_CANCEL);
    broad[idx] = ok ? make_float4(fx, fy, q, (float)prob[om]) : make_float4(0.f, 0.f, FAR_AWAY, 0.f);
    const size_t gi = (size_t)t * OM + om;
    narrow[gi] = make_float4((float)cs, (float)sn, (float)a2, (float)b2);
    double* q64 = obs64 + gi * 4;
    q64[0] = s[0]; q64[1] = s[1]; q64[2] = cs; q64[3] = sn;
  }
  // Arc-length interval of the reference line from which this obstacle state can be reached by any candidate footprint:
  // every curve point within (lateral cap + R) of the obstacle lies within cull_reach of one of its segment's knots
  // (cull_reach includes the longest knot-to-knot arc), so the hull of those knots widened by cull_ds is conservative.
  if (!sint) return;
  double lo = 1e300, hi = -1e300;
  if (ok && K > 0) {
    const double reach = cull_reach + (double)r;
    const double r2 = reach * reach;
    for (int k = sub; k < K; k += PACK_LANES) {
      const double dx = knot_xy[2 * k] - s[0], dy = knot_xy[2 * k + 1] - s[1];
      if (dx * dx + dy * dy <= r2) { lo = fmin(lo, knot_s[k]); hi = fmax(hi, knot_s[k]); }
    }
  }
  const unsigned int gmask = ((1u << PACK_LANES) - 1u) << ((threadIdx.x & 31) & ~(PACK_LANES - 1));
#pragma unroll
  for (int d = PACK_LANES / 2; d > 0; d >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(gmask, lo, d));
    hi = fmax(hi, __shfl_xor_sync(gmask, hi, d));
  }
  if (sub == 0) sint[idx] = hi >= lo ? make_float2(__double2float_rd(lo - cull_ds), __double2float_ru(hi + cull_ds)) : make_float2(FAR_AWAY, -FAR_AWAY);
}

// ------------------------------------------------------------------------------------------------------------------
// K2: candidate evaluation
// ------------------------------------------------------------------------------------------------------------------
struct EvalParams {
  // cycle
  double s0, v0, a0;
  double quintic[JSSP_MAX_WIDTHS][6];
  int W, P, now, tcap;            // tcap = final_time_step - time_step_now
  int n_dict, check_res;
  // constants
  double kmax, ea, eb, p_min, w_risk, vmax, inv_vmax, max_distance, half_w, thr_lon_a, thr_lon_j, thr_lat_a, thr_lat_j;
  double cw[8];
  // tables
  const double* time;
  const double* knot;
  const double* coef;
  int K, kpad;
  double ox, oy;
  const double* seeds;
  const int* n_seeds_ptr;         // device count (enumerated seeds) or NULL
  int n_seeds_fixed, seed_cap;
  // obstacles
  const float4* broad;
  const float4* narrow;
  const double* obs64;
  const double* ext64;
  int OM, OMpad, Tt, win_rows, smem_obs;
  const float2* sint;             // [Tt][OMpad] reachable arc-length interval of every obstacle state (block-level culling)
  const float* probf;             // [OM]
  int list_cap;                   // entries of the per-block compacted obstacle lists (mode 2)
  // outputs
  uint8_t* feasible;
  uint8_t* collision;
  float* cost;
  double* blk_cost;
  int* blk_idx;
  unsigned int* ticket;
  BestRecord* best;
  double* best_traj;              // [P][16] export of the selected trajectory (NULL = skip)
  // batched replay (NULL / unused for a single plan cycle)
#ifdef JSSP_PROFILE_PHASES
  unsigned long long* phase_ts;   // [blocks][8] globaltimer stamps (tools/phase_probe.py; profiling builds only)
#endif
  struct BatchSlot* batch;        // per-scenario replay state, updated by the hand-off at the end of every cycle
  EvalParams* self;               // this scenario's parameter record in device memory (the hand-off writes the next cycle's inputs)
};

// Replay state of one scenario of a batch (device memory).  The hand-off at the end of a plan cycle - the last block of
// the evaluation kernel - moves the ego to sample 1 of the selected trajectory (code/onsite_trans/env.py:24-30), makes
// that sample the next cycle's Frenet start (intersection_planner.py:312), advances the clock (controller.py:303) and
// evaluates the end-of-scenario status (controller.py:338-380), so that a whole replay runs without host round trips.
struct BatchSlot {
  int live;                       // 1 while the scenario is running
  int cycle;                      // plan cycles done
  int final_step;                 // int(max_t / dt)
  int gt_O, gt_Tt;                // ground-truth background vehicles and steps of their table
  int n_cycles_cap;               // the per-cycle tables hold n_cycles_cap + 1 entries: the replay pauses after that many cycles
  double end;                     // end status: -1 running, 1 max_t, 1.5 goal, 2 collision, -2 no solution, -3 short trajectory
  double max_t, horizon;
  double goal_x, goal_y, goal_cos, goal_sin;
  double ego_l, ego_w;
  double targets[JSSP_MAX_WIDTHS];
  const double* cycle_time;       // [C+1] clock at the start of cycle c (host-accumulated float additions)
  const int* cycle_now;           // [C+1] int(t / dt)
  const int* cycle_end_step;      // [C+1] table step whose key equals str(around(t, 3)), -1 = none
  const double* gt_state;         // [gt_O][gt_Tt][3] x, y, yaw of the background vehicles
  const uint8_t* gt_valid;        // [gt_O][gt_Tt]
  const double* gt_size;          // [gt_O][2] length, width
  jssp_cycle_record* records;     // [C]
};

__device__ __forceinline__ SeedState seed_state_of(const EvalParams* slots, int sc) { return SeedState{slots[sc].s0, slots[sc].v0, slots[sc].a0}; }
__device__ __forceinline__ bool slot_live(const EvalParams* slots, int sc) { return slots[sc].batch->live != 0; }

struct BlockTables {
  const float4* broad;   // window rows [win_rows][OMpad] (shared) or the global table [Tt][OMpad]
  int broad_row0;        // 0 for the shared window, time_step_now for the global table
  int broad_stride;      // rows advance by check_res in the global table, 1 in the window
  SplineView sp;
  const double* time;
  const double* lat;     // [W][P][4]
  // mode 2: per-row lists of the obstacle states this block's candidates can reach: (x', y', |o'|^2 - R^2, om bits)
  const float4* list;
  const int* row_off;
  const int* row_cnt;
  bool use_lists;
};

// Walks one candidate through the horizon and hands every derived quantity to `sink`, in the order the reference
// derives them (get_one_frenet_trajectory -> calc_global_paths).  Used by the evaluation, export and dump kernels so
// that all three see the same positions.
//   sink.sample(i, t, s, sd, sdd, sddd)                      every i in [0, P)
//   sink.point(i, x, y)                                      i < n_pts
//   sink.segment(i, x_i, y_i, dx, dy, ds_i, cos, sin)        i < n_pts - 1: forward difference p_{i+1} - p_i, its length
//                                                            and unit vector ((1, 0) when the length is 0: atan2(0,0) = 0)
//   sink.finish(n_pts, x_L, y_L)                             once; L = n_pts - 1
// `lon.eval(t, s, sd, sdd, sddd)` is the longitudinal model (the 5-segment jerk profile of the JSSP seeds, the quartic of
// the FOP baseline), `lat` the candidate's lateral samples [P][4], P its number of samples.
template <class Sink, class Lon>
__device__ __forceinline__ void walk_core(const BlockTables& tb, Lon& lon, const double* lat, int P, int seg, Sink& sink) {
  int n_pts = P;
  double xp = 0, yp = 0;
  int i = 0;
  // live part: ends at the first sample whose s is off the reference line (calc_position returned None -> break,
  // jerk_space_sampling_planner.py:138-139)
  for (; i < P; ++i) {
    double s, sd, sdd, sddd;
    lon.eval(tb.time[i], s, sd, sdd, sddd);
    sink.sample(i, tb.time[i], s, sd, sdd, sddd);
    if (!spline_in_range(tb.sp, s)) { n_pts = i; ++i; break; }
    seg = spline_walk(tb.sp, s, seg);
    double x, y;
    frenet_point(tb.sp, seg, s, lat[i * 4], x, y);
    sink.point(i, x, y);
    if (i >= 1) {
      const double dx = x - xp, dy = y - yp;
      const double q = dx * dx + dy * dy;
      double ds, c = 1.0, sn = 0.0;
      if (q > 1e-200 && q < 1e200) { const double inv = rsqrt(q); ds = q * inv; c = dx * inv; sn = dy * inv; }
      else { ds = hypot(dx, dy); if (ds > 0.0) { c = dx / ds; sn = dy / ds; } }
      sink.segment(i - 1, xp, yp, dx, dy, ds, c, sn);
    }
    xp = x; yp = y;
  }
  // Frenet-only tail (the cost still sums over every sample)
  for (; i < P; ++i) {
    double s, sd, sdd, sddd;
    lon.eval(tb.time[i], s, sd, sdd, sddd);
    sink.sample(i, tb.time[i], s, sd, sdd, sddd);
  }
  sink.finish(n_pts, xp, yp);
}

struct JerkLon {
  LonWalk& w;
  __device__ __forceinline__ void eval(double t, double& s, double& sd, double& sdd, double& sddd) { lon_walk_eval(w, t, s, sd, sdd, sddd); }
};

template <class Sink>
__device__ __forceinline__ void walk_candidate(const EvalParams& p, const BlockTables& tb, LonWalk& lp, int w,
                                               int seg, Sink& sink) {
  JerkLon lon{lp};
  walk_core(tb, lon, tb.lat + (size_t)w * p.P * 4, p.P, seg, sink);
}

__device__ __forceinline__ double heading_of(double dx, double dy) { return atan2(dy, dx); }   // np.arctan2(y_d, x_d)

// c_{i-1} = (yaw_i - yaw_{i-1}) / ds_{i-1} with yaw = arctan2 of the forward differences and NO unwrapping
// (jerk_space_sampling_planner.py:155-163); the candidate fails when |c| > kmax (:175).  (p*) = segment i-1, the others
// segment i, each as difference vector, length and unit vector.  Fast path: sin / cos of the wrapped angle from the unit
// vectors, compared against the series of sin(kmax * ds); the +-2*pi jump of the un-unwrapped difference is detected from
// the exact signs; inside a 1e-6 relative guard band, for zero-length segments and for jumps the float64 atan2 decides.
__device__ __forceinline__ bool curvature_fails(double kmax, double pdx, double pdy, double pds, double pc, double psn, double dx,
                                                double dy, double ds, double c, double sn) {
  const double phi = kmax * pds;
  bool exact = !(pds > 0.0) || !(ds > 0.0) || !(phi < 0.5);
  if (!exact) {
    const double cr = pc * sn - psn * c, dt = pc * c + psn * sn;      // sin / cos of the wrapped angle (unit vectors)
    const bool nb = signbit(dy);
    if ((signbit(pdy) != nb) && (nb ? (cr > 0.0) : (cr < 0.0))) exact = true;   // the difference jumps by 2*pi: rare
    else if (!(dt > 0.0)) return true;                                           // |angle| >= pi/2 > phi
    else {
      // |angle| < pi/2:  |angle| > phi  <=>  |sin angle| > sin phi;  sin phi by its series (phi < 0.5: 1e-11 relative)
      const double p2 = phi * phi;
      const double sp = phi * fma(p2, fma(p2, fma(p2, fma(p2, 1.0 / 362880.0, -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
      const double a = fabs(cr);
      if (fabs(a - sp) > 1e-6 * fmax(a, sp)) return a > sp;
      exact = true;
    }
  }
  const double dyaw = heading_of(dx, dy) - heading_of(pdx, pdy);
  return fabs(dyaw) > phi;      // pds == 0: inf fails (dyaw != 0), nan passes (dyaw == 0)
}

// -- the evaluation sink ---------------------------------------------------------------------------------------------
struct EvalSink {
  const EvalParams& p;
  const BlockTables& tb;
  double sum_a = 0, sum_j = 0, sum_v = 0, s_first = 0, s_last = 0, risk = 0;
  double pdx = 0, pdy = 0, pds = 0, pc = 1, psn = 0;   // previous segment (difference, length, unit vector)
  int next_check = 0, row = 0;                         // next sample index to collision-check and its window row
  bool curv_fail = false, collided = false;
  __device__ __forceinline__ EvalSink(const EvalParams& p_, const BlockTables& tb_) : p(p_), tb(tb_) {}

  __device__ __forceinline__ void sample(int i, double, double s, double sd, double sdd, double sddd) {
    if (i == 0) s_first = s;
    s_last = s;
    const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
    const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
    const double ev = (p.vmax - sd) * p.inv_vmax;
    sum_a = sum_a + ea * ea;
    sum_j = sum_j + ej * ej;
    sum_v = sum_v + ev * ev;
  }
  // the same accumulation from per-sample cost terms computed elsewhere (group kernel: by the lane that owns the sample)
  __device__ __forceinline__ void sample_terms(double ea2, double ej2, double ev2) {
    sum_a = sum_a + ea2;
    sum_j = sum_j + ej2;
    sum_v = sum_v + ev2;
  }
  __device__ __forceinline__ void point(int, double, double) {}

  __device__ __forceinline__ void curvature_step(double dx, double dy, double ds, double c, double sn) {
    if (curvature_fails(p.kmax, pdx, pdy, pds, pc, psn, dx, dy, ds, c, sn)) curv_fail = true;
  }

  // separating-axis test of the ego rectangle at (x, y, heading (c, sn)) against obstacle mode om at abs_step
  __device__ __forceinline__ bool narrow_hit(int abs_step, int om, float dxf, float dyf, double x, double y, double c, double sn) const {
    const size_t gi = (size_t)abs_step * p.OM + om;
    const float4 nb = __ldg(p.narrow + gi);
    const int cls = sat_classify_f32(dxf, dyf, (float)c, (float)sn, (float)p.ea, (float)p.eb, nb.x, nb.y, nb.z, nb.w, SAT_EPS);
    if (cls != 0) return cls < 0;
    const double* q = p.obs64 + gi * 4;
    return sat_intersects_f64(q[0] - x, q[1] - y, c, sn, p.ea, p.eb, q[2], q[3], p.ext64[2 * om], p.ext64[2 * om + 1]);
  }

  __device__ __forceinline__ void collide(int i, double x, double y, double c, double sn) {
    if (i != next_check) return;               // i % check_res == 0 without the division
    next_check += p.check_res;
    const int r = row++;
    if (collided || p.n_dict <= 0 || i >= p.tcap) return;
    const int abs_step = p.now + i;
    if (abs_step >= p.Tt) return;
    const float4* row = tb.broad + (size_t)(tb.broad_row0 + r * tb.broad_stride) * p.OMpad;
    const float fx = (float)(x - p.ox), fy = (float)(y - p.oy);
    const float m2x = -2.0f * fx, m2y = -2.0f * fy, ne2 = -fmaf(fx, fx, fy * fy);
    if (tb.use_lists) {      // block-level culled rows: a handful of entries instead of OMpad
      const int cnt = tb.row_cnt[r];
      const float4* e = tb.list + tb.row_off[r];
      for (int k = 0; k < cnt; ++k) {
        const float4 b = e[k];
        if (fmaf(b.x, m2x, fmaf(b.y, m2y, b.z)) <= ne2) {
          const int om = __float_as_int(b.w);
          if (narrow_hit(abs_step, om, b.x - fx, b.y - fy, x, y, c, sn)) {
            const float pr = __ldg(p.probf + om);
            if ((double)pr >= p.p_min) { collided = true; return; }
            risk = risk + (double)pr;
          }
        }
      }
      return;
    }
    for (int base = 0; base < p.OM; base += 32) {
      unsigned int mask = 0u;
      if (p.OM - base >= 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 b = row[base + j];
          const float t = fmaf(b.x, m2x, fmaf(b.y, m2y, b.z));
          mask |= (t <= ne2) ? (1u << j) : 0u;
        }
      } else {                                  // last, partial chunk: the padding up to OMpad is never touched
        const int nj = p.OM - base;
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
          const float4 b = row[base + j];
          const float t = fmaf(b.x, m2x, fmaf(b.y, m2y, b.z));
          mask |= (t <= ne2) ? (1u << j) : 0u;
        }
      }
      while (mask) {      // rare: bounding circles overlap -> separating-axis test
        const int om = base + __ffs(mask) - 1;
        mask &= mask - 1u;
        const float4 b = row[om];
        if (narrow_hit(abs_step, om, b.x - fx, b.y - fy, x, y, c, sn)) {
          if ((double)b.w >= p.p_min) { collided = true; return; }
          risk = risk + (double)b.w;
        }
      }
    }
  }
  __device__ __forceinline__ void segment(int i, double x, double y, double dx, double dy, double ds, double c, double sn) {
    if (i >= 1) curvature_step(dx, dy, ds, c, sn);
    collide(i, x, y, c, sn);
    pdx = dx; pdy = dy; pds = ds; pc = c; psn = sn;
  }
  __device__ __forceinline__ void finish(int n_pts, double x, double y) {
    if (n_pts >= 2) collide(n_pts - 1, x, y, pc, psn);                      // last point reuses the previous heading (:158-160)
    else if (n_pts == 1 && p.n_dict > 0 && p.tcap >= 1) collided = true;    // bare except -> collision (:250-254)
  }
};

// ------------------------------------------------------------------------------------------------------------------
// Group form of the walk: the W candidates that share a longitudinal seed (one per lateral target) sit in W consecutive
// lanes.  Everything that depends on the seed only - the unrolled s, s_d, s_dd, s_ddd, the cost terms, the spline segment,
// the reference point and its unit normal - is computed ONCE per sample by the lane that owns the sample (lane j owns
// samples j, j + W, ...) and handed to the group by warp shuffles; every lane then adds its own lateral offset and runs the
// per-candidate part (difference, length, curvature test, collision rows).  Same expressions, same order of the cost
// sums: results are bit-identical to walk_candidate.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_f64(unsigned int mask, double v, int src) { return __shfl_sync(mask, v, src); }

__device__ __forceinline__ void walk_group(const EvalParams& p, const BlockTables& tb, LonWalk& lp, int w, int W, unsigned int gmask,
                                           int glane0, EvalSink& sink) {
  const int P = p.P;
  const double* lat = tb.lat + (size_t)w * P * 4;
  int n_pts = P;
  bool live = true;
  double xp = 0, yp = 0;
  for (int base = 0; base < P; base += W) {
    // the sample this lane owns in this round
    const int io = base + w;
    double o_s = 0, o_ea = 0, o_ej = 0, o_ev = 0, o_ix = 0, o_iy = 0, o_u = 0, o_v = 0;
    int o_in = 0;
    if (io < P) {
      double sd, sdd, sddd;
      lon_walk_eval(lp, tb.time[io], o_s, sd, sdd, sddd);
      const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
      const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
      const double ev = (p.vmax - sd) * p.inv_vmax;
      o_ea = ea * ea; o_ej = ej * ej; o_ev = ev * ev;
      if (live && spline_in_range(tb.sp, o_s)) {
        o_in = 1;
        frenet_base(tb.sp, spline_walk(tb.sp, o_s, 0), o_s, o_ix, o_iy, o_u, o_v);
      }
    }
    const int cnt = min(W, P - base);
    for (int jj = 0; jj < cnt; ++jj) {
      const int i = base + jj, src = glane0 + jj;
      sink.sample_terms(shfl_f64(gmask, o_ea, src), shfl_f64(gmask, o_ej, src), shfl_f64(gmask, o_ev, src));
      if (i == 0) sink.s_first = shfl_f64(gmask, o_s, src);
      if (i == P - 1) sink.s_last = shfl_f64(gmask, o_s, src);
      if (!live) continue;
      if (!__shfl_sync(gmask, o_in, src)) { n_pts = i; live = false; continue; }   // first sample off the reference line (:138-139)
      const double ix = shfl_f64(gmask, o_ix, src), iy = shfl_f64(gmask, o_iy, src);
      const double u = shfl_f64(gmask, o_u, src), v = shfl_f64(gmask, o_v, src);
      const double d = lat[i * 4];
      const double x = ix - d * u, y = iy + d * v;
      if (i >= 1) {
        const double dx = x - xp, dy = y - yp;
        const double q = dx * dx + dy * dy;
        double ds, c = 1.0, sn = 0.0;
        if (q > 1e-200 && q < 1e200) { const double inv = rsqrt(q); ds = q * inv; c = dx * inv; sn = dy * inv; }
        else { ds = hypot(dx, dy); if (ds > 0.0) { c = dx / ds; sn = dy / ds; } }
        sink.segment(i - 1, xp, yp, dx, dy, ds, c, sn);
      }
      xp = x; yp = y;
    }
  }
  sink.finish(n_pts, xp, yp);
}

__device__ __forceinline__ double candidate_cost(const EvalParams& p, double sum_a, double sum_j, double sum_v, double s_first,
                                                 double s_last, double risk, const double* latcost) {
  const double P = (double)p.P;
  double cost = p.cw[1] * (1.0 - (s_last - s_first) / p.max_distance);
  cost = cost + p.cw[2] * (latcost[0] / P / (p.half_w * p.half_w));
  cost = cost + p.cw[3] * (latcost[1] / P);
  cost = cost + p.cw[4] * (latcost[2] / P);
  cost = cost + p.cw[5] * (sum_a / P);
  cost = cost + p.cw[6] * (sum_j / P);
  cost = cost + p.cw[7] * (sum_v / P);
  cost = cost + p.w_risk * risk;
  return cost;
}
__device__ __forceinline__ double candidate_cost(const EvalParams& p, const EvalSink& k, const double* latcost) {
  return candidate_cost(p, k.sum_a, k.sum_j, k.sum_v, k.s_first, k.s_last, k.risk, latcost);
}

// dynamic shared memory carve-up (the host uses the same function to size the launch)
struct SmemLayout {
  size_t off_broad, off_rows, off_knot, off_coef, off_time, off_lat, off_latcost, total;
};
// mode: 0 = obstacle table read from global memory, 1 = the cycle's window staged by TMA, 2 = per-block culled lists
__host__ __device__ inline SmemLayout eval_smem_layout(int win_rows, int OMpad, int mode, int K, int kpad, int P, int W, int list_cap = 0) {
  SmemLayout L;
  size_t o = 16;   // mbarrier / list cursor
  L.off_broad = o; o += mode == 1 ? (size_t)win_rows * OMpad * 16 : (mode == 2 ? (size_t)list_cap * 16 : 0);
  L.off_rows = o; o += mode == 2 ? (size_t)((4 * win_rows + 3) & ~3) * 4 : 0;   // row_off, row_cnt, row_smin, row_smax
  L.off_knot = o; o += (size_t)((K + 1) & ~1) * 8;
  L.off_coef = o; o += (size_t)8 * kpad * 8;
  L.off_time = o; o += (size_t)((P + 1) & ~1) * 8;
  L.off_lat = o; o += (size_t)W * P * 4 * 8;
  L.off_latcost = o; o += (size_t)W * 4 * 8;
  L.total = o;
  return L;
}

// Stage the per-block tables: obstacle window by bulk copy (TMA engine; one row per issued copy, all completing on one
// mbarrier), spline, time grid and the lateral quintic samples (jerk_space_sampling_planner.py:118-122) by ordinary loads.
template <int kMode>
__device__ __forceinline__ void stage_tables(const EvalParams& p, unsigned char* smem, BlockTables& tb, const double*& latcost) {
  constexpr bool kSmem = kMode == 1;
  const SmemLayout L = eval_smem_layout(p.win_rows, p.OMpad, kMode, p.K, p.kpad, p.P, p.W, p.list_cap);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float4* sm_broad = reinterpret_cast<float4*>(smem + L.off_broad);
  double* sm_knot = reinterpret_cast<double*>(smem + L.off_knot);
  double* sm_coef = reinterpret_cast<double*>(smem + L.off_coef);
  double* sm_time = reinterpret_cast<double*>(smem + L.off_time);
  double* sm_lat = reinterpret_cast<double*>(smem + L.off_lat);
  double* sm_latcost = reinterpret_cast<double*>(smem + L.off_latcost);
  const int tid = threadIdx.x, nt = blockDim.x;
  const uint32_t row_bytes = (uint32_t)p.OMpad * 16u;
  // Everything that is copied verbatim goes through the TMA engine (1-D bulk copies completing on one mbarrier): the
  // reference-line knots and coefficients, the time grid and - mode 1 - the rows of the obstacle window.
  int rows_avail = 0;
  if (kSmem) for (int r = 0; r < p.win_rows; ++r) rows_avail += (p.now + r * p.check_res < p.Tt) ? 1 : 0;   // rows that exist in the scenario table
  const uint32_t knot_bytes = (uint32_t)p.kpad * 8u, coef_bytes = (uint32_t)p.kpad * 64u, time_bytes = (uint32_t)((p.P + 1) & ~1) * 8u;
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid < 32) {
    if (tid == 0) {
      mbar_arrive_expect_tx(bar, knot_bytes + coef_bytes + time_bytes + row_bytes * (uint32_t)rows_avail);
      bulk_copy_g2s(sm_knot, p.knot, knot_bytes, bar);
      bulk_copy_g2s(sm_coef, p.coef, coef_bytes, bar);
      bulk_copy_g2s(sm_time, p.time, time_bytes, bar);
    }
    __syncwarp();
    for (int r = tid; r < rows_avail; r += 32)
      bulk_copy_g2s(sm_broad + (size_t)r * p.OMpad, p.broad + (size_t)(p.now + r * p.check_res) * p.OMpad, row_bytes, bar);
  }
  if (kSmem) for (int e = rows_avail * p.OMpad + tid; e < p.win_rows * p.OMpad; e += nt) sm_broad[e] = make_float4(0.f, 0.f, FAR_AWAY, 0.f);
  mbar_wait(bar, 0);
  for (int e = tid; e < p.W * p.P; e += nt) {
    const int w = e / p.P, i = e - w * p.P;
    double d, d1, d2, d3;
    quintic_eval(p.quintic[w], sm_time[i], d, d1, d2, d3);
    double* o = sm_lat + (size_t)e * 4;
    o[0] = d; o[1] = d1; o[2] = d2; o[3] = d3;
  }
  __syncthreads();
  // per-width lateral cost sums: one warp per width, lane l adds samples l, l + 32, ... in time order, then a fixed
  // shuffle tree (deterministic; every kernel stages its tables through this function, so all of them see the same sums)
  for (int w = tid >> 5; w < p.W; w += nt >> 5) {
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = tid & 31; i < p.P; i += 32) {
      const double* o = sm_lat + ((size_t)w * p.P + i) * 4;
      const double e1 = fmax(fabs(o[2]) - p.thr_lat_a, 0.0), e2 = fmax(fabs(o[3]) - p.thr_lat_j, 0.0);
      s0 = s0 + o[0] * o[0]; s1 = s1 + e1 * e1; s2 = s2 + e2 * e2;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s0 = s0 + __shfl_down_sync(0xffffffffu, s0, off);
      s1 = s1 + __shfl_down_sync(0xffffffffu, s1, off);
      s2 = s2 + __shfl_down_sync(0xffffffffu, s2, off);
    }
    if ((tid & 31) == 0) { sm_latcost[w * 4] = s0; sm_latcost[w * 4 + 1] = s1; sm_latcost[w * 4 + 2] = s2; }
  }
  __syncthreads();
  if (kSmem) {
    tb.broad = sm_broad; tb.broad_row0 = 0; tb.broad_stride = 1;
  } else {
    tb.broad = p.broad; tb.broad_row0 = p.now; tb.broad_stride = p.check_res;
  }
  tb.sp.knot = sm_knot; tb.sp.coef = sm_coef; tb.sp.K = p.K; tb.sp.kpad = p.kpad;
  tb.sp.inv_h = (double)(p.K - 1) / sm_knot[p.K - 1];
  tb.time = sm_time; tb.lat = sm_lat;
  tb.list = sm_broad; tb.row_off = reinterpret_cast<const int*>(smem + L.off_rows); tb.row_cnt = tb.row_off + p.win_rows;
  tb.use_lists = false;
  latcost = sm_latcost;
}

// Mode 2, after stage_tables: cull the obstacle states of this cycle against the arc-length range the block's candidates
// cover at every collision row, and compact the survivors into shared memory (order within a row = obstacle-mode order).
//   phase A  every thread unrolls its longitudinal profile at the collision steps; warp REDUX + shared atomics give
//            [s_min, s_max] per row (fp32, rounded outwards, clamped to the reference line);
//   phase B  one warp per row: ballot-compaction of the states whose reachable interval (pack_obstacles_kernel)
//            overlaps the row's range.
// Falls back to the global table (tb.use_lists stays false) when the lists would not fit.
// Phase A lets every thread scan the rows row_first, row_first + row_stride, ... of ITS seed: (0, 1) when every thread has a
// seed of its own, (lane-in-group, W) when the W lanes of a group share one.
__device__ __forceinline__ void build_lists(const EvalParams& p, unsigned char* smem, BlockTables& tb, LonWalk lp, bool active,
                                            int row_first = 0, int row_stride = 1) {
  const SmemLayout L = eval_smem_layout(p.win_rows, p.OMpad, 2, p.K, p.kpad, p.P, p.W, p.list_cap);
  float4* list = reinterpret_cast<float4*>(smem + L.off_broad);
  int* row_off = reinterpret_cast<int*>(smem + L.off_rows);
  int* row_cnt = row_off + p.win_rows;
  unsigned int* row_min = reinterpret_cast<unsigned int*>(row_cnt + p.win_rows);
  unsigned int* row_max = row_min + p.win_rows;
  int* cursor = reinterpret_cast<int*>(smem);        // [0] next free list entry, [1] overflow flag
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  for (int r = tid; r < p.win_rows; r += blockDim.x) { row_min[r] = 0x7f800000u; row_max[r] = 0u; row_cnt[r] = 0; row_off[r] = 0; }
  if (tid == 0) { cursor[0] = 0; cursor[1] = 0; }
  __syncthreads();
  const double s_end = tb.sp.knot[p.K - 1];
  if (row_stride == 1) {
    for (int r = 0; r < p.win_rows; ++r) {
      unsigned int lo = 0x7f800000u, hi = 0u;
      if (active) {
        double s, sd, sdd, sddd;
        lon_walk_eval(lp, tb.time[r * p.check_res], s, sd, sdd, sddd);
        s = fmin(fmax(s, 0.0), s_end);                  // off-line samples are never collision-checked
        lo = __float_as_uint(__double2float_rd(s));     // s >= 0: the bit patterns order like the values
        hi = __float_as_uint(__double2float_ru(s));
      }
      lo = __reduce_min_sync(0xffffffffu, lo);
      hi = __reduce_max_sync(0xffffffffu, hi);
      if (lane == 0) { atomicMin(&row_min[r], lo); atomicMax(&row_max[r], hi); }
    }
  } else {
    // the lanes of a group share the seed: lane j of the group evaluates the rows r = j (mod W); in one round every lane
    // with the same j holds the same row, so a stride-W shuffle tree leaves the warp's range in the lanes of group 0
    const int W = row_stride;
    for (int r0 = 0; r0 < p.win_rows; r0 += W) {
      const int r = r0 + row_first;
      unsigned int lo = 0x7f800000u, hi = 0u;
      if (active && r < p.win_rows) {
        double s, sd, sdd, sddd;
        lon_walk_eval(lp, tb.time[r * p.check_res], s, sd, sdd, sddd);
        s = fmin(fmax(s, 0.0), s_end);
        lo = __float_as_uint(__double2float_rd(s));
        hi = __float_as_uint(__double2float_ru(s));
      }
      for (int off = W; off < 32; off <<= 1) {
        const unsigned int lo2 = __shfl_down_sync(0xffffffffu, lo, off), hi2 = __shfl_down_sync(0xffffffffu, hi, off);
        if (lane + off < 32) { lo = min(lo, lo2); hi = max(hi, hi2); }
      }
      if (lane < W && r < p.win_rows) { atomicMin(&row_min[r], lo); atomicMax(&row_max[r], hi); }
    }
  }
  __syncthreads();
  constexpr int CH = 8;                               // chunks of 32 obstacle states held in registers per pass
  for (int r = wid; r < p.win_rows; r += nw) {
    const int t = p.now + r * p.check_res;
    if (t >= p.Tt) continue;                          // no obstacle state exists beyond the scenario
    const float smin = __uint_as_float(row_min[r]) - 1e-3f, smax = __uint_as_float(row_max[r]) + 1e-3f;
    const float2* iv = p.sint + (size_t)t * p.OMpad;
    const float4* src = p.broad + (size_t)t * p.OMpad;
    if (p.OMpad <= 32 * CH) {
      // single pass: every load of the row is in flight at once
      float2 v[CH];
      float4 bb[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int e = c * 32 + lane;
        v[c] = e < p.OMpad ? __ldg(iv + e) : make_float2(FAR_AWAY, -FAR_AWAY);
        bb[c] = e < p.OMpad ? __ldg(src + e) : make_float4(0.f, 0.f, FAR_AWAY, 0.f);
      }
      unsigned int m[CH];
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c) { m[c] = __ballot_sync(0xffffffffu, v[c].y >= smin && v[c].x <= smax); cnt += __popc(m[c]); }
      int off = 0;
      if (lane == 0) off = atomicAdd(&cursor[0], cnt);
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off + cnt > p.list_cap) { if (lane == 0) cursor[1] = 1; continue; }
      if (lane == 0) { row_off[r] = off; row_cnt[r] = cnt; }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if ((m[c] >> lane) & 1u) {
          float4 b = bb[c];
          b.w = __int_as_float(c * 32 + lane);
          list[off + __popc(m[c] & ((1u << lane) - 1u))] = b;
        }
        off += __popc(m[c]);
      }
      continue;
    }
    int cnt = 0;
    for (int base = 0; base < p.OMpad; base += 32) {
      const float2 v = iv[base + lane];
      cnt += __popc(__ballot_sync(0xffffffffu, v.y >= smin && v.x <= smax));
    }
    int off = 0;
    if (lane == 0) off = atomicAdd(&cursor[0], cnt);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + cnt > p.list_cap) { if (lane == 0) cursor[1] = 1; continue; }
    if (lane == 0) { row_off[r] = off; row_cnt[r] = cnt; }
    for (int base = 0; base < p.OMpad; base += 32) {
      const float2 v = iv[base + lane];
      const bool keep = v.y >= smin && v.x <= smax;
      const unsigned int m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        float4 b = src[base + lane];
        b.w = __int_as_float(base + lane);
        list[off + __popc(m & ((1u << lane) - 1u))] = b;
      }
      off += __popc(m);
    }
  }
  __syncthreads();
  tb.use_lists = cursor[1] == 0;
  tb.broad = p.broad; tb.broad_row0 = p.now; tb.broad_stride = p.check_res;   // fallback path
}

__device__ __forceinline__ int resolve_n_seeds(const EvalParams& p) {
  return p.n_seeds_ptr ? min(p.n_seeds_ptr[2], p.seed_cap) : p.n_seeds_fixed;
}

__device__ __forceinline__ bool lex_less(double ca, int ia, double cb, int ib) {
  // (cost, index) lexicographic; index -1 = empty
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  return ca < cb || (ca == cb && ia < ib);
}

// Export of the selected trajectory by one thread block: thread i owns sample i.  Positions come from the same
// lon_eval / frenet_point arithmetic as the evaluation walk (bisect and walk select the same spline segment), headings are
// the float64 arctan2 of the forward differences exactly as jerk_space_sampling_planner.py:155-165 derives them.
// `out` (global, [P][16]) doubles as the exchange buffer between the phases.
__device__ __forceinline__ void export_block_global(const EvalParams& p, const BlockTables& tb, int best, double* __restrict__ out,
                                                    int* sm_npts);

// Second half of the export, common to every candidate model: thread i holds row[0..8] of sample i (t, s.., d..) and
// *sm_npts the index of the first sample off the reference line (P if none).  Positions, forward-difference headings with
// the last one repeated, ds, c, c_d, c_dd exactly as jerk_space_sampling_planner.py:131-165; neighbours through `ex`.
constexpr int EXPORT_MAX_P = 128;
__device__ __forceinline__ void export_tail(const BlockTables& tb, double (&row)[JSSP_TRAJ_COLS], int P, double d_i, double dt,
                                            double* __restrict__ out, int* sm_npts, BestRecord* best_rec, double (*ex)[EXPORT_MAX_P]) {
  const int i = threadIdx.x;
  __syncthreads();
  const int np = *sm_npts;
  if (i < np) {
    frenet_point(tb.sp, spline_bisect(tb.sp, row[1]), row[1], d_i, row[9], row[10]);
    ex[0][i] = row[9]; ex[1][i] = row[10];
  }
  __syncthreads();
  if (i + 1 < np) {
    const double dx = ex[0][i + 1] - row[9], dy = ex[1][i + 1] - row[10];
    const double q = dx * dx + dy * dy;
    row[11] = heading_of(dx, dy);
    row[12] = (q > 1e-200 && q < 1e200) ? q * rsqrt(q) : hypot(dx, dy);
    ex[2][i] = row[11];
  }
  __syncthreads();
  if (np >= 2 && i == np - 1) { row[11] = ex[2][np - 2]; ex[2][i] = row[11]; }     // yaw[-1] repeated (:158-160)
  __syncthreads();
  if (i + 1 < np) { row[13] = (ex[2][i + 1] - row[11]) / row[12]; ex[4][i] = row[13]; }
  __syncthreads();
  if (i + 2 < np) { row[14] = (ex[4][i + 1] - row[13]) / dt; ex[5][i] = row[14]; }
  __syncthreads();
  if (i + 3 < np) row[15] = (ex[5][i + 1] - row[14]) / dt;
  if (i < P) {
    double* r = out + (size_t)i * JSSP_TRAJ_COLS;
#pragma unroll
    for (int c = 0; c < JSSP_TRAJ_COLS; ++c) r[c] = row[c];
  }
  if (i == 0) best_rec->n_pts = np;
}

// Fast form for P <= EXPORT_MAX_P samples and blockDim >= P: thread i owns sample i, neighbours are exchanged through
// shared memory and every row is written to global memory once.  Same expressions as export_block_global.
__device__ __forceinline__ void export_block(const EvalParams& p, const BlockTables& tb, int best, double* __restrict__ out,
                                             int* sm_npts) {
  const int P = p.P, i = threadIdx.x;
  if (P > EXPORT_MAX_P || (int)blockDim.x < P) { export_block_global(p, tb, best, out, sm_npts); return; }
  __shared__ double ex[6][EXPORT_MAX_P];          // x, y, yaw, ds, c, c_d
  const int Ns = resolve_n_seeds(p);
  const int w = best / Ns, k = best - w * Ns;
  const double* lat = tb.lat + (size_t)w * P * 4;
  if (i == 0) *sm_npts = P;
  __syncthreads();
  const double nanv = nan("");
  double row[JSSP_TRAJ_COLS];
  if (i < P) {
    LonProfile lp;
    lon_profile_init(lp, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
    lon_eval(lp, p.s0, p.v0, p.a0, tb.time[i], row[1], row[2], row[3], row[4]);
    row[0] = tb.time[i];
    row[5] = lat[i * 4]; row[6] = lat[i * 4 + 1]; row[7] = lat[i * 4 + 2]; row[8] = lat[i * 4 + 3];
#pragma unroll
    for (int c = 9; c < JSSP_TRAJ_COLS; ++c) row[c] = nanv;
    if (!spline_in_range(tb.sp, row[1])) atomicMin(sm_npts, i);   // first sample off the reference line ends the Cartesian path (:138-139)
  }
  export_tail(tb, row, P, i < P ? lat[i * 4] : 0.0, tb.time[1] - tb.time[0], out, sm_npts, p.best, ex);
}

__device__ __forceinline__ void export_block_global(const EvalParams& p, const BlockTables& tb, int best, double* __restrict__ out,
                                                    int* sm_npts) {
  const int P = p.P, i = threadIdx.x;
  const int Ns = resolve_n_seeds(p);
  const int w = best / Ns, k = best - w * Ns;
  const double* lat = tb.lat + (size_t)w * P * 4;
  if (i == 0) *sm_npts = P;
  __syncthreads();
  double s = 0.0;
  bool inr = false;
  for (int j = i; j < P; j += blockDim.x) {      // blockDim >= P in practice; the loop keeps small blocks correct
    LonProfile lp;
    lon_profile_init(lp, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
    double sd, sdd, sddd;
    lon_eval(lp, p.s0, p.v0, p.a0, tb.time[j], s, sd, sdd, sddd);
    double* r = out + (size_t)j * JSSP_TRAJ_COLS;
    r[0] = tb.time[j]; r[1] = s; r[2] = sd; r[3] = sdd; r[4] = sddd;
    r[5] = lat[j * 4]; r[6] = lat[j * 4 + 1]; r[7] = lat[j * 4 + 2]; r[8] = lat[j * 4 + 3];
    for (int c = 9; c < JSSP_TRAJ_COLS; ++c) r[c] = nan("");
    inr = spline_in_range(tb.sp, s);
    if (!inr) atomicMin(sm_npts, j);             // first sample off the reference line ends the Cartesian path (:138-139)
  }
  __syncthreads();
  const int np = *sm_npts;
  for (int j = i; j < np; j += blockDim.x) {
    const double sj = out[(size_t)j * JSSP_TRAJ_COLS + 1];
    double x, y;
    frenet_point(tb.sp, spline_bisect(tb.sp, sj), sj, lat[j * 4], x, y);
    out[(size_t)j * JSSP_TRAJ_COLS + 9] = x; out[(size_t)j * JSSP_TRAJ_COLS + 10] = y;
  }
  __syncthreads();
  for (int j = i; j + 1 < np; j += blockDim.x) {
    const double dx = out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 9] - out[(size_t)j * JSSP_TRAJ_COLS + 9];
    const double dy = out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 10] - out[(size_t)j * JSSP_TRAJ_COLS + 10];
    const double q = dx * dx + dy * dy;
    out[(size_t)j * JSSP_TRAJ_COLS + 11] = heading_of(dx, dy);
    out[(size_t)j * JSSP_TRAJ_COLS + 12] = (q > 1e-200 && q < 1e200) ? q * rsqrt(q) : hypot(dx, dy);
  }
  __syncthreads();
  if (i == 0 && np >= 2) out[(size_t)(np - 1) * JSSP_TRAJ_COLS + 11] = out[(size_t)(np - 2) * JSSP_TRAJ_COLS + 11];   // yaw[-1] repeated
  __syncthreads();
  for (int j = i; j + 1 < np; j += blockDim.x)
    out[(size_t)j * JSSP_TRAJ_COLS + 13] = (out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 11] - out[(size_t)j * JSSP_TRAJ_COLS + 11]) / out[(size_t)j * JSSP_TRAJ_COLS + 12];
  __syncthreads();
  const double dt = tb.time[1] - tb.time[0];
  for (int j = i; j + 2 < np; j += blockDim.x)
    out[(size_t)j * JSSP_TRAJ_COLS + 14] = (out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 13] - out[(size_t)j * JSSP_TRAJ_COLS + 13]) / dt;
  __syncthreads();
  for (int j = i; j + 3 < np; j += blockDim.x)
    out[(size_t)j * JSSP_TRAJ_COLS + 15] = (out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 14] - out[(size_t)j * JSSP_TRAJ_COLS + 14]) / dt;
  if (i == 0) p.best->n_pts = np;
}

// Closed-set overlap of two centred rectangles, float64 (ReplayController._collision_detect, controller.py:382-415).
__device__ __forceinline__ bool rect_overlap_f64(double x1, double y1, double yaw1, double l1, double w1, double x2, double y2,
                                                 double yaw2, double l2, double w2) {
  double s1, c1, s2, c2;
  sincos(yaw1, &s1, &c1);
  sincos(yaw2, &s2, &c2);
  return sat_intersects_f64(x2 - x1, y2 - y1, c1, s1, l1 / 2.0, w1 / 2.0, c2, s2, l2 / 2.0, w2 / 2.0);
}

// Batched replay: end-of-cycle hand-off of one scenario, run by the block that exported the selected trajectory.
__device__ __forceinline__ void batch_handoff(const EvalParams& p, int best, int N, int* sm_flag) {
  BatchSlot* b = p.batch;
  __syncthreads();                                   // the exported rows (global) are visible to the whole block
  if (threadIdx.x == 0) *sm_flag = 0;
  __syncthreads();
  const int c = b->cycle;
  const int np = best >= 0 ? p.best->n_pts : 0;
  const bool ok = best >= 0 && np >= 2;              // "No solution" / env.py:26 needs best_traj.x[1]
  const double* row1 = p.best_traj + JSSP_TRAJ_COLS;
  double tn = 0.0, x = 0.0, y = 0.0, yaw = 0.0;
  if (ok) {
    tn = b->cycle_time[c + 1];                       // t + dt (controller.py:303)
    x = row1[9]; y = row1[10]; yaw = row1[11];
    const int k = b->cycle_end_step[c + 1];
    if (tn > 0.5 && k >= 0 && k < b->gt_Tt) {        // collision with a background vehicle (controller.py:343-349)
      for (int o = threadIdx.x; o < b->gt_O; o += blockDim.x) {
        if (!b->gt_valid[(size_t)o * b->gt_Tt + k]) continue;
        const double* st = b->gt_state + ((size_t)o * b->gt_Tt + k) * 3;
        if (rect_overlap_f64(x, y, yaw, b->ego_l, b->ego_w, st[0], st[1], st[2], b->gt_size[2 * o], b->gt_size[2 * o + 1])) atomicOr(sm_flag, 1);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  jssp_cycle_record* rec = b->records + c;
  rec->best_idx = best; rec->n_candidates = N; rec->best_cost = best >= 0 ? p.best->cost : nan("");
  if (!ok) {
    rec->t = rec->x = rec->y = rec->v = rec->a = rec->yaw = nan("");
    b->end = best < 0 ? -2.0 : -3.0;
    b->live = 0;
  } else {
    rec->t = tn; rec->x = x; rec->y = y; rec->v = row1[2]; rec->a = row1[3]; rec->yaw = yaw;
    double status = -1.0;
    if (*sm_flag) status = 2.0;
    if (tn >= b->max_t) status = fmax(status, 1.0);                                          // controller.py:352-353
    const double dx = x - b->goal_x, dy = y - b->goal_y;
    if (b->goal_sin * dy + b->goal_cos * dx >= 0.0) status = fmax(status, 1.5);               // goal line passed
    b->end = status;
    if (status != -1.0) b->live = 0;
    // next cycle starts from sample 1 of the selected trajectory (intersection_planner.py:312)
    EvalParams* q = p.self;
    q->s0 = row1[1]; q->v0 = row1[2]; q->a0 = row1[3];
    for (int w = 0; w < p.W; ++w) quintic_coeffs(row1[5], row1[6], row1[7], b->targets[w], 0.0, 0.0, b->horizon, q->quintic[w]);
    const int now = b->cycle_now[c + 1];
    q->now = now; q->tcap = b->final_step - now;
    int rows = 0;
    if (q->tcap > 0) rows = (min(p.P, q->tcap) + p.check_res - 1) / p.check_res;
    q->win_rows = p.OM > 0 ? rows : 0;
  }
  b->cycle = c + 1;
  if (c + 1 >= b->n_cycles_cap) b->live = 0;         // table end: stays "running" (end = -1) like a max_cycles stop
}

// Selection shared by both evaluation kernels: block argmin of (cost, index) by warp shuffles, one record per block,
// and the last block to finish (ticket counter) reduces the per-block winners - deterministic, lexicographic, no
// floating-point atomics - and exports the winning trajectory in the same launch.
template <int kMode>
__device__ __forceinline__ void select_and_export(const EvalParams& p, unsigned char* smem, BlockTables& tb, const double*& latcost,
                                                  bool staged, double my_cost, int my_idx, int N, int Ns, double* red_cost,
                                                  int* red_idx, bool* is_last_p, int* sm_npts_p) {
  const int nwarps = blockDim.x >> 5;
  // block argmin (cost, index) - warp shuffles, then one value per warp through shared memory
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, o);
    const int oi = __shfl_down_sync(0xffffffffu, my_idx, o);
    if (lex_less(oc, oi, my_cost, my_idx)) { my_cost = oc; my_idx = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red_cost[wid] = my_cost; red_idx[wid] = my_idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < nwarps; ++q)
      if (lex_less(red_cost[q], red_idx[q], my_cost, my_idx)) { my_cost = red_cost[q]; my_idx = red_idx[q]; }
    p.blk_cost[blockIdx.x] = my_cost;
    p.blk_idx[blockIdx.x] = my_idx;
    __threadfence();
    const unsigned int t = atomicAdd(p.ticket, 1u);
    *is_last_p = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!*is_last_p) return;
  // the last block to finish reduces the per-block winners: deterministic, no floating-point atomics
  __threadfence();
  my_cost = INFINITY; my_idx = -1;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
    const double c = __ldcg(p.blk_cost + b);
    const int i = __ldcg(p.blk_idx + b);
    if (lex_less(c, i, my_cost, my_idx)) { my_cost = c; my_idx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, o);
    const int oi = __shfl_down_sync(0xffffffffu, my_idx, o);
    if (lex_less(oc, oi, my_cost, my_idx)) { my_cost = oc; my_idx = oi; }
  }
  if (lane == 0) { red_cost[wid] = my_cost; red_idx[wid] = my_idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < nwarps; ++q)
      if (lex_less(red_cost[q], red_idx[q], my_cost, my_idx)) { my_cost = red_cost[q]; my_idx = red_idx[q]; }
    BestRecord r;
    r.key = orderable_bits(my_idx >= 0 ? my_cost : INFINITY);
    r.idx = my_idx;
    r.cost = my_cost;
    r.n_pts = 0;
    r.n_candidates = N;
    r.n_seeds = Ns;
    r.pad = 0;
    *p.best = r;
    *p.ticket = 0u;   // re-arm for the next launch (stream-ordered)
    red_idx[0] = my_idx;
  }
  // the same block exports the winner: no second launch, no host round trip in between
  if (p.best_traj == nullptr) return;
  __syncthreads();
  const int best = red_idx[0];
  if (best >= 0) {
    if (!staged) stage_tables<kMode>(p, smem, tb, latcost);
    export_block(p, tb, best, p.best_traj, sm_npts_p);
  }
  if (p.batch) batch_handoff(p, best, N, sm_npts_p);
}


template <int kMode>
__device__ __forceinline__ void evaluate_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx, bool* is_last,
                                              int* sm_npts) {
  const int Ns = resolve_n_seeds(p);
  const int N = Ns * p.W;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double my_cost = INFINITY;
  int my_idx = -1;
  BlockTables tb;
  const double* latcost;
  const bool staged = blockIdx.x * blockDim.x < N;   // block-uniform
  if (staged) {
    stage_tables<kMode>(p, smem, tb, latcost);
    const bool active = n < N;
    const int w = active ? n / Ns : 0, k = active ? n - w * Ns : 0;
    LonWalk lp;
    lon_walk_init(lp, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
    if (kMode == 2) build_lists(p, smem, tb, lp, active);
    if (active) {
      EvalSink sink(p, tb);
      const int seg0 = spline_bisect(tb.sp, fmin(fmax(p.s0, tb.sp.knot[0]), tb.sp.knot[p.K - 1]));
      walk_candidate(p, tb, lp, w, seg0, sink);
      const double cost = candidate_cost(p, sink, latcost + w * 4);
      if (p.feasible) p.feasible[n] = sink.curv_fail ? 0 : 1;
      if (p.collision) p.collision[n] = sink.collided ? 1 : 0;
      if (p.cost) p.cost[n] = (float)cost;
      if (!sink.curv_fail && !sink.collided && isfinite(cost)) { my_cost = cost; my_idx = n; }
    }
  }
  select_and_export<kMode>(p, smem, tb, latcost, staged, my_cost, my_idx, N, Ns, red_cost, red_idx, is_last, sm_npts);
}

template <int kMode>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  evaluate_body<kMode>(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

// K2 (default for large cycles): the group form - W consecutive lanes evaluate the W lateral targets of one seed and share
// the seed-only work (walk_group).  Seeds per block = warps * floor(32 / W); lanes beyond floor(32 / W) * W idle.
__host__ __device__ inline int group_seeds_per_block(int W) { return (EVAL_THREADS / 32) * (32 / W); }

template <int kMode>
__device__ __forceinline__ void evaluate_group_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx,
                                                    bool* is_last, int* sm_npts) {
  const int Ns = resolve_n_seeds(p);
  const int W = p.W;
  const int N = Ns * W;
  const int gpw = 32 / W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane / W, j = lane - g * W;
  const int spb = (EVAL_THREADS / 32) * gpw;
  const int k = blockIdx.x * spb + wid * gpw + g;
  const bool active = g < gpw && k < Ns;
  double my_cost = INFINITY;
  int my_idx = -1;
  BlockTables tb;
  const double* latcost;
  const bool staged = blockIdx.x * spb < Ns;   // block-uniform
  PHASE(p, 0);
  if (staged) {
    stage_tables<kMode>(p, smem, tb, latcost);
    PHASE(p, 1);
    LonWalk lp;
    lon_walk_init(lp, p.seeds + (size_t)(active ? k : 0) * 10, p.s0, p.v0, p.a0);
    if (kMode == 2) build_lists(p, smem, tb, lp, active, j, W);
    PHASE(p, 2);
    if (active) {
      const unsigned int gmask = (W >= 32 ? 0xffffffffu : ((1u << W) - 1u)) << (g * W);
      EvalSink sink(p, tb);
      walk_group(p, tb, lp, j, W, gmask, g * W, sink);
      const int n = j * Ns + k;                // candidate index: width-major (jerk_space_sampling_planner.py:106,124-125)
      const double cost = candidate_cost(p, sink, latcost + j * 4);
      if (p.feasible) p.feasible[n] = sink.curv_fail ? 0 : 1;
      if (p.collision) p.collision[n] = sink.collided ? 1 : 0;
      if (p.cost) p.cost[n] = (float)cost;
      if (!sink.curv_fail && !sink.collided && isfinite(cost)) { my_cost = cost; my_idx = n; }
    }
  }
  __syncthreads();
  PHASE(p, 3);
  select_and_export<kMode>(p, smem, tb, latcost, staged, my_cost, my_idx, N, Ns, red_cost, red_idx, is_last, sm_npts);
  PHASE(p, 4);
}

template <int kMode>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_group_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  evaluate_group_body<kMode>(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

// Batched replay: blockIdx.y = scenario; the scenario's parameter record is copied into shared memory once per block.
__device__ __forceinline__ void load_slot(EvalParams* dst, const EvalParams* src) {
  static_assert(sizeof(EvalParams) % 8 == 0, "EvalParams is copied as 8-byte words");
  const unsigned long long* s = reinterpret_cast<const unsigned long long*>(src);
  unsigned long long* d = reinterpret_cast<unsigned long long*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(EvalParams) / 8); i += blockDim.x) d[i] = s[i];
  __syncthreads();
}

template <int kMode>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_batch_kernel(const EvalParams* __restrict__ slots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  if (!slot_live(slots, blockIdx.y)) return;           // block-uniform; stable until the scenario's last block hands off
  load_slot(&sp, slots + blockIdx.y);
  evaluate_body<kMode>(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

template <int kMode>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_group_batch_kernel(const EvalParams* __restrict__ slots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  if (!slot_live(slots, blockIdx.y)) return;
  load_slot(&sp, slots + blockIdx.y);
  evaluate_group_body<kMode>(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

// ------------------------------------------------------------------------------------------------------------------
// K2b: the same evaluation with ONE WARP PER CANDIDATE for small cycles (the default grid yields ~10^3 candidates: with a
// thread per candidate the kernel's duration is one thread walking 51 samples serially, ~100 us).  Lanes own samples
// i = lane, lane + 32, ...; positions, segments and per-sample cost terms go through shared memory, curvature is tested in
// parallel over segments, each collision row tests 32 obstacle states at once (ballot), and lane 0 adds the cost terms
// in time order so that the cost is bit-identical to the thread-per-candidate kernel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SMALL_WARPS = 8;                 // candidates per block
constexpr int SMALL_ROW = 10;                  // doubles per sample in the warp scratch: x, y | dx, dy, ds, c, sn | ea2, ej2, ev2
__host__ __device__ inline size_t small_scratch_bytes(int P) { return (size_t)SMALL_WARPS * ((size_t)P * SMALL_ROW + 2) * 8; }

__device__ __forceinline__ void evaluate_small_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx,
                                                    bool* is_last_p, int* sm_npts_p) {
  const int Ns = resolve_n_seeds(p);
  const int N = Ns * p.W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = blockIdx.x * SMALL_WARPS + wid;
  const int P = p.P;
  double my_cost = INFINITY;
  int my_idx = -1;
  BlockTables tb;
  const double* latcost;
  const bool staged = blockIdx.x * SMALL_WARPS < N;   // block-uniform
  if (staged) {
    stage_tables<0>(p, smem, tb, latcost);
    const SmemLayout L = eval_smem_layout(0, p.OMpad, 0, p.K, p.kpad, p.P, p.W);
    double* scr = reinterpret_cast<double*>(smem + L.total) + (size_t)wid * ((size_t)P * SMALL_ROW + 2);
    double* ends = scr + (size_t)P * SMALL_ROW;        // s at the first / last sample
    if (n < N) {                                       // warp-uniform
      const int w = n / Ns, k = n - w * Ns;
      const double* lat = tb.lat + (size_t)w * P * 4;
      LonProfile lp;
      lon_profile_init(lp, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
      // 1) samples: Frenet state, cost terms, position; n_pts = first sample off the reference line
      int n_pts = P;
      for (int base = 0; base < P; base += 32) {
        const int i = base + lane;
        bool off = false;
        if (i < P) {
          double s, sd, sdd, sddd;
          lon_eval(lp, p.s0, p.v0, p.a0, tb.time[i], s, sd, sdd, sddd);
          const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0), ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0), ev = (p.vmax - sd) * p.inv_vmax;
          double* r = scr + (size_t)i * SMALL_ROW;
          r[7] = ea * ea; r[8] = ej * ej; r[9] = ev * ev;
          if (i == 0) ends[0] = s;
          if (i == P - 1) ends[1] = s;
          off = !spline_in_range(tb.sp, s);
          if (!off) frenet_point(tb.sp, spline_walk(tb.sp, s, 0), s, lat[i * 4], r[0], r[1]);
        }
        const unsigned int m = __ballot_sync(0xffffffffu, off);
        if (m != 0u && n_pts == P) n_pts = base + __ffs(m) - 1;
      }
      __syncwarp();
      // 2) forward differences p_{i+1} - p_i: length and unit vector (same expressions as walk_candidate)
      for (int base = 0; base + 1 < n_pts; base += 32) {
        const int i = base + lane;
        if (i + 1 < n_pts) {
          double* r = scr + (size_t)i * SMALL_ROW;
          const double dx = r[SMALL_ROW] - r[0], dy = r[SMALL_ROW + 1] - r[1];
          const double q = dx * dx + dy * dy;
          double ds, c = 1.0, sn = 0.0;
          if (q > 1e-200 && q < 1e200) { const double inv = rsqrt(q); ds = q * inv; c = dx * inv; sn = dy * inv; }
          else { ds = hypot(dx, dy); if (ds > 0.0) { c = dx / ds; sn = dy / ds; } }
          r[2] = dx; r[3] = dy; r[4] = ds; r[5] = c; r[6] = sn;
        }
      }
      __syncwarp();
      // 3) curvature of every interior segment pair
      bool fail = false;
      for (int base = 0; base + 2 < n_pts; base += 32) {
        const int i = base + lane + 1;                 // segment i against segment i - 1
        if (i + 1 < n_pts) {
          const double* a = scr + (size_t)(i - 1) * SMALL_ROW;
          const double* b = scr + (size_t)i * SMALL_ROW;
          fail |= curvature_fails(p.kmax, a[2], a[3], a[4], a[5], a[6], b[2], b[3], b[4], b[5], b[6]);
        }
      }
      fail = __any_sync(0xffffffffu, fail);
      // 4) collision rows, 32 obstacle states at a time
      bool collided = false;
      double risk = 0.0;
      if (p.n_dict > 0) {
        if (n_pts == 1 && p.tcap >= 1) collided = true;                 // bare except -> collision (:250-254)
        const int lim = n_pts >= 2 ? min(n_pts, p.tcap) : 0;
        EvalSink probe(p, tb);                                          // for narrow_hit()
        for (int i = 0; i < lim && !collided; i += p.check_res) {
          const int abs_step = p.now + i;
          if (abs_step >= p.Tt) break;
          const double* r = scr + (size_t)i * SMALL_ROW;
          const double* h = scr + (size_t)(i == n_pts - 1 ? i - 1 : i) * SMALL_ROW;   // last point reuses the previous heading
          const double x = r[0], y = r[1], c = h[5], sn = h[6];
          const float fx = (float)(x - p.ox), fy = (float)(y - p.oy);
          const float m2x = -2.0f * fx, m2y = -2.0f * fy, ne2 = -fmaf(fx, fx, fy * fy);
          const float4* row = p.broad + (size_t)abs_step * p.OMpad;
          for (int base = 0; base < p.OMpad; base += 32) {
            const float4 b = __ldg(row + base + lane);
            bool hit = fmaf(b.x, m2x, fmaf(b.y, m2y, b.z)) <= ne2;
            if (hit) hit = probe.narrow_hit(abs_step, base + lane, b.x - fx, b.y - fy, x, y, c, sn);
            const bool hard = hit && (double)b.w >= p.p_min;
            if (__any_sync(0xffffffffu, hard)) { collided = true; break; }
            unsigned int soft = __ballot_sync(0xffffffffu, hit);
            while (soft) {                                              // risk in obstacle order, like the serial walk
              const int j = __ffs(soft) - 1;
              soft &= soft - 1u;
              risk = risk + (double)__shfl_sync(0xffffffffu, b.w, j);
            }
          }
        }
      }
      // 5) cost: lane 0 adds the per-sample terms in time order
      if (lane == 0) {
        double sum_a = 0, sum_j = 0, sum_v = 0;
        for (int i = 0; i < P; ++i) {
          const double* r = scr + (size_t)i * SMALL_ROW;
          sum_a = sum_a + r[7]; sum_j = sum_j + r[8]; sum_v = sum_v + r[9];
        }
        const double cost = candidate_cost(p, sum_a, sum_j, sum_v, ends[0], ends[1], risk, latcost + w * 4);
        if (p.feasible) p.feasible[n] = fail ? 0 : 1;
        if (p.collision) p.collision[n] = collided ? 1 : 0;
        if (p.cost) p.cost[n] = (float)cost;
        if (!fail && !collided && isfinite(cost)) { my_cost = cost; my_idx = n; }
      }
    }
  }
  select_and_export<0>(p, smem, tb, latcost, staged, my_cost, my_idx, N, Ns, red_cost, red_idx, is_last_p, sm_npts_p);
}

__global__ void __launch_bounds__(SMALL_WARPS * 32) evaluate_small_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[SMALL_WARPS];
  __shared__ int red_idx[SMALL_WARPS];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  evaluate_small_body(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

__global__ void __launch_bounds__(SMALL_WARPS * 32) evaluate_small_batch_kernel(const EvalParams* __restrict__ slots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[SMALL_WARPS];
  __shared__ int red_idx[SMALL_WARPS];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  if (!slot_live(slots, blockIdx.y)) return;
  load_slot(&sp, slots + blockIdx.y);
  evaluate_small_body(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

// ------------------------------------------------------------------------------------------------------------------
// K6: the FOP baseline planner's plan cycle through the same stages (SURVEY.md section 8 row f3;
// frenet_optimal_planner.py:97-144,146-183,185-225,249-299,328-333).  Candidate n = (w * num_t + it) * num_speed + iv
// (the reference's loop order): lateral quintic to target w over horizon T_it, longitudinal QUARTIC from (s, s_d, s_dd) to
// (speed_iv, 0) over T_it, samples t = arange(0, T_it, dt) (per-candidate length), Frenet->Cartesian as for JSSP, feasibility
// = |s_dd| <= max_accel and |s_ddd| <= max_jerk at every sample (the curvature test is commented out in the reference),
// rectangle collision every 2nd sample (the handle is created with obstacle_inflation = 0, :280), cost = the JSSP terms over
// the candidate's own samples + w_time * (1 - (T - t_min) / (t_max - t_min)), best = LAST minimum (`min_cost >= cost`).
// One thread per candidate; QuarticPolynomial / the cost class are absent from the reference tree (builder-specified,
// oracle/fop_oracle.py).
// ------------------------------------------------------------------------------------------------------------------
struct FopParams {
  EvalParams base;            // constants, tables, obstacle pointers, outputs; W = num_width, P = longest sample count
  int num_t, num_speed;
  const double* horizon;      // [num_t]  T_it = linspace(min_t, max_t, num_t)
  const int* n_samples;       // [num_t]  len(arange(0, T_it, dt))
  const double* speed;        // [num_speed] target speeds = linspace(lowest, highest, num_speed)
  double d0, d_d0, d_dd0;
  double targets[JSSP_MAX_WIDTHS];
  double dt, max_accel, max_jerk, t_min, t_max;
};

struct QuarticLon {           // QuarticPolynomial(xs, vxs, axs, vxe, axe, T), closed form of the PythonRobotics 2x2 system
  double a0, a1, a2, a3, a4;
  __device__ __forceinline__ void init(double xs, double vxs, double axs, double vxe, double axe, double T) {
    a0 = xs; a1 = vxs; a2 = axs / 2.0;
    const double b0 = vxe - a1 - 2.0 * a2 * T, b1 = axe - 2.0 * a2;
    a3 = b0 / (T * T) - b1 / (3.0 * T);
    a4 = b1 / (4.0 * T * T) - b0 / (2.0 * T * T * T);
  }
  __device__ __forceinline__ void eval(double t, double& s, double& sd, double& sdd, double& sddd) const {
    const double t2 = t * t, t3 = t2 * t, t4 = t3 * t;
    s = a0 + a1 * t + a2 * t2 + a3 * t3 + a4 * t4;
    sd = a1 + 2 * a2 * t + 3 * a3 * t2 + 4 * a4 * t3;
    sdd = 2 * a2 + 6 * a3 * t + 12 * a4 * t2;
    sddd = 6 * a3 + 24 * a4 * t;
  }
};

struct FopSink : EvalSink {   // same collision / cost accumulation; accel and jerk limits instead of the curvature test
  double max_accel, max_jerk;
  __device__ __forceinline__ FopSink(const EvalParams& p_, const BlockTables& tb_, double ma, double mj) : EvalSink(p_, tb_), max_accel(ma), max_jerk(mj) {}
  __device__ __forceinline__ void sample(int i, double t, double s, double sd, double sdd, double sddd) {
    EvalSink::sample(i, t, s, sd, sdd, sddd);
    if (fabs(sdd) > max_accel || fabs(sddd) > max_jerk) curv_fail = true;                  // :211-216
  }
  __device__ __forceinline__ void segment(int i, double x, double y, double dx, double dy, double ds, double c, double sn) {
    collide(i, x, y, c, sn);
    pdx = dx; pdy = dy; pds = ds; pc = c; psn = sn;
  }
};

struct FopSmem { size_t off_knot, off_coef, off_time, off_quint, off_lat, off_latcost, total; };
__host__ __device__ inline FopSmem fop_smem_layout(int K, int kpad, int P, int W, int num_t) {
  FopSmem L;
  size_t o = 16;
  L.off_knot = o; o += (size_t)kpad * 8;
  L.off_coef = o; o += (size_t)8 * kpad * 8;
  L.off_time = o; o += (size_t)((P + 1) & ~1) * 8;
  L.off_quint = o; o += (size_t)W * num_t * 6 * 8;
  L.off_lat = o; o += (size_t)W * num_t * P * 4 * 8;
  L.off_latcost = o; o += (size_t)W * num_t * 4 * 8;
  L.total = o;
  return L;
}

// tables of a FOP block: spline, time samples i * dt, one lateral quintic per (width, horizon) sampled over its own length
__device__ __forceinline__ void fop_stage(const FopParams& fp, unsigned char* smem, BlockTables& tb, const double*& latcost, const double*& quint) {
  const EvalParams& p = fp.base;
  const FopSmem L = fop_smem_layout(p.K, p.kpad, p.P, p.W, fp.num_t);
  double* sm_knot = reinterpret_cast<double*>(smem + L.off_knot);
  double* sm_coef = reinterpret_cast<double*>(smem + L.off_coef);
  double* sm_time = reinterpret_cast<double*>(smem + L.off_time);
  double* sm_quint = reinterpret_cast<double*>(smem + L.off_quint);
  double* sm_lat = reinterpret_cast<double*>(smem + L.off_lat);
  double* sm_latcost = reinterpret_cast<double*>(smem + L.off_latcost);
  const int tid = threadIdx.x, nt = blockDim.x, P = p.P, nq = p.W * fp.num_t;
  for (int e = tid; e < p.K; e += nt) sm_knot[e] = p.knot[e];
  for (int e = tid; e < 8 * p.kpad; e += nt) sm_coef[e] = p.coef[e];
  for (int e = tid; e < P; e += nt) sm_time[e] = 0.0 + (double)e * fp.dt;                  // np.arange(0.0, T, dt)
  for (int e = tid; e < nq; e += nt) {
    const int w = e / fp.num_t, it = e - w * fp.num_t;
    quintic_coeffs(fp.d0, fp.d_d0, fp.d_dd0, fp.targets[w], 0.0, 0.0, fp.horizon[it], sm_quint + (size_t)e * 6);
  }
  __syncthreads();
  for (int e = tid; e < nq * P; e += nt) {
    const int q = e / P, i = e - q * P;
    double* o = sm_lat + (size_t)e * 4;
    if (i < fp.n_samples[q % fp.num_t]) quintic_eval(sm_quint + (size_t)q * 6, sm_time[i], o[0], o[1], o[2], o[3]);
    else { o[0] = o[1] = o[2] = o[3] = 0.0; }
  }
  __syncthreads();
  for (int q = tid >> 5; q < nq; q += nt >> 5) {           // lateral cost sums over the candidate's own samples (stage_tables' order)
    const int n = fp.n_samples[q % fp.num_t];
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = tid & 31; i < n; i += 32) {
      const double* o = sm_lat + ((size_t)q * P + i) * 4;
      const double e1 = fmax(fabs(o[2]) - p.thr_lat_a, 0.0), e2 = fmax(fabs(o[3]) - p.thr_lat_j, 0.0);
      s0 = s0 + o[0] * o[0]; s1 = s1 + e1 * e1; s2 = s2 + e2 * e2;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s0 = s0 + __shfl_down_sync(0xffffffffu, s0, off);
      s1 = s1 + __shfl_down_sync(0xffffffffu, s1, off);
      s2 = s2 + __shfl_down_sync(0xffffffffu, s2, off);
    }
    if ((tid & 31) == 0) { sm_latcost[q * 4] = s0; sm_latcost[q * 4 + 1] = s1; sm_latcost[q * 4 + 2] = s2; }
  }
  __syncthreads();
  tb.broad = p.broad; tb.broad_row0 = p.now; tb.broad_stride = p.check_res;
  tb.sp.knot = sm_knot; tb.sp.coef = sm_coef; tb.sp.K = p.K; tb.sp.kpad = p.kpad;
  tb.sp.inv_h = (double)(p.K - 1) / sm_knot[p.K - 1];
  tb.time = sm_time; tb.lat = sm_lat;
  tb.list = nullptr; tb.row_off = nullptr; tb.row_cnt = nullptr; tb.use_lists = false;
  latcost = sm_latcost; quint = sm_quint;
}

__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_fop_kernel(const __grid_constant__ FopParams fp) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  const EvalParams& p = fp.base;
  const int N = p.W * fp.num_t * fp.num_speed;
  BlockTables tb;
  const double *latcost, *quint;
  fop_stage(fp, smem, tb, latcost, quint);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double my_cost = INFINITY;
  int my_key = -1;                                          // N - 1 - n: the LAST minimum wins (:328-333)
  if (n < N) {
    const int iv = n % fp.num_speed, q = n / fp.num_speed, it = q % fp.num_t;
    const int Pc = fp.n_samples[it];
    const double T = fp.horizon[it];
    QuarticLon lon;
    lon.init(p.s0, p.v0, p.a0, fp.speed[iv], 0.0, T);                                      // :125
    FopSink sink(p, tb, fp.max_accel, fp.max_jerk);
    const int seg0 = spline_bisect(tb.sp, fmin(fmax(p.s0, tb.sp.knot[0]), tb.sp.knot[p.K - 1]));
    walk_core(tb, lon, tb.lat + (size_t)q * p.P * 4, Pc, seg0, sink);
    const double Pd = (double)Pc;
    const double* lc = latcost + q * 4;
    double cost = p.cw[1] * (1.0 - (sink.s_last - sink.s_first) / p.max_distance);
    cost = cost + p.cw[2] * (lc[0] / Pd / (p.half_w * p.half_w));
    cost = cost + p.cw[3] * (lc[1] / Pd);
    cost = cost + p.cw[4] * (lc[2] / Pd);
    cost = cost + p.cw[5] * (sink.sum_a / Pd);
    cost = cost + p.cw[6] * (sink.sum_j / Pd);
    cost = cost + p.cw[7] * (sink.sum_v / Pd);
    if (fp.t_max > fp.t_min) cost = cost + p.cw[0] * (1.0 - (T - fp.t_min) / (fp.t_max - fp.t_min));
    if (p.feasible) p.feasible[n] = sink.curv_fail ? 0 : 1;
    if (p.collision) p.collision[n] = sink.collided ? 1 : 0;
    if (p.cost) p.cost[n] = (float)cost;
    if (!sink.curv_fail && !sink.collided && isfinite(cost)) { my_cost = cost; my_key = N - 1 - n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, o);
    const int oi = __shfl_down_sync(0xffffffffu, my_key, o);
    if (lex_less(oc, oi, my_cost, my_key)) { my_cost = oc; my_key = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red_cost[wid] = my_cost; red_idx[wid] = my_key; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
      if (lex_less(red_cost[k], red_idx[k], my_cost, my_key)) { my_cost = red_cost[k]; my_key = red_idx[k]; }
    p.blk_cost[blockIdx.x] = my_cost;
    p.blk_idx[blockIdx.x] = my_key;
  }
}

// one block: reduce the per-block winners, write the selection record and export the selected trajectory
__global__ void __launch_bounds__(EVAL_THREADS) fop_select_export_kernel(const __grid_constant__ FopParams fp, int n_blocks) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ int sm_best, sm_npts;
  __shared__ double ex[6][EXPORT_MAX_P];
  const EvalParams& p = fp.base;
  const int N = p.W * fp.num_t * fp.num_speed;
  double my_cost = INFINITY;
  int my_key = -1;
  for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
    const double c = p.blk_cost[b];
    const int k = p.blk_idx[b];
    if (lex_less(c, k, my_cost, my_key)) { my_cost = c; my_key = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, o);
    const int oi = __shfl_down_sync(0xffffffffu, my_key, o);
    if (lex_less(oc, oi, my_cost, my_key)) { my_cost = oc; my_key = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red_cost[wid] = my_cost; red_idx[wid] = my_key; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
      if (lex_less(red_cost[k], red_idx[k], my_cost, my_key)) { my_cost = red_cost[k]; my_key = red_idx[k]; }
    const int best = my_key >= 0 ? N - 1 - my_key : -1;
    BestRecord r;
    r.key = orderable_bits(best >= 0 ? my_cost : INFINITY);
    r.idx = best; r.cost = my_cost; r.n_pts = 0; r.n_candidates = N; r.n_seeds = fp.num_t * fp.num_speed; r.pad = 0;
    *p.best = r;
    sm_best = best;
  }
  __syncthreads();
  const int best = sm_best;
  if (best < 0 || p.best_traj == nullptr || p.P > EXPORT_MAX_P) return;
  BlockTables tb;
  const double *latcost, *quint;
  fop_stage(fp, smem, tb, latcost, quint);
  const int iv = best % fp.num_speed, q = best / fp.num_speed, it = q % fp.num_t;
  const int Pc = fp.n_samples[it], i = threadIdx.x;
  if (i == 0) sm_npts = Pc;
  __syncthreads();
  double row[JSSP_TRAJ_COLS];
  const double nanv = nan("");
#pragma unroll
  for (int c = 0; c < JSSP_TRAJ_COLS; ++c) row[c] = nanv;
  const double* lat = tb.lat + (size_t)q * p.P * 4;
  if (i < Pc) {
    QuarticLon lon;
    lon.init(p.s0, p.v0, p.a0, fp.speed[iv], 0.0, fp.horizon[it]);
    lon.eval(tb.time[i], row[1], row[2], row[3], row[4]);
    row[0] = tb.time[i];
    row[5] = lat[i * 4]; row[6] = lat[i * 4 + 1]; row[7] = lat[i * 4 + 2]; row[8] = lat[i * 4 + 3];
    if (!spline_in_range(tb.sp, row[1])) atomicMin(&sm_npts, i);
  }
  // rows beyond the candidate's own length stay NaN; the exported block always has P rows
  export_tail(tb, row, p.P, i < Pc ? lat[i * 4] : 0.0, fp.dt, p.best_traj, &sm_npts, p.best, ex);
}

// ------------------------------------------------------------------------------------------------------------------
// K3: export of the selected trajectory (float64 rows) and optional fp32 state dump of every candidate
// ------------------------------------------------------------------------------------------------------------------
struct DumpSink {
  float* out;   // [P][8] of this candidate: s, s_d, s_dd, d, x, y, yaw, c
  const double* lat;
  double pyaw, pds;
  __device__ __forceinline__ void sample(int i, double, double s, double sd, double sdd, double) {
    float4* r = reinterpret_cast<float4*>(out + (size_t)i * JSSP_STATE_COLS);
    const float nanv = __int_as_float(0x7fc00000);
    r[0] = make_float4((float)s, (float)sd, (float)sdd, (float)lat[i * 4]);
    r[1] = make_float4(nanv, nanv, nanv, nanv);
  }
  __device__ __forceinline__ void point(int i, double x, double y) {
    out[(size_t)i * JSSP_STATE_COLS + 4] = (float)x; out[(size_t)i * JSSP_STATE_COLS + 5] = (float)y;
  }
  __device__ __forceinline__ void segment(int i, double, double, double dx, double dy, double ds, double, double) {
    const double yaw = heading_of(dx, dy);
    out[(size_t)i * JSSP_STATE_COLS + 6] = (float)yaw;
    if (i >= 1) out[(size_t)(i - 1) * JSSP_STATE_COLS + 7] = (float)((yaw - pyaw) / pds);
    pyaw = yaw; pds = ds;
  }
  __device__ __forceinline__ void finish(int n_pts, double, double) {
    if (n_pts >= 2) {
      out[(size_t)(n_pts - 1) * JSSP_STATE_COLS + 6] = (float)pyaw;
      out[(size_t)(n_pts - 2) * JSSP_STATE_COLS + 7] = (float)((pyaw - pyaw) / pds);
    }
  }
};

__global__ void __launch_bounds__(EVAL_THREADS) dump_states_kernel(const __grid_constant__ EvalParams p, float* __restrict__ states) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int Ns = resolve_n_seeds(p);
  const int N = Ns * p.W;
  if (blockIdx.x * blockDim.x >= N) return;
  BlockTables tb;
  const double* latcost;
  stage_tables<0>(p, smem, tb, latcost);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int w = n / Ns, k = n - w * Ns;
  LonWalk lp;
  lon_walk_init(lp, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
  DumpSink sink{states + (size_t)n * p.P * JSSP_STATE_COLS, tb.lat + (size_t)w * p.P * 4, 0.0, 0.0};
  const int seg0 = spline_bisect(tb.sp, fmin(fmax(p.s0, tb.sp.knot[0]), tb.sp.knot[p.K - 1]));
  walk_candidate(p, tb, lp, w, seg0, sink);
}

__global__ void best_record_kernel(const BestRecord* __restrict__ best, long long offset, unsigned long long* __restrict__ rec) {
  rec[0] = best->key;
  rec[1] = best->idx >= 0 ? (unsigned long long)(best->idx + offset) : 0xffffffffffffffffull;
}

// ------------------------------------------------------------------------------------------------------------------
// K5: batched Cartesian -> Frenet projection: FrenetState.from_state(state, polyline) as the
// driver uses it at t <= 0.05 and in kinematics mode (intersection_planner.py:140-142, 308-310).  The class is absent from
// the reference tree; the specification is dlii-jssp_b200/frenet.py::FrenetState.from_state: nearest point of the
// reference line re-discretised every `step` metres (generate_frenet_frame, jssp.py:366-371), arc length = sum of the chord
// lengths up to it plus the tangential offset, signed lateral offset by the left-hand normal, speed / acceleration split
// by the heading error.  One block per state: threads stride over the polyline samples.
// ------------------------------------------------------------------------------------------------------------------
constexpr int PROJECT_THREADS = 256;
__global__ void __launch_bounds__(PROJECT_THREADS) project_states_kernel(const double* __restrict__ knot, const double* __restrict__ coef,
                                                                         int K, int kpad, double step, int n_samples,
                                                                         const double* __restrict__ state, double* __restrict__ frenet) {
  __shared__ double red_d[PROJECT_THREADS / 32];
  __shared__ int red_i[PROJECT_THREADS / 32];
  __shared__ double red_s[PROJECT_THREADS / 32];
  __shared__ int best_i;
  const double* st = state + (size_t)blockIdx.x * 5;        // x, y, yaw, v, a
  const double px = st[0], py = st[1];
  SplineView sp{knot, coef, K, kpad, (double)(K - 1) / knot[K - 1]};
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // 1) nearest sample (first occurrence on ties, like numpy.argmin)
  double bd = INFINITY;
  int bi = -1;
  for (int m = tid; m < n_samples; m += blockDim.x) {
    const double s = (double)m * step;
    double x, y;
    frenet_point(sp, spline_bisect(sp, s), s, 0.0, x, y);
    const double d2 = (x - px) * (x - px) + (y - py) * (y - py);
    if (d2 < bd) { bd = d2; bi = m; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_down_sync(0xffffffffu, bd, o);
    const int oi = __shfl_down_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (bi < 0 || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
  }
  if (lane == 0) { red_d[wid] = bd; red_i[wid] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int q = 1; q < (int)(blockDim.x >> 5); ++q)
      if (red_i[q] >= 0 && (bi < 0 || red_d[q] < bd || (red_d[q] == bd && red_i[q] < bi))) { bd = red_d[q]; bi = red_i[q]; }
    best_i = bi;
  }
  __syncthreads();
  const int ib = best_i;
  // 2) arc length of the polyline up to the nearest sample
  double acc = 0.0;
  for (int m = tid; m < ib; m += blockDim.x) {
    const double s0 = (double)m * step, s1 = (double)(m + 1) * step;
    double x0, y0, x1, y1;
    frenet_point(sp, spline_bisect(sp, s0), s0, 0.0, x0, y0);
    frenet_point(sp, spline_bisect(sp, s1), s1, 0.0, x1, y1);
    acc = acc + hypot(x1 - x0, y1 - y0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = acc + __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) red_s[wid] = acc;
  __syncthreads();
  if (tid != 0) return;
  for (int q = 1; q < (int)(blockDim.x >> 5); ++q) acc = acc + red_s[q];
  // 3) projection on the tangent / left-hand normal of the nearest sample
  const double sb = (double)ib * step;
  const int seg = spline_bisect(sp, sb);
  double rx, ry;
  frenet_point(sp, seg, sb, 0.0, rx, ry);
  const double h = sb - knot[seg];
  const double* c = coef + seg;
  const double tx = (c[kpad] + (2.0 * c[2 * kpad]) * h) + (3.0 * c[3 * kpad]) * (h * h);
  const double ty = (c[5 * kpad] + (2.0 * c[6 * kpad]) * h) + (3.0 * c[7 * kpad]) * (h * h);
  const double ryaw = atan2(ty, tx);
  double sy, cy;
  sincos(ryaw, &sy, &cy);
  const double dx = px - rx, dy = py - ry;
  double sdy, cdy;
  sincos(st[2] - ryaw, &sdy, &cdy);
  double* o = frenet + (size_t)blockIdx.x * 6;              // s, s_d, s_dd, d, d_d, d_dd
  o[0] = acc + (dx * cy + dy * sy);
  o[1] = st[3] * cdy;
  o[2] = st[4] * cdy;
  o[3] = -dx * sy + dy * cy;
  o[4] = st[3] * sdy;
  o[5] = st[4] * sdy;
}

// ------------------------------------------------------------------------------------------------------------------
// K4: batched IMM lane-intention update (builder-specified; no reference code exists)
// ------------------------------------------------------------------------------------------------------------------
constexpr int IMM_MAX_MODES = 8;
__global__ void imm_update_kernel(jssp_imm_params prm, double* __restrict__ mean, double* __restrict__ cov, double* __restrict__ mu,
                                  const double* __restrict__ z, int O, int M) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  double e[IMM_MAX_MODES], ed[IMM_MAX_MODES], p00[IMM_MAX_MODES], p01[IMM_MAX_MODES], p11[IMM_MAX_MODES], m[IMM_MAX_MODES],
      zz[IMM_MAX_MODES], cbar[IMM_MAX_MODES], lik[IMM_MAX_MODES];
  for (int i = 0; i < M; ++i) {
    e[i] = mean[((size_t)o * M + i) * 2]; ed[i] = mean[((size_t)o * M + i) * 2 + 1];
    p00[i] = cov[((size_t)o * M + i) * 3]; p01[i] = cov[((size_t)o * M + i) * 3 + 1]; p11[i] = cov[((size_t)o * M + i) * 3 + 2];
    m[i] = mu[(size_t)o * M + i]; zz[i] = z[(size_t)o * M + i];
  }
  const double p_off = M > 1 ? (1.0 - prm.p_stay) / (double)(M - 1) : 0.0;
  double ne[IMM_MAX_MODES], ned[IMM_MAX_MODES], n00[IMM_MAX_MODES], n01[IMM_MAX_MODES], n11[IMM_MAX_MODES];
  for (int j = 0; j < M; ++j) {
    // 1) interaction / mixing, expressed in mode j's lane frame: e_i + (z_j - z_i)
    double cb = 0.0;
    for (int i = 0; i < M; ++i) cb = cb + (i == j ? prm.p_stay : p_off) * m[i];
    cbar[j] = cb;
    double x0 = 0.0, x1 = 0.0;
    for (int i = 0; i < M; ++i) {
      const double wgt = cb > 0.0 ? (i == j ? prm.p_stay : p_off) * m[i] / cb : (i == j ? 1.0 : 0.0);
      x0 = x0 + wgt * (e[i] + (zz[j] - zz[i]));
      x1 = x1 + wgt * ed[i];
    }
    double q00 = 0.0, q01 = 0.0, q11 = 0.0;
    for (int i = 0; i < M; ++i) {
      const double wgt = cb > 0.0 ? (i == j ? prm.p_stay : p_off) * m[i] / cb : (i == j ? 1.0 : 0.0);
      const double d0 = e[i] + (zz[j] - zz[i]) - x0, d1 = ed[i] - x1;
      q00 = q00 + wgt * (p00[i] + d0 * d0); q01 = q01 + wgt * (p01[i] + d0 * d1); q11 = q11 + wgt * (p11[i] + d1 * d1);
    }
    // 2) mode-matched prediction, F = [[1, dt], [-k dt, 1]] (lane keeping towards this mode's centre line)
    const double dt = prm.dt, kd = -prm.k_lat * dt;
    const double xp0 = x0 + dt * x1, xp1 = kd * x0 + x1;
    const double f00 = q00 + dt * q01, f01 = q01 + dt * q11;      // (F P) rows
    const double f10 = kd * q00 + q01, f11 = kd * q01 + q11;
    const double pp00 = f00 + f01 * dt + prm.q_pos;
    const double pp01 = f00 * kd + f01;
    const double pp11 = f10 * kd + f11 + prm.q_vel;
    // 3) update with z_j (H = [1 0])
    const double nu = zz[j] - xp0, S = pp00 + prm.r_pos;
    const double k0 = pp00 / S, k1 = pp01 / S;
    ne[j] = xp0 + k0 * nu; ned[j] = xp1 + k1 * nu;
    n00[j] = pp00 - k0 * pp00; n01[j] = pp01 - k0 * pp01; n11[j] = pp11 - k1 * pp01;
    lik[j] = exp(-0.5 * nu * nu / S) / sqrt(2.0 * 3.14159265358979323846 * S);
  }
  // 4) mode probabilities
  double tot = 0.0;
  for (int j = 0; j < M; ++j) { m[j] = fmax(lik[j] * cbar[j], prm.mu_floor); tot = tot + m[j]; }
  for (int j = 0; j < M; ++j) {
    mean[((size_t)o * M + j) * 2] = ne[j]; mean[((size_t)o * M + j) * 2 + 1] = ned[j];
    cov[((size_t)o * M + j) * 3] = n00[j]; cov[((size_t)o * M + j) * 3 + 1] = n01[j]; cov[((size_t)o * M + j) * 3 + 2] = n11[j];
    mu[(size_t)o * M + j] = m[j] / tot;
  }
}

// one thread per (obstacle, mode): walk the mode's lane polyline at constant speed
__global__ void predict_modes_kernel(const double* __restrict__ lanes, int L, int Pn, const int* __restrict__ lane_of,
                                     const double* __restrict__ s_start, const double* __restrict__ speed, double dt, int O,
                                     int M, int Tt, double* __restrict__ state, uint8_t* __restrict__ valid) {
  const int om = blockIdx.x * blockDim.x + threadIdx.x;
  if (om >= O * M) return;
  const int lane = lane_of[om];
  double* st = state + (size_t)om * Tt * 3;
  uint8_t* va = valid + (size_t)om * Tt;
  if (lane < 0 || lane >= L) { for (int t = 0; t < Tt; ++t) { va[t] = 0; st[t * 3] = st[t * 3 + 1] = st[t * 3 + 2] = 0.0; } return; }
  const double* pl = lanes + (size_t)lane * Pn * 2;
  const double v = speed[om / M];
  int seg = 0;
  double acc = 0.0;   // arc length at the start of segment `seg`
  double seg_len = hypot(pl[2] - pl[0], pl[3] - pl[1]);
  for (int t = 0; t < Tt; ++t) {
    const double s = s_start[om] + v * ((double)t * dt);
    while (seg < Pn - 2 && s > acc + seg_len) {
      acc = acc + seg_len; ++seg;
      seg_len = hypot(pl[(seg + 1) * 2] - pl[seg * 2], pl[(seg + 1) * 2 + 1] - pl[seg * 2 + 1]);
    }
    const bool ok = s >= 0.0 && s <= acc + seg_len && seg_len > 0.0;
    const double dx = pl[(seg + 1) * 2] - pl[seg * 2], dy = pl[(seg + 1) * 2 + 1] - pl[seg * 2 + 1];
    const double u = seg_len > 0.0 ? (s - acc) / seg_len : 0.0;
    st[t * 3] = pl[seg * 2] + u * dx; st[t * 3 + 1] = pl[seg * 2 + 1] + u * dy; st[t * 3 + 2] = atan2(dy, dx);
    va[t] = ok ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// bench utility: dense FMA throughput of the CUDA cores (the roofline denominator of the evaluation kernel)
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* __restrict__ out, int iters, T a, T b) {
  T x0 = (T)threadIdx.x, x1 = x0 + (T)1, x2 = x0 + (T)2, x3 = x0 + (T)3, x4 = x0 + (T)4, x5 = x0 + (T)5, x6 = x0 + (T)6, x7 = x0 + (T)7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace jssp
