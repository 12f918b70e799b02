// jssp_seed.cuh - K1: seed enumeration, one thread-block cluster per scenario (polyvt_sampling.py:167-305)
#pragma once
#include "jssp_common.cuh"

namespace jssp {

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
  int jpos_uniform;                           // jpos is an arithmetic progression (np.linspace): an interval of j2 values is an index range
  float jpos0, jpos_inv_step;
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
// Synchronisation policies of the enumeration: a thread-block cluster (lists in CTA 0's shared memory, reached through
// distributed shared memory) or a single CTA (the persistent replay kernel).
struct SeedClusterSync {
  cg::cluster_group c;
  __device__ __forceinline__ void sync() { c.sync(); }
  __device__ __forceinline__ int rank() const { return (int)c.block_rank(); }
  __device__ __forceinline__ int size() const { return (int)c.num_blocks(); }
};
struct SeedBlockSync {
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ int rank() const { return 0; }
  __device__ __forceinline__ int size() const { return 1; }
};

// The enumeration of one scenario by the threads of `sy` (every thread of the cluster / CTA calls it).  `seed_smem` = the list
// memory (CTA 0's for a cluster), `seeds` / `counts` = this scenario's outputs.  After the call the sorted loop positions are
// still in the key list: key_out / n_out hand them to a caller that wants to decode seeds itself.
template <class Sync>
__device__ __forceinline__ void seed_enum_body(const SeedGrids& g, const SeedState& st, const SeedWork& wk, Sync& sy, unsigned char* seed_smem,
                                               double* __restrict__ seeds, int* __restrict__ counts) {
  int* cnt = reinterpret_cast<int*>(seed_smem);                 // 0 A main, 1 A stop, 2 B, 3 keys, 4 main keys, 5 overflow
  SeedPrefixA* la_main = reinterpret_cast<SeedPrefixA*>(seed_smem + 32);
  SeedPrefixA* la_stop = la_main + wk.cap_a_main;
  SeedPrefixB* lb = reinterpret_cast<SeedPrefixB*>(la_stop + wk.cap_a_stop);
  uint32_t* key = reinterpret_cast<uint32_t*>(lb + wk.cap_b);
  const int tid = sy.rank() * blockDim.x + threadIdx.x, nt = sy.size() * blockDim.x;     // position in the cluster
  SEED_STAMP(0);
  if (tid < 8) cnt[tid] = 0;
  sy.sync();
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
  sy.sync();
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
      // The j2 > 0 that pass the band test below form an INTERVAL: v* = v1 - h / j2 (h = a1^2 / 2) lies in [-0.32 - 0.045 / j2, 0.02]
      // <=>  (h - 0.045) / (v1 + 0.32) <= j2 <= h / (v1 - 0.02)  (the right bound only for v1 > 0.02, the left only for h > 0.045;
      // v1 >= -1e-9 after check_state).  Two compares (with fp32 slack) skip the other j2 before any arithmetic.
      const float hq = 0.5f * a1f * a1f;
      float j2_lo = 0.f, j2_hi = 3.0e38f;
      if (v1f > 0.02f) j2_hi = hq / (v1f - 0.02f) * 1.001f + 1e-3f;
      if (hq > 0.045f && v1f > -0.3f) j2_lo = (hq - 0.045f) / (v1f + 0.32f) * 0.999f - 1e-3f;
      // on the uniform grid the interval is an index range (one index of slack either side; j2 <= 1e-6 - no band test - stays in)
      int ij_lo = 0, ij_hi = g.n_jpos - 1;
      if (g.jpos_uniform && dt_s > 0.f) {
        const float flo = (j2_lo - g.jpos0) * g.jpos_inv_step, fhi = (fminf(j2_hi, 1.0e30f) - g.jpos0) * g.jpos_inv_step;
        ij_lo = (int)fminf(fmaxf(floorf(flo) - 1.f, 0.f), (float)g.n_jpos);
        ij_hi = (int)fmaxf(fminf(ceilf(fhi) + 1.f, (float)(g.n_jpos - 1)), -1.f);
        if (g.jpos0 <= 1e-6f) ij_lo = 0;
      }
      for (int ij2 = ij_lo + j_first; ij2 <= ij_hi; ij2 += j_step) {
        const double j2 = g.jpos[ij2];
        int lo = 0, hi = g.n_ts - 1;
        if (j2 > 1e-6 && dt_s > 0.f) {
          if ((float)j2 > j2_hi || (float)j2 < j2_lo) continue;
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
  sy.sync();
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
  sy.sync();
  SEED_STAMP(4);
  // reference order: one sort of the loop positions (stop seeds carry the top bit) by CTA 0, then every CTA decodes
  const int n_all = min(cnt[3], wk.cap_keys), n_main = cnt[4];
  if (sy.rank() == 0) {
    int np2 = 1;
    while (np2 < n_all) np2 <<= 1;
    for (int i = n_all + (int)threadIdx.x; i < np2; i += blockDim.x) key[i] = 0xffffffffu;
    __syncthreads();
    if (np2 <= (int)blockDim.x) bitonic_sort_regs(key, np2); else bitonic_sort_smem(key, np2);
  }
  SEED_STAMP(5);
  sy.sync();
  SEED_STAMP(6);
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
    int overflow = cnt[5];
    int emitted = n_all;
    if (cnt[3] > wk.cap_keys || emitted > wk.cap_seeds) { overflow = 1; emitted = min(emitted, wk.cap_seeds); }
    counts[0] = n_main; counts[1] = cnt[3] - n_main; counts[2] = emitted; counts[3] = overflow;
  }
  SEED_STAMP(7);
}

template <bool kBatch>
__global__ void __launch_bounds__(SEED_THREADS) seed_enum_kernel(SeedGrids g, SeedState st1, const EvalParams* __restrict__ slots,
                                                                 SeedWork wk) {
  extern __shared__ __align__(16) unsigned char seed_smem_local[];
  SeedClusterSync sy{cg::this_cluster()};
  const int csize = sy.size();
  const int sc = blockIdx.x / csize;
  pdl_launch_dependents();                                      // the evaluation kernel behind may start being placed
  if (kBatch) pdl_wait();                                       // the slot records are written by the previous cycle's hand-off
  if (kBatch && !slot_live(slots, sc)) return;                  // cluster-uniform
  const SeedState st = kBatch ? seed_state_of(slots, sc) : st1;
  unsigned char* seed_smem = csize > 1 ? sy.c.map_shared_rank(seed_smem_local, 0) : seed_smem_local;
  seed_enum_body(g, st, wk, sy, seed_smem, wk.seeds + (size_t)sc * wk.cap_seeds * 10, wk.counts + (size_t)sc * 4);
  sy.sync();                                                    // CTA 0's shared memory stays alive until every CTA has read it
}

}  // namespace jssp
