// jssp_aux.cuh - K3 state dump, K5 Cartesian->Frenet projection, K4 IMM, FMA peak micro-benchmark
#pragma once
#include "jssp_eval.cuh"

namespace jssp {

// ------------------------------------------------------------------------------------------------------------------
// K3: export of the selected trajectory (float64 rows) and optional fp32 state dump of every candidate
// ------------------------------------------------------------------------------------------------------------------
// One warp per seed, lanes = samples: nothing here is a serial walk.  Per seed the unrolled s, s_d, s_dd and the reference point
// with its unit normal go to the warp's shared-memory scratch once; per lateral target every lane forms its own point and its two
// neighbours, the forward-difference headings (jerk_space_sampling_planner.py:155-160, last one repeated) and the un-unwrapped
// curvature (:161-163) - the heading difference as atan2(cross, dot) of the float64 chord vectors plus the +-2*pi jump the
// reference's plain subtraction has when the heading crosses +-pi - and writes its row [s, s_d, s_dd, d, x, y, yaw, c]: a warp
// stores 32 consecutive rows = 1 KB contiguous, so the dump is written at HBM speed (1,632 B per candidate, the HBM-bound variant
// of SURVEY.md section 8d).
constexpr int DUMP_WARPS = 8;
// per warp: segment table | [P][4] ix, iy, u, v | [P][3] float s, s_d, s_dd | [2][P] double2 points of the current lateral target
__host__ __device__ inline size_t dump_warp_doubles(int P) { return (TILE_PROF + (size_t)P * 4 + (size_t)((P * 3 + 1) / 2) + 3 + (size_t)P * 4) & ~(size_t)1; }   // even: 16-byte rows
__host__ __device__ inline size_t dump_scratch_bytes(int P) { return (size_t)DUMP_WARPS * dump_warp_doubles(P) * 8; }

// atan2 for the float32 dump rows: octant reduction + the Cephes atanf polynomial on |t| <= tan(pi/8) (relative error ~1e-7 with
// the approximate division), about a third of the instructions of the library atan2f; atan2(0, 0) = 0 like NumPy
__device__ __forceinline__ float dump_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const bool big = mn > 0.41421356f * mx;
  const float t = __fdividef(big ? mn - mx : mn, big ? mn + mx : mx);
  const float z = t * t;
  float r = fmaf(fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f) * z, t, t);
  if (big) r += 0.78539816f;
  if (!(mx > 0.f)) r = 0.f;
  if (ay > ax) r = 1.57079637f - r;
  if (signbit(x)) r = 3.14159274f - r;
  return copysignf(r, y);
}

__global__ void __launch_bounds__(DUMP_WARPS * 32) dump_states_kernel(const __grid_constant__ EvalParams p, float* __restrict__ states) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int Ns = resolve_n_seeds(p);
  const int P = p.P, W = p.W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if ((int)blockIdx.x * DUMP_WARPS >= Ns) return;
  BlockTables tb;
  const double* latcost;
  stage_tables<0>(p, smem, tb, latcost);
  const SmemLayout L = eval_smem_layout(0, p.OMpad, 0, p.K, p.kpad, p.P, p.W);
  const size_t per_warp = dump_warp_doubles(P);
  double* prof = reinterpret_cast<double*>(smem + ((L.total + 15) & ~(size_t)15)) + (size_t)wid * per_warp;
  double* base = prof + TILE_PROF;                       // [P][4] ix, iy, u, v
  float* lon = reinterpret_cast<float*>(base + (size_t)P * 4);   // [P][3] s, s_d, s_dd
  double2* pts = reinterpret_cast<double2*>(prof + ((per_warp - (size_t)P * 4) & ~(size_t)1));   // [2][P] (x, y), one buffer per parity of the lateral target
  const float nanv = __int_as_float(0x7fc00000);
  for (int k = blockIdx.x * DUMP_WARPS + wid; k < Ns; k += gridDim.x * DUMP_WARPS) {       // warp-uniform
    if (lane == 0) tile_profile_store(prof, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
    __syncwarp();
    int n_pts = P, kseg = 0;
    for (int b0 = 0; b0 < P; b0 += 32) {
      const int i = b0 + lane;
      bool off = false;
      if (i < P) {
        double s, sd, sdd, sddd;
        tile_lon_eval(prof, kseg, tb.time[i], s, sd, sdd, sddd);
        lon[i * 3] = (float)s; lon[i * 3 + 1] = (float)sd; lon[i * 3 + 2] = (float)sdd;
        off = !spline_in_range(tb.sp, s);
        if (!off) {
          double ix, iy, u, v;
          frenet_base(tb.sp, spline_walk(tb.sp, s, 0), s, ix, iy, u, v);
          double2* e = reinterpret_cast<double2*>(base + (size_t)i * 4);
          e[0] = make_double2(ix, iy); e[1] = make_double2(u, v);
        }
      }
      const unsigned int m = __ballot_sync(0xffffffffu, off);
      if (m != 0u && n_pts == P) n_pts = b0 + __ffs(m) - 1;      // first sample off the reference line (jssp.py:138-139)
    }
    __syncwarp();
    for (int w = 0; w < W; ++w) {
      const double* lat = tb.lat + (size_t)w * P * 4;
      float* out = states + ((size_t)w * Ns + k) * P * JSSP_STATE_COLS;
      // every point of this candidate once (float64), into the buffer of this target's parity; the rows then read their two
      // neighbours.  One warp barrier per target: a buffer is rewritten two targets later, after every lane passed the next barrier.
      double2* pw = pts + (size_t)(w & 1) * P;
      for (int i = lane; i < n_pts; i += 32) {
        const double2 a = *reinterpret_cast<const double2*>(base + (size_t)i * 4), b = *reinterpret_cast<const double2*>(base + (size_t)i * 4 + 2);
        const double d = lat[i * 4];
        pw[i] = make_double2(fma(-d, b.x, a.x), fma(d, b.y, a.y));
      }
      __syncwarp();
      for (int i = lane; i < P; i += 32) {
        float4 r0 = make_float4(lon[i * 3], lon[i * 3 + 1], lon[i * 3 + 2], (float)lat[i * 4]);
        float4 r1 = make_float4(nanv, nanv, nanv, nanv);
        if (i < n_pts) {
          // heading of sample i = chord hs -> hs + 1 with hs = min(i, n_pts - 2) (the last heading repeats the previous one)
          const int hs = n_pts >= 2 ? min(i, n_pts - 2) : i;
          const double2 pa = pw[hs];
          const double2 me = (hs == i) ? pa : pw[i];
          r1.x = (float)me.x; r1.y = (float)me.y;
          if (n_pts >= 2) {
            const double2 pb = pw[hs + 1];
            const double dx0 = pb.x - pa.x, dy0 = pb.y - pa.y;
            r1.z = dump_atan2((float)dy0, (float)dx0);
            if (i < n_pts - 1) {                      // c_i = (yaw_{i+1} - yaw_i) / ds_i, no unwrapping
              float raw = 0.0f;
              if (i + 1 <= n_pts - 2) {
                const double2 pc = pw[i + 2];
                const double dx1 = pc.x - pb.x, dy1 = pc.y - pb.y;
                const double cross = fma(dx0, dy1, -(dy0 * dx1)), dot = fma(dx0, dx1, dy0 * dy1);
                // the angle between two consecutive chords is small almost everywhere: atan(z) by its series for |z| < 0.1 (truncation
                // error z^7 / 7 < 2e-8 relative), the library atan2f otherwise
                const float crf = (float)cross, dtf = (float)dot;
                if (dtf > 0.f && fabsf(crf) < 0.1f * dtf) {
                  const float z = __fdividef(crf, dtf), z2 = z * z;
                  raw = z * fmaf(z2, fmaf(z2, 0.2f, -0.33333334f), 1.0f);
                } else raw = dump_atan2(crf, dtf);
                const bool nb = signbit(dy1);
                if ((signbit(dy0) != nb) && (nb ? (cross > 0.0) : (cross < 0.0))) raw += nb ? -6.2831853f : 6.2831853f;
                if (dx1 == 0.0 && dy1 == 0.0) raw = 0.0f - r1.z;      // atan2(0, 0) = 0
                else if (dx0 == 0.0 && dy0 == 0.0) raw = dump_atan2((float)dy1, (float)dx1);
              }
              // raw / ds in float32 (the row is float32): one MUFU.RSQ instead of a float64 square root and division; ds == 0 gives
              // +-inf or nan exactly as the division does
              const float q0 = (float)fma(dx0, dx0, dy0 * dy0);
              r1.w = (q0 > 1e-30f) ? raw * rsqrtf(q0) : (float)((double)raw / sqrt(dx0 * dx0 + dy0 * dy0));
            }
          }
        }
        float4* o = reinterpret_cast<float4*>(out + (size_t)i * JSSP_STATE_COLS);
        __stcs(o, r0); __stcs(o + 1, r1);               // streaming stores: the dump is written once and read by nobody on the device
      }
    }
    __syncwarp();                                      // the scratch is reused by the warp's next seed
  }
}

__global__ void best_record_kernel(const BestRecord* __restrict__ best, long long offset, unsigned long long* __restrict__ rec) {
  rec[0] = best->key;
  rec[1] = best->idx >= 0 ? (unsigned long long)(best->idx + offset) : 0xffffffffffffffffull;
}

// ------------------------------------------------------------------------------------------------------------------
// K5: batched Cartesian -> Frenet projection: FrenetState.from_state(state, polyline) as the
// driver uses it at t <= 0.05 and in kinematics mode (intersection_planner.py:140-142, 308-310).  The class is absent from
// the reference tree; the specification is dlii_jssp_b200/frenet.py::FrenetState.from_state: nearest point of the
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

// Constant-speed roll-out of every (obstacle, mode) along its lane polyline, two kernels:
//   lane_tables_kernel    one block per lane: segment lengths and headings in parallel, then the running arc length
//                         (sequential additions, the order numpy.cumsum uses) -> cum [L][Pn], yaw [L][Pn-1]
//   predict_modes_kernel  one thread per (obstacle-mode, time step): binary search of the segment with cum[i] < s <= cum[i+1]
//                         (numpy.searchsorted(cum, s, "left") - 1, clipped), linear interpolation; stores coalesced along t
__global__ void lane_tables_kernel(const double* __restrict__ lanes, int Pn, double* __restrict__ cum, double* __restrict__ len,
                                   double* __restrict__ yaw) {
  const double* pl = lanes + (size_t)blockIdx.x * Pn * 2;
  double* c = cum + (size_t)blockIdx.x * Pn;
  double* ln = len + (size_t)blockIdx.x * (Pn - 1);
  double* yw = yaw + (size_t)blockIdx.x * (Pn - 1);
  for (int i = threadIdx.x; i < Pn - 1; i += blockDim.x) {
    const double dx = pl[(i + 1) * 2] - pl[i * 2], dy = pl[(i + 1) * 2 + 1] - pl[i * 2 + 1];
    ln[i] = hypot(dx, dy);
    yw[i] = atan2(dy, dx);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    c[0] = 0.0;
    for (int i = 0; i < Pn - 1; ++i) { acc = acc + ln[i]; c[i + 1] = acc; }
  }
}

__global__ void predict_modes_kernel(const double* __restrict__ lanes, int L, int Pn, const int* __restrict__ lane_of,
                                     const double* __restrict__ s_start, const double* __restrict__ speed, double dt, int O,
                                     int M, int Tt, const double* __restrict__ cum, const double* __restrict__ len,
                                     const double* __restrict__ yaw, double* __restrict__ state, uint8_t* __restrict__ valid) {
  const int om = blockIdx.x, t = blockIdx.y * blockDim.x + threadIdx.x;      // grid (O * M, ceil(Tt / blockDim.x))
  if (t >= Tt) return;
  const size_t gid = (size_t)om * Tt + t;
  const int lane = lane_of[om];
  double* st = state + (size_t)gid * 3;
  if (lane < 0 || lane >= L) { valid[gid] = 0; st[0] = st[1] = st[2] = 0.0; return; }
  const double* pl = lanes + (size_t)lane * Pn * 2;
  const double* c = cum + (size_t)lane * Pn;
  const double s = s_start[om] + speed[om / M] * ((double)t * dt);
  int lo = 0, hi = Pn - 1;            // first index with c[idx] >= s (Pn when none), by bisection over c[1 .. Pn-1]
  if (!(c[0] >= s)) {
    lo = 1; hi = Pn;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (c[mid] >= s) hi = mid; else lo = mid + 1;
    }
  } else lo = 0;
  const int seg = min(max(lo - 1, 0), Pn - 2);
  const double seg_len = len[(size_t)lane * (Pn - 1) + seg];
  const double dx = pl[(seg + 1) * 2] - pl[seg * 2], dy = pl[(seg + 1) * 2 + 1] - pl[seg * 2 + 1];
  const double u = seg_len > 0.0 ? (s - c[seg]) / seg_len : 0.0;
  st[0] = pl[seg * 2] + u * dx; st[1] = pl[seg * 2 + 1] + u * dy; st[2] = yaw[(size_t)lane * (Pn - 1) + seg];
  valid[gid] = (s >= 0.0 && s <= c[Pn - 1] && seg_len > 0.0) ? 1 : 0;
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
