// jssp_pack.cuh - K0: obstacle dictionary tensors -> broad / narrow / float64 tables + culling intervals
#pragma once
#include "jssp_common.cuh"

namespace jssp {

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

}  // namespace jssp
