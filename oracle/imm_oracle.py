"""NumPy statement of the batched IMM lane-intention update and the per-mode roll-out (TEST INFRASTRUCTURE).

PARITY UNPINNED: the reference contains no predictor code at all (SURVEY.md section 0.3 - only README prose,
README.md:19,23,98-110, and the stub dict configuration_parameters.py:65-74), so this file *is* the specification the
CUDA kernels `imm_update_kernel` / `predict_modes_kernel` are tested against:

  per obstacle, M lane hypotheses; mode j's state is the lateral error (e, e_dot) to lane j's centre line;
  1. mixing with a Markov matrix (p_stay on the diagonal, the rest shared), states re-expressed in mode j's lane frame
     through the measured offsets: e_i -> e_i + (z_j - z_i);
  2. mode-matched prediction with the lane-keeping model F = [[1, dt], [-k_lat dt, 1]], Q = diag(q_pos, q_vel);
  3. Kalman update with the measured lateral offset z_j (H = [1 0], R = r_pos);
  4. mode probabilities mu_j ~ N(nu_j; 0, S_j) * cbar_j, floored and normalised."""
import numpy as np


def imm_update(mean, cov, mu, z, dt=0.1, q_pos=0.01, q_vel=0.05, r_pos=0.04, k_lat=0.5, p_stay=0.95, mu_floor=1e-6):
    O, M = mu.shape
    mean, cov, mu = mean.copy(), cov.copy(), mu.copy()
    out_mean, out_cov, out_mu = np.zeros_like(mean), np.zeros_like(cov), np.zeros_like(mu)
    p_off = (1.0 - p_stay) / (M - 1) if M > 1 else 0.0
    for o in range(O):
        lik, cbar = np.zeros(M), np.zeros(M)
        for j in range(M):
            pij = np.array([p_stay if i == j else p_off for i in range(M)])
            cb = 0.0
            for i in range(M):
                cb = cb + pij[i] * mu[o, i]
            cbar[j] = cb
            w = pij * mu[o] / cb if cb > 0 else (np.arange(M) == j).astype(float)
            x0 = x1 = 0.0
            for i in range(M):
                x0 = x0 + w[i] * (mean[o, i, 0] + (z[o, j] - z[o, i])); x1 = x1 + w[i] * mean[o, i, 1]
            q00 = q01 = q11 = 0.0
            for i in range(M):
                d0, d1 = mean[o, i, 0] + (z[o, j] - z[o, i]) - x0, mean[o, i, 1] - x1
                q00 = q00 + w[i] * (cov[o, i, 0] + d0 * d0); q01 = q01 + w[i] * (cov[o, i, 1] + d0 * d1); q11 = q11 + w[i] * (cov[o, i, 2] + d1 * d1)
            kd = -k_lat * dt
            xp0, xp1 = x0 + dt * x1, kd * x0 + x1
            f00, f01, f10, f11 = q00 + dt * q01, q01 + dt * q11, kd * q00 + q01, kd * q01 + q11
            pp00, pp01, pp11 = f00 + f01 * dt + q_pos, f00 * kd + f01, f10 * kd + f11 + q_vel
            nu, S = z[o, j] - xp0, pp00 + r_pos
            k0, k1 = pp00 / S, pp01 / S
            out_mean[o, j] = (xp0 + k0 * nu, xp1 + k1 * nu)
            out_cov[o, j] = (pp00 - k0 * pp00, pp01 - k0 * pp01, pp11 - k1 * pp01)
            lik[j] = np.exp(-0.5 * nu * nu / S) / np.sqrt(2.0 * np.pi * S)
        m = np.maximum(lik * cbar, mu_floor)
        out_mu[o] = m / m.sum()
    return out_mean, out_cov, out_mu


def predict_modes(lanes, lane_of, s_start, speed, dt, Tt):
    """Constant-speed roll-out along each mode's lane polyline -> state [O,M,Tt,3] (x, y, yaw), valid [O,M,Tt]."""
    O, M = lane_of.shape
    state = np.zeros((O, M, Tt, 3)); valid = np.zeros((O, M, Tt), dtype=bool)
    for o in range(O):
        for m in range(M):
            l = lane_of[o, m]
            if l < 0 or l >= len(lanes):
                continue
            pl = lanes[l]
            seg = np.hypot(*np.diff(pl, axis=0).T)
            cum = np.concatenate(([0.0], np.cumsum(seg)))
            for t in range(Tt):
                s = s_start[o, m] + speed[o] * (t * dt)
                i = int(np.clip(np.searchsorted(cum, s, side="left") - 1, 0, len(pl) - 2))
                u = (s - cum[i]) / seg[i]
                state[o, m, t, :2] = pl[i] + u * (pl[i + 1] - pl[i])
                state[o, m, t, 2] = np.arctan2(pl[i + 1, 1] - pl[i, 1], pl[i + 1, 0] - pl[i, 0])
                valid[o, m, t] = 0.0 <= s <= cum[-1]
    return state, valid
