"""Host-side reference-line model: the ``CubicSpline2D`` / ``QuinticPolynomial`` interface the reference imports from its
unpublished ``common.geometry`` package (call sites: jerk_space_sampling_planner.py:109-122,137-141,364-369;
intersection_planner.py:155-161).  Only the spline *construction* runs on the host (once per scenario, O(K)); every
per-candidate evaluation happens in the CUDA kernels from the coefficient table built here."""
from __future__ import annotations

import bisect
import math

import numpy as np


def _natural_spline_c(h: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Second-derivative coefficients c_i of the natural cubic spline (c_0 = c_{n-1} = 0) by the Thomas algorithm."""
    n = len(y)
    c = np.zeros(n)
    if n < 3:
        return c
    lower = h[:-1].copy()                       # sub-diagonal   (rows 1..n-2)
    diag = 2.0 * (h[:-1] + h[1:])               # diagonal
    upper = h[1:].copy()                        # super-diagonal
    rhs = 3.0 * (y[2:] - y[1:-1]) / h[1:] - 3.0 * (y[1:-1] - y[:-2]) / h[:-1]
    m = n - 2
    for i in range(1, m):                       # forward elimination
        w = lower[i] / diag[i - 1]
        diag[i] -= w * upper[i - 1]
        rhs[i] -= w * rhs[i - 1]
    sol = np.zeros(m)
    sol[-1] = rhs[-1] / diag[-1]
    for i in range(m - 2, -1, -1):
        sol[i] = (rhs[i] - upper[i] * sol[i + 1]) / diag[i]
    c[1:-1] = sol
    return c


class CubicSpline1D:
    def __init__(self, x, y):
        self.x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        h = np.diff(self.x)
        if np.any(h <= 0):
            raise ValueError("spline knots must be strictly increasing (remove duplicate centre-line points first)")
        self._xl = self.x.tolist()
        self.a = y.copy()
        self.c = _natural_spline_c(h, y)
        self.d = (self.c[1:] - self.c[:-1]) / (3.0 * h)
        self.b = (y[1:] - y[:-1]) / h - h / 3.0 * (2.0 * self.c[:-1] + self.c[1:])

    def _seg(self, x: float) -> int:
        return min(max(bisect.bisect_right(self._xl, x) - 1, 0), len(self.x) - 2)

    def calc_position(self, x: float):
        if x < self.x[0] or x > self.x[-1]:
            return None
        i = self._seg(x); dx = x - self.x[i]
        return ((self.a[i] + self.b[i] * dx) + self.c[i] * (dx * dx)) + self.d[i] * ((dx * dx) * dx)

    def calc_first_derivative(self, x: float):
        if x < self.x[0] or x > self.x[-1]:
            return None
        i = self._seg(x); dx = x - self.x[i]
        return (self.b[i] + (2.0 * self.c[i]) * dx) + (3.0 * self.d[i]) * (dx * dx)

    def calc_second_derivative(self, x: float):
        if x < self.x[0] or x > self.x[-1]:
            return None
        i = self._seg(x); dx = x - self.x[i]
        return 2.0 * self.c[i] + (6.0 * self.d[i]) * dx


class CubicSpline2D:
    """Natural cubic spline x(s), y(s) over cumulative chord length.  ``s_list``, ``calc_position`` (``(None, None)``
    outside the line), ``calc_yaw``, ``calc_curvature`` as used by the reference."""

    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
        self.s_list = np.concatenate(([0.0], np.cumsum(np.hypot(np.diff(x), np.diff(y)))))
        self.sx = CubicSpline1D(self.s_list, x)
        self.sy = CubicSpline1D(self.s_list, y)

    def calc_position(self, s: float):
        x = self.sx.calc_position(s)
        if x is None:
            return None, None
        return x, self.sy.calc_position(s)

    def calc_yaw(self, s: float):
        dx, dy = self.sx.calc_first_derivative(s), self.sy.calc_first_derivative(s)
        return None if dx is None else math.atan2(dy, dx)

    def calc_curvature(self, s: float):
        dx, ddx = self.sx.calc_first_derivative(s), self.sx.calc_second_derivative(s)
        dy, ddy = self.sy.calc_first_derivative(s), self.sy.calc_second_derivative(s)
        if dx is None:
            return None
        return (ddy * dx - ddx * dy) / ((dx ** 2 + dy ** 2) ** 1.5)

    def sample(self, s: np.ndarray) -> np.ndarray:
        """[len(s), 4] = x, y, yaw, curvature at the arc lengths ``s`` (all inside the line): the vectorised form of the
        per-point calls of generate_frenet_frame (jerk_space_sampling_planner.py:366-371), same expressions per element."""
        s = np.asarray(s, dtype=np.float64)
        i = np.clip(np.searchsorted(self.s_list, s, side="right") - 1, 0, len(self.s_list) - 2)
        dx = s - self.s_list[i]
        out = np.empty((len(s), 4))
        d1, d2 = [], []
        for col, sp in enumerate((self.sx, self.sy)):
            out[:, col] = ((sp.a[i] + sp.b[i] * dx) + sp.c[i] * (dx * dx)) + sp.d[i] * ((dx * dx) * dx)
            d1.append((sp.b[i] + (2.0 * sp.c[i]) * dx) + (3.0 * sp.d[i]) * (dx * dx))
            d2.append(2.0 * sp.c[i] + (6.0 * sp.d[i]) * dx)
        out[:, 2] = np.arctan2(d1[1], d1[0])
        out[:, 3] = (d2[1] * d1[0] - d2[0] * d1[1]) / ((d1[0] ** 2 + d1[1] ** 2) ** 1.5)
        return out

    def coefficient_table(self) -> np.ndarray:
        """[K-1, 8] = ax, bx, cx, dx, ay, by, cy, dy per segment: the layout jssp_set_refline uploads."""
        return np.stack([self.sx.a[:-1], self.sx.b, self.sx.c[:-1], self.sx.d, self.sy.a[:-1], self.sy.b, self.sy.c[:-1], self.sy.d], 1)


def spline_tables(spline) -> tuple:
    """(knot_s [K], coef [K-1, 8]) from any object with the PythonRobotics-style attributes
    ``s_list``/``s``, ``sx.{a,b,c,d}``, ``sy.{a,b,c,d}`` - i.e. also from the reference's own CubicSpline2D."""
    s = np.asarray(getattr(spline, "s_list", getattr(spline, "s", None)), dtype=np.float64)
    K = len(s)
    cols = []
    for sp in (spline.sx, spline.sy):
        cols += [np.asarray(sp.a, dtype=np.float64)[:K - 1], np.asarray(sp.b, dtype=np.float64)[:K - 1],
                 np.asarray(sp.c, dtype=np.float64)[:K - 1], np.asarray(sp.d, dtype=np.float64)[:K - 1]]
    return s, np.stack(cols, 1)


class QuinticPolynomial:
    """QuinticPolynomial(xs, vxs, axs, xe, vxe, axe, time) with calc_point / calc_*_derivative
    (jerk_space_sampling_planner.py:109-122); closed-form coefficients (same as the kernels)."""

    def __init__(self, xs, vxs, axs, xe, vxe, axe, time):
        T = float(time)
        T2, T3, T4, T5 = T * T, T ** 3, T ** 4, T ** 5
        h = xe - xs
        self.a0, self.a1, self.a2 = xs, vxs, axs / 2.0
        self.a3 = (20.0 * h - (8.0 * vxe + 12.0 * vxs) * T - (3.0 * axs - axe) * T2) / (2.0 * T3)
        self.a4 = (-30.0 * h + (14.0 * vxe + 16.0 * vxs) * T + (3.0 * axs - 2.0 * axe) * T2) / (2.0 * T4)
        self.a5 = (12.0 * h - 6.0 * (vxe + vxs) * T + (axe - axs) * T2) / (2.0 * T5)

    def calc_point(self, t):
        return self.a0 + self.a1 * t + self.a2 * t ** 2 + self.a3 * t ** 3 + self.a4 * t ** 4 + self.a5 * t ** 5

    def calc_first_derivative(self, t):
        return self.a1 + 2 * self.a2 * t + 3 * self.a3 * t ** 2 + 4 * self.a4 * t ** 3 + 5 * self.a5 * t ** 4

    def calc_second_derivative(self, t):
        return 2 * self.a2 + 6 * self.a3 * t + 12 * self.a4 * t ** 2 + 20 * self.a5 * t ** 3

    def calc_third_derivative(self, t):
        return 6 * self.a3 + 24 * self.a4 * t + 60 * self.a5 * t ** 2
