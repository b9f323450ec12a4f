"""oracle/numpy_ukf.py -- CPU ORACLE cross-check.  TEST INFRASTRUCTURE ONLY.

An independent NumPy / SciPy-LAPACK restatement of the same path as
oracle/ukf_oracle.hpp (SURVEY.md Appendix A + the reference's in-tree models),
written separately (vectorised over sigma points, LAPACK dpotrf through SciPy,
numpy.linalg.inv instead of cofactors) so that the C++ oracle is not its own only
witness.  PARITY UNPINNED: the reference has no golden vectors for this path.

Only tests/ may import this module; the product never does.

Reference sites: PoseUKF.cpp:7-196, OrientationUKF.cpp:12-89,
UnscentedKalmanFilter.hpp:83-125; upstream ukfom/ukf.hpp, mtk/types/SOn.hpp,
mtk/src/mtkmath.hpp for the engine.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import cholesky

MEAN_TOL = 1e-5
MEAN_MAX_IT = 10000
MTK_TOL = 1e-11
TAYLOR_BOUND = 2.0 ** -13
EARTHW = 2.0 * math.pi / 86164.0


# ---- quaternions, arrays (..., 4) stored x, y, z, w ---------------------------
def qmul(a, b):
    ax, ay, az, aw = np.moveaxis(np.asarray(a, float), -1, 0)
    bx, by, bz, bw = np.moveaxis(np.asarray(b, float), -1, 0)
    return np.stack(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz,
        ],
        axis=-1,
    )


def qconj(a):
    a = np.asarray(a, float)
    return a * np.array([-1.0, -1.0, -1.0, 1.0])


def qrot(q, v):
    """Rotate v by unit quaternion q via the rotation matrix (independent of the
    oracle's cross-product form)."""
    R = qmat(q)
    return np.einsum("...ij,...j->...i", R, np.asarray(v, float))


def qmat(q):
    x, y, z, w = np.moveaxis(np.asarray(q, float), -1, 0)
    R = np.empty(np.shape(x) + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def so3_exp(v, scale=1.0):
    v = np.asarray(v, float)
    half = scale / 2.0
    x2 = half * half * np.sum(v * v, axis=-1)
    x = np.sqrt(x2)
    big = x2 >= TAYLOR_BOUND
    xs = np.where(big, x, 1.0)
    c = np.where(big, np.cos(xs), 1 - x2 / 2 + x2**2 / 24 - x2**3 / 720)
    sinc = np.where(big, np.sin(xs) / xs, 1 - x2 / 6 + x2**2 / 120 - x2**3 / 5040)
    return np.concatenate([(sinc * half)[..., None] * v, c[..., None]], axis=-1)


def so3_log(q):
    q = np.asarray(q, float)
    nv = np.sqrt(np.sum(q[..., :3] ** 2, axis=-1))
    nv = np.maximum(nv, MTK_TOL)
    s = 2.0 / nv * np.arctan(nv / q[..., 3])
    return s[..., None] * q[..., :3]


class Manifold:
    """Compound manifold described by a list of (kind, mu_slice, tangent_slice)."""

    def __init__(self, parts, left=True):
        self.parts = parts
        self.left = left
        self.dof = max(t.stop for _, _, t in parts)
        self.mu = max(m.stop for _, m, _ in parts)

    def boxplus(self, x, d, scale=1.0):
        x = np.array(x, float, copy=True)
        d = np.asarray(d, float)
        out = np.broadcast_to(x, np.broadcast_shapes(x.shape[:-1], d.shape[:-1]) + (self.mu,)).copy()
        for kind, ms, ts in self.parts:
            if kind == "vec":
                out[..., ms] = x[..., ms] + scale * d[..., ts]
            else:
                e = so3_exp(d[..., ts], scale)
                out[..., ms] = qmul(e, x[..., ms]) if self.left else qmul(x[..., ms], e)
        return out

    def boxminus(self, x, y):
        x = np.asarray(x, float)
        y = np.asarray(y, float)
        shape = np.broadcast_shapes(x.shape[:-1], y.shape[:-1])
        out = np.empty(shape + (self.dof,))
        for kind, ms, ts in self.parts:
            if kind == "vec":
                out[..., ts] = x[..., ms] - y[..., ms]
            else:
                rel = qmul(x[..., ms], qconj(y[..., ms])) if self.left else qmul(qconj(y[..., ms]), x[..., ms])
                out[..., ts] = so3_log(rel)
        return out


def pose_manifold(left=True):
    return Manifold(
        [("vec", slice(0, 3), slice(0, 3)), ("so3", slice(3, 7), slice(3, 6)), ("vec", slice(7, 10), slice(6, 9)),
         ("vec", slice(10, 13), slice(9, 12))], left)


def orientation_manifold(left=True):
    return Manifold(
        [("so3", slice(0, 4), slice(0, 3)), ("vec", slice(4, 7), slice(3, 6)), ("vec", slice(7, 10), slice(6, 9)),
         ("vec", slice(10, 13), slice(9, 12)), ("vec", slice(13, 14), slice(12, 13))], left)


def vec_manifold(m):
    return Manifold([("vec", slice(0, m), slice(0, m))])


def rot_manifold(left=True):
    return Manifold([("so3", slice(0, 4), slice(0, 3))], left)


class Ukf:
    """ukfom::ukf."""

    def __init__(self, man, mu, sigma):
        self.man = man
        self.mu = np.array(mu, float)
        self.sigma = np.array(sigma, float)
        self.passes = []
        self.accept_max_d2 = np.inf  # ukfom::accept_any_mahalanobis_distance (PoseUKF.cpp:116)
        self.rejected = 0

    def sigma_points(self, delta=None):
        n = self.man.dof
        L = cholesky(self.sigma, lower=True)  # LAPACK dpotrf('L')
        d = np.zeros(n) if delta is None else np.asarray(delta, float)
        D = np.empty((2 * n + 1, n))
        D[0] = d
        D[1::2] = d + L.T
        D[2::2] = d - L.T
        return self.man.boxplus(self.mu, D)

    def mean(self, man, X):
        ref = X[0].copy()
        it = 0
        passes = 0
        while True:
            md = man.boxminus(X, ref).sum(axis=0) / X.shape[0]
            ref = man.boxplus(ref, md)
            passes += 1
            it_ok = True
            if np.linalg.norm(md) > MEAN_TOL:
                it += 1
                it_ok = it < MEAN_MAX_IT
                if it_ok:
                    continue
            break
        if man is self.man:
            self.passes.append(passes)
        return ref

    @staticmethod
    def cov(man, mean, V):
        d = man.boxminus(V, mean)
        return 0.5 * d.T @ d

    def predict(self, g, Q):
        X = self.sigma_points()
        X = g(X)
        self.mu = self.mean(self.man, X)
        self.sigma = self.cov(self.man, self.mu, X) + Q

    def apply_delta(self, delta):
        X = self.sigma_points(delta)
        self.mu = self.mean(self.man, X)
        self.sigma = self.cov(self.man, self.mu, X)

    def update(self, zman, z, h, R):
        X = self.sigma_points()
        Z = h(X)
        zbar = self.mean(zman, Z)
        S = self.cov(zman, zbar, Z) + R
        dx = self.man.boxminus(X, self.mu)
        dz = zman.boxminus(Z, zbar)
        Sxz = 0.5 * dx.T @ dz
        Sinv = np.linalg.inv(S)
        K = Sxz @ Sinv
        innov = zman.boxminus(np.asarray(z, float), zbar)
        if float(innov @ Sinv @ innov) > self.accept_max_d2:  # the accept functor slot
            self.rejected += 1
            return
        self.sigma = self.sigma - K @ S @ K.T
        self.apply_delta(K @ innov)


class Shell:
    """UnscentedKalmanFilter<Manifold> time guards (UnscentedKalmanFilter.hpp:83-125)."""

    def __init__(self):
        self.t_last = 0
        self.min_dt = 1e-9
        self.max_dt = np.finfo(float).max

    def predict_time(self, ts_us):
        if self.t_last == 0:
            self.t_last = ts_us
            return
        dt = (ts_us - self.t_last) / 1e6
        if dt > self.min_dt:
            self.t_last = ts_us
        self.predict_dt(dt)

    def predict_dt(self, dt):
        if dt < 0:
            raise RuntimeError("Delta time is negative!")
        if dt <= self.min_dt:
            return
        if dt > self.max_dt:
            raise RuntimeError("Delta time is greater then the allowed maximum!")
        self.predict_impl(dt)


POSE_SELECT = {
    0: [0, 1, 2], 1: [0, 1], 2: [2], 4: [7, 8, 9], 5: [7, 8], 6: [9], 7: [7, 12], 8: [10, 11, 12],
}


class PoseUKF(Shell):
    def __init__(self, mu, sigma, left=True):
        super().__init__()
        self.left = left
        self.man = pose_manifold(left)
        self.ukf = Ukf(self.man, mu, sigma)
        self.Q = np.diag([0.01] * 3 + [0.001] * 3 + [1e-5] * 3 + [1e-5] * 3)
        self.acc_mu = np.full(3, np.nan)
        self.acc_cov = np.eye(3)

    def set_acceleration(self, mu, cov):
        self.acc_mu = np.array(mu, float)
        self.acc_cov = np.array(cov, float)

    def _model(self, X, dt, acc=None):
        X = X.copy()
        q = X[:, 3:7].copy()
        if acc is not None:
            X[:, 7:10] = X[:, 7:10] + dt * acc
        X[:, 0:3] = X[:, 0:3] + dt * qrot(q, X[:, 7:10])
        e = so3_exp(qrot(q, X[:, 10:13]), dt)
        X[:, 3:7] = qmul(e, q) if self.left else qmul(q, e)
        return X

    def predict_impl(self, dt):
        R = qmat(self.ukf.mu[3:7])
        pn = self.Q.copy()
        pn[0:3, 0:3] = R @ self.Q[0:3, 0:3] @ R.T
        pn[3:6, 3:6] = R @ self.Q[3:6, 3:6] @ R.T
        pn = dt * pn
        if np.all(np.isfinite(self.acc_mu)):
            pn = self.Q.copy()  # PoseUKF.cpp:190 shadows: unrotated, unscaled
            pn[6:9, 6:9] = 2.0 * self.acc_cov
            self.ukf.predict(lambda X: self._model(X, dt, self.acc_mu), pn)
        else:
            self.ukf.predict(lambda X: self._model(X, dt), pn)

    def update(self, kind, zmu, zcov):
        zmu = np.atleast_1d(np.asarray(zmu, float))
        zcov = np.atleast_2d(np.asarray(zcov, float))
        if kind == 3:
            z = so3_exp(zmu, 1.0)
            self.ukf.update(rot_manifold(self.left), z, lambda X: X[:, 3:7], zcov)
        else:
            sel = POSE_SELECT[kind]
            self.ukf.update(vec_manifold(len(sel)), zmu, lambda X: X[:, sel], zcov)


class OrientationUKF(Shell):
    def __init__(self, mu, sigma, tau_g, tau_a, latitude, left=True):
        super().__init__()
        self.left = left
        self.man = orientation_manifold(left)
        self.ukf = Ukf(self.man, mu, sigma)
        self.Q = np.zeros((13, 13))
        self.tau_g, self.tau_a = tau_g, tau_a
        self.earth = np.array([EARTHW * math.cos(latitude), 0.0, EARTHW * math.sin(latitude)])
        self.gyro = np.zeros(3)
        self.acc = np.array([0.0, 0.0, float(mu[13])])

    def _model(self, X, dt):
        X = X.copy()
        w = qrot(X[:, 0:4], self.gyro - X[:, 7:10]) - self.earth
        e = so3_exp(w, dt)
        X[:, 0:4] = qmul(e, X[:, 0:4]) if self.left else qmul(X[:, 0:4], e)
        a = qrot(X[:, 0:4], self.acc - X[:, 10:13])
        a[:, 2] -= X[:, 13]
        X[:, 4:7] = X[:, 4:7] + dt * a
        X[:, 7:10] = X[:, 7:10] + dt * ((-1.0 / self.tau_g) * X[:, 7:10])
        X[:, 10:13] = X[:, 10:13] + dt * ((-1.0 / self.tau_a) * X[:, 10:13])
        return X

    def predict_impl(self, dt):
        R = qmat(self.ukf.mu[0:4])
        pn = self.Q.copy()
        pn[0:3, 0:3] = R @ self.Q[0:3, 0:3] @ R.T
        pn[3:6, 3:6] = R @ self.Q[3:6, 3:6] @ R.T
        pn = dt**2 * pn
        self.ukf.predict(lambda X: self._model(X, dt), pn)

    def update_velocity(self, zmu, zcov):
        self.ukf.update(vec_manifold(3), np.asarray(zmu, float),
                        lambda X: qrot(qconj(X[:, 0:4]), X[:, 4:7]), np.asarray(zcov, float))

    def rotation_rate(self):
        return self.gyro - self.ukf.mu[7:10] - qrot(qconj(self.ukf.mu[0:4]), self.earth)
