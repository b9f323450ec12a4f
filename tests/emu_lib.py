"""ctypes driver for the SIMT emulation build of the device code (tests/simt_emu).

TEST INFRASTRUCTURE ONLY.  Lets the GPU-less container run the product's kernel source
on host threads so its shared-memory indexing and phase logic can be compared with the
oracle before any GPU time is spent.  The GPU parity tests (-m gpu) go through the real
C ABI instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt_emu")
_LIB = None


Q_DIAGONAL_FLAG = 2  # 2: as the engine sets StepParams::q_diagonal; 1: never the isotropic-block shortcut; 0: always the general code


class StepParams(C.Structure):
    _fields_ = [
        ("state", C.c_void_p), ("Q", C.c_void_p), ("q_stride", C.c_longlong), ("B", C.c_longlong),
        ("status", C.c_void_p), ("t_last", C.c_void_p), ("hist", C.c_void_p),
        ("do_predict", C.c_int), ("time_mode", C.c_int), ("dt", C.c_void_p), ("dt_stride", C.c_longlong),
        ("ts", C.c_void_p), ("ts_stride", C.c_longlong), ("min_dt", C.c_double), ("max_dt", C.c_double),
        ("acc_mu", C.c_void_p), ("acc_cov", C.c_void_p), ("gyro_mu", C.c_void_p),
        ("neg_inv_tau_g", C.c_double), ("neg_inv_tau_a", C.c_double), ("earth", C.c_double * 3),
        ("do_update", C.c_int), ("kind", C.c_int), ("kinds", C.c_void_p), ("z", C.c_void_p), ("z_stride", C.c_int),
        ("R", C.c_void_p), ("r_stride", C.c_longlong), ("r_ld", C.c_int), ("mask", C.c_void_p),
        ("K", C.c_int), ("dt_kstride", C.c_longlong), ("ts_kstride", C.c_longlong), ("z_kstride", C.c_longlong),
        ("r_kstride", C.c_longlong), ("kinds_kstride", C.c_longlong), ("mask_kstride", C.c_longlong),
        ("tick_kinds", C.c_void_p), ("imu", C.c_void_p), ("imu_kstride", C.c_longlong),
        ("events", C.c_int), ("r_kind_stride", C.c_longlong), ("gate_d2", C.c_double),
        ("ori_params", C.c_void_p), ("prefetch_tiles", C.c_longlong), ("prefetch_bytes", C.c_int), ("q_diagonal", C.c_int),
        ("tile_done", C.c_void_p), ("launch_seq", C.c_ulonglong),
    ]


def load(sanitize: bool = False):
    global _LIB
    if _LIB is not None:
        return _LIB
    # UKFB_EMU_SANITIZE=1 (tools/run_emu_sanitized.sh, which also preloads libasan): the same sources under ASan + UBSan
    sanitize = sanitize or os.environ.get("UKFB_EMU_SANITIZE") == "1"
    out = os.path.join(_DIR, "libukfb_emu_san.so" if sanitize else "libukfb_emu.so")
    srcs = [os.path.join(_DIR, "emu_harness.cpp"), os.path.join(_DIR, "simt_emu_rt.hpp"),
            os.path.join(_DIR, "../../slam_pose_estimation_b200/csrc/ukf_device.cuh"),
            os.path.join(_DIR, "../../slam_pose_estimation_b200/csrc/so3.cuh"),
            os.path.join(_DIR, "../../slam_pose_estimation_b200/csrc/simt.cuh"),
            os.path.join(_DIR, "../../slam_pose_estimation_b200/csrc/ukf_thread.cuh"),
            os.path.join(_DIR, "../../slam_pose_estimation_b200/csrc/ukf_pose_fast.cuh"),
            os.path.join(_DIR, "../../slam_pose_estimation_b200/csrc/ukf_ori_fast.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        cmd = ["/usr/bin/g++", "-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", _DIR,
               "-o", out, srcs[0]]
        if sanitize:
            cmd[3:3] = ["-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
            cmd[2] = "-O1"
        subprocess.run(cmd, check=True)
    _LIB = C.CDLL(out)
    assert _LIB.emu_sizeof_params() == C.sizeof(StepParams), (_LIB.emu_sizeof_params(), C.sizeof(StepParams))
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p).value


class EmuBatch:
    """Host-side state + calls into the emulated kernel; mirrors OracleBatch method names."""

    def __init__(self, kind: int, B: int, G: int = 4, kernel: str = "warp"):
        """kernel: 'warp' (ukf_device.cuh, AoS records), 'thread' (ukf_thread.cuh, 32-filter entry-major tiles) or
        'fast' (ukf_pose_fast.cuh, PoseUKF only, same tiles)"""
        self.lib = load()
        self.kind, self.B, self.G, self.tiled = kind, B, G, kernel in ("thread", "fast")
        self.fast = kernel == "fast"
        self.n, self.MU, self.REC = (12, 13, 91) if kind == 0 else (13, 14, 105)
        self.LP = self.n * (self.n + 1) // 2
        self.tril = np.tril_indices(self.n)
        self.Bpad = (B + 31) // 32 * 32
        self.state = np.zeros((self.Bpad, self.REC))  # logical [filter][entry] view; see _to_device / _from_device
        self.status = np.zeros(B, np.uint32)
        self.t_last = np.zeros(B, np.int64)
        self.hist = np.zeros(64 * 8, np.uint64)
        self.Q = np.zeros(self.LP)
        if kind == 0:
            Qd = np.diag([0.01] * 3 + [0.001] * 3 + [1e-5] * 3 + [1e-5] * 3)
            self.Q = Qd[self.tril].copy()
        self.q_stride = 0
        self.acc_mu = np.full((B, 3), np.nan) if kind == 0 else np.zeros((B, 3))
        self.acc_cov = np.tile(np.eye(3).ravel(), (B, 1))
        self.gyro_mu = np.zeros((B, 3))
        self.min_dt, self.max_dt = 1e-9, np.finfo(float).max
        self.tau_g = self.tau_a = np.inf
        self.earth = np.array([2.0 * np.pi / 86164.0, 0.0, 0.0])  # latitude 0 until set_orientation_params
        self._first_init = True
        self.gate_d2 = np.inf
        self.ori_params = None
        self.tile_done = np.zeros(self.Bpad // 32, np.uint64)  # fast kernels: lanes finished per tile, 32 per launch
        self.fast_launches = 0

    def initialize(self, mu, sigma):
        mu = np.asarray(mu, float).reshape(self.B, self.MU)
        sigma = np.asarray(sigma, float).reshape(self.B, self.n, self.n)
        self.state[:] = 0
        self.state[: self.B, : self.MU] = mu
        self.state[: self.B, self.MU : self.MU + self.LP] = sigma[:, self.tril[0], self.tril[1]]
        self.t_last[:] = 0
        if self.kind == 1 and self._first_init:
            self.acc_mu[:] = 0
            self.acc_mu[:, 2] = mu[:, 13]
        self._first_init = False

    def get_state(self):
        mu = self.state[: self.B, : self.MU].copy()
        sg = np.zeros((self.B, self.n, self.n))
        sg[:, self.tril[0], self.tril[1]] = self.state[: self.B, self.MU : self.MU + self.LP]
        sg = sg + np.transpose(np.tril(sg, -1), (0, 2, 1))
        return mu, sg

    def set_process_noise(self, Q):
        Q = np.asarray(Q, float)
        if Q.ndim == 3:
            self.Q = np.ascontiguousarray(Q[:, self.tril[0], self.tril[1]])
            self.q_stride = self.LP
        else:
            self.Q = np.ascontiguousarray(Q[self.tril])
            self.q_stride = 0

    def set_time_bounds(self, a, b):
        self.min_dt, self.max_dt = a, b

    def set_mahalanobis_gate(self, max_d2):
        self.gate_d2 = float(max_d2)

    def set_orientation_params(self, tau_g, tau_a, lat):
        w = 2.0 * np.pi / 86164.0
        if np.ndim(tau_g) or np.ndim(tau_a) or np.ndim(lat):
            tg, ta, la = (np.broadcast_to(np.asarray(a, float), (self.B,)) for a in (tau_g, tau_a, lat))
            self.ori_params = np.ascontiguousarray(np.stack([-1.0 / tg, -1.0 / ta, w * np.cos(la), np.zeros(self.B), w * np.sin(la)], axis=1))
            return
        self.ori_params = None
        self.tau_g, self.tau_a = tau_g, tau_a
        self.earth = np.array([w * np.cos(lat), 0.0, w * np.sin(lat)])

    def set_acceleration(self, mu, cov=None, mask=None):
        sel = slice(None) if mask is None else np.asarray(mask, bool)
        self.acc_mu[sel] = np.asarray(mu, float).reshape(self.B, 3)[sel]
        if cov is not None:
            cov = np.asarray(cov, float)
            c = cov.reshape(self.B, 9) if cov.ndim == 3 else np.tile(cov.ravel(), (self.B, 1))
            self.acc_cov[sel] = c[sel]

    def set_rotation_rate(self, mu, cov=None, mask=None):
        sel = slice(None) if mask is None else np.asarray(mask, bool)
        self.gyro_mu[sel] = np.asarray(mu, float).reshape(self.B, 3)[sel]

    def _launch(self, **kw):
        p = StepParams()
        p.state = _ptr(self.state)
        self._Q = np.ascontiguousarray(self.Q)
        p.Q = _ptr(self._Q)
        if self.q_stride == 0 and Q_DIAGONAL_FLAG:  # as the engine does for a broadcast Q without off-diagonal entries
            full = np.zeros((self.n, self.n))
            full[self.tril] = self._Q
            dg = np.diag(full)
            p.q_diagonal = int(not np.any(full - np.diag(dg)))
            if p.q_diagonal and Q_DIAGONAL_FLAG == 2 and dg[0] == dg[1] == dg[2] and dg[3] == dg[4] == dg[5]:
                p.q_diagonal = 2
        p.q_stride = self.q_stride
        p.B = self.B
        p.K = 1
        p.status = _ptr(self.status)
        p.t_last = _ptr(self.t_last)
        p.hist = _ptr(self.hist)
        p.min_dt, p.max_dt = self.min_dt, self.max_dt
        p.gate_d2 = self.gate_d2
        p.ori_params = _ptr(self.ori_params)
        p.acc_mu, p.acc_cov, p.gyro_mu = _ptr(self.acc_mu), _ptr(self.acc_cov), _ptr(self.gyro_mu)
        p.neg_inv_tau_g, p.neg_inv_tau_a = -1.0 / self.tau_g, -1.0 / self.tau_a
        p.earth = (C.c_double * 3)(*self.earth)
        keep = []
        for k, v in kw.items():
            if isinstance(v, np.ndarray):
                keep.append(v)
                setattr(p, k, _ptr(v))
            else:
                setattr(p, k, v)
        if self.tiled:  # [tile][entry][lane] in memory
            dev = np.ascontiguousarray(self.state.reshape(-1, 32, self.REC).transpose(0, 2, 1))
            p.state = _ptr(dev)
            if self.fast:  # as the engine launches them: each launch waits, tile by tile, for the one before
                p.tile_done = _ptr(self.tile_done)
                self.fast_launches += 1
                p.launch_seq = self.fast_launches
            if self.fast and self.kind == 1:
                rc = self.lib.emu_ori_fast_step(C.byref(p))
            elif self.fast:
                rc = self.lib.emu_pose_fast_step(C.byref(p))
            else:
                rc = self.lib.emu_thread_step(C.c_int(self.kind), C.byref(p))
            self.state[:] = dev.transpose(0, 2, 1).reshape(self.Bpad, self.REC)
            if self.fast:
                assert (self.tile_done == 32 * self.fast_launches).all(), "a lane did not report its tile done"
        else:
            rc = self.lib.emu_step(C.c_int(self.kind), C.c_int(self.G), C.byref(p))
        assert rc == 0

    def predict_dt(self, dt):
        dt = np.ascontiguousarray(np.atleast_1d(np.asarray(dt, float)))
        self._launch(do_predict=1, time_mode=0, dt=dt, dt_stride=1 if dt.size == self.B else 0)

    def predict_time(self, ts):
        ts = np.ascontiguousarray(np.atleast_1d(np.asarray(ts, np.int64)))
        self._launch(do_predict=1, time_mode=1, ts=ts, ts_stride=1 if ts.size == self.B else 0)

    def _upd_args(self, kind, mu, cov, mask):
        m = {1: 2, 5: 2, 7: 2, 2: 1, 6: 1}.get(kind, 3)
        mu = np.ascontiguousarray(np.asarray(mu, float).reshape(self.B, m))
        cov = np.ascontiguousarray(np.asarray(cov, float))
        a = dict(do_update=1, kind=kind, z=mu, z_stride=m, R=cov, r_stride=m * m if cov.ndim == 3 else 0, r_ld=m)
        if mask is not None:
            a["mask"] = np.ascontiguousarray(np.asarray(mask, np.uint8))
        return a

    def update(self, kind, mu, cov, mask=None):
        self._launch(**self._upd_args(kind, mu, cov, mask))

    def update_mixed(self, kinds, mu3, cov33):
        self._launch(do_update=1, kind=-2, kinds=np.ascontiguousarray(np.asarray(kinds, np.int8)),
                     z=np.ascontiguousarray(np.asarray(mu3, float)), z_stride=3,
                     R=np.ascontiguousarray(np.asarray(cov33, float)), r_stride=9, r_ld=3)

    def step(self, dt, kind, mu, cov, mask=None):
        dt = np.ascontiguousarray(np.atleast_1d(np.asarray(dt, float)))
        a = self._upd_args(kind, mu, cov, mask)
        self._launch(do_predict=1, time_mode=0, dt=dt, dt_stride=1 if dt.size == self.B else 0, **a)

    def run_events(self, ts, kinds, mu3, cov):
        """same arguments as UkfBatch.run_events"""
        ts = np.ascontiguousarray(np.asarray(ts, np.int64))
        K, B = ts.size // self.B, self.B
        cov = np.ascontiguousarray(np.asarray(cov, float))
        per_event = cov.shape != (13, 3, 3)
        assert not per_event or cov.size == K * B * 9
        self._launch(K=K, events=1, do_predict=1, time_mode=1, ts=ts, ts_stride=1, ts_kstride=B, do_update=1, kind=-2,
                     kinds=np.ascontiguousarray(np.asarray(kinds, np.int8)), kinds_kstride=B,
                     z=np.ascontiguousarray(np.asarray(mu3, float)), z_stride=3, z_kstride=3 * B, R=cov, r_ld=3,
                     r_stride=9 if per_event else 0, r_kstride=9 * B if per_event else 0,
                     r_kind_stride=0 if per_event else 9)

    def fallbacks(self):
        """(literal predict, literal update, literal apply_delta, predicts served by the any-angle instance of the
        structured code, apply_deltas / updates served by it) calls made so far by the fast kernel's lanes"""
        out = (C.c_ulonglong * 5)()
        (self.lib.emu_ori_fast_fallbacks if self.kind == 1 else self.lib.emu_pose_fast_fallbacks)(out)
        return np.array(list(out), np.int64)

    def get_status(self):
        return self.status.copy()

    def get_mean_iter_hist(self):
        return self.hist.reshape(64, 8).sum(axis=0)
