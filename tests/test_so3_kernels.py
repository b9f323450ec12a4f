"""The SO(3) exp / log kernels (csrc/so3.cuh: near-minimax polynomials inside their ranges, libm outside) and the
Newton reciprocal / square root of csrc/simt.cuh against 50-digit mpmath values, over the whole range the filters can
reach: rotation angles from 1e-12 to just below pi, across the polynomial boundaries (half angle 0.5 rad in exp, 33 degrees
in log), the fast kernels' own pair (degree-5 exp, reciprocal-free log) on its range of 0.58 rad, and their any-angle pair
(eighth-angle polynomial + three quaternion squarings; three quaternion square roots + the asin form) over all angles.
CPU: the host build of the same source; GPU: ukfb_selftest_so3."""
from __future__ import annotations

import ctypes as C

import mpmath as mp
import numpy as np
import pytest

mp.mp.dps = 50


def inputs():
    rng = np.random.default_rng(9)
    ang = np.concatenate([10.0 ** rng.uniform(-12, -1, 300), rng.uniform(0.1, 3.1, 500), [0.999, 1.0, 1.001, 0.59, 0.6, 0.61, 3.1]])
    axis = rng.normal(size=(ang.size, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    v = axis * ang[:, None]
    x = 10.0 ** rng.uniform(-12, 12, ang.size)
    return v, x


def reference(v, x):
    out = np.empty((len(x), 10))
    for i in range(len(x)):
        vv = [mp.mpf(float(c)) for c in v[i]]
        n = mp.sqrt(sum(c * c for c in vv))
        h = n / 2
        s = mp.sin(h) / n if n > 0 else mp.mpf(0.5)
        q = [float(c * s) for c in vv] + [float(mp.cos(h))]
        out[i, :4] = q
        out[i, 4:7] = v[i]  # log(exp(v)) = v for |v| < pi
        xi = mp.mpf(float(x[i]))
        out[i, 7], out[i, 8], out[i, 9] = float(1 / xi), float(mp.sqrt(xi)), float(1 / mp.sqrt(xi))
    return out


def check(got, v, x):
    ref = reference(v, x)
    ang = np.linalg.norm(v, axis=1)
    eq = np.abs(got[:, :4] - ref[:, :4]).max(axis=1)
    assert eq.max() < 3e-16, f"exp: {eq.max():.2e} at angle {ang[eq.argmax()]}"
    # log(exp(v)): relative to the angle; near pi the conditioning of atan(nv / w) grows like 1 / (pi - angle)
    el = np.abs(got[:, 4:7] - ref[:, 4:7]).max(axis=1) / ang
    bound = 1e-15 * np.maximum(1.0, 0.2 / (np.pi - ang))
    assert (el < bound).all(), f"log: {el.max():.2e} at angle {ang[(el / bound).argmax()]}"
    # the fast kernels' pair: polynomial exp, reciprocal-free log (valid below 0.58 rad; flagged above)
    fast = got[:, 13] == 0.0
    assert fast[ang < 0.57].all() and not fast[ang > 0.6].any()
    ef = np.abs(got[fast, 10:13] - ref[fast, 4:7]).max(axis=1) / ang[fast]
    assert ef.max() < 8e-16, f"fast log: {ef.max():.2e} at angle {ang[fast][ef.argmax()]}"
    # the any-angle pair: every angle below pi (log(exp(v)) = v there), same conditioning near pi
    ew = np.abs(got[:, 14:18] - ref[:, :4]).max(axis=1)
    assert ew.max() < 2e-15, f"wide exp: {ew.max():.2e} at angle {ang[ew.argmax()]}"
    lw = np.abs(got[:, 18:21] - ref[:, 4:7]).max(axis=1) / ang
    boundw = 4e-15 * np.maximum(1.0, 0.2 / (np.pi - ang))
    assert (lw < boundw).all(), f"wide log: {lw.max():.2e} at angle {ang[(lw / boundw).argmax()]}"
    for col, name in ((7, "rcp"), (8, "sqrt"), (9, "rsqrt")):
        rel = np.abs(got[:, col] / ref[:, col] - 1.0)
        assert rel.max() < 3e-16, f"{name}: {rel.max():.2e}"


def test_host_build_of_the_so3_kernels():
    import emu_lib

    lib = emu_lib.load()
    v, x = inputs()
    v, x = np.ascontiguousarray(v), np.ascontiguousarray(x)
    out = np.empty((x.size, 21))
    lib.emu_selftest_so3(C.c_longlong(x.size), v.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    check(out, v, x)


@pytest.mark.gpu
def test_device_so3_kernels():
    from slam_pose_estimation_b200 import UkfBatch

    v, x = inputs()
    check(UkfBatch(0, 1).selftest_so3(v, x), v, x)
