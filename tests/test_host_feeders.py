"""Host-side feeders (include/pose_estimation_b200/{GeographicProjection,GravitationalModel,StreamAlignmentVerifier}.hpp):
mirrors of the reference helpers that feed measurement buffers.  Pinned by the
reference's own test (test/test_coordinate_projection.cpp:8-54) and independent NumPy / SciPy computations."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest
from scipy.integrate import quad

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAT0, LON0 = 0.92698121, 0.154595663
A_WGS, F_WGS, K0 = 6378137.0, 1.0 / 298.257223563, 0.9996
E2 = F_WGS * (2 - F_WGS)


@pytest.fixture(scope="module")
def out(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "feeders_demo")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "feeders_demo.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, check=True)
    rows = {}
    for line in r.stdout.splitlines():
        tag, *rest = line.split()
        rows.setdefault(tag, []).append([float(v) for v in rest])
    return rows


def meridian_arc(phi):
    return quad(lambda t: A_WGS * (1 - E2) / (1 - E2 * np.sin(t) ** 2) ** 1.5, 0.0, phi, epsabs=1e-7, epsrel=1e-13)[0]


def snyder_tm(phi, lam):
    """Snyder, Map Projections (USGS PP 1395) eq. 8-9 / 8-10, with the meridian arc by quadrature"""
    ep2 = E2 / (1 - E2)
    N = A_WGS / np.sqrt(1 - E2 * np.sin(phi) ** 2)
    T, C, A = np.tan(phi) ** 2, ep2 * np.cos(phi) ** 2, (lam - LON0) * np.cos(phi)
    x = K0 * N * (A + (1 - T + C) * A**3 / 6 + (5 - 18 * T + T * T + 72 * C - 58 * ep2) * A**5 / 120)
    y = K0 * (meridian_arc(phi) - meridian_arc(LAT0) + N * np.tan(phi) * (A * A / 2 + (5 - T + 9 * C + 4 * C * C) * A**4 / 24
              + (61 - 58 * T + T * T + 600 * C - 330 * ep2) * A**6 / 720))
    return x, y  # easting, northing


def test_reference_projection_test_sequence(out):
    """test/test_coordinate_projection.cpp:8-54 -- identity, offset, inverse and the sign checks"""
    assert out["ref_identity_ok"] == [[1.0]] and out["ref_inverse_ok"] == [[1.0]]
    assert out["ref_identity"][0] == [0.0, 0.0]                      # pos.x() == 0, pos.y() == 0
    assert np.allclose(out["ref_inverse"][0], [LAT0, LON0], rtol=0, atol=2e-16)
    assert out["ref_offset"][0] == [-500.0, 1234.0]                  # pos.x() == -500, pos.y() == 1234
    assert np.allclose(out["ref_offset_inverse"][0], [LAT0, LON0], rtol=0, atol=2e-16)
    assert out["ref_plus"][0][0] > -500.0 and out["ref_plus"][0][1] < 1234.0   # north up, west positive
    assert out["ref_minus"][0][0] < LAT0 and out["ref_minus"][0][1] > LON0


def test_projection_against_meridian_arc_and_snyder_series(out):
    g = np.array(out["grid"])
    for la, lo, x, y, la2, lo2 in g:
        east, north = snyder_tm(la, lo)
        assert abs(x - north) < 2e-5 and abs(-y - east) < 2e-5, (la, lo, x - north, -y - east)
        assert abs(la2 - la) < 1e-15 and abs(lo2 - lo) < 1e-15      # round trip to an ulp
    # on the central meridian the northing is exactly k0 x the meridian arc
    on_cm = g[np.abs(g[:, 1] - LON0) < 1e-15]
    for la, lo, x, y, *_ in on_cm:
        assert abs(x - K0 * (meridian_arc(la) - meridian_arc(LAT0))) < 2e-8 and abs(y) < 1e-9
    assert out["domain"] == [[0.0, 0.0]]


def test_gravity_model(out):
    for la, alt, g in out["gravity"]:
        s2 = np.sin(la) ** 2
        ref = 9.7803267714 * (1 + 0.00193185138639 * s2) / np.sqrt(1 - 0.0818191908426**2 * s2) * (6378137.0 / (6378137.0 + alt)) ** 2
        assert abs(g - ref) < 1e-14
    eq, pole = out["gravity"][0][2], out["gravity"][9][2]
    assert abs(eq - 9.7803267714) < 1e-15 and abs(pole - 9.8321849) < 1e-5   # WGS-84 equator / pole values
    assert out["earthw"][0][0] == 2 * np.pi / 86164.0


def test_stream_alignment_verifier(out):
    v = {int(t): (int(f), int(c)) for t, f, c in out["verifier"]}
    assert v[1000000] == (99, 99)      # interval not elapsed: outputs untouched
    assert v[2500000] == (0, 0)        # first check only records
    assert v[3000000] == (0, 0)
    assert v[5000000] == (1, 0)        # dvl: 20 of 30 dropped
    assert v[7500000] == (0, 2)        # dvl 30/30, gps 16/16
    assert v[10000000] == (1, 0)       # imu 5 % with a 1 % warning threshold
    assert out["verifier_log_lines"][0][0] == 5   # dvl failure, gps too few, dvl + gps critical, imu failure
    assert out["config"][0] == [LAT0, 26.0]  # 2 x (3 + 3 + 3 + 1) + 3 + 3 doubles
