"""Generates the polynomial kernels of so3.cuh (near-minimax: interpolation at Chebyshev nodes, 60-digit
arithmetic):  atan(sqrt(u))/sqrt(u) on [0, 0.09],  cos(sqrt(v)) and sin(sqrt(v))/sqrt(v) on [0, 0.25] (general) and [0, 0.09] (fast kernels)."""
import mpmath as mp

mp.mp.dps = 60


def fit(f, b, deg):
    n = deg + 1
    nodes = [b / 2 + b / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x**j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def emit(name, c):
    print(f"/* {name} */")
    print("{" + ", ".join(float(x).hex() for x in c) + "}")
    print("{" + ", ".join(repr(float(x)) for x in c) + "}")


emit("ATAN_OVER_T, u = t^2 in [0, 0.09], degree 9", fit(lambda u: mp.atan(mp.sqrt(u)) / mp.sqrt(u), mp.mpf("0.09"), 9))
emit("COS_SQRT, v in [0, 0.25], degree 6", fit(lambda v: mp.cos(mp.sqrt(v)), mp.mpf("0.25"), 6))
emit("SINC_SQRT, v in [0, 0.25], degree 6", fit(lambda v: mp.sin(mp.sqrt(v)) / mp.sqrt(v), mp.mpf("0.25"), 6))
# fast kernels: half angle <= 0.3 rad, the same rotation angles (0.6 rad) as the log kernels; degree 5 is 2e-18 / 4e-19 from the functions
emit("COS_SQRT, v in [0, 0.09], degree 5", fit(lambda v: mp.cos(mp.sqrt(v)), mp.mpf("0.09"), 5))
emit("SINC_SQRT, v in [0, 0.09], degree 5", fit(lambda v: mp.sin(mp.sqrt(v)) / mp.sqrt(v), mp.mpf("0.09"), 5))
# log of a UNIT quaternion without a reciprocal: 2 asin(|vec|)/|vec| as a function of y = |vec|^2 = sin^2(theta/2),
# on the same angle range as the atan kernel (tan^2(theta/2) <= 0.09  <=>  y <= 0.09 / 1.09)
emit("TWO_ASIN_OVER_S, y = s^2 in [0, 0.09/1.09], degree 8",
     fit(lambda y: 2 * mp.asin(mp.sqrt(y)) / mp.sqrt(y), mp.mpf("0.09") / mp.mpf("1.09"), 8))
