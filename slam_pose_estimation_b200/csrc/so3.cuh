/*
 * so3.cuh -- per-lane FP64 SO(3) / quaternion primitives of the batched UKF engine.
 *
 * Semantics follow MTK::SO3 (mtk/types/SOn.hpp, mtk/src/mtkmath.hpp) and the Eigen
 * quaternion operations the reference calls (PoseUKF.cpp:80-81,135,182;
 * OrientationUKF.cpp:19-22,38,81); conventions come from include/ukfb_constants.h.
 * Quaternions are double[4] stored x,y,z,w.
 *
 * exp and log are the hot special functions of the filter (about 100 exp and 125 log
 * per predict+update).  On the GPU the common case -- rotation vectors of at most about
 * one radian in exp, relative rotations of at most about 33 degrees in log -- is
 * evaluated by near-minimax polynomials in the SQUARED argument (tools/gen_poly.py;
 * truncation below 1e-19), which needs no square root, no sin/cos/atan call and a
 * single reciprocal.  They evaluate the same functions as MTK's expressions
 * (cos x, sin x / x, 2 atan(|v|/w)/|v|) to within an ulp or two, well inside the
 * 1e-9 parity tolerance; outside that range the literal MTK expressions are used.
 * The two structure-exploiting kernels (ukf_pose_fast.cuh, ukf_ori_fast.cuh) use a
 * leaner pair on rotations of at most 0.58 rad: degree-5 cos / sinc and a log of a
 * unit quaternion without the reciprocal (2 asin(s)/s in s^2, SO3_ASIN_C below).
 */
#ifndef UKFB_SO3_CUH
#define UKFB_SO3_CUH

#include "simt.cuh"
#include "../../include/ukfb_constants.h"

namespace ukfb {

/* r = a * b (Hamilton product).  r may alias neither a nor b. */
UKFB_D void quat_mul(const double* a, const double* b, double* r)
{
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

/* r = a * conj(b) */
UKFB_D void quat_mul_conj(const double* a, const double* b, double* r)
{
    r[3] = a[3] * b[3] + a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    r[0] = -a[3] * b[0] + a[0] * b[3] - a[1] * b[2] + a[2] * b[1];
    r[1] = -a[3] * b[1] + a[1] * b[3] - a[2] * b[0] + a[0] * b[2];
    r[2] = -a[3] * b[2] + a[2] * b[3] - a[0] * b[1] + a[1] * b[0];
}

/* r = conj(a) * b */
UKFB_D void quat_conj_mul(const double* a, const double* b, double* r)
{
    r[3] = a[3] * b[3] + a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    r[0] = a[3] * b[0] - a[0] * b[3] - a[1] * b[2] + a[2] * b[1];
    r[1] = a[3] * b[1] - a[1] * b[3] - a[2] * b[0] + a[0] * b[2];
    r[2] = a[3] * b[2] - a[2] * b[3] - a[0] * b[1] + a[1] * b[0];
}

/* out = q * v, Eigen _transformVector: uv = 2 (q.vec x v); v + w uv + q.vec x uv */
UKFB_D void quat_rotate(const double* q, const double* v, double* out)
{
    double ux = q[1] * v[2] - q[2] * v[1];
    double uy = q[2] * v[0] - q[0] * v[2];
    double uz = q[0] * v[1] - q[1] * v[0];
    ux += ux;
    uy += uy;
    uz += uz;
    out[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    out[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    out[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}

/* out = q.inverse() * v with Eigen's inverse() = conj / squaredNorm (OrientationUKF.cpp:38) */
UKFB_D void quat_inv_rotate(const double* q, const double* v, double* out)
{
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double r = 1.0 / n2;
    const double qi[4] = {-q[0] * r, -q[1] * r, -q[2] * r, q[3] * r};
    quat_rotate(qi, v, out);
}

/* Eigen toRotationMatrix, row-major */
UKFB_D void quat_matrix(const double* q, double* R)
{
    const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1.0 - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = 1.0 - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = 1.0 - (txx + tyy);
}

/* ---- polynomial kernels (tools/gen_poly.py), Estrin form for instruction-level parallelism ---- */
constexpr double SO3_EXP_FAST_X2 = 0.25; /* (half angle)^2 bound of the cos/sinc polynomials */
constexpr double SO3_LOG_FAST_U = 0.09;  /* (|q.vec| / w)^2 bound of the atan polynomial */

/* coefficients live in constant memory so that each FMA takes its coefficient as a
 * constant-bank operand instead of two 32-bit immediates moved into registers */
UKFB_CONSTANT double SO3_COS_C[7] = {1.0, -0x1.fffffffffffffp-2, 0x1.5555555555421p-5, -0x1.6c16c16bdd04ep-10,
                                     0x1.a01a00fb1bc6fp-16, -0x1.27e40964b47d4p-22, 0x1.1d8d32755f8fbp-29};
UKFB_CONSTANT double SO3_SINC_C[7] = {1.0, -0x1.5555555555555p-3, 0x1.11111111110bfp-7, -0x1.a01a019ffb337p-13,
                                      0x1.71de39fd64b1fp-19, -0x1.ae63543245d1fp-26, 0x1.5fac6e0083f22p-33};
UKFB_CONSTANT double SO3_ATAN_C[10] = {1.0, -0x1.5555555555500p-2, 0x1.999999998a517p-3, -0x1.2492491c0c582p-3,
                                       0x1.c71c6cf04ff82p-4, -0x1.745c4ca68d45dp-4, 0x1.3aff6b481f0f7p-4,
                                       -0x1.0fcd05c851591p-4, 0x1.c90783e417298p-5, -0x1.229f36308eeefp-5};

/* cos / sinc for the fast kernels: half angle <= 0.3 rad, i.e. the same rotation angles (0.6 rad) as their log; degree 5
 * is 2e-18 / 4e-19 from the functions on that range */
constexpr double SO3_EXP5_FAST_X2 = 0.09;
UKFB_CONSTANT double SO3_COS5_C[6] = {0x1.0000000000000p+0, -0x1.ffffffffffff8p-2, 0x1.55555555535c1p-5, -0x1.6c16c160642f9p-10,
                                      0x1.a019c2f382291p-16, -0x1.274a2f4d3cb99p-22};
UKFB_CONSTANT double SO3_SINC5_C[6] = {0x1.0000000000000p+0, -0x1.5555555555554p-3, 0x1.1111111110759p-7, -0x1.a01a0198e6dbbp-13,
                                       0x1.71de13c236539p-19, -0x1.ada5cb577b3e1p-26};
#define UKFB_POLY5(C, v, v2) fma(fma(C[5], v, C[4]), (v2) * (v2), fma(fma(C[3], v, C[2]), v2, fma(C[1], v, C[0])))

/* 2 asin(s)/s as a function of y = s*s, y <= SO3_LOG_FAST_Y (the same angles as the atan kernel: tan^2 <= 0.09), degree 8:
 * 4e-17 from the function, and log(exp(v)) through pf_exp / pf_log within 7e-16 |v| of v (tests/test_so3_kernels.py) */
constexpr double SO3_LOG_FAST_Y = 0.09 / 1.09;
/* for a quaternion with | |q|^2 - 1 | <= 1e-7:  w >= SO3_LOG_FAST_W  implies  w > 0 and |vec|^2 <= SO3_LOG_FAST_Y
 * (sqrt(1 - Y) (1 + 1e-6): W^2 = 0.917433027, so |vec|^2 <= 1.0000001 - W^2 = 0.0825671 < Y = 0.0825688) */
constexpr double SO3_LOG_FAST_W = 0.957827243;
UKFB_CONSTANT double SO3_ASIN_C[9] = {0x1.0000000000000p+1, 0x1.555555555503dp-2, 0x1.333333340075ap-3, 0x1.6db6daa72c4afp-4,
                                      0x1.f1c77c63fd8f8p-5, 0x1.6e7ea9f45b674p-5, 0x1.1d54d58e79ab9p-5, 0x1.b1cba0f288f34p-6,
                                      0x1.05955d9d4e360p-5};

/* the coefficient arrays are named directly (not passed as pointers) so that the compiler can address them
 * as constant-bank operands */
#define UKFB_POLY6(C, v, v2) \
    fma(fma(C[6], v2, fma(C[5], v, C[4])), (v2) * (v2), fma(fma(C[3], v, C[2]), v2, fma(C[1], v, C[0])))

/* atan(t)/t as a function of u = t*t, u <= 0.09 */
UKFB_D double atan_over_t_poly(double u)
{
    const double u2 = u * u, u4 = u2 * u2;
    const double p01 = fma(SO3_ATAN_C[1], u, SO3_ATAN_C[0]), p23 = fma(SO3_ATAN_C[3], u, SO3_ATAN_C[2]);
    const double p45 = fma(SO3_ATAN_C[5], u, SO3_ATAN_C[4]), p67 = fma(SO3_ATAN_C[7], u, SO3_ATAN_C[6]);
    const double p89 = fma(SO3_ATAN_C[9], u, SO3_ATAN_C[8]);
    const double q0 = fma(p23, u2, p01), q1 = fma(p67, u2, p45);
    return fma(fma(p89, u4, q1), u4, q0);
}

/* 2 asin(sqrt y)/sqrt y, y <= SO3_LOG_FAST_Y */
UKFB_D double two_asin_over_s_poly(double y)
{
    const double y2 = y * y, y4 = y2 * y2;
    const double p01 = fma(SO3_ASIN_C[1], y, SO3_ASIN_C[0]), p23 = fma(SO3_ASIN_C[3], y, SO3_ASIN_C[2]);
    const double p45 = fma(SO3_ASIN_C[5], y, SO3_ASIN_C[4]), p67 = fma(SO3_ASIN_C[7], y, SO3_ASIN_C[6]);
    const double q0 = fma(p23, y2, p01), q1 = fma(p67, y2, p45);
    return fma(fma(SO3_ASIN_C[8], y4, q1), y4, q0);
}

/* MTK cos_sinc_sqrt: (cos sqrt(x2), sinc sqrt(x2)), the literal expressions (Taylor pair below 2^-13).
 * Out of line: only reached for half angles above 0.5 rad. */
UKFB_DNI void cos_sinc_sqrt(double x2, double* c_out, double* sinc_out)
{
    if (x2 >= UKFB_TAYLOR_N_BOUND) {
        const double x = sqrt(x2);
        double s, c;
        ukfb_sincos(x, &s, &c);
        *c_out = c;
        *sinc_out = s / x;
    } else {
        /* 1 - x2/2 + x2^2/24 - x2^3/720 and 1 - x2/6 + x2^2/120 - x2^3/5040, term by term
         * as mtkmath.hpp does */
        double cosi = 1.0, si = 1.0;
        double term = -0.5 * x2;
        cosi += term;
        term *= (1.0 / 3.0);
        si += term;
        term *= (-(1.0 / 4.0) * x2);
        cosi += term;
        term *= (1.0 / 5.0);
        si += term;
        term *= (-(1.0 / 6.0) * x2);
        cosi += term;
        term *= (1.0 / 7.0);
        si += term;
        *c_out = cosi;
        *sinc_out = si;
    }
}

/* the literal MTK::SO3::log scale (2/nv) atan(nv/w), nv floored at MTK::tolerance.  Out of line: only
 * reached for relative rotations above about 33 degrees. */
UKFB_DNI double so3_log_scale_slow(double nv2, double w)
{
    double nv = sqrt(nv2);
    if (nv < UKFB_MTK_TOLERANCE) nv = UKFB_MTK_TOLERANCE;
    return 2.0 / nv * atan(nv / w);
}

/* MTK::SO3::exp(v, scale): w = cos(|v| scale/2), vec = sinc(|v| scale/2) (scale/2) v */
UKFB_D void so3_exp(const double* v, double scale, double* q)
{
    const double half = scale / 2.0;
    const double norm2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double x2 = half * half * norm2;
    double c, sinc;
    if (x2 <= SO3_EXP_FAST_X2) {
        const double x4 = x2 * x2;
        c = UKFB_POLY6(SO3_COS_C, x2, x4);
        sinc = UKFB_POLY6(SO3_SINC_C, x2, x4);
    } else {
        cos_sinc_sqrt(x2, &c, &sinc);
    }
    const double mult = sinc * half;
    q[0] = mult * v[0];
    q[1] = mult * v[1];
    q[2] = mult * v[2];
    q[3] = c;
}

/* MTK::SO3::log(q) = (2/nv) atan(nv/w) q.vec, nv = |q.vec| floored at MTK::tolerance.
 * Fast path: (2/nv) atan(nv/w) = 2 P((nv/w)^2) / w. */
UKFB_D void so3_log(const double* q, double* out)
{
    const double nv2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const double w = q[3];
    double s;
    if (nv2 <= SO3_LOG_FAST_U * (w * w)) {
        const double rw = fast_rcp(w);
        const double t = nv2 * rw;
        s = (2.0 * rw) * atan_over_t_poly(t * rw);
    } else {
        s = so3_log_scale_slow(nv2, w);
    }
    out[0] = s * q[0];
    out[1] = s * q[1];
    out[2] = s * q[2];
}

/* ---- optimistic variants for straight-line code -------------------------------------------------------
 * FAST = true evaluates only the polynomial path, with no branch, and ORs `slow` when the argument was outside
 * the polynomial's range (the result is then meaningless); the caller discards the whole pass and redoes it with
 * FAST = false, which is the branching code above.  This keeps calls and branches out of the sigma-point loops,
 * so the scheduler can overlap independent sigma points. */
template <bool FAST>
UKFB_D void so3_exp_t(const double* v, double scale, double* q, bool& slow)
{
    if (!FAST) {
        so3_exp(v, scale, q);
        return;
    }
    const double half = scale / 2.0;
    const double norm2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double x2 = half * half * norm2;
    slow = slow || !(x2 <= SO3_EXP_FAST_X2);
    const double x4 = x2 * x2;
    const double c = UKFB_POLY6(SO3_COS_C, x2, x4);
    const double mult = UKFB_POLY6(SO3_SINC_C, x2, x4) * half;
    q[0] = mult * v[0];
    q[1] = mult * v[1];
    q[2] = mult * v[2];
    q[3] = c;
}

template <bool FAST>
UKFB_D void so3_log_t(const double* q, double* out, bool& slow)
{
    if (!FAST) {
        so3_log(q, out);
        return;
    }
    const double nv2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const double w = q[3];
    slow = slow || !(nv2 <= SO3_LOG_FAST_U * (w * w));
    const double rw = fast_rcp(w);
    const double t = nv2 * rw;
    const double s = (2.0 * rw) * atan_over_t_poly(t * rw);
    out[0] = s * q[0];
    out[1] = s * q[1];
    out[2] = s * q[2];
}

template <bool FAST>
UKFB_D void so3_boxplus_t(double* q, const double* v, double scale, bool& slow)
{
    double e[4], r[4];
    so3_exp_t<FAST>(v, scale, e, slow);
#if UKFB_SO3_BOXPLUS_LEFT
    quat_mul(e, q, r);
#else
    quat_mul(q, e, r);
#endif
    q[0] = r[0], q[1] = r[1], q[2] = r[2], q[3] = r[3];
}

template <bool FAST>
UKFB_D void so3_boxminus_t(const double* q, const double* o, double* res, bool& slow)
{
    double r[4];
#if UKFB_SO3_BOXPLUS_LEFT
    quat_mul_conj(q, o, r);
#else
    quat_conj_mul(o, q, r);
#endif
    so3_log_t<FAST>(r, res, slow);
}

/* q <- q [+] v*scale */
UKFB_D void so3_boxplus(double* q, const double* v, double scale)
{
    double e[4], r[4];
    so3_exp(v, scale, e);
#if UKFB_SO3_BOXPLUS_LEFT
    quat_mul(e, q, r);
#else
    quat_mul(q, e, r);
#endif
    q[0] = r[0], q[1] = r[1], q[2] = r[2], q[3] = r[3];
}

/* res = q [-] o */
UKFB_D void so3_boxminus(const double* q, const double* o, double* res)
{
    double r[4];
#if UKFB_SO3_BOXPLUS_LEFT
    quat_mul_conj(q, o, r);
#else
    quat_conj_mul(o, q, r);
#endif
    so3_log(r, res);
}

} /* namespace ukfb */

#endif /* UKFB_SO3_CUH */
