/*
 * so3.cuh -- per-lane FP64 SO(3) / quaternion primitives of the batched UKF engine.
 *
 * Semantics follow MTK::SO3 (mtk/types/SOn.hpp, mtk/src/mtkmath.hpp) and the Eigen
 * quaternion operations the reference calls (PoseUKF.cpp:80-81,135,182;
 * OrientationUKF.cpp:19-22,38,81); conventions come from include/ukfb_constants.h.
 * Quaternions are double[4] stored x,y,z,w.
 */
#ifndef UKFB_SO3_CUH
#define UKFB_SO3_CUH

#include "simt.cuh"
#include "../../include/ukfb_constants.h"

namespace ukfb {

/* r = a * b (Hamilton product).  r may alias neither a nor b. */
UKFB_HD void quat_mul(const double* a, const double* b, double* r)
{
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}

/* r = a * conj(b) */
UKFB_HD void quat_mul_conj(const double* a, const double* b, double* r)
{
    r[3] = a[3] * b[3] + a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    r[0] = -a[3] * b[0] + a[0] * b[3] - a[1] * b[2] + a[2] * b[1];
    r[1] = -a[3] * b[1] + a[1] * b[3] - a[2] * b[0] + a[0] * b[2];
    r[2] = -a[3] * b[2] + a[2] * b[3] - a[0] * b[1] + a[1] * b[0];
}

/* r = conj(a) * b */
UKFB_HD void quat_conj_mul(const double* a, const double* b, double* r)
{
    r[3] = a[3] * b[3] + a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    r[0] = a[3] * b[0] - a[0] * b[3] - a[1] * b[2] + a[2] * b[1];
    r[1] = a[3] * b[1] - a[1] * b[3] - a[2] * b[0] + a[0] * b[2];
    r[2] = a[3] * b[2] - a[2] * b[3] - a[0] * b[1] + a[1] * b[0];
}

/* out = q * v, Eigen _transformVector: uv = 2 (q.vec x v); v + w uv + q.vec x uv */
UKFB_HD void quat_rotate(const double* q, const double* v, double* out)
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
UKFB_HD void quat_inv_rotate(const double* q, const double* v, double* out)
{
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double r = 1.0 / n2;
    const double qi[4] = {-q[0] * r, -q[1] * r, -q[2] * r, q[3] * r};
    quat_rotate(qi, v, out);
}

/* Eigen toRotationMatrix, row-major */
UKFB_HD void quat_matrix(const double* q, double* R)
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

/* MTK cos_sinc_sqrt: (cos sqrt(x2), sinc sqrt(x2)); Taylor pair below 2^-13. */
UKFB_HD void cos_sinc_sqrt(double x2, double& c, double& sinc)
{
    if (x2 >= UKFB_TAYLOR_N_BOUND) {
        const double x = sqrt(x2);
        double s;
        ukfb_sincos(x, &s, &c);
        sinc = s / x;
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
        c = cosi;
        sinc = si;
    }
}

/* MTK::SO3::exp(v, scale) */
UKFB_HD void so3_exp(const double* v, double scale, double* q)
{
    const double half = scale / 2.0;
    const double norm2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double c, sinc;
    cos_sinc_sqrt(half * half * norm2, c, sinc);
    const double mult = sinc * half;
    q[0] = mult * v[0];
    q[1] = mult * v[1];
    q[2] = mult * v[2];
    q[3] = c;
}

/* MTK::SO3::log(q) = (2/nv) atan(nv/w) q.vec, nv floored at MTK::tolerance */
UKFB_HD void so3_log(const double* q, double* out)
{
    double nv = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (nv < UKFB_MTK_TOLERANCE) nv = UKFB_MTK_TOLERANCE;
    const double s = 2.0 / nv * atan(nv / q[3]);
    out[0] = s * q[0];
    out[1] = s * q[1];
    out[2] = s * q[2];
}

/* q <- q [+] v*scale */
UKFB_HD void so3_boxplus(double* q, const double* v, double scale)
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
UKFB_HD void so3_boxminus(const double* q, const double* o, double* res)
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
