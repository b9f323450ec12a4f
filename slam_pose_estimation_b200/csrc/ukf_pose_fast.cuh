/*
 * ukf_pose_fast.cuh -- the structure-exploiting lane-per-filter step kernel for PoseUKF (sm_100a).
 *
 * Same StepParams, same HBM tile layout and same results (to rounding) as ukf_thread.cuh, which stays the general
 * ("literal") kernel: every lane owns one filter.  This kernel evaluates the SAME estimator -- ukfom predict /
 * update / apply_delta over the 2n+1 sigma points mu [+] +-L[:,j] with the PoseUKF models
 * (PoseUKF.cpp:7-97,180-196) -- but uses what is known at compile time about those models, so that sums the
 * reference evaluates term by term are taken in closed form wherever they are EXACT identities:
 *
 *   predict (PoseUKF.cpp:75-97): velocity and angular velocity pass through the process model unchanged, hence
 *     their deviations from the mean are exactly +-L[6:12,j]:
 *       - the velocity/angular-velocity block of the new covariance is  Sigma[6:12,6:12] + Q[6:12,6:12];
 *       - its cross block with position/orientation is  1/2 sum_j L[6:12,j] (d+_j - d-_j)^T;
 *       - only the 21 position/orientation entries are accumulated point by point.
 *     L is lower triangular, so columns j >= 6 perturb neither position nor orientation: those 12 sigma points
 *     share the rotation of the prior orientation (one 3x3 matrix) and need no sigma-point boxplus.
 *     The +/- points of a column share exp(L_ori), conj for the minus point.
 *   update with a component-selector measurement (all PoseUKF models but the orientation one, PoseUKF.cpp:7-69):
 *     Z_p = mu[sel] +- L[sel,j], so zbar = mu[sel], S = Sigma[sel,sel] + R and Sigma_xz = Sigma[:,sel] exactly
 *     (valid while every |L_ori[:,j]| < pi: guaranteed by trace(Sigma_ori) < 9 in the hot path, checked column by column
 *     out of line beyond that -- only a column next to pi sends the filter's update to the literal code);
 *     gain and delta = K innov follow the reference's expression order; in Sigma - K S K^T the product K S is taken as
 *     Sigma_xz (K = Sigma_xz S^-1), which the reference recomputes.
 *   update with the orientation measurement (PoseUKF.cpp:28-33,133-138): Z_p = exp(+-L_ori[:,j]) q for columns 0..5, q
 *     otherwise; iterative SO(3) mean, S and Sigma_xz from the 12 evaluated points (pf_update<true>, reached through the
 *     out-of-line slow-path call of the kernel instance ukf_pose_fast_kernel<true>).
 *   apply_delta: the Euclidean components of mu [+] (delta +- L[:,j]) have mean mu + delta and deviations +-L, so
 *     the Euclidean block of the new covariance is that of Sigma - K S K^T; only the orientation rows/columns are
 *     recomputed, from the 12 points of columns 0..5 (the others carry the orientation of X_0), and only the first
 *     six columns of the second Cholesky factor are needed.
 *
 * All SO(3) exp/log calls here are branch-free polynomial kernels (pf_exp: degree-5 cos / sinc; pf_log: no reciprocal, for
 * quaternions of unit norm, which pf_unit checks once per phase; their range checks are integer comparisons).  A lane
 * whose argument leaves the polynomial range -- a filter that barely knows its attitude -- is served by a second,
 * out-of-line instance of the same structured code with an any-angle exp / log pair (pf_predict_slow / pf_update_slow);
 * what is beyond even that, or fails a guard, falls back to the literal code of ukf_thread.cuh (cold), which is also
 * what UKFB_KERNEL=thread runs for every filter.  The kernel has a second instance (OVERLAP) for handles whose
 * consecutive launches are ordered tile by tile instead of launch by launch (StepParams::tile_done).
 * Covariance accumulators (57 / 33 doubles) and the state stay in registers, both Cholesky factorisations run in
 * registers on statically indexed arrays; shared memory ([entry][lane], conflict free) only holds the factor
 * columns, which the sigma-point loops index dynamically.
 */
#ifndef UKFB_POSE_FAST_CUH
#define UKFB_POSE_FAST_CUH

#include "ukf_thread.cuh"

namespace ukfb {

#define UKFB_PS(e) sm[(e) * TILE + lane]

/* tuning knobs of the sigma-point loops (column pairs in flight per iteration) and of the register budget; the defaults
 * are what measured best on B200 (profiles/r02_kernel_experiments.txt) */
#ifndef UKFB_PF_MEAN_UNROLL
#define UKFB_PF_MEAN_UNROLL 1
#endif
#ifndef UKFB_PF_COV_UNROLL
#define UKFB_PF_COV_UNROLL 1
#endif
#ifndef UKFB_PF_MIN_BLOCKS
#define UKFB_PF_MIN_BLOCKS 2 /* x 128 threads: 255 registers per thread */
#endif
#ifndef UKFB_PF_MAX_THREADS
#define UKFB_PF_MAX_THREADS (4 * TILE)
#endif
#define UKFB_PRAGMA_(x) _Pragma(#x)
#ifdef UKFB_SIMT_EMU
#define UKFB_LOOP_UNROLL(n)
#else
#define UKFB_LOOP_UNROLL(n) UKFB_PRAGMA_(unroll n)
#endif


/* shared-memory slots (doubles per lane).  The factor is stored as two square blocks with explicit zeros above the
 * diagonal so that the column loops need no triangular index tests:
 *   LA(j, i) = j * 12 + i        columns 0..5,  rows 0..11
 *   LB(j, i) = 72 + (j-6)*6 + (i-6)   columns 6..11, rows 6..11
 * Between predict and update the first 78 slots hold the predicted covariance, packed lower (dynamic selector
 * indexing).  The literal fallback runs on a per-thread local array instead (cold path). */
constexpr int PF_LA = 0;
constexpr int PF_LB = 72;
constexpr int PF_PER_LANE = 108;
constexpr double PF_PI2_GUARD = 9.0; /* trace(Sigma_ori) below this (< pi^2) makes (mu [+] L_j) [-] mu = L_j exact */
constexpr double PF_PI2_COLUMN = 9.6; /* what that needs: every |L_ori[:, j]|^2 below pi^2 (checked out of line when the trace is not) */

/* The range checks of the polynomial pair, on the upper words (simt.cuh, hi_word): integer instructions instead of two
 * FP64-pipe comparisons per sigma point.  Conservative by less than 1e-6 of the bound: x2 is in range when its upper word
 * is BELOW that of the bound (a NaN or a negative value is not), w when its upper word lies above that of its bound and
 * not above that of 1 + 2e-6 (a NaN, an infinity or a negative w does not). */
UKFB_D bool pf_exp_in_range(double x2) { return unsigned(hi_word(x2)) < unsigned(hi_word(SO3_EXP5_FAST_X2)); }
UKFB_D bool pf_log_in_range(double w)
{
    const unsigned first = unsigned(hi_word(SO3_LOG_FAST_W)) + 1u, span = unsigned(hi_word(1.000002)) - first;
    return unsigned(hi_word(w)) - first <= span;
}

/* a Cholesky pivot the factorisation can go on with: positive, normal and below 1e300 (dpotf2 stops at ajj <= 0 or NaN; an
 * infinite or subnormal pivot is treated the same here), again on the upper word */
UKFB_D bool pivot_ok(double ajj) { return unsigned(hi_word(ajj)) - 0x00100000u < unsigned(hi_word(1.0e300)) - 0x00100000u; }

/* ---- branch-free SO(3) kernels: polynomial path only, `slow` collects range violations ------------------------ */
UKFB_D void pf_exp(const double* v, double scale, double* q, bool& slow)
{
    const double half = scale * 0.5;
    const double norm2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double x2 = half * half * norm2;
    slow = slow || !pf_exp_in_range(x2);
    const double x4 = x2 * x2;
    const double c = UKFB_POLY5(SO3_COS5_C, x2, x4);
    const double mult = UKFB_POLY5(SO3_SINC5_C, x2, x4) * half;
    q[0] = mult * v[0];
    q[1] = mult * v[1];
    q[2] = mult * v[2];
    q[3] = c;
}

/* MTK::SO3::log(q) = (2 / nv) atan(nv / w) q.vec without the reciprocal of w, for a q of norm 1 + O(1e-7) (pf_unit guards
 * the state's quaternion once per phase; every other factor is an exp).  With |q|^2 = 1 + d and y = nv^2 / |q|^2 =
 * sin^2(theta / 2) the scale-invariant value is [2 asin(sqrt y) / sqrt y] / |q| with y = nv^2 (1 - d) and 1 / |q| = 1 - d / 2
 * to first order in d: the neglected d^2 is below 1e-14, and d itself is 1e-13 after 10 000 steps. */
UKFB_D void pf_log(const double* q, double* out, bool& slow)
{
    const double nv2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const double w = q[3];
    const double d = fma(w, w, nv2) - 1.0;
    slow = slow || !pf_log_in_range(w); /* one comparison: w > 0 and nv2 within the polynomial's range (|q| = 1 to 1e-7) */
    const double s0 = two_asin_over_s_poly(fma(-nv2, d, nv2));
    const double s = fma(s0, -0.5 * d, s0);
    out[0] = s * q[0];
    out[1] = s * q[1];
    out[2] = s * q[2];
}

/* is |q|^2 within PF_QNORM_TOL of 1?  (A state initialised with an unnormalised quaternion, which the reference's
 * scale-invariant log tolerates, runs the literal code.) */
constexpr double PF_QNORM_TOL = 2.5e-8;
UKFB_D bool pf_unit(const double* q)
{
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    return fabs(n2 - 1.0) <= PF_QNORM_TOL;
}

/* q * v as quat_rotate (Eigen _transformVector: v + 2 (w u + q.vec x u), u = q.vec x v), with the sums contracted into
 * fused multiply-adds: 18 instead of 21 FP64 instructions */
UKFB_D void pf_rotate(const double* q, const double* v, double* out)
{
    const double ux = fma(q[1], v[2], -(q[2] * v[1]));
    const double uy = fma(q[2], v[0], -(q[0] * v[2]));
    const double uz = fma(q[0], v[1], -(q[1] * v[0]));
    out[0] = fma(2.0, fma(q[3], ux, fma(q[1], uz, -(q[2] * uy))), v[0]);
    out[1] = fma(2.0, fma(q[3], uy, fma(q[2], ux, -(q[0] * uz))), v[1]);
    out[2] = fma(2.0, fma(q[3], uz, fma(q[0], uy, -(q[1] * ux))), v[2]);
}

UKFB_D void pf_matvec(const double* Rm, const double* v, double* out)
{
    out[0] = Rm[0] * v[0] + Rm[1] * v[1] + Rm[2] * v[2];
    out[1] = Rm[3] * v[0] + Rm[4] * v[1] + Rm[5] * v[2];
    out[2] = Rm[6] * v[0] + Rm[7] * v[1] + Rm[8] * v[2];
}

/* ---- two sigma points at a time -----------------------------------------------------------------------------------
 * A dependent FP64 instruction can issue 15 cycles after its producer on B200 while the pipe accepts one warp
 * instruction every 2 cycles (tools/microbench.cu, profiles/r02d_microbench.txt): a scheduler needs ~8 independent FP64
 * instructions in flight, and with the two resident warps per scheduler that 255 registers allow each warp has to bring
 * about four.  The +/- points of a sigma-point pair are independent, but written one after the other they are also
 * SCHEDULED one after the other (ptxas keeps the source order of equal-priority chains).  D2 carries both points through
 * the same expressions, so that every source line emits the two points' instructions next to each other and a chain's
 * next instruction finds twice as many independent ones in between. */
struct D2 {
    double a, b;
    UKFB_D D2() {}
    UKFB_D explicit D2(double x) : a(x), b(x) {}
    UKFB_D D2(double x, double y) : a(x), b(y) {}
};
UKFB_D D2 operator+(D2 x, D2 y) { return D2(x.a + y.a, x.b + y.b); }
UKFB_D D2 operator-(D2 x, D2 y) { return D2(x.a - y.a, x.b - y.b); }
UKFB_D D2 operator*(D2 x, D2 y) { return D2(x.a * y.a, x.b * y.b); }
UKFB_D D2 operator-(D2 x) { return D2(-x.a, -x.b); }
UKFB_D D2 operator+(D2 x, double y) { return D2(x.a + y, x.b + y); }
UKFB_D D2 operator-(D2 x, double y) { return D2(x.a - y, x.b - y); }
UKFB_D D2 operator*(D2 x, double y) { return D2(x.a * y, x.b * y); }
UKFB_D D2 operator*(double x, D2 y) { return D2(x * y.a, x * y.b); }
UKFB_D D2 operator+(double x, D2 y) { return D2(x + y.a, x + y.b); }
UKFB_D double tfma(double x, double y, double z) { return fma(x, y, z); }
UKFB_D D2 tfma(D2 x, D2 y, D2 z) { return D2(fma(x.a, y.a, z.a), fma(x.b, y.b, z.b)); }
UKFB_D D2 tfma(double x, D2 y, D2 z) { return D2(fma(x, y.a, z.a), fma(x, y.b, z.b)); }
UKFB_D D2 tfma(D2 x, double y, D2 z) { return D2(fma(x.a, y, z.a), fma(x.b, y, z.b)); }
UKFB_D D2 tfma(D2 x, D2 y, double z) { return D2(fma(x.a, y.a, z), fma(x.b, y.b, z)); }
UKFB_D D2 tfma(double x, D2 y, double z) { return D2(fma(x, y.a, z), fma(x, y.b, z)); }
UKFB_D D2 tfma(D2 x, double y, double z) { return D2(fma(x.a, y, z), fma(x.b, y, z)); }
UKFB_D D2 tfma(double x, double y, D2 z) { return D2(fma(x, y, z.a), fma(x, y, z.b)); }
UKFB_D bool all_le(D2 x, double lim) { return (x.a <= lim) && (x.b <= lim); }
#define UKFB_TPOLY5(C, v, v2) tfma(tfma(C[5], v, C[4]), (v2) * (v2), tfma(tfma(C[3], v, C[2]), v2, tfma(C[1], v, C[0])))

/* pf_exp / pf_log / pf_rotate and the quaternion products of so3.cuh, the same expressions, on a pair */
UKFB_D void pf_exp(const D2* v, double scale, D2* q, bool& slow)
{
    const double half = scale * 0.5;
    const D2 norm2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const D2 x2 = (half * half) * norm2;
    slow = slow || !(pf_exp_in_range(x2.a) && pf_exp_in_range(x2.b));
    const D2 x4 = x2 * x2;
    const D2 c = UKFB_TPOLY5(SO3_COS5_C, x2, x4);
    const D2 mult = UKFB_TPOLY5(SO3_SINC5_C, x2, x4) * half;
    q[0] = mult * v[0];
    q[1] = mult * v[1];
    q[2] = mult * v[2];
    q[3] = c;
}

UKFB_D D2 two_asin_over_s_poly(D2 y)
{
    const D2 y2 = y * y, y4 = y2 * y2;
    const D2 p01 = tfma(SO3_ASIN_C[1], y, SO3_ASIN_C[0]), p23 = tfma(SO3_ASIN_C[3], y, SO3_ASIN_C[2]);
    const D2 p45 = tfma(SO3_ASIN_C[5], y, SO3_ASIN_C[4]), p67 = tfma(SO3_ASIN_C[7], y, SO3_ASIN_C[6]);
    const D2 q0 = tfma(p23, y2, p01), q1 = tfma(p67, y2, p45);
    return tfma(tfma(SO3_ASIN_C[8], y4, q1), y4, q0);
}

UKFB_D void pf_log(const D2* q, D2* out, bool& slow)
{
    const D2 nv2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const D2 w = q[3];
    const D2 d = tfma(w, w, nv2) - 1.0;
    slow = slow || !(pf_log_in_range(w.a) && pf_log_in_range(w.b));
    const D2 s0 = two_asin_over_s_poly(tfma(-nv2, d, nv2));
    const D2 s = tfma(s0, -0.5 * d, s0);
    out[0] = s * q[0];
    out[1] = s * q[1];
    out[2] = s * q[2];
}

UKFB_D void pf_rotate(const D2* q, const D2* v, D2* out)
{
    const D2 ux = tfma(q[1], v[2], -(q[2] * v[1]));
    const D2 uy = tfma(q[2], v[0], -(q[0] * v[2]));
    const D2 uz = tfma(q[0], v[1], -(q[1] * v[0]));
    out[0] = tfma(2.0, tfma(q[3], ux, tfma(q[1], uz, -(q[2] * uy))), v[0]);
    out[1] = tfma(2.0, tfma(q[3], uy, tfma(q[2], ux, -(q[0] * uz))), v[1]);
    out[2] = tfma(2.0, tfma(q[3], uz, tfma(q[0], uy, -(q[1] * ux))), v[2]);
}

/* r = a * b and r = a * conj(b); B is D2 or double (a factor both points share) */
template <class B>
UKFB_D void quat_mul(const D2* a, const B* b, D2* r)
{
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
template <class B>
UKFB_D void quat_mul_conj(const D2* a, const B* b, D2* r)
{
    r[3] = a[3] * b[3] + a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    r[0] = -a[3] * b[0] + a[0] * b[3] - a[1] * b[2] + a[2] * b[1];
    r[1] = -a[3] * b[1] + a[1] * b[3] - a[2] * b[0] + a[0] * b[2];
    r[2] = -a[3] * b[2] + a[2] * b[3] - a[0] * b[1] + a[1] * b[0];
}

/* ---- the second tier: SO(3) exp / log for ANY rotation angle, still branch-free, still no libm ----------------------
 * A filter whose heading is barely known (sigma_ori of a radian or more: the normal start-up state of a pose filter) has
 * sigma points far outside the 0.58 rad of the short polynomials above.  Columns whose points leave that range are redone
 * with this pair instead of sending the whole filter to the literal code:
 *   exp: the polynomial pair at an eighth of the angle (degree 6, half angle / 8 <= 0.4 rad), then three quaternion
 *        squarings q <- q q (w <- w^2 - |v|^2, v <- 2 w v); rotations up to 6.4 rad;
 *   log: q == -q folded to w >= 0, three quaternion square roots ((v, w) -> (v / (2 c), c), c = sqrt((1 + w) / 2): Goldschmidt
 *        from the hardware reciprocal-square-root seed), which leaves a half angle <= pi / 16, then the reciprocal-free
 *        asin form of pf_log, times 8.
 * Against 50-digit values both are within 2e-15 of MTK's cos / sinc / atan expressions (tests/test_so3_kernels.py). */
#ifndef UKFB_PF_WIDE_CALLS
#define UKFB_PF_WIDE_CALLS 1
#endif
constexpr double PF_EXP_WIDE_X2 = 10.24; /* (half angle)^2 bound of pf_exp_wide: 3.2 rad */
#ifndef UKFB_PF_WIDE_TRACE
#define UKFB_PF_WIDE_TRACE 1.0
#endif
constexpr double PF_WIDE_TRACE = UKFB_PF_WIDE_TRACE;    /* trace(Sigma_ori) above this (0.58 rad per axis): sigma points certainly outside the short polynomials */
#define UKFB_TPOLY6(C, v, v2) \
    tfma(tfma(C[6], v2, tfma(C[5], v, C[4])), (v2) * (v2), tfma(tfma(C[3], v, C[2]), v2, tfma(C[1], v, C[0])))
UKFB_D bool all_le(double x, double lim) { return x <= lim; }
UKFB_D double v_sign(double w) { return w < 0.0 ? -1.0 : 1.0; }
UKFB_D D2 v_sign(D2 w) { return D2(v_sign(w.a), v_sign(w.b)); }
UKFB_D void v_sqrt_rsqrt(double x, double& sq, double& rs) { fast_sqrt_rsqrt(x, sq, rs); }
UKFB_D void v_sqrt_rsqrt(D2 x, D2& sq, D2& rs)
{
    fast_sqrt_rsqrt(x.a, sq.a, rs.a);
    fast_sqrt_rsqrt(x.b, sq.b, rs.b);
}

template <class V>
UKFB_D void pf_exp_wide(const V* v, double scale, V* q, bool& hard)
{
    const double half = scale * 0.5;
    const V norm2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const V x2 = (half * half) * norm2;
    hard = hard || !all_le(x2, PF_EXP_WIDE_X2);
    const V y2 = x2 * (1.0 / 64.0), y4 = y2 * y2;
    V w = UKFB_TPOLY6(SO3_COS_C, y2, y4);
    const V mult = UKFB_TPOLY6(SO3_SINC_C, y2, y4) * (half * 0.125);
    V a = mult * v[0], b = mult * v[1], c = mult * v[2];
    UKFB_UNROLL
    for (int k = 0; k < 3; ++k) {
        const V t = w + w, n = a * a + b * b + c * c;
        w = tfma(w, w, -n);
        a = t * a, b = t * b, c = t * c;
    }
    q[0] = a, q[1] = b, q[2] = c, q[3] = w;
}

template <class V>
UKFB_D void pf_log_wide(const V* q, V* out)
{
    V x = q[0], y = q[1], z = q[2], w = q[3];
    /* |q| = 1 + O(1e-7) (pf_unit): normalised to first order, then q == -q folded to w >= 0 (MTK's atan(nv / w) form) */
    const V n2 = tfma(w, w, x * x + y * y + z * z);
    const V k = v_sign(w) * tfma(n2 - 1.0, -0.5, 1.0);
    x = x * k, y = y * k, z = z * k, w = w * k;
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) { /* (v, w) <- its square root */
        V c, rc;
        v_sqrt_rsqrt((w + 1.0) * 0.5, c, rc);
        const V h = rc * 0.5;
        x = x * h, y = y * h, z = z * h, w = c;
    }
    const V nv2 = x * x + y * y + z * z;
    const V d = tfma(w, w, nv2) - 1.0;
    const V s0 = two_asin_over_s_poly(tfma(-nv2, d, nv2));
    const V sc = tfma(s0, -0.5 * d, s0) * 8.0;
    out[0] = sc * x, out[1] = sc * y, out[2] = sc * z;
}

/* The any-angle pair is only ever reached from the out-of-line instance, whose code the warps of an SM stream through
 * the instruction cache at different places: one copy of each (a call) instead of one per use keeps that instance's
 * footprint near the hot path's (profiles/r02_kernel_experiments.txt: instruction-fetch stalls). */
template <class V> struct PfV3 { V v[3]; };
template <class V> struct PfQ4 { V q[4]; bool hard; };
#if UKFB_PF_WIDE_CALLS
template <class V>
UKFB_DNI PfQ4<V> pf_exp_wide_call(PfV3<V> v, double scale)
{
    PfQ4<V> r;
    r.hard = false;
    pf_exp_wide<V>(v.v, scale, r.q, r.hard);
    return r;
}
template <class V>
UKFB_DNI PfV3<V> pf_log_wide_call(PfQ4<V> q)
{
    PfV3<V> r;
    pf_log_wide<V>(q.q, r.v);
    return r;
}
template <class V>
UKFB_D void pf_exp_wide_shared(const V* v, double scale, V* q, bool& hard)
{
    PfV3<V> a;
    a.v[0] = v[0], a.v[1] = v[1], a.v[2] = v[2];
    const PfQ4<V> r = pf_exp_wide_call<V>(a, scale);
    q[0] = r.q[0], q[1] = r.q[1], q[2] = r.q[2], q[3] = r.q[3];
    hard = hard || r.hard;
}
template <class V>
UKFB_D void pf_log_wide_shared(const V* q, V* out)
{
    PfQ4<V> a;
    a.q[0] = q[0], a.q[1] = q[1], a.q[2] = q[2], a.q[3] = q[3], a.hard = false;
    const PfV3<V> r = pf_log_wide_call<V>(a);
    out[0] = r.v[0], out[1] = r.v[1], out[2] = r.v[2];
}
#else
template <class V>
UKFB_D void pf_exp_wide_shared(const V* v, double scale, V* q, bool& hard) { pf_exp_wide<V>(v, scale, q, hard); }
template <class V>
UKFB_D void pf_log_wide_shared(const V* q, V* out) { pf_log_wide<V>(q, out); }
#endif

/* single values in the two instances of the structured code: the hot one (WIDE = false) uses the short polynomials and
 * flags a range exit, the out-of-line one (WIDE = true) the any-angle pair (`slow`: beyond even its range) */
template <bool WIDE>
UKFB_D void pf_exp1(const double* v, double scale, double* q, bool& slow)
{
    if (WIDE)
        pf_exp_wide_shared<double>(v, scale, q, slow);
    else
        pf_exp(v, scale, q, slow);
}
template <bool WIDE>
UKFB_D void pf_log1(const double* q, double* out, bool& slow)
{
    if (WIDE)
        pf_log_wide_shared<double>(q, out);
    else
        pf_log(q, out, slow);
}

/* ---- Cholesky of a packed lower 12x12 in registers, first NCOL columns (LAPACK dpotf2('L') order) ------------- */
template <int NCOL>
UKFB_D bool pf_cholesky(double* a)
{
    bool ok = true;
    UKFB_UNROLL
    for (int j = 0; j < NCOL; ++j) {
        double ajj = a[tri(j, j)];
        UKFB_UNROLL
        for (int k = 0; k < j; ++k) ajj -= a[tri(j, k)] * a[tri(j, k)];
        if (!pivot_ok(ajj)) {
            ok = false;
            ajj = 1.0;
        }
        double d, rinv;
        fast_sqrt_rsqrt(ajj, d, rinv);
        a[tri(j, j)] = d;
        UKFB_UNROLL
        for (int i = j + 1; i < 12; ++i) {
            double s = a[tri(i, j)];
            UKFB_UNROLL
            for (int k = 0; k < j; ++k) s -= a[tri(i, k)] * a[tri(j, k)];
            a[tri(i, j)] = s * rinv;
        }
    }
    return ok;
}

struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

struct PoseMu {
    double p[3], q[4], v[3], w[3];
};

/* A warp that starts on `tile` asks L2 for the record of tile + p.prefetch_tiles, the one a warp of the next wave starts
 * on about when this one retires: that warp's first loads then cost an L2 hit instead of a DRAM access.  Lane l requests
 * the addresses (l + 32 i) * prefetch_bytes of the record. */
template <class F>
UKFB_D void prefetch_next_wave(const StepParams& p, long long tile, int lane)
{
    if (p.prefetch_tiles <= 0) return;
    const long long nt = tile + p.prefetch_tiles;
    if ((nt + 1) * TILE > p.B) return;
    const char* rec = reinterpret_cast<const char*>(p.state + nt * (TILE * F::REC));
    const int bytes = TILE * F::REC * int(sizeof(double));
    UKFB_NOUNROLL
    for (int o = lane * p.prefetch_bytes; o < bytes; o += TILE * p.prefetch_bytes) prefetch_l2(rec + o);
    /* the per-filter inputs of that tile's first tick: stored IMU sample, measurement, time step (a few lines each) */
    const long long nb = nt * TILE;
    const int o = lane * p.prefetch_bytes;
    if (p.acc_mu && o < TILE * 24) prefetch_l2(reinterpret_cast<const char*>(p.acc_mu + nb * 3) + o);
    if (F::KIND == 1 && p.gyro_mu && o < TILE * 24) prefetch_l2(reinterpret_cast<const char*>(p.gyro_mu + nb * 3) + o);
    if (p.do_update && p.z && o < TILE * 8 * p.z_stride) prefetch_l2(reinterpret_cast<const char*>(p.z + nb * p.z_stride) + o);
    if (p.do_update && p.R && p.r_stride > 0 && o < TILE * 8 * p.r_stride) /* one measurement covariance per filter */
        prefetch_l2(reinterpret_cast<const char*>(p.R + nb * p.r_stride) + o);
    if (p.do_predict && !p.time_mode && p.dt && o < TILE * 8 * p.dt_stride) prefetch_l2(reinterpret_cast<const char*>(p.dt + nb * p.dt_stride) + o);
}

/* exp / log of a pair with the short polynomials (WIDE = false: `flag` collects range exits, the values are then meaningless
 * and the caller redoes the column with WIDE = true) or with the any-angle pair (`flag`: beyond even its range) */
template <bool WIDE, class V>
UKFB_D void pf_exp_t(const V* v, double scale, V* q, bool& flag)
{
    if (WIDE)
        pf_exp_wide_shared<V>(v, scale, q, flag);
    else
        pf_exp(v, scale, q, flag);
}
template <bool WIDE, class V>
UKFB_D void pf_log_t(const V* q, V* out, bool& flag)
{
    if (WIDE)
        pf_log_wide_shared<V>(q, out);
    else
        pf_log(q, out, flag);
}


/* a pair of propagated sigma points: g(x) [-] ref for the position and orientation components (PoseUKF.cpp:75-83) */
template <bool WIDE>
UKFB_D void pf_point2(const D2* qs, const D2* ps, const D2* vs, const D2* ws, double dt, const double* ref_p, const double* ref_q,
                      D2* d, bool& slow)
{
    D2 rv[3], e[4], qn[4], r[4];
    pf_rotate(qs, vs, rv);
    /* q [+] (q w) dt = exp((q w) dt) q = q exp(w dt) q^-1 q = q exp(w dt) for a unit q: w needs no rotation */
    pf_exp_t<WIDE>(ws, dt, e, slow);
    quat_mul(qs, e, qn);
    d[0] = tfma(dt, rv[0], ps[0]) - ref_p[0];
    d[1] = tfma(dt, rv[1], ps[1]) - ref_p[1];
    d[2] = tfma(dt, rv[2], ps[2]) - ref_p[2];
    quat_mul_conj(qn, ref_q, r);
    pf_log_t<WIDE>(r, d + 3, slow);
}

/* the +/- sigma points of a column j < 6 through the process model; L receives the column (12 entries) */
template <bool WIDE>
UKFB_D void pf_pair_a(const double* sm, int lane, int j, const PoseMu& m, const double* vm, double dt, const double* ref_p,
                      const double* ref_q, double* L, double* dpl, double* dmi, bool& slow)
{
    UKFB_UNROLL
    for (int i = 0; i < 12; ++i) L[i] = UKFB_PS(PF_LA + j * 12 + i);
    double e[4];
    pf_exp_t<WIDE>(L + 3, 1.0, e, slow);
    /* exp(+-Lo) * q = e.w q +- t,  t = (e.vec, 0) * q */
    const double* q = m.q;
    const double t0 = e[0] * q[3] + e[1] * q[2] - e[2] * q[1];
    const double t1 = e[1] * q[3] + e[2] * q[0] - e[0] * q[2];
    const double t2 = e[2] * q[3] + e[0] * q[1] - e[1] * q[0];
    const double t3 = -(e[0] * q[0] + e[1] * q[1] + e[2] * q[2]);
    const D2 qs[4] = {D2(fma(e[3], q[0], t0), fma(e[3], q[0], -t0)), D2(fma(e[3], q[1], t1), fma(e[3], q[1], -t1)),
                      D2(fma(e[3], q[2], t2), fma(e[3], q[2], -t2)), D2(fma(e[3], q[3], t3), fma(e[3], q[3], -t3))};
    const D2 ps[3] = {D2(m.p[0] + L[0], m.p[0] - L[0]), D2(m.p[1] + L[1], m.p[1] - L[1]), D2(m.p[2] + L[2], m.p[2] - L[2])};
    const D2 vs[3] = {D2(vm[0] + L[6], vm[0] - L[6]), D2(vm[1] + L[7], vm[1] - L[7]), D2(vm[2] + L[8], vm[2] - L[8])};
    const D2 ws[3] = {D2(m.w[0] + L[9], m.w[0] - L[9]), D2(m.w[1] + L[10], m.w[1] - L[10]), D2(m.w[2] + L[11], m.w[2] - L[11])};
    D2 d[6];
    pf_point2<WIDE>(qs, ps, vs, ws, dt, ref_p, ref_q, d, slow);
    UKFB_UNROLL
    for (int i = 0; i < 6; ++i) dpl[i] = d[i].a, dmi[i] = d[i].b;
}

/* the +/- sigma points of a column j >= 6: position and orientation unperturbed.
 * rw0 = R w, c = q * conj(ref_q).  Outputs the orientation deviations and wv = (R Lv) dt, the +- offset of the
 * propagated position; L receives rows 6..11 of the column. */
template <bool WIDE>
UKFB_D void pf_pair_b(const double* sm, int lane, int j, const double* Rm, const double* rw0, const double* c, double dt,
                      double* L, double* dopl, double* domi, double* wv, bool& slow)
{
    UKFB_UNROLL
    for (int i = 0; i < 6; ++i) L[i] = UKFB_PS(PF_LB + (j - 6) * 6 + i);
    double u[3], rv[3];
    pf_matvec(Rm, L + 3, u);
    pf_matvec(Rm, L, rv);
    wv[0] = dt * rv[0], wv[1] = dt * rv[1], wv[2] = dt * rv[2];
    const D2 rw[3] = {D2(rw0[0] + u[0], rw0[0] - u[0]), D2(rw0[1] + u[1], rw0[1] - u[1]), D2(rw0[2] + u[2], rw0[2] - u[2])};
    D2 e[4], r[4], d[3];
    pf_exp_t<WIDE>(rw, dt, e, slow);
    quat_mul(e, c, r);
    pf_log_t<WIDE>(r, d, slow);
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) dopl[i] = d[i].a, domi[i] = d[i].b;
}

/* the +/- points of column j < 6 of apply_delta: exp(delta_ori +- L_ori[:, j]) c, deviations from the reference */
template <bool WIDE>
UKFB_D void pf_pair_d(const double* sm, int lane, int j, const double* delta, const double* c, double* dp, double* dn, bool& slow)
{
    D2 v[3], e[4], r2[4], d[3]; /* the + and - point of the column, side by side */
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        const double l = UKFB_PS(PF_LA + j * 12 + 3 + i);
        v[i] = D2(delta[3 + i] + l, delta[3 + i] - l);
    }
    pf_exp_t<WIDE>(v, 1.0, e, slow);
    quat_mul(e, c, r2);
    pf_log_t<WIDE>(r2, d, slow);
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) dp[i] = d[i].a, dn[i] = d[i].b;
}

/* The column pairs as the two instances of the structured code call them.  The hot instance (WIDE = false) knows only the
 * short polynomials.  The out-of-line instance (WIDE = true) picks per column: orientation entries this large go to the
 * any-angle pair at once; any other column (the position columns, the velocity / angular-velocity columns of a filter
 * that is unsure of its attitude only) tries the short polynomials first and is redone if a point left their range. */
#ifndef UKFB_PF_WIDE_COLUMN_N2
#define UKFB_PF_WIDE_COLUMN_N2 0.2
#endif
constexpr double PF_WIDE_COLUMN_N2 = UKFB_PF_WIDE_COLUMN_N2; /* |L_ori[:, j]|^2 above this: do not bother with the short polynomials */
template <bool WIDE>
UKFB_D void pf_pair_a_sel(const double* sm, int lane, int j, const PoseMu& m, const double* vm, double dt, const double* ref_p,
                          const double* ref_q, double* L, double* dpl, double* dmi, bool& slow)
{
    if (!WIDE) {
        pf_pair_a<false>(sm, lane, j, m, vm, dt, ref_p, ref_q, L, dpl, dmi, slow);
        return;
    }
    const double l3 = UKFB_PS(PF_LA + j * 12 + 3), l4 = UKFB_PS(PF_LA + j * 12 + 4), l5 = UKFB_PS(PF_LA + j * 12 + 5);
    bool redo = l3 * l3 + l4 * l4 + l5 * l5 > PF_WIDE_COLUMN_N2;
    if (!redo) pf_pair_a<false>(sm, lane, j, m, vm, dt, ref_p, ref_q, L, dpl, dmi, redo);
    if (redo) pf_pair_a<true>(sm, lane, j, m, vm, dt, ref_p, ref_q, L, dpl, dmi, slow);
}
template <bool WIDE>
UKFB_D void pf_pair_b_sel(const double* sm, int lane, int j, const double* Rm, const double* rw0, const double* c, double dt,
                          double* L, double* dopl, double* domi, double* wv, bool& slow)
{
    if (!WIDE) {
        pf_pair_b<false>(sm, lane, j, Rm, rw0, c, dt, L, dopl, domi, wv, slow);
        return;
    }
    bool redo = false;
    pf_pair_b<false>(sm, lane, j, Rm, rw0, c, dt, L, dopl, domi, wv, redo);
    if (redo) pf_pair_b<true>(sm, lane, j, Rm, rw0, c, dt, L, dopl, domi, wv, slow);
}
template <bool WIDE>
UKFB_D void pf_pair_d_sel(const double* sm, int lane, int j, const double* delta, const double* c, double* dp, double* dn, bool& slow)
{
    if (!WIDE) {
        pf_pair_d<false>(sm, lane, j, delta, c, dp, dn, slow);
        return;
    }
    const double l3 = UKFB_PS(PF_LA + j * 12 + 3), l4 = UKFB_PS(PF_LA + j * 12 + 4), l5 = UKFB_PS(PF_LA + j * 12 + 5);
    bool redo = l3 * l3 + l4 * l4 + l5 * l5 > PF_WIDE_COLUMN_N2;
    if (!redo) pf_pair_d<false>(sm, lane, j, delta, c, dp, dn, redo);
    if (redo) pf_pair_d<true>(sm, lane, j, delta, c, dp, dn, slow);
}

#ifdef UKFB_SIMT_EMU
/* host emulation only: how often each fallback ran (predict, update, apply), so the tests can tell that they did */
inline unsigned long long pf_fallbacks[5] = {0, 0, 0, 0, 0}; /* [3], [4]: predicts / apply_deltas served by the any-angle instance */
#define UKFB_PF_COUNT(i) __atomic_fetch_add(&pf_fallbacks[i], 1ull, __ATOMIC_RELAXED)
#else
#define UKFB_PF_COUNT(i)
#endif
/* ---- literal fallbacks (cold, out of line): the general code of ukf_thread.cuh on this lane's filter ---------- */
struct PfLit { /* result of a literal fallback, returned by value so that the caller's state stays in registers */
    PoseMu m;
    uint32_t status;
    int passes;
};

UKFB_D void pf_mu_to_slots(double* loc, const PoseMu& m)
{
    typedef TSmem<PoseF> TS;
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        loc[TS::OFF_MU + i] = m.p[i];
        loc[TS::OFF_MU + 7 + i] = m.v[i];
        loc[TS::OFF_MU + 10 + i] = m.w[i];
    }
    UKFB_UNROLL
    for (int i = 0; i < 4; ++i) loc[TS::OFF_MU + 3 + i] = m.q[i];
}

UKFB_D void pf_mu_from_slots(const double* loc, PoseMu& m)
{
    typedef TSmem<PoseF> TS;
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        m.p[i] = loc[TS::OFF_MU + i];
        m.v[i] = loc[TS::OFF_MU + 7 + i];
        m.w[i] = loc[TS::OFF_MU + 10 + i];
    }
    UKFB_UNROLL
    for (int i = 0; i < 4; ++i) m.q[i] = loc[TS::OFF_MU + 3 + i];
}

/* The literal functions of ukf_thread.cuh index their scratch as [entry * ST + lane]; here they run on a per-thread
 * local array (ST = 1, lane = 0), so the fast kernel's shared memory only has to hold the factor blocks. */
UKFB_DNI PfLit pf_literal_predict(double* sig, const double* Qp, const double* acov, ModelArgs ma, PoseMu m)
{
    typedef TSmem<PoseF> TS;
    UKFB_PF_COUNT(0);
    double loc[TS::PER_LANE];
    PfLit r;
    r.m = m, r.passes = 0;
    pf_mu_to_slots(loc, m);
    if (!cholesky_thread<PoseF, 1>(sig, loc, 0)) {
        r.status = UKFB_STATUS_NOT_SPD;
        return r;
    }
    store_noise<PoseF, 1>(loc, 0, sig, Qp, acov, ma);
    r.status = mean_and_cov<PoseF, true, 1>(loc, 0, sig, ma, &r.passes);
    UKFB_UNROLL
    for (int i = 0; i < PoseF::MU; ++i) loc[TS::OFF_MU + i] = loc[TS::OFF_REF + i];
    pf_mu_from_slots(loc, r.m);
    return r;
}

/* apply_delta (second half of ukfom update) from the record's covariance: first = true runs the first half too */
struct PfDelta {
    double d[12];
};

UKFB_DNI PfLit pf_literal_update(double* sig, int kind, const double* zm, const double* Rm, int r_ld, ModelArgs ma, PoseMu m,
                                 PfDelta delta, bool first, double gate_d2)
{
    typedef TSmem<PoseF> TS;
    UKFB_PF_COUNT(first ? 1 : 2);
    double loc[TS::PER_LANE];
    PfLit r;
    r.m = m, r.passes = 0, r.status = 0;
    pf_mu_to_slots(loc, m);
    if (first) {
        if (!cholesky_thread<PoseF, 1>(sig, loc, 0)) {
            r.status = UKFB_STATUS_NOT_SPD;
            return r;
        }
        r.status = update_first_half<PoseF, 1>(loc, 0, sig, kind, zm, Rm, r_ld, gate_d2);
        if (r.status & UKFB_STATUS_MEAS_REJECTED) return r; /* gated out: nothing was modified */
    } else {
        UKFB_UNROLL
        for (int i = 0; i < 12; ++i) loc[TS::OFF_DELTA + i] = delta.d[i];
    }
    if (!cholesky_thread<PoseF, 1>(sig, loc, 0)) {
        r.status |= UKFB_STATUS_NOT_SPD;
        return r;
    }
    r.status |= mean_and_cov<PoseF, false, 1>(loc, 0, sig, ma, &r.passes);
    UKFB_UNROLL
    for (int i = 0; i < PoseF::MU; ++i) loc[TS::OFF_MU + i] = loc[TS::OFF_REF + i];
    pf_mu_from_slots(loc, r.m);
    return r;
}

/* ---- structured predict.  Returns false when a polynomial range was left (nothing has been modified then) ------ */
/* On success: m holds the new mean, the record and (when to_smem) slots 0..77 hold the new covariance.
 * `a` (the prior covariance, packed lower) is destroyed. */
/* WIDE = false: the hot instance, inlined into the kernel: short polynomials only, gives up (false) at the first sigma
 * point outside their range.  WIDE = true: the instance behind pf_predict_wide (out of line), which serves filters of
 * any orientation uncertainty by running the passes that need it through pf_predict_pass_wide. */
template <bool WIDE>
UKFB_D bool pf_predict(int q_diagonal, double* sm, int lane, double* sig, double* a, const double* Qp, const double* acov,
                       const ModelArgs& ma, PoseMu& m, bool to_smem, uint32_t& status, int& passes_out, bool& spd)
{
    const double dt = ma.dt;
    /* a filter whose orientation uncertainty is this large certainly has sigma points outside the range of the short
     * polynomials: the hot instance does not even start (nothing has been modified) */
    if (!WIDE && a[tri(3, 3)] + a[tri(4, 4)] + a[tri(5, 5)] > PF_WIDE_TRACE) return false;
    /* Cholesky of the covariance (a: loaded from the record by the caller), in registers; the factor goes to the
     * LA / LB blocks */
    {
        spd = pf_cholesky<12>(a);
        if (!spd) return true;
        UKFB_UNROLL
        for (int j = 0; j < 6; ++j) {
            UKFB_UNROLL
            for (int i = 0; i < 12; ++i) UKFB_PS(PF_LA + j * 12 + i) = i >= j ? a[tri(i, j)] : 0.0;
        }
        UKFB_UNROLL
        for (int j = 6; j < 12; ++j) {
            UKFB_UNROLL
            for (int i = 6; i < 12; ++i) UKFB_PS(PF_LB + (j - 6) * 6 + (i - 6)) = i >= j ? a[tri(i, j)] : 0.0;
        }
    }
    bool slow = !pf_unit(m.q);
    double Rm[9];
    quat_matrix(m.q, Rm);
    double vm[3] = {m.v[0], m.v[1], m.v[2]};
    if (ma.has_acc) {
        vm[0] = fma(dt, ma.acc[0], vm[0]);
        vm[1] = fma(dt, ma.acc[1], vm[1]);
        vm[2] = fma(dt, ma.acc[2], vm[2]);
    }
    /* X0' = g(mu) */
    double rw0[3], p0n[3], q0n[4];
    {
        double rv0[3], e0[4];
        pf_matvec(Rm, vm, rv0);
        pf_matvec(Rm, m.w, rw0);
        p0n[0] = fma(dt, rv0[0], m.p[0]);
        p0n[1] = fma(dt, rv0[1], m.p[1]);
        p0n[2] = fma(dt, rv0[2], m.p[2]);
        pf_exp1<WIDE>(rw0, dt, e0, slow);
        quat_mul(e0, m.q, q0n);
    }
    double ref_p[3] = {p0n[0], p0n[1], p0n[2]};
    double ref_q[4] = {q0n[0], q0n[1], q0n[2], q0n[3]};

    /* ---- manifold mean (ukfom sigma_points_mean): only position and orientation can move */
    int it = 0, passes = 0;
    while (true) {
        double md[6];
        double d0[6];
        {
            double r[4];
            d0[0] = p0n[0] - ref_p[0], d0[1] = p0n[1] - ref_p[1], d0[2] = p0n[2] - ref_p[2];
            quat_mul_conj(q0n, ref_q, r);
            pf_log1<WIDE>(r, d0 + 3, slow);
        }
        /* X0 and the 12 points of columns 6..11 deviate by d0 in position (the +- offsets cancel) */
        md[0] = 13.0 * d0[0], md[1] = 13.0 * d0[1], md[2] = 13.0 * d0[2];
        md[3] = d0[3], md[4] = d0[4], md[5] = d0[5];
        UKFB_LOOP_UNROLL(UKFB_PF_MEAN_UNROLL)
        for (int j = 0; j < 6; ++j) {
            double L[12], dpl[6], dmi[6];
            pf_pair_a_sel<WIDE>(sm, lane, j, m, vm, dt, ref_p, ref_q, L, dpl, dmi, slow);
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) md[i] += dpl[i] + dmi[i];
        }
        double c[4];
        quat_mul_conj(m.q, ref_q, c);
        UKFB_LOOP_UNROLL(UKFB_PF_MEAN_UNROLL)
        for (int j = 6; j < 12; ++j) {
            double L[6], dopl[3], domi[3], wv[3];
            pf_pair_b_sel<WIDE>(sm, lane, j, Rm, rw0, c, dt, L, dopl, domi, wv, slow);
            UKFB_UNROLL
            for (int i = 0; i < 3; ++i) md[3 + i] += dopl[i] + domi[i];
        }
        double n2 = 0.0;
        UKFB_UNROLL
        for (int i = 0; i < 6; ++i) {
            md[i] = div_ns<PoseF::NS>(md[i]);
            n2 += md[i] * md[i];
        }
        ref_p[0] += md[0], ref_p[1] += md[1], ref_p[2] += md[2];
        {
            double e[4], r[4];
            pf_exp1<WIDE>(md + 3, 1.0, e, slow);
            quat_mul(e, ref_q, r);
            ref_q[0] = r[0], ref_q[1] = r[1], ref_q[2] = r[2], ref_q[3] = r[3];
        }
        ++passes;
        if (slow || !(n2 > UKFB_MEAN_TOL * UKFB_MEAN_TOL)) break;
        if (++it >= UKFB_MEAN_MAX_IT) {
            status |= UKFB_STATUS_MEAN_NO_CONVERGE;
            break;
        }
    }
    if (slow) return false;

    /* ---- covariance: C = position/orientation block, X = cross block (rows 6..11 x columns 0..5).
     * X = sum_j L[6:12, j] (d+_j - d-_j)^T is NOT accumulated point by point: 36 accumulators next to the 21 of C and the
     * ~46 doubles of per-filter context pushed the loops over the register file (spills reloaded on every column).  The
     * differences d+_j - d-_j of columns 0..5 go to the six slots L[0:6, j] of the column they came from (consumed once the
     * column's sigma points exist), the columns 6..11 accumulate only their orientation part (18 values; their position
     * part is the closed form 2 dt R L[6:9, j]), and X is formed after the loops from the factor rows still in place. */
    double C[21], X[36];
    {
        double d0[6], r[4];
        d0[0] = p0n[0] - ref_p[0], d0[1] = p0n[1] - ref_p[1], d0[2] = p0n[2] - ref_p[2];
        quat_mul_conj(q0n, ref_q, r);
        pf_log1<WIDE>(r, d0 + 3, slow);
        UKFB_UNROLL
        for (int i = 0; i < 6; ++i) {
            UKFB_UNROLL
            for (int k = 0; k <= i; ++k) C[tri(i, k)] = d0[i] * d0[k];
        }
        UKFB_LOOP_UNROLL(UKFB_PF_COV_UNROLL)
        for (int j = 0; j < 6; ++j) {
            double L[12], dpl[6], dmi[6];
            pf_pair_a_sel<WIDE>(sm, lane, j, m, vm, dt, ref_p, ref_q, L, dpl, dmi, slow);
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) {
                UKFB_UNROLL
                for (int k = 0; k <= i; ++k) C[tri(i, k)] = fma(dpl[i], dpl[k], fma(dmi[i], dmi[k], C[tri(i, k)]));
            }
            UKFB_UNROLL
            for (int k = 0; k < 6; ++k) UKFB_PS(PF_LA + j * 12 + k) = dpl[k] - dmi[k];
        }
        double Xb[18]; /* columns 6..11: sum_j L[6:12, j] (do+_j - do-_j)^T, the orientation part of their differences */
        UKFB_UNROLL
        for (int i = 0; i < 18; ++i) Xb[i] = 0.0;
        double c[4];
        quat_mul_conj(m.q, ref_q, c);
        UKFB_LOOP_UNROLL(UKFB_PF_COV_UNROLL)
        for (int j = 6; j < 12; ++j) {
            double L[6], dpl[6], dmi[6], wv[3];
            pf_pair_b_sel<WIDE>(sm, lane, j, Rm, rw0, c, dt, L, dpl + 3, dmi + 3, wv, slow);
            UKFB_UNROLL
            for (int i = 0; i < 3; ++i) dpl[i] = d0[i] + wv[i], dmi[i] = d0[i] - wv[i];
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) {
                UKFB_UNROLL
                for (int k = 0; k <= i; ++k) C[tri(i, k)] = fma(dpl[i], dpl[k], fma(dmi[i], dmi[k], C[tri(i, k)]));
            }
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) {
                const double dd = dpl[3 + k] - dmi[3 + k];
                UKFB_UNROLL
                for (int i = 0; i < 6; ++i) Xb[i * 3 + k] = fma(L[i], dd, Xb[i * 3 + k]);
            }
        }
        if (slow) return false;
        /* X, after the loops */
        UKFB_UNROLL
        for (int i = 0; i < 6; ++i) {
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) X[i * 6 + k] = 0.0, X[i * 6 + 3 + k] = Xb[i * 3 + k];
        }
        UKFB_UNROLL
        for (int j = 0; j < 6; ++j) { /* columns 0..5: the stored differences against rows 6..11 of the column */
            double dd[6];
            UKFB_UNROLL
            for (int k = 0; k < 6; ++k) dd[k] = UKFB_PS(PF_LA + j * 12 + k);
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) {
                const double l = UKFB_PS(PF_LA + j * 12 + 6 + i);
                UKFB_UNROLL
                for (int k = 0; k < 6; ++k) X[i * 6 + k] = fma(l, dd[k], X[i * 6 + k]);
            }
        }
        UKFB_UNROLL
        for (int j = 6; j < 9; ++j) { /* columns 6..8: position differences 2 (R L[6:9, j]) dt; columns 9..11 have L[6:9, j] = 0 */
            double Lc[6], rv[3];
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) Lc[i] = i >= j - 6 ? UKFB_PS(PF_LB + (j - 6) * 6 + i) : 0.0;
            pf_matvec(Rm, Lc, rv);
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) {
                const double dd = 2.0 * (dt * rv[k]);
                UKFB_UNROLL
                for (int i = j - 6; i < 6; ++i) X[i * 6 + k] = fma(Lc[i], dd, X[i * 6 + k]);
            }
        }
    }

    /* ---- new covariance = 1/2 C + process noise (PoseUKF.cpp:182-191), committed to the record.  qv(i, k) = Q[i][k], i >= k:
     * a load, or for a broadcast diagonal Q (the reference's default and the usual configuration) a load on the diagonal
     * and a literal zero elsewhere, which leaves 12 loads of the 78 */
    auto commit = [&](auto qv, auto iso) {
        const double scale = ma.has_acc ? 1.0 : dt;
        /* the entries of the noise that are not plain scale * Q: the two rotated blocks, or 2 acc.cov */
        double nb[12], na[6];
        UKFB_UNROLL
        for (int i = 0; i < 12; ++i) nb[i] = 0.0;
        UKFB_UNROLL
        for (int i = 0; i < 6; ++i) na[i] = 0.0;
        if (!ma.has_acc && decltype(iso)::value) {
            /* R (q I) R^T = q I for both rotated blocks (the reference's default Q: the rotation changes nothing) */
            nb[tri(0, 0)] = nb[tri(1, 1)] = nb[tri(2, 2)] = scale * qv(0, 0);
            nb[6 + tri(0, 0)] = nb[6 + tri(1, 1)] = nb[6 + tri(2, 2)] = scale * qv(3, 3);
        } else if (!ma.has_acc) {
            UKFB_UNROLL
            for (int blk = 0; blk < 2; ++blk) {
                const int off = blk * 3;
                double t[9];
                UKFB_UNROLL
                for (int r = 0; r < 3; ++r) {
                    UKFB_UNROLL
                    for (int k = 0; k < 3; ++k) {
                        double s = 0.0;
                        UKFB_UNROLL
                        for (int l = 0; l < 3; ++l) s += Rm[r * 3 + l] * (l >= k ? qv(off + l, off + k) : qv(off + k, off + l));
                        t[r * 3 + k] = s;
                    }
                }
                UKFB_UNROLL
                for (int r = 0; r < 3; ++r) {
                    UKFB_UNROLL
                    for (int cc = 0; cc <= r; ++cc) {
                        double s = 0.0;
                        UKFB_UNROLL
                        for (int k = 0; k < 3; ++k) s += t[r * 3 + k] * Rm[cc * 3 + k];
                        nb[blk * 6 + tri(r, cc)] = scale * s;
                    }
                }
            }
        } else {
            UKFB_UNROLL
            for (int r = 0; r < 3; ++r) {
                UKFB_UNROLL
                for (int cc = 0; cc <= r; ++cc) na[tri(r, cc)] = 2.0 * acov[r * 3 + cc];
            }
        }
        UKFB_UNROLL
        for (int i = 0; i < 12; ++i) {
            UKFB_UNROLL
            for (int k = 0; k <= i; ++k) {
                const int e = tri(i, k);
                double nz = scale * qv(i, k);
                if (i < 6 && i / 3 == k / 3) nz = ma.has_acc ? nz : nb[(i / 3) * 6 + tri(i % 3, k % 3)];
                if (i >= 6 && i < 9 && k >= 6) nz = ma.has_acc ? na[tri(i - 6, k - 6)] : nz;
                double s;
                if (i < 6)
                    s = fma(0.5, C[e], nz);
                else if (k < 6)
                    s = fma(0.5, X[(i - 6) * 6 + k], nz);
                else
                    s = sig[e * TILE] + nz;
                sig[e * TILE] = s;
                if (to_smem) UKFB_PS(e) = s;
            }
        }
    };
    if (q_diagonal == 2) /* diagonal, and a multiple of the identity in each of the two rotated 3 x 3 blocks */
        commit([&](int i, int k) { return i == k ? UKFB_LDG(Qp + tri(i, i)) : 0.0; }, TrueT());
    else if (q_diagonal)
        commit([&](int i, int k) { return i == k ? UKFB_LDG(Qp + tri(i, i)) : 0.0; }, FalseT());
    else
        commit([&](int i, int k) { return UKFB_LDG(Qp + tri(i, k)); }, FalseT());
    m.p[0] = ref_p[0], m.p[1] = ref_p[1], m.p[2] = ref_p[2];
    m.q[0] = ref_q[0], m.q[1] = ref_q[1], m.q[2] = ref_q[2], m.q[3] = ref_q[3];
    m.v[0] = vm[0], m.v[1] = vm[1], m.v[2] = vm[2];
    passes_out = passes;
    return true;
}

/* Everything the hot predict does not do, out of line (one call site in the kernel): a filter whose sigma points leave the
 * range of the short polynomials (known from its orientation variance, or found out by the hot instance, which then has
 * modified nothing) is served by the any-angle instance of the same structured code; one beyond that too (a quaternion
 * off the unit sphere, a rotation above 6.4 rad) by the literal code of ukf_thread.cuh. */
UKFB_DNI PfLit pf_predict_slow(double* sm, int lane, double* sig, const double* Qp, const double* acov, ModelArgs ma, PoseMu m,
                               int to_smem, int q_diagonal, int* in_smem)
{
    {
        double a[PoseF::LP];
        UKFB_UNROLL
        for (int e = 0; e < PoseF::LP; ++e) a[e] = sig[e * TILE];
        PfLit r;
        r.m = m, r.status = 0, r.passes = 0;
        bool spd = true;
        if (pf_predict<true>(q_diagonal, sm, lane, sig, a, Qp, acov, ma, r.m, to_smem != 0, r.status, r.passes, spd)) {
            UKFB_PF_COUNT(3);
            if (!spd) r.status |= UKFB_STATUS_NOT_SPD;
            *in_smem = spd ? to_smem : 0;
            return r;
        }
    }
    *in_smem = 0;
    return pf_literal_predict(sig, Qp, acov, ma, m);
}

/* tangent index of measurement component c of a selector kind (PoseUKF.cpp:7-69), -1 = unused component */
UKFB_D int pf_sel(int kind, int c)
{
    switch (kind) {
        case 0: return c;
        case 1: return c < 2 ? c : -1;
        case 2: return c == 0 ? 2 : -1;
        case 4: return 6 + c;
        case 5: return c < 2 ? 6 + c : -1;
        case 6: return c == 0 ? 8 : -1;
        case 7: return c == 0 ? 6 : (c == 1 ? 11 : -1);
        default: return 9 + c; /* 8 */
    }
}

UKFB_D double pf_mu_tangent(const PoseMu& m, int t)
{
    /* Euclidean tangent component t (0..2 position, 6..8 velocity, 9..11 angular velocity) of the mean */
    double r = m.p[0];
    r = t == 1 ? m.p[1] : r;
    r = t == 2 ? m.p[2] : r;
    r = t == 6 ? m.v[0] : r;
    r = t == 7 ? m.v[1] : r;
    r = t == 8 ? m.v[2] : r;
    r = t == 9 ? m.w[0] : r;
    r = t == 10 ? m.w[1] : r;
    r = t == 11 ? m.w[2] : r;
    return r;
}

/* ---- apply_delta of the structured update: orientation rows only.  Slots LA hold the first six factor columns of the
 * downdated covariance, which is in the record.  WIDE = false: the hot instance (short polynomials; gives up with
 * stage = 2 when a sigma point leaves their range before anything was overwritten, stage = 1 otherwise);
 * WIDE = true: the out-of-line instance for any orientation uncertainty (gives up with stage = 1: literal code). */
template <bool WIDE>
UKFB_D bool pf_apply_delta(double* sm, int lane, double* sig, PoseMu& m, const double* delta, uint32_t& status, int& passes_out, int& stage)
{
    bool slow = !pf_unit(m.q);
    const bool off_sphere = slow;
    bool lost_factor = false;
    double e0[4], q0n[4];
    pf_exp1<WIDE>(delta + 3, 1.0, e0, slow);
    quat_mul(e0, m.q, q0n);
    double ref_q[4] = {q0n[0], q0n[1], q0n[2], q0n[3]};
    int it = 0, passes = 0;
    while (true) {
        double c[4], r[4], d0[3], md[3];
        quat_mul_conj(m.q, ref_q, c);
        quat_mul(e0, c, r);
        pf_log1<WIDE>(r, d0, slow);
        md[0] = 13.0 * d0[0], md[1] = 13.0 * d0[1], md[2] = 13.0 * d0[2];
        UKFB_NOUNROLL
        for (int j = 0; j < 6; ++j) {
            double dp[3], dn[3];
            pf_pair_d_sel<WIDE>(sm, lane, j, delta, c, dp, dn, slow);
            md[0] += dp[0] + dn[0], md[1] += dp[1] + dn[1], md[2] += dp[2] + dn[2];
        }
        double n2 = 0.0;
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) {
            md[i] = div_ns<PoseF::NS>(md[i]);
            n2 += md[i] * md[i];
        }
        {
            double e[4], rr[4];
            pf_exp1<WIDE>(md, 1.0, e, slow);
            quat_mul(e, ref_q, rr);
            ref_q[0] = rr[0], ref_q[1] = rr[1], ref_q[2] = rr[2], ref_q[3] = rr[3];
        }
        ++passes;
        if (slow || !(n2 > UKFB_MEAN_TOL * UKFB_MEAN_TOL)) break;
        if (++it >= UKFB_MEAN_MAX_IT) {
            status |= UKFB_STATUS_MEAN_NO_CONVERGE;
            break;
        }
    }
    /* covariance of the orientation rows: Coo (6) and the cross block with the 9 Euclidean components */
    double Coo[6], Xc[27];
    if (!slow) {
        double c[4], r[4], d0[3];
        quat_mul_conj(m.q, ref_q, c);
        quat_mul(e0, c, r);
        pf_log1<WIDE>(r, d0, slow);
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) {
            UKFB_UNROLL
            for (int k = 0; k <= i; ++k) Coo[tri(i, k)] = 13.0 * d0[i] * d0[k];
        }
        /* the cross block is formed after the loop (as X in the predict): the differences dp - dn of column j go to the slots
         * L[3:6, j] the column's two points were just generated from */
        bool late = false;
        UKFB_NOUNROLL
        for (int j = 0; j < 6; ++j) {
            double dp[3], dn[3];
            pf_pair_d_sel<WIDE>(sm, lane, j, delta, c, dp, dn, late);
            UKFB_UNROLL
            for (int i = 0; i < 3; ++i) {
                UKFB_UNROLL
                for (int k = 0; k <= i; ++k) Coo[tri(i, k)] = fma(dp[i], dp[k], fma(dn[i], dn[k], Coo[tri(i, k)]));
            }
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) UKFB_PS(PF_LA + j * 12 + 3 + k) = dp[k] - dn[k];
        }
        slow = slow || late;
        lost_factor = late; /* the column slots hold differences now */
        UKFB_UNROLL
        for (int i = 0; i < 27; ++i) Xc[i] = 0.0;
        UKFB_UNROLL
        for (int j = 0; j < 6; ++j) {
            double dd[3];
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) dd[k] = UKFB_PS(PF_LA + j * 12 + 3 + k);
            UKFB_UNROLL
            for (int t = 0; t < 9; ++t) {
                const int row = t < 3 ? t : t + 3;
                if (row < j) continue; /* L is lower triangular */
                const double l = UKFB_PS(PF_LA + j * 12 + row);
                UKFB_UNROLL
                for (int k = 0; k < 3; ++k) Xc[t * 3 + k] = fma(l, dd[k], Xc[t * 3 + k]);
            }
        }
    }
    if (slow) { /* Sigma - K S K^T is in the record, mu untouched: the caller hands mu and delta on */
        stage = (WIDE || lost_factor || off_sphere) ? 1 : 2; /* 2: the any-angle instance can take over (factor columns intact) */
        return false;
    }
    /* commit */
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        UKFB_UNROLL
        for (int k = 0; k <= i; ++k) sig[tri(3 + i, 3 + k) * TILE] = 0.5 * Coo[tri(i, k)];
        UKFB_UNROLL
        for (int t = 0; t < 3; ++t) sig[tri(3 + i, t) * TILE] = 0.5 * Xc[t * 3 + i];         /* orientation x position */
        UKFB_UNROLL
        for (int t = 3; t < 9; ++t) sig[tri(t + 3, 3 + i) * TILE] = 0.5 * Xc[t * 3 + i];     /* (vel, angvel) x orientation */
    }
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        m.p[i] += delta[i];
        m.v[i] += delta[6 + i];
        m.w[i] += delta[9 + i];
    }
    m.q[0] = ref_q[0], m.q[1] = ref_q[1], m.q[2] = ref_q[2], m.q[3] = ref_q[3];
    passes_out = passes;
    return true;
}

/* ---- structured update with a selector measurement.  Slots 0..77 hold the prior covariance (packed lower), which
 * is also in the record.  Returns false when apply_delta left the polynomial range: the record then holds
 * Sigma - K S K^T, `delta` = K innov, m is untouched. */
template <bool ORI_MEAS>
UKFB_D bool pf_update(double* sm, int lane, double* sig, int kind, const double* zm, const double* Rmeas, int r_ld, PoseMu& m,
                      double* delta, uint32_t& status, int& passes_out, bool& spd, double gate_d2, int& stage)
{
    stage = 0; /* nothing modified yet */
    const int m_dim = meas_dim(kind);
    int sel[3];
    UKFB_UNROLL
    for (int c = 0; c < 3; ++c) sel[c] = pf_sel(kind, c);

    double Sxz[36], S[9], innov[3];
    if (ORI_MEAS) {
        /* h(X) = orientation (PoseUKF.cpp:28-33), a manifold-valued measurement: Z_p = exp(+-L_ori[:,j]) q for the 12
         * points of columns 0..5, q for X0 and the other 12.  zbar by the iterative SO(3) mean, S = 1/2 sum dz dz^T + R,
         * Sxz = 1/2 sum_{j<6} L[:,j] (dz+_j - dz-_j)^T (deviations from the prior mu are +-L[:,j] exactly under the
         * trace guard), innovation = exp(z) [-] zbar (PoseUKF.cpp:135). */
        bool slow = !pf_unit(m.q);
        double a[PoseF::LP];
        UKFB_UNROLL
        for (int e = 0; e < PoseF::LP; ++e) a[e] = UKFB_PS(e);
        spd = pf_cholesky<12>(a); /* all of it, as the reference does before an update: NOT_SPD is reported here */
        if (!spd) return true;
        constexpr int PF_E = 78; /* exp(L_ori[:,j]), then dz+_j - dz-_j, in the slots above the packed covariance */
        UKFB_UNROLL
        for (int j = 0; j < 6; ++j) {
            double e[4];
            const double Lo[3] = {j <= 3 ? a[tri(3, j <= 3 ? j : 3)] : 0.0, j <= 4 ? a[tri(4, j <= 4 ? j : 4)] : 0.0, a[tri(5, j)]};
            pf_exp(Lo, 1.0, e, slow);
            UKFB_UNROLL
            for (int i = 0; i < 4; ++i) UKFB_PS(PF_E + 4 * j + i) = e[i];
        }
        double zin[4];
        {
            const double v[3] = {zm[0], zm[1], zm[2]};
            so3_exp(v, 1.0, zin); /* the measured rotation vector may be any angle: the general exp (PoseUKF.cpp:135) */
        }
        double zref[4] = {m.q[0], m.q[1], m.q[2], m.q[3]};
        int it = 0;
        while (!slow) {
            double c[4], d0[3], md[3];
            quat_mul_conj(m.q, zref, c);
            pf_log(c, d0, slow);
            md[0] = 13.0 * d0[0], md[1] = 13.0 * d0[1], md[2] = 13.0 * d0[2];
            UKFB_NOUNROLL
            for (int j = 0; j < 6; ++j) {
                D2 e[4], r2[4], d[3]; /* exp(L_ori[:, j]) and its conjugate, side by side */
                UKFB_UNROLL
                for (int i = 0; i < 4; ++i) {
                    const double x = UKFB_PS(PF_E + 4 * j + i);
                    e[i] = D2(x, i < 3 ? -x : x);
                }
                quat_mul(e, c, r2);
                pf_log(r2, d, slow);
                md[0] += d[0].a + d[0].b, md[1] += d[1].a + d[1].b, md[2] += d[2].a + d[2].b;
            }
            UKFB_UNROLL
            for (int i = 0; i < 3; ++i) md[i] = div_ns<PoseF::NS>(md[i]);
            const double n2 = md[0] * md[0] + md[1] * md[1] + md[2] * md[2];
            {
                double e[4], rr[4];
                pf_exp(md, 1.0, e, slow);
                quat_mul(e, zref, rr);
                zref[0] = rr[0], zref[1] = rr[1], zref[2] = rr[2], zref[3] = rr[3];
            }
            if (!(n2 > UKFB_MEAN_TOL * UKFB_MEAN_TOL)) break;
            if (++it >= UKFB_MEAN_MAX_IT) {
                status |= UKFB_STATUS_MEAN_NO_CONVERGE;
                break;
            }
        }
        {
            double c[4], d0[3];
            quat_mul_conj(m.q, zref, c);
            pf_log(c, d0, slow);
            UKFB_UNROLL
            for (int r = 0; r < 3; ++r) {
                UKFB_UNROLL
                for (int cc = 0; cc < 3; ++cc) S[r * 3 + cc] = 13.0 * d0[r] * d0[cc];
            }
            UKFB_NOUNROLL
            for (int j = 0; j < 6; ++j) {
                D2 e[4], r2[4], d[3];
                UKFB_UNROLL
                for (int i = 0; i < 4; ++i) {
                    const double x = UKFB_PS(PF_E + 4 * j + i);
                    e[i] = D2(x, i < 3 ? -x : x);
                }
                quat_mul(e, c, r2);
                pf_log(r2, d, slow);
                UKFB_UNROLL
                for (int r = 0; r < 3; ++r) {
                    UKFB_UNROLL
                    for (int cc = 0; cc < 3; ++cc) S[r * 3 + cc] = fma(d[r].a, d[cc].a, fma(d[r].b, d[cc].b, S[r * 3 + cc]));
                    UKFB_PS(PF_E + 4 * j + r) = d[r].a - d[r].b;
                }
            }
            double r4[4];
            quat_mul_conj(zin, zref, r4);
            pf_log(r4, innov, slow);
        }
        if (slow) return false; /* stage 0: the caller runs the whole literal update */
        UKFB_UNROLL
        for (int r = 0; r < 3; ++r) {
            UKFB_UNROLL
            for (int cc = 0; cc < 3; ++cc) S[r * 3 + cc] = fma(0.5, S[r * 3 + cc], Rmeas[r * r_ld + cc]);
        }
        UKFB_UNROLL
        for (int i = 0; i < 36; ++i) Sxz[i] = 0.0;
        UKFB_UNROLL
        for (int j = 0; j < 6; ++j) {
            UKFB_UNROLL
            for (int cc = 0; cc < 3; ++cc) {
                const double ddz = UKFB_PS(PF_E + 4 * j + cc);
                UKFB_UNROLL
                for (int i = j; i < 12; ++i) Sxz[i * 3 + cc] = fma(a[tri(i, j)], ddz, Sxz[i * 3 + cc]);
            }
        }
        UKFB_UNROLL
        for (int i = 0; i < 36; ++i) Sxz[i] *= 0.5;
    } else {
    /* S = Sigma[sel,sel] + R, Sxz = Sigma[:,sel] (identity / zero padded to 3 like the reference's 3-vectors) */
    UKFB_UNROLL
    for (int c = 0; c < 3; ++c) {
        const int s = sel[c] < 0 ? 0 : sel[c];
        const bool used = c < m_dim;
        UKFB_UNROLL
        for (int i = 0; i < 12; ++i) {
            const int e = i >= s ? tri(i, 0) + s : tri(s, 0) + i;
            const double x = UKFB_PS(e);
            Sxz[i * 3 + c] = used ? x : 0.0;
        }
        innov[c] = used ? zm[c] - pf_mu_tangent(m, s) : 0.0;
    }
    UKFB_UNROLL
    for (int a = 0; a < 3; ++a) {
        UKFB_UNROLL
        for (int c = 0; c < 3; ++c) {
            const int sa = sel[a] < 0 ? 0 : sel[a], sc = sel[c] < 0 ? 0 : sel[c];
            const int e = sa >= sc ? tri(sa, 0) + sc : tri(sc, 0) + sa;
            const bool used = a < m_dim && c < m_dim;
            const double x = UKFB_PS(e);
            S[a * 3 + c] = used ? x + Rmeas[a * r_ld + c] : (a == c ? 1.0 : 0.0);
        }
    }
    } /* selector kinds */
    /* S^-1 by cofactors (Eigen fixed-size inverse) */
    double Si[9];
    {
        const double c00 = S[4] * S[8] - S[5] * S[7];
        const double c10 = S[7] * S[2] - S[8] * S[1];
        const double c20 = S[1] * S[5] - S[2] * S[4];
        const double det = c00 * S[0] + c10 * S[3] + c20 * S[6];
        const double invdet = 1.0 / det;
        Si[0] = c00 * invdet;
        Si[1] = c10 * invdet;
        Si[2] = c20 * invdet;
        Si[3] = (S[5] * S[6] - S[3] * S[8]) * invdet;
        Si[4] = (S[8] * S[0] - S[6] * S[2]) * invdet;
        Si[5] = (S[2] * S[3] - S[0] * S[5]) * invdet;
        Si[6] = (S[3] * S[7] - S[4] * S[6]) * invdet;
        Si[7] = (S[6] * S[1] - S[7] * S[0]) * invdet;
        Si[8] = (S[0] * S[4] - S[1] * S[3]) * invdet;
    }
    /* the accept functor: squared Mahalanobis distance against the gate (never taken with the reference's accept_any) */
    {
        double d2 = 0.0;
        UKFB_UNROLL
        for (int a = 0; a < 3; ++a) d2 += innov[a] * (Si[a * 3] * innov[0] + Si[a * 3 + 1] * innov[1] + Si[a * 3 + 2] * innov[2]);
        if (d2 > gate_d2) {
            status |= UKFB_STATUS_MEAS_REJECTED;
            return true; /* nothing was modified */
        }
    }
    /* Row by row: K[i,:] = Sxz[i,:] S^-1 (in place of Sxz), delta_i = K[i,:] innov, and row i of
     * Sigma <- Sigma - (K S) K^T -- to the record, and kept in registers for the factorisation.  (K S)[i,:] = Sxz[i,:]
     * (the reference multiplies K by S again; the two differ by rounding only). */
    double a[PoseF::LP];
    UKFB_UNROLL
    for (int i = 0; i < 12; ++i) {
        double k3[3], ks3[3];
        UKFB_UNROLL
        for (int cc = 0; cc < 3; ++cc) {
            double s = 0.0;
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) s += Sxz[i * 3 + k] * Si[k * 3 + cc];
            k3[cc] = s;
        }
        double dl = 0.0;
        UKFB_UNROLL
        for (int c = 0; c < 3; ++c) {
            ks3[c] = Sxz[i * 3 + c]; /* (K S)[i,:] = (Sxz S^-1 S)[i,:] = Sxz[i,:] */
            Sxz[i * 3 + c] = k3[c];
            dl += k3[c] * innov[c];
        }
        delta[i] = dl;
        UKFB_UNROLL
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) s += ks3[k] * Sxz[j * 3 + k];
            const double x = UKFB_PS(tri(i, j)) - s;
            a[tri(i, j)] = x;
            sig[tri(i, j) * TILE] = x;
        }
    }
    stage = 1; /* the record holds Sigma - K S K^T */
    const bool wide_trace = a[tri(3, 3)] + a[tri(4, 4)] + a[tri(5, 5)] > PF_WIDE_TRACE; /* as in the predict */
    /* first six columns of the factor of the updated covariance (the reference factorises all of it: a failure in the
     * later columns shows at the next factorisation of this filter instead) */
    spd = pf_cholesky<6>(a);
    if (!spd) return true;
    UKFB_UNROLL
    for (int j = 0; j < 6; ++j) {
        UKFB_UNROLL
        for (int i = 0; i < 12; ++i) UKFB_PS(PF_LA + j * 12 + i) = i >= j ? a[tri(i, j)] : 0.0;
    }

    /* ---- apply_delta: orientation rows only */
    if (wide_trace) {
        stage = 2; /* a wide filter: straight to the out-of-line instance; the factor columns are in place */
        return false;
    }
    return pf_apply_delta<false>(sm, lane, sig, m, delta, status, passes_out, stage);
}

/* Everything the selector kinds' structured update does not do, out of line (one call site, where few registers are
 * live): the update with the orientation measurement -- its own structured instance pf_update<true>, tried first -- and
 * the literal code of ukf_thread.cuh for lanes that failed a guard or left a polynomial range.
 * first = true: nothing has been modified yet; false: the record holds Sigma - K S K^T and `delta` = K innov. */
template <bool WITH_ORI>
UKFB_DNI PfLit pf_update_slow(double* sm, int lane, double* sig, int kind, const double* zm, const double* Rm, int r_ld, ModelArgs ma,
                              PoseMu m, PfDelta delta, int stage, double gate_d2)
{
    if (WITH_ORI && stage == 0 && kind == UKFB_MEAS_POSE_ORIENTATION) {
        /* the prior covariance into slots 0..77, as the structured update expects it; under the trace guard
         * (mu [+] L_j) [-] mu = L_j holds and the structured instance applies */
        UKFB_UNROLL
        for (int e = 0; e < PoseF::LP; ++e) UKFB_PS(e) = sig[e * TILE];
        const double tr = UKFB_PS(tri(3, 3)) + UKFB_PS(tri(4, 4)) + UKFB_PS(tri(5, 5));
        if (tr < PF_PI2_GUARD) {
            PfLit r;
            r.m = m, r.status = 0, r.passes = 0;
            bool spd = true;
            const bool done = pf_update<true>(sm, lane, sig, kind, zm, Rm, r_ld, r.m, delta.d, r.status, r.passes, spd, gate_d2, stage);
            if (done) {
                if (!spd) r.status |= UKFB_STATUS_NOT_SPD;
                return r;
            }
            /* stage now tells where the structured code stopped */
        }
    }
    if (stage == 0 && kind != UKFB_MEAS_POSE_ORIENTATION) {
        /* sent here by the caller's trace guard, which bounds every |L_ori[:, j]| by sqrt(trace): sufficient, not necessary.
         * What the structured update rests on, (mu [+] L_j) [-] mu = L_j, needs each column below pi, so look at the
         * columns themselves (a filter that does not know its attitude at all: 1.8 rad and more per axis).  The prior
         * covariance is in slots 0..77 and has been shown to be positive definite. */
        double a[PoseF::LP];
        UKFB_UNROLL
        for (int e = 0; e < PoseF::LP; ++e) a[e] = UKFB_PS(e);
        bool ok = pf_cholesky<6>(a);
        UKFB_UNROLL
        for (int j = 0; j < 6; ++j) {
            double n2 = a[tri(5, j)] * a[tri(5, j)];
            if (j <= 4) n2 += a[tri(4, j)] * a[tri(4, j)];
            if (j <= 3) n2 += a[tri(3, j)] * a[tri(3, j)];
            ok = ok && n2 < PF_PI2_COLUMN;
        }
        if (ok) {
            PfLit r;
            r.m = m, r.status = 0, r.passes = 0;
            bool spd = true;
            if (pf_update<false>(sm, lane, sig, kind, zm, Rm, r_ld, r.m, delta.d, r.status, r.passes, spd, gate_d2, stage)) {
                if (!spd) r.status |= UKFB_STATUS_NOT_SPD;
                return r;
            }
            /* stage 2: the downdate is done, apply_delta is next */
        }
    }
    if (stage == 2) { /* Sigma - K S K^T in the record, its factor columns in place, delta = K innov: apply_delta with the any-angle pair */
        PfLit r;
        r.m = m, r.status = 0, r.passes = 0;
        if (pf_apply_delta<true>(sm, lane, sig, r.m, delta.d, r.status, r.passes, stage)) {
            UKFB_PF_COUNT(4);
            return r;
        }
        stage = 1; /* beyond that too: the literal apply_delta (nothing was committed) */
    }
    return pf_literal_update(sig, kind, zm, Rm, r_ld, ma, m, delta, stage == 0, gate_d2);
}

/* ---- the kernel: one warp per block, one filter per lane -------------------------------------------------------- */
/* WITH_ORI = false: the instance the host launches when no orientation measurement can occur in the launch (a uniform
 * kind other than 3): an orientation measurement would run the literal code.  WITH_ORI = true: its structured instance
 * is compiled into the slow-path call; the larger callee costs the hot path about 3 % (register allocation around the
 * call), which is why there are two instances. */
/* OVERLAP = true: the instance for handles whose launches overlap at their ends (StepParams::tile_done).  A separate
 * instance because the handshake code costs the other one 1.8 % even when it is skipped at run time. */
template <bool WITH_ORI, bool OVERLAP = false>
UKFB_GLOBAL void UKFB_LAUNCH_BOUNDS(UKFB_PF_MAX_THREADS, UKFB_PF_MIN_BLOCKS) ukf_pose_fast_kernel(const UKFB_GRID_CONSTANT StepParams p)
{
    typedef PoseF F;
    UKFB_SMEM_DECL
    /* a block is 1..4 independent warps (no barrier between them: warps of one block merely start together, which keeps
     * their instruction fetches close); each warp owns one tile of 32 filters and its own slice of shared memory */
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    double* sm = ukfb_smem + wib * (PF_PER_LANE * TILE);
    /* overlapped launches (StepParams::tile_done): the next launch may take the slots this grid's last wave leaves empty,
     * and every warp waits for its own tile of the previous launch */
    if (OVERLAP) pdl_launch_dependents();
    if (tile * TILE >= p.B) return; /* a warp past the last tile (no barriers in this kernel) */
    if (OVERLAP) tile_done_wait(p.tile_done + tile, 32ull * (p.launch_seq - 1));
    const long long b = tile * TILE + lane;
    const bool valid = b < p.B;
    const long long bb = valid ? b : p.B - 1; /* lanes past the end shadow the last filter and never store */
    double* rec = p.state + tile * (TILE * F::REC) + lane; /* entry e at rec[e * TILE] */
    double* sig = rec + F::MU * TILE;

    PoseMu m;
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) m.p[i] = rec[i * TILE], m.v[i] = rec[(7 + i) * TILE], m.w[i] = rec[(10 + i) * TILE];
    UKFB_UNROLL
    for (int i = 0; i < 4; ++i) m.q[i] = rec[(3 + i) * TILE];
    prefetch_next_wave<F>(p, tile, lane);

    ModelArgs ma;
    ma.dt = 0.0;
    ma.has_acc = false;
    ma.neg_inv_tau_g = 0.0;
    ma.neg_inv_tau_a = 0.0;
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        ma.earth[i] = 0.0;
        ma.acc[i] = p.acc_mu[bb * 3 + i];
        ma.omega[i] = 0.0;
    }
    const double big = 1.79769313486231570e308;
    uint32_t status = 0;
    bool dirty_mu = false;
    int hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    UKFB_NOUNROLL
    for (int tick = 0; tick < p.K; ++tick) {
        /* the covariance is needed first by whichever phase runs: issue its loads before the control code so that
         * their latency overlaps it (and the loads of mu above) */
        double a[F::LP];
        UKFB_UNROLL
        for (int e = 0; e < F::LP; ++e) a[e] = sig[e * TILE];

        /* ---- control: time guards (UnscentedKalmanFilter.hpp:83-125), masks */
        bool do_pred = false, do_upd = false;
        int kind = -1, store = -1;
        const double* Rmeas = p.R + tick * p.r_kstride + b * p.r_stride;
        if (valid) {
            bool idle = false;
            if (p.events) { /* one queued sample per filter and slot; UKFB_EVENT_IDLE: nothing happens */
                kind = int(p.kinds[tick * p.kinds_kstride + b]);
                idle = kind == UKFB_EVENT_IDLE;
                if (kind < UKFB_EVENT_IDLE || kind == UKFB_MEAS_ORI_VELOCITY || kind > UKFB_EVENT_POSE_ACCELERATION) {
                    status |= UKFB_STATUS_BAD_EVENT;
                    idle = true;
                }
                if (idle) kind = -1;
                if (kind >= 0) Rmeas += kind * p.r_kind_stride;
                if (kind == UKFB_EVENT_POSE_ACCELERATION) store = kind, kind = -1;
            }
            if (p.do_predict && !idle) {
                double dt;
                bool have_dt = true;
                if (p.time_mode) {
                    const long long ts = p.ts[tick * p.ts_kstride + b * p.ts_stride];
                    const long long tl = p.t_last[b];
                    if (tl == 0) { /* first call: latch only (:86-90) */
                        p.t_last[b] = ts;
                        have_dt = false;
                        dt = 0.0;
                    } else {
                        dt = double(ts - tl) / UKFB_US_PER_S;
                        if (dt > p.min_dt) p.t_last[b] = ts; /* :96-97 */
                    }
                } else {
                    dt = p.dt[tick * p.dt_kstride + b * p.dt_stride];
                }
                if (have_dt) {
                    if (dt < 0.0) {
                        status |= UKFB_STATUS_NEG_DT;
                        if (p.events) kind = -1, store = -1; /* the reference's callback leaves here (the throw): this sample is neither integrated nor stored */
                    } else if (dt <= p.min_dt) {
                        /* delta time is zero or close to zero: no-op */
                    } else if (dt > p.max_dt) {
                        status |= UKFB_STATUS_DT_TOO_LARGE;
                        if (p.events) kind = -1, store = -1;
                    }
                    else {
                        do_pred = true;
                        ma.dt = dt;
                    }
                }
            }
            if (p.do_update && !idle) {
                if (!p.events) {
                    kind = p.tick_kinds ? int(p.tick_kinds[tick]) : (p.kind == -2 ? int(p.kinds[tick * p.kinds_kstride + b]) : p.kind);
                    if (p.kind == -2 && !p.tick_kinds && kind != UKFB_MEAS_NONE && !meas_kind_of_class<F>(kind)) {
                        status |= UKFB_STATUS_BAD_EVENT; /* per-filter kinds on the device: a kind of the other filter class is ignored */
                        kind = -1;
                    }
                    if (p.mask && !p.mask[tick * p.mask_kstride + b]) kind = -1;
                }
                do_upd = kind >= 0; /* PoseUKF never finite-checks a measurement (PoseUKF.cpp:112-173) */
            }
        }

        int passes_a = 0, passes_b = 0;
        bool sigma_in_smem = false; /* slots 0..77 hold this lane's current covariance */
        const double* Qp = p.Q + b * p.q_stride;
        const double* acov = p.acc_cov ? p.acc_cov + b * 9 : nullptr;

        /* ---- predict (ukfom predict, App. A.3) ------------------------------------------------------------- */
        if (do_pred) {
            ma.has_acc = (fabs(ma.acc[0]) <= big) && (fabs(ma.acc[1]) <= big) && (fabs(ma.acc[2]) <= big);
            bool spd = true;
            const bool want_smem = do_upd && kind != UKFB_MEAS_POSE_ORIENTATION;
            if (pf_predict<false>(p.q_diagonal, sm, lane, sig, a, Qp, acov, ma, m, want_smem, status, passes_a, spd)) {
                if (!spd) {
                    status |= UKFB_STATUS_NOT_SPD;
                    do_upd = false; /* every later factorisation of this covariance fails too */
                } else {
                    sigma_in_smem = want_smem;
                    dirty_mu = true;
                }
            } else { /* a wide filter, or a sigma point left the range of the short polynomials: nothing was modified */
                int in_smem = 0;
                const PfLit r = pf_predict_slow(sm, lane, sig, Qp, acov, ma, m, want_smem ? 1 : 0, p.q_diagonal, &in_smem);
                status |= r.status;
                passes_a = r.passes;
                if (r.status & UKFB_STATUS_NOT_SPD)
                    do_upd = false;
                else {
                    m = r.m;
                    sigma_in_smem = in_smem != 0;
                    dirty_mu = true;
                }
            }
        }

        /* ---- AccelerationMeasurement event: kept for the next predict, unchecked (PoseUKF.cpp:175-178) ---------------- */
        if (store >= 0) {
            const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
            UKFB_UNROLL
            for (int r = 0; r < 3; ++r) {
                ma.acc[r] = zm[r];
                UKFB_UNROLL
                for (int cc = 0; cc < 3; ++cc) p.acc_cov[b * 9 + r * 3 + cc] = Rmeas[r * p.r_ld + cc];
            }
        }

        /* ---- update (ukfom update + apply_delta, App. A.4) --------------------------------------------------- */
        if (do_upd) {
            const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
            bool literal = kind == UKFB_MEAS_POSE_ORIENTATION; /* handled out of line, below */
            if (!literal) {
                if (!sigma_in_smem) {
                    /* no fast predict ran on this covariance in this tick: it has not been shown to be SPD yet, and the
                     * reference's update factorises it first */
                    /* read from the record again (it is current whichever predict ran, or none): keeping the
                     * tick's first copy alive across the predict only costs spills */
                    double a2[F::LP];
                    UKFB_UNROLL
                    for (int e = 0; e < F::LP; ++e) a2[e] = sig[e * TILE];
                    UKFB_UNROLL
                    for (int e = 0; e < F::LP; ++e) UKFB_PS(e) = a2[e];
                    if (!pf_cholesky<12>(a2)) {
                        status |= UKFB_STATUS_NOT_SPD;
                        do_upd = false;
                    }
                }
                const double tr = UKFB_PS(tri(3, 3)) + UKFB_PS(tri(4, 4)) + UKFB_PS(tri(5, 5));
                literal = !(tr < PF_PI2_GUARD);
            }
            if (do_upd) {
                bool spd = true, fast_done = false;
                int stage = 0;
                double delta[12];
                UKFB_UNROLL
                for (int i = 0; i < 12; ++i) delta[i] = 0.0;
                if (!literal) {
                    fast_done = pf_update<false>(sm, lane, sig, kind, zm, Rmeas, p.r_ld, m, delta, status, passes_b, spd, p.gate_d2, stage);
                    if (fast_done) {
                        if (!spd)
                            status |= UKFB_STATUS_NOT_SPD;
                        else
                            dirty_mu = true;
                    }
                }
                if (!fast_done) { /* orientation measurement, failed guard, or a polynomial range was left */
                    PfDelta dl;
                    UKFB_UNROLL
                    for (int i = 0; i < 12; ++i) dl.d[i] = delta[i];
                    /* stage: 0 = nothing done yet, 1 / 2 = where the selector instance stopped after the downdate */
                    const PfLit r = pf_update_slow<WITH_ORI>(sm, lane, sig, kind, zm, Rmeas, p.r_ld, ma, m, dl, literal ? 0 : stage, p.gate_d2);
                    status |= r.status;
                    passes_b = r.passes;
                    if (!(r.status & UKFB_STATUS_NOT_SPD)) {
                        m = r.m;
                        dirty_mu = true;
                    }
                }
            }
        }
        {
            const int pa = passes_a < 7 ? passes_a : 7, pb = passes_b < 7 ? passes_b : 7;
            UKFB_UNROLL
            for (int k = 1; k < 8; ++k) hist[k] += (pa == k) + (pb == k);
        }
    }

    /* ---- write back the mean: the covariance is already in the record ---------------------------------------------- */
    if (valid && dirty_mu) {
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) rec[i * TILE] = m.p[i], rec[(7 + i) * TILE] = m.v[i], rec[(10 + i) * TILE] = m.w[i];
        UKFB_UNROLL
        for (int i = 0; i < 4; ++i) rec[(3 + i) * TILE] = m.q[i];
    }
    if (valid && p.events) {
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) p.acc_mu[b * 3 + i] = ma.acc[i];
    }
    if (valid && status) p.status[b] |= status;
    if (p.hist && valid) {
        unsigned long long* hs = p.hist + (tile % HIST_SLOTS) * 8;
        UKFB_UNROLL
        for (int k = 1; k < 8; ++k)
            if (hist[k]) atomicAdd(hs + k, (unsigned long long)hist[k]);
    }
    if (OVERLAP) tile_done_add(p.tile_done + tile); /* everything this lane stores for the tile has been issued */
}

#undef UKFB_PS

} /* namespace ukfb */

#endif /* UKFB_POSE_FAST_CUH */
