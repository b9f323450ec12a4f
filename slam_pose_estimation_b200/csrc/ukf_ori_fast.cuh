/*
 * ukf_ori_fast.cuh -- the structure-exploiting lane-per-filter step kernel for OrientationUKF (sm_100a).
 *
 * Same StepParams, HBM tile layout and results (to rounding) as ukf_thread_kernel<OriF>, the literal kernel, and the
 * same construction as ukf_pose_fast.cuh: every lane owns one filter and evaluates the reference's estimator -- ukfom
 * predict / update / apply_delta over the 2n+1 = 27 sigma points mu [+] +-L[:,j] with the
 * OrientationUKF models (OrientationUKF.cpp:12-39, 79-89) -- taking sums in closed form where they are exact identities
 * of that sequence.  Tangent order: orientation 0..2, velocity 3..5, gyro bias 6..8, acc bias 9..11, gravity 12.
 *
 *   predict (OrientationUKF.cpp:12-32): the biases decay linearly (b' = b + dt (-1/tau) b) and gravity passes through,
 *     so their deviations are exactly +-D L[6:13,j], D = diag(cg, cg, cg, ca, ca, ca, 1):
 *       - the bias/gravity block of the new covariance is  D Sigma[6:13,6:13] D + Q';
 *       - its cross block with orientation/velocity is  1/2 sum_j D L[6:13,j] (d+_j - d-_j)^T;
 *       - only the 21 orientation/velocity entries are accumulated point by point.
 *     L is lower triangular: columns 3..8 do not perturb the orientation (one prior rotation matrix serves them),
 *     columns 9..12 perturb neither the orientation nor the gyro bias, so their propagated orientation is that of
 *     X0 and their velocity differs from X0's by -+dt (R' L_ba + L_g e3): no sigma-point evaluation at all.
 *   update with the body-velocity measurement h = q^-1 v (OrientationUKF.cpp:34-39, 65-72): columns 6..12 leave h
 *     unchanged and columns 3..5 move it linearly (+- R^T L_v), so only the 6 points of columns 0..2 are evaluated;
 *     (mu [+] +-L_j) [-] mu = +-L_j exactly while |L_ori| < pi (guard: trace(Sigma_ori) < 9, else the literal code),
 *     hence Sigma_xz = 1/2 sum_{j<6} L[:,j] (z+_j - z-_j)^T.  S^-1 by cofactors, K, Sigma - (K S) K^T and
 *     delta = K innov follow the reference's expression order.
 *   apply_delta: the Euclidean components of mu [+] (delta +- L[:,j]) have mean mu + delta and deviations +-L, so the
 *     Euclidean block of the new covariance is that of Sigma - K S K^T; only the orientation rows are recomputed,
 *     from the 6 points of columns 0..2, and only three columns of the second Cholesky factor are needed.
 *
 * SO(3) exp/log are the branch-free polynomial kernels of ukf_pose_fast.cuh; a lane that leaves their range -- a filter
 * that barely knows its attitude -- is served by a second, out-of-line instance of the same structured code with the
 * any-angle exp / log pair (of_predict_slow / of_update_slow), and what is beyond even that, or fails a guard, runs the
 * literal code of ukf_thread.cuh on a per-thread local array.
 */
#ifndef UKFB_ORI_FAST_CUH
#define UKFB_ORI_FAST_CUH

#include "ukf_pose_fast.cuh"

namespace ukfb {

#define UKFB_OS(e) sm[(e) * TILE + lane]

/* shared-memory slots (doubles per lane) holding factor columns with explicit zeros above the diagonal:
 *   predict:  OA(j, i) = j * 13 + i              columns 0..2,  rows 0..12
 *             OB(j, i) = 39 + (j - 3) * 10 + i - 3   columns 3..8,  rows 3..12
 *             OC       = 99 + packed lower 4 x 4  columns 9..12, rows 9..12 (column-major: 4 + 3 + 2 + 1)
 * The update needs no shared memory: all its factor accesses are statically indexed and stay in registers. */
constexpr int OF_OB = 39;
constexpr int OF_OC = 99;
constexpr int OF_PER_LANE = 109;

/* Cholesky of a packed lower N x N in registers, first NCOL columns.  Right-looking: as soon as column j is final the
 * columns to its right are updated with it, so every entry receives the subtractions a_ik a_jk in the order
 * k = 0, 1, ... of LAPACK dpotf2('L') -- bit for bit the same factor -- while finished columns can leave the register
 * file early (STORE(j) is called right after column j is final). */
template <int N, int NCOL, class Store>
UKFB_D bool reg_cholesky(double* a, Store store)
{
    bool ok = true;
    UKFB_UNROLL
    for (int j = 0; j < NCOL; ++j) {
        double ajj = a[tri(j, j)];
        if (!pivot_ok(ajj)) {
            ok = false;
            ajj = 1.0;
        }
        double d, rinv;
        fast_sqrt_rsqrt(ajj, d, rinv);
        a[tri(j, j)] = d;
        UKFB_UNROLL
        for (int i = j + 1; i < N; ++i) a[tri(i, j)] *= rinv;
        UKFB_UNROLL
        for (int l = j + 1; l < NCOL; ++l) {
            UKFB_UNROLL
            for (int i = l; i < N; ++i) a[tri(i, l)] -= a[tri(i, j)] * a[tri(l, j)];
        }
        store(j);
    }
    return ok;
}

template <int N, int NCOL>
UKFB_D bool reg_cholesky(double* a)
{
    return reg_cholesky<N, NCOL>(a, [](int) {});
}

struct OriMu {
    double q[4], v[3], bg[3], ba[3], g;
};

/* inputs of the process model that are the same for every sigma point of a predict */
struct OriCtx {
    double dt;
    double omega[3], acc[3], earth[3];
};

/* a pair of propagated sigma points (OrientationUKF.cpp:12-32), the + and the - point of a column side by side (D2, as in
 * ukf_pose_fast.cuh): orientation and velocity deviations from the reference.
 * qs, vs: the points' orientation and velocity; wb = omega - bg, ab = acc - ba, gs = gravity of the points. */
template <bool WIDE> /* WIDE: the any-angle exp / log pair of ukf_pose_fast.cuh (pf_exp_t / pf_log_t) */
UKFB_D void of_point2(const D2* qs, const D2* vs, const D2* wb, const D2* ab, D2 gs, const OriCtx& cx, const double* ref_q,
                      const double* ref_v, D2* d, bool& slow)
{
    D2 av[3], e[4], qn[4], an[3], r[4];
    pf_rotate(qs, wb, av);
    av[0] = av[0] - cx.earth[0], av[1] = av[1] - cx.earth[1], av[2] = av[2] - cx.earth[2];
    pf_exp_t<WIDE>(av, cx.dt, e, slow);
    quat_mul(e, qs, qn);
    pf_rotate(qn, ab, an); /* with the UPDATED orientation (:22-23) */
    an[2] = an[2] - gs;
    d[3] = tfma(cx.dt, an[0], vs[0]) - ref_v[0];
    d[4] = tfma(cx.dt, an[1], vs[1]) - ref_v[1];
    d[5] = tfma(cx.dt, an[2], vs[2]) - ref_v[2];
    quat_mul_conj(qn, ref_q, r);
    pf_log_t<WIDE>(r, d, slow);
}

/* the +/- sigma points of a column j < 3 (everything perturbed); L receives the column (13 entries) */
template <bool WIDE>
UKFB_D void of_pair_a(const double* sm, int lane, int j, const OriMu& m, const OriCtx& cx, const double* ref_q, const double* ref_v,
                      double* L, double* dpl, double* dmi, bool& slow)
{
    UKFB_UNROLL
    for (int i = 0; i < 13; ++i) L[i] = UKFB_OS(j * 13 + i);
    double e[4];
    pf_exp1<WIDE>(L, 1.0, e, slow);
    const double* q = m.q;
    const double t0 = e[0] * q[3] + e[1] * q[2] - e[2] * q[1];
    const double t1 = e[1] * q[3] + e[2] * q[0] - e[0] * q[2];
    const double t2 = e[2] * q[3] + e[0] * q[1] - e[1] * q[0];
    const double t3 = -(e[0] * q[0] + e[1] * q[1] + e[2] * q[2]);
    const D2 qs[4] = {D2(fma(e[3], q[0], t0), fma(e[3], q[0], -t0)), D2(fma(e[3], q[1], t1), fma(e[3], q[1], -t1)),
                      D2(fma(e[3], q[2], t2), fma(e[3], q[2], -t2)), D2(fma(e[3], q[3], t3), fma(e[3], q[3], -t3))};
    const D2 vs[3] = {D2(m.v[0] + L[3], m.v[0] - L[3]), D2(m.v[1] + L[4], m.v[1] - L[4]), D2(m.v[2] + L[5], m.v[2] - L[5])};
    const D2 wb[3] = {D2(cx.omega[0] - (m.bg[0] + L[6]), cx.omega[0] - (m.bg[0] - L[6])),
                      D2(cx.omega[1] - (m.bg[1] + L[7]), cx.omega[1] - (m.bg[1] - L[7])),
                      D2(cx.omega[2] - (m.bg[2] + L[8]), cx.omega[2] - (m.bg[2] - L[8]))};
    const D2 ab[3] = {D2(cx.acc[0] - (m.ba[0] + L[9]), cx.acc[0] - (m.ba[0] - L[9])),
                      D2(cx.acc[1] - (m.ba[1] + L[10]), cx.acc[1] - (m.ba[1] - L[10])),
                      D2(cx.acc[2] - (m.ba[2] + L[11]), cx.acc[2] - (m.ba[2] - L[11]))};
    D2 d[6];
    of_point2<WIDE>(qs, vs, wb, ab, D2(m.g + L[12], m.g - L[12]), cx, ref_q, ref_v, d, slow);
    UKFB_UNROLL
    for (int i = 0; i < 6; ++i) dpl[i] = d[i].a, dmi[i] = d[i].b;
}

/* the +/- sigma points of a column 3 <= j < 9: orientation unperturbed.  Rm = R(q), w0 = R (omega - bg) - earth,
 * c = q * conj(ref_q); L receives rows 3..12 of the column (10 entries: velocity, gyro bias, acc bias, gravity). */
template <bool WIDE>
UKFB_D void of_pair_b(const double* sm, int lane, int j, const OriMu& m, const OriCtx& cx, const double* Rm, const double* w0,
                      const double* c, const double* ref_v, double* L, double* dpl, double* dmi, bool& slow)
{
    UKFB_UNROLL
    for (int i = 0; i < 10; ++i) L[i] = UKFB_OS(OF_OB + (j - 3) * 10 + i);
    double u[3];
    pf_matvec(Rm, L + 3, u);
    const D2 av[3] = {D2(w0[0] - u[0], w0[0] + u[0]), D2(w0[1] - u[1], w0[1] + u[1]), D2(w0[2] - u[2], w0[2] + u[2])};
    D2 e[4], qn[4], r[4], an[3], d[6];
    pf_exp_t<WIDE>(av, cx.dt, e, slow);
    quat_mul(e, m.q, qn);
    const D2 ab[3] = {D2(cx.acc[0] - (m.ba[0] + L[6]), cx.acc[0] - (m.ba[0] - L[6])),
                      D2(cx.acc[1] - (m.ba[1] + L[7]), cx.acc[1] - (m.ba[1] - L[7])),
                      D2(cx.acc[2] - (m.ba[2] + L[8]), cx.acc[2] - (m.ba[2] - L[8]))};
    pf_rotate(qn, ab, an);
    an[2] = an[2] - D2(m.g + L[9], m.g - L[9]);
    d[3] = tfma(cx.dt, an[0], D2(m.v[0] + L[0], m.v[0] - L[0])) - ref_v[0];
    d[4] = tfma(cx.dt, an[1], D2(m.v[1] + L[1], m.v[1] - L[1])) - ref_v[1];
    d[5] = tfma(cx.dt, an[2], D2(m.v[2] + L[2], m.v[2] - L[2])) - ref_v[2];
    quat_mul(e, c, r);
    pf_log_t<WIDE>(r, d, slow);
    UKFB_UNROLL
    for (int i = 0; i < 6; ++i) dpl[i] = d[i].a, dmi[i] = d[i].b;
}

/* The column pairs as the two instances of the structured code call them (as pf_pair_*_sel in ukf_pose_fast.cuh): the hot
 * instance knows only the short polynomials; the out-of-line one sends an orientation column this large to the any-angle
 * pair at once and redoes any other column whose points left the short range. */
template <bool WIDE>
UKFB_D void of_pair_a_sel(const double* sm, int lane, int j, const OriMu& m, const OriCtx& cx, const double* ref_q, const double* ref_v,
                          double* L, double* dpl, double* dmi, bool& slow)
{
    if (!WIDE) {
        of_pair_a<false>(sm, lane, j, m, cx, ref_q, ref_v, L, dpl, dmi, slow);
        return;
    }
    const double l0 = UKFB_OS(j * 13), l1 = UKFB_OS(j * 13 + 1), l2 = UKFB_OS(j * 13 + 2);
    bool redo = l0 * l0 + l1 * l1 + l2 * l2 > PF_WIDE_COLUMN_N2;
    if (!redo) of_pair_a<false>(sm, lane, j, m, cx, ref_q, ref_v, L, dpl, dmi, redo);
    if (redo) of_pair_a<true>(sm, lane, j, m, cx, ref_q, ref_v, L, dpl, dmi, slow);
}
template <bool WIDE>
UKFB_D void of_pair_b_sel(const double* sm, int lane, int j, const OriMu& m, const OriCtx& cx, const double* Rm, const double* w0,
                          const double* c, const double* ref_v, double* L, double* dpl, double* dmi, bool& slow)
{
    if (!WIDE) {
        of_pair_b<false>(sm, lane, j, m, cx, Rm, w0, c, ref_v, L, dpl, dmi, slow);
        return;
    }
    bool redo = false;
    of_pair_b<false>(sm, lane, j, m, cx, Rm, w0, c, ref_v, L, dpl, dmi, redo);
    if (redo) of_pair_b<true>(sm, lane, j, m, cx, Rm, w0, c, ref_v, L, dpl, dmi, slow);
}


/* ---- literal fallbacks (cold, out of line): the general code of ukf_thread.cuh on this lane's filter ---------- */
#ifdef UKFB_SIMT_EMU
inline unsigned long long of_fallbacks[5] = {0, 0, 0, 0, 0}; /* [3], [4]: predicts / updates served by the any-angle instance */
#define UKFB_OF_COUNT(i) __atomic_fetch_add(&of_fallbacks[i], 1ull, __ATOMIC_RELAXED)
#else
#define UKFB_OF_COUNT(i)
#endif
struct OfLit {
    OriMu m;
    uint32_t status;
    int passes;
};

UKFB_D void of_mu_to_slots(double* loc, const OriMu& m)
{
    typedef TSmem<OriF> TS;
    UKFB_UNROLL
    for (int i = 0; i < 4; ++i) loc[TS::OFF_MU + i] = m.q[i];
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        loc[TS::OFF_MU + 4 + i] = m.v[i];
        loc[TS::OFF_MU + 7 + i] = m.bg[i];
        loc[TS::OFF_MU + 10 + i] = m.ba[i];
    }
    loc[TS::OFF_MU + 13] = m.g;
}

UKFB_D void of_mu_from_slots(const double* loc, OriMu& m)
{
    typedef TSmem<OriF> TS;
    UKFB_UNROLL
    for (int i = 0; i < 4; ++i) m.q[i] = loc[TS::OFF_MU + i];
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        m.v[i] = loc[TS::OFF_MU + 4 + i];
        m.bg[i] = loc[TS::OFF_MU + 7 + i];
        m.ba[i] = loc[TS::OFF_MU + 10 + i];
    }
    m.g = loc[TS::OFF_MU + 13];
}

UKFB_DNI OfLit of_literal_predict(double* sig, const double* Qp, ModelArgs ma, OriMu m)
{
    typedef TSmem<OriF> TS;
    UKFB_OF_COUNT(0);
    double loc[TS::PER_LANE];
    OfLit r;
    r.m = m, r.passes = 0;
    of_mu_to_slots(loc, m);
    if (!cholesky_thread<OriF, 1>(sig, loc, 0)) {
        r.status = UKFB_STATUS_NOT_SPD;
        return r;
    }
    store_noise<OriF, 1>(loc, 0, sig, Qp, nullptr, ma);
    r.status = mean_and_cov<OriF, true, 1>(loc, 0, sig, ma, &r.passes);
    UKFB_UNROLL
    for (int i = 0; i < OriF::MU; ++i) loc[TS::OFF_MU + i] = loc[TS::OFF_REF + i];
    of_mu_from_slots(loc, r.m);
    return r;
}

struct OfDelta {
    double d[13];
};

/* first = true: the whole ukfom update; false: apply_delta only, from the record's Sigma - K S K^T and `delta` */
UKFB_DNI OfLit of_literal_update(double* sig, int kind, const double* zm, const double* Rm, int r_ld, ModelArgs ma, OriMu m,
                                 OfDelta delta, bool first, double gate_d2)
{
    typedef TSmem<OriF> TS;
    UKFB_OF_COUNT(first ? 1 : 2);
    double loc[TS::PER_LANE];
    OfLit r;
    r.m = m, r.passes = 0, r.status = 0;
    of_mu_to_slots(loc, m);
    if (first) {
        if (!cholesky_thread<OriF, 1>(sig, loc, 0)) {
            r.status = UKFB_STATUS_NOT_SPD;
            return r;
        }
        r.status = update_first_half<OriF, 1>(loc, 0, sig, kind, zm, Rm, r_ld, gate_d2);
        if (r.status & UKFB_STATUS_MEAS_REJECTED) return r;
    } else {
        UKFB_UNROLL
        for (int i = 0; i < 13; ++i) loc[TS::OFF_DELTA + i] = delta.d[i];
    }
    if (!cholesky_thread<OriF, 1>(sig, loc, 0)) {
        r.status |= UKFB_STATUS_NOT_SPD;
        return r;
    }
    r.status |= mean_and_cov<OriF, false, 1>(loc, 0, sig, ma, &r.passes);
    UKFB_UNROLL
    for (int i = 0; i < OriF::MU; ++i) loc[TS::OFF_MU + i] = loc[TS::OFF_REF + i];
    of_mu_from_slots(loc, r.m);
    return r;
}

/* ---- structured predict.  Returns false when a polynomial range was left (nothing has been modified then) ------ */
/* On success: m holds the new mean, the record holds the new covariance.  `a` (prior covariance) is destroyed. */
/* WIDE = false: the hot instance (short polynomials; does not start on a filter whose orientation uncertainty certainly
 * leaves their range); WIDE = true: the out-of-line instance behind of_predict_slow, any orientation uncertainty. */
template <bool WIDE>
UKFB_D bool of_predict(int q_diagonal, double* sm, int lane, double* sig, double* a, const double* Qp, const ModelArgs& ma, OriMu& m,
                       uint32_t& status, int& passes_out, bool& spd)
{
    if (!WIDE && a[tri(0, 0)] + a[tri(1, 1)] + a[tri(2, 2)] > PF_WIDE_TRACE) return false; /* nothing has been modified */
    OriCtx cx;
    cx.dt = ma.dt;
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) cx.omega[i] = ma.omega[i], cx.acc[i] = ma.acc[i], cx.earth[i] = ma.earth[i];
    const double dt = ma.dt;
    /* the factor goes to the OA / OB / OC blocks column by column as it is produced (nothing outside shared memory is
     * modified before the factorisation has succeeded) */
    spd = reg_cholesky<13, 13>(a, [&](int j) {
        if (j < 3) {
            UKFB_UNROLL
            for (int i = 0; i < 13; ++i) UKFB_OS(j * 13 + i) = i >= j ? a[tri(i >= j ? i : j, j)] : 0.0;
        } else if (j < 9) {
            UKFB_UNROLL
            for (int i = 3; i < 13; ++i) UKFB_OS(OF_OB + (j - 3) * 10 + (i - 3)) = i >= j ? a[tri(i >= j ? i : j, j)] : 0.0;
        } else {
            const int base = OF_OC + (j == 9 ? 0 : (j == 10 ? 4 : (j == 11 ? 7 : 9)));
            UKFB_UNROLL
            for (int i = j; i < 13; ++i) UKFB_OS(base + i - j) = a[tri(i, j)];
        }
    });
    if (!spd) return true;
    bool slow = !pf_unit(m.q);
    double Rm[9];
    quat_matrix(m.q, Rm);
    /* X0' = g(mu) */
    double w0[3], q0n[4], v0n[3], R0n[9];
    {
        const double wb[3] = {cx.omega[0] - m.bg[0], cx.omega[1] - m.bg[1], cx.omega[2] - m.bg[2]};
        double e0[4], an[3];
        pf_matvec(Rm, wb, w0);
        w0[0] -= cx.earth[0], w0[1] -= cx.earth[1], w0[2] -= cx.earth[2];
        pf_exp1<WIDE>(w0, dt, e0, slow);
        quat_mul(e0, m.q, q0n);
        quat_matrix(q0n, R0n);
        const double ab[3] = {cx.acc[0] - m.ba[0], cx.acc[1] - m.ba[1], cx.acc[2] - m.ba[2]};
        pf_matvec(R0n, ab, an);
        an[2] -= m.g;
        v0n[0] = fma(dt, an[0], m.v[0]);
        v0n[1] = fma(dt, an[1], m.v[1]);
        v0n[2] = fma(dt, an[2], m.v[2]);
    }
    double ref_q[4] = {q0n[0], q0n[1], q0n[2], q0n[3]};
    double ref_v[3] = {v0n[0], v0n[1], v0n[2]};

    /* ---- manifold mean (ukfom sigma_points_mean): only orientation and velocity can move */
    int it = 0, passes = 0;
    while (true) {
        double md[6], d0[6];
        {
            double r[4];
            quat_mul_conj(q0n, ref_q, r);
            pf_log1<WIDE>(r, d0, slow);
            d0[3] = v0n[0] - ref_v[0], d0[4] = v0n[1] - ref_v[1], d0[5] = v0n[2] - ref_v[2];
        }
        /* X0 and the 8 points of columns 9..12 share X0's orientation; their velocity offsets cancel in pairs */
        UKFB_UNROLL
        for (int i = 0; i < 6; ++i) md[i] = 9.0 * d0[i];
        UKFB_NOUNROLL
        for (int j = 0; j < 3; ++j) {
            double L[13], dpl[6], dmi[6];
            of_pair_a_sel<WIDE>(sm, lane, j, m, cx, ref_q, ref_v, L, dpl, dmi, slow);
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) md[i] += dpl[i] + dmi[i];
        }
        double c[4];
        quat_mul_conj(m.q, ref_q, c);
        UKFB_NOUNROLL
        for (int j = 3; j < 9; ++j) {
            double L[10], dpl[6], dmi[6];
            of_pair_b_sel<WIDE>(sm, lane, j, m, cx, Rm, w0, c, ref_v, L, dpl, dmi, slow);
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) md[i] += dpl[i] + dmi[i];
        }
        double n2 = 0.0;
        UKFB_UNROLL
        for (int i = 0; i < 6; ++i) {
            md[i] = div_ns<OriF::NS>(md[i]);
            n2 += md[i] * md[i];
        }
        ref_v[0] += md[3], ref_v[1] += md[4], ref_v[2] += md[5];
        {
            double e[4], r[4];
            pf_exp1<WIDE>(md, 1.0, e, slow);
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

    /* ---- covariance: C = orientation/velocity block (21), X = cross block (rows 6..12 x columns 0..5) */
    const double cg = fma(dt, ma.neg_inv_tau_g, 1.0), ca = fma(dt, ma.neg_inv_tau_a, 1.0);
    /* X = sum_j L[6:13, j] (d+_j - d-_j)^T is only partly accumulated point by point (as in ukf_pose_fast.cuh: 42 accumulators
     * next to C and the per-filter context are more than the register file holds).  The differences of columns 0..2 go to the
     * slots of rows 0..5 of their column, the orientation part of the differences of columns 3..8 to the three slots of their
     * velocity rows (both consumed once the column's sigma points exist); only the velocity part of columns 3..8 is accumulated
     * in the loop (Xv, 21 values), and X is formed after the loops from the factor rows still in place. */
    double C[21], X[42];
    {
        double d0[6], r[4];
        quat_mul_conj(q0n, ref_q, r);
        pf_log1<WIDE>(r, d0, slow);
        d0[3] = v0n[0] - ref_v[0], d0[4] = v0n[1] - ref_v[1], d0[5] = v0n[2] - ref_v[2];
        UKFB_UNROLL
        for (int i = 0; i < 6; ++i) {
            UKFB_UNROLL
            for (int k = 0; k <= i; ++k) C[tri(i, k)] = d0[i] * d0[k];
        }
        UKFB_NOUNROLL
        for (int j = 0; j < 3; ++j) {
            double L[13], dpl[6], dmi[6];
            of_pair_a_sel<WIDE>(sm, lane, j, m, cx, ref_q, ref_v, L, dpl, dmi, slow);
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) {
                UKFB_UNROLL
                for (int k = 0; k <= i; ++k) C[tri(i, k)] = fma(dpl[i], dpl[k], fma(dmi[i], dmi[k], C[tri(i, k)]));
            }
            UKFB_UNROLL
            for (int k = 0; k < 6; ++k) UKFB_OS(j * 13 + k) = dpl[k] - dmi[k];
        }
        double Xv[21]; /* columns 3..8: sum_j L[6:13, j] (dv+_j - dv-_j)^T, the velocity part of their differences */
        UKFB_UNROLL
        for (int i = 0; i < 21; ++i) Xv[i] = 0.0;
        double c[4];
        quat_mul_conj(m.q, ref_q, c);
        UKFB_NOUNROLL
        for (int j = 3; j < 9; ++j) {
            double L[10], dpl[6], dmi[6];
            of_pair_b_sel<WIDE>(sm, lane, j, m, cx, Rm, w0, c, ref_v, L, dpl, dmi, slow);
            UKFB_UNROLL
            for (int i = 0; i < 6; ++i) {
                UKFB_UNROLL
                for (int k = 0; k <= i; ++k) C[tri(i, k)] = fma(dpl[i], dpl[k], fma(dmi[i], dmi[k], C[tri(i, k)]));
            }
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) UKFB_OS(OF_OB + (j - 3) * 10 + k) = dpl[k] - dmi[k];
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) {
                const double dd = dpl[3 + k] - dmi[3 + k];
                UKFB_UNROLL
                for (int i = 0; i < 7; ++i) Xv[i * 3 + k] = fma(L[3 + i], dd, Xv[i * 3 + k]);
            }
        }
        /* X, after the loops */
        UKFB_UNROLL
        for (int i = 0; i < 7; ++i) {
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) X[i * 6 + k] = 0.0, X[i * 6 + 3 + k] = Xv[i * 3 + k];
        }
        UKFB_NOUNROLL
        for (int j = 0; j < 3; ++j) {
            double dd[6];
            UKFB_UNROLL
            for (int k = 0; k < 6; ++k) dd[k] = UKFB_OS(j * 13 + k);
            UKFB_UNROLL
            for (int i = 0; i < 7; ++i) {
                const double l = UKFB_OS(j * 13 + 6 + i);
                UKFB_UNROLL
                for (int k = 0; k < 6; ++k) X[i * 6 + k] = fma(l, dd[k], X[i * 6 + k]);
            }
        }
        UKFB_NOUNROLL
        for (int j = 3; j < 9; ++j) {
            double dd[3];
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) dd[k] = UKFB_OS(OF_OB + (j - 3) * 10 + k);
            UKFB_UNROLL
            for (int i = 0; i < 7; ++i) {
                const double l = UKFB_OS(OF_OB + (j - 3) * 10 + 3 + i);
                UKFB_UNROLL
                for (int k = 0; k < 3; ++k) X[i * 6 + k] = fma(l, dd[k], X[i * 6 + k]);
            }
        }
        /* columns 9..12: orientation deviation d0, velocity deviation d0 -+ wv, wv = dt (R' L_ba + L_g e3) */
        {
            int s = OF_OC;
            UKFB_UNROLL
            for (int j = 9; j < 13; ++j) {
                double Lc[4] = {0.0, 0.0, 0.0, 0.0}; /* rows 9..12 of the column */
                UKFB_UNROLL
                for (int i = j; i < 13; ++i) Lc[i - 9] = UKFB_OS(s++);
                double wv[3];
                pf_matvec(R0n, Lc, wv);
                wv[0] *= dt, wv[1] *= dt, wv[2] = dt * (wv[2] + Lc[3]);
                double dpl[6], dmi[6];
                UKFB_UNROLL
                for (int i = 0; i < 3; ++i) {
                    dpl[i] = dmi[i] = d0[i];
                    dpl[3 + i] = d0[3 + i] - wv[i];
                    dmi[3 + i] = d0[3 + i] + wv[i];
                }
                UKFB_UNROLL
                for (int i = 0; i < 6; ++i) {
                    UKFB_UNROLL
                    for (int k = 0; k <= i; ++k) C[tri(i, k)] = fma(dpl[i], dpl[k], fma(dmi[i], dmi[k], C[tri(i, k)]));
                }
                UKFB_UNROLL
                for (int k = 3; k < 6; ++k) {
                    const double dd = -2.0 * wv[k - 3];
                    UKFB_UNROLL
                    for (int i = 3; i < 7; ++i) X[i * 6 + k] = fma(Lc[i - 3], dd, X[i * 6 + k]);
                }
            }
        }
    }
    if (slow) return false;

    /* ---- new covariance = 1/2 C + process noise dt^2 Q' (OrientationUKF.cpp:81-86), committed to the record.
     * qv(i, k) = Q[i][k], i >= k: a load, or for a broadcast diagonal Q a load on the diagonal and a literal zero elsewhere */
    auto commit = [&](auto qv, auto iso) {
        const double scale = dt * dt;
        double nb[12]; /* the two rotated blocks of the noise; every other entry is scale * Q */
        UKFB_UNROLL
        for (int i = 0; i < 12; ++i) nb[i] = 0.0;
        if (decltype(iso)::value) {
            /* R (q I) R^T = q I for both rotated blocks: the rotation changes nothing */
            nb[tri(0, 0)] = nb[tri(1, 1)] = nb[tri(2, 2)] = scale * qv(0, 0);
            nb[6 + tri(0, 0)] = nb[6 + tri(1, 1)] = nb[6 + tri(2, 2)] = scale * qv(3, 3);
        } else {
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
        }
        UKFB_UNROLL
        for (int i = 0; i < 13; ++i) {
            const double di = i < 9 ? cg : (i < 12 ? ca : 1.0);
            UKFB_UNROLL
            for (int k = 0; k <= i; ++k) {
                const int e = tri(i, k);
                const double dk = k < 9 ? cg : (k < 12 ? ca : 1.0);
                const double nz = (i < 6 && i / 3 == k / 3) ? nb[(i / 3) * 6 + tri(i % 3, k % 3)] : scale * qv(i, k);
                double s;
                if (i < 6)
                    s = fma(0.5, C[e], nz);
                else if (k < 6)
                    s = fma(0.5 * di, X[(i - 6) * 6 + k], nz);
                else
                    s = fma(di * dk, sig[e * TILE], nz);
                sig[e * TILE] = s;
            }
        }
    };
    if (q_diagonal == 2) /* diagonal, and a multiple of the identity in each of the two rotated 3 x 3 blocks */
        commit([&](int i, int k) { return i == k ? UKFB_LDG(Qp + tri(i, i)) : 0.0; }, TrueT());
    else if (q_diagonal)
        commit([&](int i, int k) { return i == k ? UKFB_LDG(Qp + tri(i, i)) : 0.0; }, FalseT());
    else
        commit([&](int i, int k) { return UKFB_LDG(Qp + tri(i, k)); }, FalseT());
    m.q[0] = ref_q[0], m.q[1] = ref_q[1], m.q[2] = ref_q[2], m.q[3] = ref_q[3];
    m.v[0] = ref_v[0], m.v[1] = ref_v[1], m.v[2] = ref_v[2];
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        const double dg = ma.neg_inv_tau_g * m.bg[i];
        m.bg[i] += dt * dg;
        const double da = ma.neg_inv_tau_a * m.ba[i];
        m.ba[i] += dt * da;
    }
    passes_out = passes;
    return true;
}

/* ---- structured update with the body-velocity measurement.  `a`: the covariance (also in the record), destroyed.
 * Returns true when done (spd = false: a factorisation failed; a rejected measurement sets its status bit).
 * Returns false when a polynomial range was left: stage = 0: nothing has been modified; stage = 1: the record holds
 * Sigma - K S K^T, `delta` = K innov, m is untouched (the caller runs the literal apply_delta). */
template <bool WIDE>
UKFB_D bool of_apply_delta(double* sig, double* a, OriMu& m, const double* delta, uint32_t& status, int& passes_out, bool& spd, bool slow);

/* WIDE as in of_predict.  The closed forms rest on (mu [+] L_j) [-] mu = L_j, i.e. on every |L_ori[:, j]| < pi: the caller
 * of the hot instance guarantees it through trace(Sigma_ori) < 9, the out-of-line instance looks at the three columns
 * themselves (a filter that does not know its attitude at all) and leaves a column next to pi to the literal code. */
template <bool WIDE>
UKFB_D bool of_update(double* sm, int lane, double* sig, double* a, const double* zm, const double* Rmeas, int r_ld, OriMu& m,
                      double* delta, uint32_t& status, int& passes_out, bool& spd, int& stage, double gate_d2)
{
    stage = 0;
    if (!WIDE && a[tri(0, 0)] + a[tri(1, 1)] + a[tri(2, 2)] > PF_WIDE_TRACE) return false; /* nothing has been modified */
    spd = reg_cholesky<13, 13>(a);
    if (!spd) return true;
    bool slow = !pf_unit(m.q);
    if (WIDE) {
        UKFB_UNROLL
        for (int j = 0; j < 3; ++j) {
            double n2 = a[tri(2, j)] * a[tri(2, j)];
            if (j <= 1) n2 += a[tri(1, j)] * a[tri(1, j)];
            if (j == 0) n2 += a[tri(0, 0)] * a[tri(0, 0)];
            slow = slow || !(n2 < PF_PI2_COLUMN);
        }
        if (slow) return false;
    }
    /* Z of X0, of the 6 points of columns 0..2, and the linear offsets of columns 3..5.  The factor stays in registers:
     * every index below is static */
#define OF_L(i, j) ((i) >= (j) ? a[tri((i) >= (j) ? (i) : (j), (j))] : 0.0)
    double z0[3], zp[9], zn[9], u[9];
    quat_inv_rotate(m.q, m.v, z0);
    UKFB_UNROLL
    for (int j = 0; j < 3; ++j) {
        double e[4];
        const double Lo[3] = {OF_L(0, j), OF_L(1, j), OF_L(2, j)};
        pf_exp1<WIDE>(Lo, 1.0, e, slow);
        const double* q = m.q;
        const double t0 = e[0] * q[3] + e[1] * q[2] - e[2] * q[1];
        const double t1 = e[1] * q[3] + e[2] * q[0] - e[0] * q[2];
        const double t2 = e[2] * q[3] + e[0] * q[1] - e[1] * q[0];
        const double t3 = -(e[0] * q[0] + e[1] * q[1] + e[2] * q[2]);
        const double Lv[3] = {a[tri(3, j)], a[tri(4, j)], a[tri(5, j)]};
        {
            const double qs[4] = {fma(e[3], q[0], t0), fma(e[3], q[1], t1), fma(e[3], q[2], t2), fma(e[3], q[3], t3)};
            const double vs[3] = {m.v[0] + Lv[0], m.v[1] + Lv[1], m.v[2] + Lv[2]};
            quat_inv_rotate(qs, vs, zp + 3 * j);
        }
        {
            const double qs[4] = {fma(e[3], q[0], -t0), fma(e[3], q[1], -t1), fma(e[3], q[2], -t2), fma(e[3], q[3], -t3)};
            const double vs[3] = {m.v[0] - Lv[0], m.v[1] - Lv[1], m.v[2] - Lv[2]};
            quat_inv_rotate(qs, vs, zn + 3 * j);
        }
    }
    UKFB_UNROLL
    for (int j = 3; j < 6; ++j) {
        const double Lv[3] = {OF_L(3, j), OF_L(4, j), OF_L(5, j)};
        quat_inv_rotate(m.q, Lv, u + 3 * (j - 3));
    }
    if (slow) return false;

    /* mean of Z (ukfom sigma_points_mean on the measurement space) */
    double zref[3] = {z0[0], z0[1], z0[2]};
    {
#if UKFB_EUCLID_MEAS_DIRECT_MEAN
        UKFB_UNROLL
        for (int cc = 0; cc < 3; ++cc) {
            double s = 21.0 * z0[cc];
            UKFB_UNROLL
            for (int j = 0; j < 3; ++j) s += zp[3 * j + cc] + zn[3 * j + cc];
            zref[cc] = div_ns<OriF::NS>(s);
        }
#else
        int it = 0;
        while (true) {
            double md[3];
            UKFB_UNROLL
            for (int cc = 0; cc < 3; ++cc) {
                double s = 21.0 * (z0[cc] - zref[cc]);
                UKFB_UNROLL
                for (int j = 0; j < 3; ++j) s += (zp[3 * j + cc] - zref[cc]) + (zn[3 * j + cc] - zref[cc]);
                md[cc] = div_ns<OriF::NS>(s);
            }
            const double n2 = md[0] * md[0] + md[1] * md[1] + md[2] * md[2];
            zref[0] += md[0], zref[1] += md[1], zref[2] += md[2];
            if (!(n2 > UKFB_MEAN_TOL * UKFB_MEAN_TOL)) break;
            if (++it >= UKFB_MEAN_MAX_IT) {
                status |= UKFB_STATUS_MEAN_NO_CONVERGE;
                break;
            }
        }
#endif
    }

    /* S = 1/2 sum dz dz^T + R,  Sxz = 1/2 sum_{j<6} L[:,j] (z+_j - z-_j)^T */
    double S[9], Sxz[39];
    {
        const double dc[3] = {z0[0] - zref[0], z0[1] - zref[1], z0[2] - zref[2]};
        UKFB_UNROLL
        for (int r = 0; r < 3; ++r) {
            UKFB_UNROLL
            for (int cc = 0; cc < 3; ++cc) {
                double s = 15.0 * dc[r] * dc[cc]; /* X0 and the 14 points of columns 6..12 */
                UKFB_UNROLL
                for (int j = 0; j < 3; ++j) {
                    s = fma(zp[3 * j + r] - zref[r], zp[3 * j + cc] - zref[cc], s);
                    s = fma(zn[3 * j + r] - zref[r], zn[3 * j + cc] - zref[cc], s);
                    s = fma(dc[r] + u[3 * j + r], dc[cc] + u[3 * j + cc], s);
                    s = fma(dc[r] - u[3 * j + r], dc[cc] - u[3 * j + cc], s);
                }
                S[r * 3 + cc] = fma(0.5, s, Rmeas[r * r_ld + cc]);
            }
        }
        UKFB_UNROLL
        for (int i = 0; i < 39; ++i) Sxz[i] = 0.0;
        UKFB_UNROLL
        for (int j = 0; j < 3; ++j) {
            UKFB_UNROLL
            for (int cc = 0; cc < 3; ++cc) {
                const double ddz = zp[3 * j + cc] - zn[3 * j + cc];
                UKFB_UNROLL
                for (int i = j; i < 13; ++i) Sxz[i * 3 + cc] = fma(a[tri(i, j)], ddz, Sxz[i * 3 + cc]);
            }
        }
        UKFB_UNROLL
        for (int j = 3; j < 6; ++j) {
            UKFB_UNROLL
            for (int cc = 0; cc < 3; ++cc) {
                const double ddz = 2.0 * u[3 * (j - 3) + cc];
                UKFB_UNROLL
                for (int i = j; i < 13; ++i) Sxz[i * 3 + cc] = fma(a[tri(i, j)], ddz, Sxz[i * 3 + cc]);
            }
        }
        UKFB_UNROLL
        for (int i = 0; i < 39; ++i) Sxz[i] *= 0.5;
    }
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
    const double innov[3] = {zm[0] - zref[0], zm[1] - zref[1], zm[2] - zref[2]};
    {
        double d2 = 0.0;
        UKFB_UNROLL
        for (int r = 0; r < 3; ++r) d2 += innov[r] * (Si[r * 3] * innov[0] + Si[r * 3 + 1] * innov[1] + Si[r * 3 + 2] * innov[2]);
        if (d2 > gate_d2) {
            status |= UKFB_STATUS_MEAS_REJECTED;
            return true;
        }
    }
    /* Row by row: K[i,:] = Sxz[i,:] S^-1 (in place of Sxz), (K S)[i,:], delta_i, and row i of
     * Sigma <- Sigma - (K S) K^T -- to the record, and kept in registers for the factorisation */
    UKFB_UNROLL
    for (int i = 0; i < 13; ++i) {
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
        for (int cc = 0; cc < 3; ++cc) {
            ks3[cc] = Sxz[i * 3 + cc]; /* (K S)[i,:] = (Sxz S^-1 S)[i,:] = Sxz[i,:] */
            Sxz[i * 3 + cc] = k3[cc];
            dl += k3[cc] * innov[cc];
        }
        delta[i] = dl;
        UKFB_UNROLL
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) s += ks3[k] * Sxz[j * 3 + k];
            const double x = sig[tri(i, j) * TILE] - s;
            a[tri(i, j)] = x;
            sig[tri(i, j) * TILE] = x;
        }
    }
    stage = 1;
    return of_apply_delta<WIDE>(sig, a, m, delta, status, passes_out, spd, slow);
#undef OF_L
}

/* ---- apply_delta of the structured update: `a` = Sigma - K S K^T (also in the record), `delta` = K innov.  Returns false
 * when a polynomial range was left (nothing more has been modified: the caller runs the any-angle instance or the
 * literal apply_delta on the record). */
template <bool WIDE>
UKFB_D bool of_apply_delta(double* sig, double* a, OriMu& m, const double* delta, uint32_t& status, int& passes_out, bool& spd, bool slow)
{
#define OF_L(i, j) ((i) >= (j) ? a[tri((i) >= (j) ? (i) : (j), (j))] : 0.0)
    /* first three columns of the factor of the updated covariance (the reference factorises all of it: a failure in the
     * later columns shows at the next factorisation of this filter instead) */
    spd = reg_cholesky<13, 3>(a);
    if (!spd) return true;

    /* ---- apply_delta: orientation rows only */
    double e0[4], q0n[4];
    pf_exp1<WIDE>(delta, 1.0, e0, slow);
    quat_mul(e0, m.q, q0n);
    double ref_q[4] = {q0n[0], q0n[1], q0n[2], q0n[3]};
    double vp[9], vn[9]; /* delta_ori +- L'_ori of columns 0..2 */
    UKFB_UNROLL
    for (int j = 0; j < 3; ++j) {
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) {
            const double l = OF_L(i, j);
            vp[3 * j + i] = delta[i] + l;
            vn[3 * j + i] = delta[i] - l;
        }
    }
    int it = 0, passes = 0;
    while (true) {
        double c[4], r[4], d0[3], md[3];
        quat_mul_conj(m.q, ref_q, c);
        quat_mul(e0, c, r);
        pf_log1<WIDE>(r, d0, slow);
        md[0] = 21.0 * d0[0], md[1] = 21.0 * d0[1], md[2] = 21.0 * d0[2];
        UKFB_UNROLL
        for (int j = 0; j < 3; ++j) {
            double ep[4], en[4], rp[4], rn[4], dp[3], dn[3];
            pf_exp1<WIDE>(vp + 3 * j, 1.0, ep, slow);
            pf_exp1<WIDE>(vn + 3 * j, 1.0, en, slow);
            quat_mul(ep, c, rp);
            quat_mul(en, c, rn);
            pf_log1<WIDE>(rp, dp, slow);
            pf_log1<WIDE>(rn, dn, slow);
            md[0] += dp[0] + dn[0], md[1] += dp[1] + dn[1], md[2] += dp[2] + dn[2];
        }
        double n2 = 0.0;
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) {
            md[i] = div_ns<OriF::NS>(md[i]);
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
    /* covariance of the orientation rows: Coo (6) and the cross block with the 10 Euclidean components */
    double Coo[6], Xc[30];
    if (!slow) {
        double c[4], r[4], d0[3];
        quat_mul_conj(m.q, ref_q, c);
        quat_mul(e0, c, r);
        pf_log1<WIDE>(r, d0, slow);
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) {
            UKFB_UNROLL
            for (int k = 0; k <= i; ++k) Coo[tri(i, k)] = 21.0 * d0[i] * d0[k];
        }
        UKFB_UNROLL
        for (int i = 0; i < 30; ++i) Xc[i] = 0.0;
        UKFB_UNROLL
        for (int j = 0; j < 3; ++j) {
            double ep[4], en[4], rp[4], rn[4], dp[3], dn[3];
            pf_exp1<WIDE>(vp + 3 * j, 1.0, ep, slow);
            pf_exp1<WIDE>(vn + 3 * j, 1.0, en, slow);
            quat_mul(ep, c, rp);
            quat_mul(en, c, rn);
            pf_log1<WIDE>(rp, dp, slow);
            pf_log1<WIDE>(rn, dn, slow);
            UKFB_UNROLL
            for (int i = 0; i < 3; ++i) {
                UKFB_UNROLL
                for (int k = 0; k <= i; ++k) Coo[tri(i, k)] = fma(dp[i], dp[k], fma(dn[i], dn[k], Coo[tri(i, k)]));
            }
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) {
                const double dd = dp[k] - dn[k];
                UKFB_UNROLL
                for (int t = 0; t < 10; ++t) Xc[t * 3 + k] = fma(a[tri(3 + t, j)], dd, Xc[t * 3 + k]);
            }
        }
    }
    if (slow) return false; /* the caller hands mu and delta to the literal apply_delta; Sigma - K S K^T is in the record */
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        UKFB_UNROLL
        for (int k = 0; k <= i; ++k) sig[tri(i, k) * TILE] = 0.5 * Coo[tri(i, k)];
        UKFB_UNROLL
        for (int t = 0; t < 10; ++t) sig[tri(3 + t, i) * TILE] = 0.5 * Xc[t * 3 + i];
    }
    m.q[0] = ref_q[0], m.q[1] = ref_q[1], m.q[2] = ref_q[2], m.q[3] = ref_q[3];
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        m.v[i] += delta[3 + i];
        m.bg[i] += delta[6 + i];
        m.ba[i] += delta[9 + i];
    }
    m.g += delta[12];
    passes_out = passes;
    return true;
#undef OF_L
}

/* ---- everything the hot instances do not do, out of line (one call site per phase, as in ukf_pose_fast.cuh): the
 * any-angle instance of the same structured code first, the literal code of ukf_thread.cuh for what is beyond that too */
UKFB_DNI OfLit of_predict_slow(double* sm, int lane, double* sig, const double* Qp, ModelArgs ma, OriMu m, int q_diagonal)
{
    OfLit r;
    r.m = m, r.status = 0, r.passes = 0;
    double a[OriF::LP];
    UKFB_UNROLL
    for (int e = 0; e < OriF::LP; ++e) a[e] = sig[e * TILE];
    bool spd = true;
    if (of_predict<true>(q_diagonal, sm, lane, sig, a, Qp, ma, r.m, r.status, r.passes, spd)) {
        if (!spd) r.status |= UKFB_STATUS_NOT_SPD;
        UKFB_OF_COUNT(3);
        return r;
    }
    return of_literal_predict(sig, Qp, ma, m);
}

/* stage 0: nothing has been done (the hot instance refused the filter, or the caller's trace guard did); 1: the record
 * holds Sigma - K S K^T and `delta` = K innov (the hot instance left a polynomial range in its apply_delta) */
UKFB_DNI OfLit of_update_slow(double* sm, int lane, double* sig, int kind, const double* zm, const double* Rm, int r_ld, ModelArgs ma,
                              OriMu m, OfDelta delta, int stage, double gate_d2)
{
    if (stage == 0) {
        OfLit r;
        r.m = m, r.status = 0, r.passes = 0;
        double a[OriF::LP];
        UKFB_UNROLL
        for (int e = 0; e < OriF::LP; ++e) a[e] = sig[e * TILE];
        bool spd = true;
        if (of_update<true>(sm, lane, sig, a, zm, Rm, r_ld, r.m, delta.d, r.status, r.passes, spd, stage, gate_d2)) {
            if (!spd) r.status |= UKFB_STATUS_NOT_SPD;
            UKFB_OF_COUNT(4);
            return r;
        }
        /* stage now tells where the any-angle instance stopped */
    } else { /* stage 1: the hot instance left a polynomial range in its apply_delta; Sigma - K S K^T is in the record */
        OfLit r;
        r.m = m, r.status = 0, r.passes = 0;
        double a[OriF::LP];
        UKFB_UNROLL
        for (int e = 0; e < OriF::LP; ++e) a[e] = sig[e * TILE];
        bool spd = true;
        if (of_apply_delta<true>(sig, a, r.m, delta.d, r.status, r.passes, spd, !pf_unit(m.q))) {
            if (!spd) r.status |= UKFB_STATUS_NOT_SPD;
            UKFB_OF_COUNT(4);
            return r;
        }
    }
    return of_literal_update(sig, kind, zm, Rm, r_ld, ma, m, delta, stage == 0, gate_d2);
}

/* ---- the kernel: one warp per block, one filter per lane -------------------------------------------------------- */
/* PER_FILTER_PARAMS = false: time constants and earth rotation are launch constants (constant-bank operands, no
 * registers); true: each filter's own set is loaded (ukfb_set_orientation_params_per_filter).  Two instances because the
 * ten extra live registers cost the common case 3 %. */
template <bool PER_FILTER_PARAMS, bool OVERLAP = false> /* OVERLAP: as in ukf_pose_fast_kernel */
UKFB_GLOBAL void UKFB_LAUNCH_BOUNDS(4 * TILE, 2) ukf_ori_fast_kernel(const UKFB_GRID_CONSTANT StepParams p)
{
    typedef OriF F;
    UKFB_SMEM_DECL
    /* a block is 1..4 independent warps (no barrier between them: warps of one block merely start together, which keeps
     * their instruction fetches close); each warp owns one tile of 32 filters and its own slice of shared memory */
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    double* sm = ukfb_smem + wib * (OF_PER_LANE * TILE);
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

    OriMu m;
    UKFB_UNROLL
    for (int i = 0; i < 4; ++i) m.q[i] = rec[i * TILE];
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) m.v[i] = rec[(4 + i) * TILE], m.bg[i] = rec[(7 + i) * TILE], m.ba[i] = rec[(10 + i) * TILE];
    m.g = rec[13 * TILE];
    prefetch_next_wave<F>(p, tile, lane);

    ModelArgs ma;
    ma.dt = 0.0;
    ma.has_acc = false;
    ma.neg_inv_tau_g = p.neg_inv_tau_g;
    ma.neg_inv_tau_a = p.neg_inv_tau_a;
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        ma.earth[i] = p.earth[i];
        ma.acc[i] = p.acc_mu[bb * 3 + i];
        ma.omega[i] = p.gyro_mu[bb * 3 + i];
    }
    if (PER_FILTER_PARAMS) { /* this filter's own constructor arguments (OrientationUKF.cpp:41-47) */
        ma.neg_inv_tau_g = p.ori_params[bb * 5], ma.neg_inv_tau_a = p.ori_params[bb * 5 + 1];
        ma.earth[0] = p.ori_params[bb * 5 + 2], ma.earth[1] = p.ori_params[bb * 5 + 3], ma.earth[2] = p.ori_params[bb * 5 + 4];
    }
    const double big = 1.79769313486231570e308;
    uint32_t status = 0;
    bool dirty_mu = false;
    int hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    UKFB_NOUNROLL
    for (int tick = 0; tick < p.K; ++tick) {
        /* ---- control: time guards (UnscentedKalmanFilter.hpp:83-125), masks, finite checks */
        bool do_pred = false, do_upd = false;
        int kind = -1, store = -1;
        const double* Rm = p.R + tick * p.r_kstride + b * p.r_stride;
        if (valid) {
            bool idle = false;
            if (p.events) { /* one queued sample per filter and slot; UKFB_EVENT_IDLE: nothing happens */
                kind = int(p.kinds[tick * p.kinds_kstride + b]);
                idle = kind == UKFB_EVENT_IDLE;
                if (kind >= UKFB_EVENT_KIND_COUNT || kind < UKFB_EVENT_IDLE
                    || (kind >= 0 && (kind < UKFB_MEAS_ORI_VELOCITY || kind == UKFB_EVENT_POSE_ACCELERATION))) {
                    status |= UKFB_STATUS_BAD_EVENT;
                    idle = true;
                }
                if (idle) kind = -1;
                if (kind >= 0) Rm += kind * p.r_kind_stride;
                if (kind >= UKFB_EVENT_POSE_ACCELERATION) store = kind, kind = -1;
            }
            if (p.imu) { /* integrateMeasurement(RotationRate / Acceleration): check, store (OrientationUKF.cpp:53-63) */
                const double* s6 = p.imu + tick * p.imu_kstride + b * 6;
                const double g0 = s6[0], g1 = s6[1], g2 = s6[2], a0 = s6[3], a1 = s6[4], a2 = s6[5];
                if (fabs(g0) <= big && fabs(g1) <= big && fabs(g2) <= big)
                    ma.omega[0] = g0, ma.omega[1] = g1, ma.omega[2] = g2;
                else
                    status |= UKFB_STATUS_NONFINITE_MEAS;
                if (fabs(a0) <= big && fabs(a1) <= big && fabs(a2) <= big)
                    ma.acc[0] = a0, ma.acc[1] = a1, ma.acc[2] = a2;
                else
                    status |= UKFB_STATUS_NONFINITE_MEAS;
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
                if (kind >= 0) { /* checkMeasurment (OrientationUKF.cpp:67) */
                    bool ok = true;
                    const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
                    for (int r = 0; r < 3; ++r) ok = ok && (fabs(zm[r]) <= big);
                    for (int r = 0; r < 3; ++r)
                        for (int c = 0; c < 3; ++c) ok = ok && (fabs(Rm[r * p.r_ld + c]) <= big);
                    if (ok)
                        do_upd = true;
                    else {
                        status |= UKFB_STATUS_NONFINITE_MEAS;
                        kind = -1;
                    }
                }
            }
        }

        int passes_a = 0, passes_b = 0;
        const double* Qp = p.Q + b * p.q_stride;

        /* ---- predict (ukfom predict, App. A.3) ------------------------------------------------------------- */
        if (do_pred) {
            double a[F::LP];
            UKFB_UNROLL
            for (int e = 0; e < F::LP; ++e) a[e] = sig[e * TILE];
            bool spd = true;
            if (of_predict<false>(p.q_diagonal, sm, lane, sig, a, Qp, ma, m, status, passes_a, spd)) {
                if (!spd) {
                    status |= UKFB_STATUS_NOT_SPD;
                    do_upd = false; /* every later factorisation of this covariance fails too */
                } else
                    dirty_mu = true;
            } else { /* a polynomial range was left: nothing was modified; the any-angle instance, then the literal code */
                const OfLit r = of_predict_slow(sm, lane, sig, Qp, ma, m, p.q_diagonal);
                status |= r.status;
                passes_a = r.passes;
                if (r.status & UKFB_STATUS_NOT_SPD)
                    do_upd = false;
                else {
                    m = r.m;
                    dirty_mu = true;
                }
            }
        }

        /* ---- storing events: the sample is kept for the next predict (OrientationUKF.cpp:53-63) -------------- */
        if (store >= 0) {
            const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
            bool ok = true;
            for (int r = 0; r < 3; ++r) ok = ok && (fabs(zm[r]) <= big);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) ok = ok && (fabs(Rm[r * p.r_ld + c]) <= big);
            if (!ok)
                status |= UKFB_STATUS_NONFINITE_MEAS;
            else if (store == UKFB_EVENT_ORI_ROTATION_RATE)
                ma.omega[0] = zm[0], ma.omega[1] = zm[1], ma.omega[2] = zm[2];
            else
                ma.acc[0] = zm[0], ma.acc[1] = zm[1], ma.acc[2] = zm[2];
        }

        /* ---- update (ukfom update + apply_delta, App. A.4) --------------------------------------------------- */
        if (do_upd) {
            const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
            double a[F::LP];
            UKFB_UNROLL
            for (int e = 0; e < F::LP; ++e) a[e] = sig[e * TILE];
            const double tr = a[tri(0, 0)] + a[tri(1, 1)] + a[tri(2, 2)];
            bool literal = !(tr < PF_PI2_GUARD);
            bool spd = true, fast_done = false;
            int stage = 0;
            double delta[13];
            UKFB_UNROLL
            for (int i = 0; i < 13; ++i) delta[i] = 0.0;
            if (!literal) {
                fast_done = of_update<false>(sm, lane, sig, a, zm, Rm, p.r_ld, m, delta, status, passes_b, spd, stage, p.gate_d2);
                if (fast_done) {
                    if (!spd)
                        status |= UKFB_STATUS_NOT_SPD;
                    else
                        dirty_mu = true;
                }
            }
            if (!fast_done) {
                OfDelta dl;
                UKFB_UNROLL
                for (int i = 0; i < 13; ++i) dl.d[i] = delta[i];
                const OfLit r = of_update_slow(sm, lane, sig, kind, zm, Rm, p.r_ld, ma, m, dl, stage, p.gate_d2);
                status |= r.status;
                passes_b = r.passes;
                if (!(r.status & UKFB_STATUS_NOT_SPD)) {
                    m = r.m;
                    dirty_mu = true;
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
        for (int i = 0; i < 4; ++i) rec[i * TILE] = m.q[i];
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) rec[(4 + i) * TILE] = m.v[i], rec[(7 + i) * TILE] = m.bg[i], rec[(10 + i) * TILE] = m.ba[i];
        rec[13 * TILE] = m.g;
    }
    if (valid && (p.imu || p.events)) {
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) {
            p.acc_mu[b * 3 + i] = ma.acc[i];
            p.gyro_mu[b * 3 + i] = ma.omega[i];
        }
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

#undef UKFB_OS
#undef OF_L

} /* namespace ukfb */

#endif /* UKFB_ORI_FAST_CUH */
