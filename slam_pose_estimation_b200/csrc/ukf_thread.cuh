/*
 * ukf_thread.cuh -- the lane-per-filter step kernel of the batched UKF engine (sm_100a).
 *
 * Same arithmetic and same StepParams as ukf_device.cuh (the warp-per-group kernel), different mapping:
 * EVERY LANE OWNS ONE FILTER and runs the whole predict / update / apply_delta sequence for it as scalar
 * FP64 code, looping over the 2n+1 sigma points.  All 32 lanes do useful work in every instruction, nothing
 * is computed redundantly across lanes, there are no shuffles, no warp barriers and no padded tensor tiles,
 * so the FP64 pipe sees close to the algorithmic operation count of the filter.  What makes it fit:
 *   - the covariance accumulators (78 doubles of the lower triangle, rows 0..11) live in REGISTERS across the
 *     sigma-point loop (the loop body is compiled once, with static accumulator indices);
 *   - the Cholesky factor, the prior mean, the running reference mean and K*innovation live in SHARED memory,
 *     entry-major ([entry][lane]: conflict free), 928 B per PoseUKF filter;
 *   - the covariance itself lives in the filter's HBM record for the whole launch (L2 resident): the Cholesky
 *     reads it from there and the new covariance is accumulated in registers and written back, so a failed
 *     factorisation leaves it untouched exactly like the reference's early return;
 *   - sigma points are regenerated for the second sweep (deviations from the converged mean) instead of being
 *     stored: 25 x 13 doubles per filter would not fit on chip.
 * HBM layout: tiles of 32 filters, entry-major inside a tile (record entry e of filter b at
 * (b / 32) * 32 * REC + e * 32 + b % 32), so every access of a warp to "entry e of my filter" is one
 * coalesced 256-byte request.
 *
 * Reference sites: UnscentedKalmanFilter.hpp:83-125 (guards), PoseUKF.cpp:7-196, OrientationUKF.cpp:12-89,
 * ukfom::ukf predict / update / apply_delta.
 */
#ifndef UKFB_THREAD_CUH
#define UKFB_THREAD_CUH

#include "ukf_device.cuh"

namespace ukfb {

constexpr int TILE = 32; /* filters per HBM tile = lanes per warp = threads per block of this kernel */

/* element index of record entry e of filter b in the tile-interleaved layout */
UKFB_HD long long tile_index(long long b, int e, int REC) { return (b / TILE) * (long long)(TILE * REC) + (long long)e * TILE + (b % TILE); }

template <class F>
struct TSmem { /* per-lane doubles, stored entry-major: entry e of lane l at e * TILE + l */
    static constexpr int OFF_L = 0;                  /* Cholesky factor, packed lower */
    static constexpr int OFF_MU = F::LP;             /* prior mean of the current pass */
    static constexpr int OFF_REF = OFF_MU + F::MU;   /* running reference / new mean */
    static constexpr int OFF_DELTA = OFF_REF + F::MU;/* K * innovation */
    static constexpr int PER_LANE = OFF_DELTA + F::N;
    static constexpr int TOTAL = PER_LANE * TILE;
};

/* ST: element stride between consecutive entries of one lane: TILE in shared memory ([entry][lane]); the fast PoseUKF
 * kernel runs these functions as its cold fallback on a per-thread local array with ST = 1, lane = 0 */
#define UKFB_TS(e) sm[(e) * ST + lane]

/* ---- Cholesky of the covariance in the HBM record into shared memory ------------------------------------ */
/* LAPACK dpotf2('L') order.  sig: this lane's covariance in its record (entry stride TILE).  Out of line: one
 * copy of the unrolled factorisation serves the three call sites. */
template <class F, int ST = TILE>
UKFB_DNI bool cholesky_thread(const double* sig, double* sm, int lane)
{
    typedef TSmem<F> TS;
    bool ok = true;
    UKFB_UNROLL
    for (int j = 0; j < F::N; ++j) {
        double pj[F::N];
        double ajj = sig[tri(j, j) * TILE];
        UKFB_UNROLL
        for (int k = 0; k < j; ++k) {
            pj[k] = UKFB_TS(TS::OFF_L + tri(j, k));
            ajj -= pj[k] * pj[k];
        }
        if (!(ajj > 0.0) || !(ajj < 1.0e300)) {
            ok = false;
            ajj = 1.0;
        }
        double d, rinv;
        fast_sqrt_rsqrt(ajj, d, rinv);
        UKFB_TS(TS::OFF_L + tri(j, j)) = d;
        UKFB_UNROLL
        for (int i = j + 1; i < F::N; ++i) {
            double s = sig[tri(i, j) * TILE];
            UKFB_UNROLL
            for (int k = 0; k < j; ++k) s -= UKFB_TS(TS::OFF_L + tri(i, k)) * pj[k];
            UKFB_TS(TS::OFF_L + tri(i, j)) = s * rinv;
        }
    }
    return ok;
}

/* ---- sigma point p of this lane's filter: X0 = mu [+] delta, X(2j+1) = mu [+] (delta + L[:,j]),
 * X(2j+2) = mu [+] (delta - L[:,j]).  p is uniform over the warp. */
template <class F, bool WITH_DELTA, int ST = TILE>
UKFB_D void sigma_point(const double* sm, int lane, int p, double* x)
{
    typedef TSmem<F> TS;
    const int j = p > 0 ? (p - 1) >> 1 : 0;
    const double sgn = p > 0 ? ((p & 1) ? 1.0 : -1.0) : 0.0;
    double d[F::N];
    UKFB_UNROLL
    for (int i = 0; i < F::N; ++i) {
        const double l = (i >= j) ? UKFB_TS(TS::OFF_L + tri(i, 0) + j) : 0.0; /* tri(i,0) + j stays inside the packed array */
        d[i] = sgn * l;
        if (WITH_DELTA) d[i] += UKFB_TS(TS::OFF_DELTA + i);
    }
    UKFB_UNROLL
    for (int i = 0; i < F::MU; ++i) x[i] = UKFB_TS(TS::OFF_MU + i);
    state_boxplus<F>(x, d, 1.0);
}

struct ModelArgs {
    double dt;
    bool has_acc;
    double acc[3], omega[3];
    double neg_inv_tau_g, neg_inv_tau_a;
    double earth[3];
};

template <class F>
UKFB_D void apply_model(double* x, const ModelArgs& a)
{
    if (F::KIND == 0)
        process_model_pose(x, a.dt, a.has_acc, a.acc);
    else
        process_model_ori(x, a.dt, a.acc, a.omega, a.neg_inv_tau_g, a.neg_inv_tau_a, a.earth);
}

/* ---- ukfom sigma_points_mean + sigma_points_cov over regenerated sigma points ------------------------------
 * PREDICT = true : X_p = g(mu [+] +-L[:,j])        (ukfom predict, App. A.3); out = 0.5 C + noise already stored in `sig`
 * PREDICT = false: X_p = mu [+] (delta +- L[:,j])  (apply_delta, App. A.4);   out = 0.5 C
 * Leaves the new mean in the REF slots and writes the new covariance to `sig` (HBM record).  Returns the number of
 * mean passes through *passes. */
template <class F, bool PREDICT, int ST = TILE>
UKFB_D uint32_t mean_and_cov(double* sm, int lane, double* sig, const ModelArgs& ma, int* passes_out)
{
    typedef TSmem<F> TS;
    uint32_t st = 0;
    int it = 0, passes = 0;
    /* mean: ref = X0; loop { md = mean(X_p [-] ref); ref [+]= md } while (|md| > tol && ++it < max_it) */
    while (true) {
        double md[F::N];
        UKFB_UNROLL
        for (int i = 0; i < F::N; ++i) md[i] = 0.0;
        UKFB_NOUNROLL
        for (int p = 0; p < F::NS; ++p) {
            double x[F::MU];
            sigma_point<F, !PREDICT, ST>(sm, lane, p, x);
            if (PREDICT) apply_model<F>(x, ma);
            if (p == 0 && passes == 0) {
                UKFB_UNROLL
                for (int i = 0; i < F::MU; ++i) UKFB_TS(TS::OFF_REF + i) = x[i];
            }
            double ref[F::MU], d[F::N];
            UKFB_UNROLL
            for (int i = 0; i < F::MU; ++i) ref[i] = UKFB_TS(TS::OFF_REF + i);
            state_boxminus<F>(x, ref, d);
            UKFB_UNROLL
            for (int i = 0; i < F::N; ++i) md[i] += d[i];
        }
        double n2 = 0.0;
        UKFB_UNROLL
        for (int i = 0; i < F::N; ++i) {
            md[i] = div_ns<F::NS>(md[i]);
            n2 += md[i] * md[i];
        }
        double ref[F::MU];
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) ref[i] = UKFB_TS(TS::OFF_REF + i);
        state_boxplus<F>(ref, md, 1.0);
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) UKFB_TS(TS::OFF_REF + i) = ref[i];
        ++passes;
        if (!(n2 > UKFB_MEAN_TOL * UKFB_MEAN_TOL)) break;
        if (++it >= UKFB_MEAN_MAX_IT) {
            st = UKFB_STATUS_MEAN_NO_CONVERGE;
            break;
        }
    }
    *passes_out = passes;

    /* covariance: C = sum_p d d^T with d = X_p [-] mean; accumulators in registers, lower triangle */
    double C[F::LP];
    UKFB_UNROLL
    for (int e = 0; e < F::LP; ++e) C[e] = 0.0;
    UKFB_NOUNROLL
    for (int p = 0; p < F::NS; ++p) {
        double d[F::N];
        {
            double x[F::MU];
            sigma_point<F, !PREDICT, ST>(sm, lane, p, x);
            if (PREDICT) apply_model<F>(x, ma);
            double ref[F::MU];
            UKFB_UNROLL
            for (int i = 0; i < F::MU; ++i) ref[i] = UKFB_TS(TS::OFF_REF + i);
            state_boxminus<F>(x, ref, d);
        }
        UKFB_UNROLL
        for (int i = 0; i < F::N; ++i) {
            UKFB_UNROLL
            for (int j = 0; j <= i; ++j) C[tri(i, j)] = fma(d[i], d[j], C[tri(i, j)]);
        }
    }
    UKFB_UNROLL
    for (int e = 0; e < F::LP; ++e) {
        if (PREDICT)
            sig[e * TILE] = fma(0.5, C[e], sig[e * TILE]); /* the noise was stored there after the factorisation */
        else
            sig[e * TILE] = 0.5 * C[e];
    }
    return st;
}

/* ---- this step's process noise, packed lower, into the (now free) covariance slots of the record ------------
 *   no acceleration: scale * Q with the two rotated blocks rot * Q[blk] * rot^T, rot from the PRIOR orientation;
 *     scale = dt (PoseUKF.cpp:182-186) or dt^2 (OrientationUKF.cpp:81-86);
 *   acceleration (the shadowing local of PoseUKF.cpp:190-191): Q unrotated and unscaled, velocity block = 2 acc.cov */
template <class F, int ST = TILE>
UKFB_D void store_noise(const double* sm, int lane, double* sig, const double* Qp, const double* acov, const ModelArgs& ma)
{
    typedef TSmem<F> TS;
    const double scale = ma.has_acc ? 1.0 : (F::KIND == 0 ? ma.dt : ma.dt * ma.dt);
    UKFB_UNROLL
    for (int e = 0; e < F::LP; ++e) sig[e * TILE] = scale * UKFB_LDG(Qp + e);
    if (!ma.has_acc) {
        double q[4], Rm[9];
        UKFB_UNROLL
        for (int i = 0; i < 4; ++i) q[i] = UKFB_TS(TS::OFF_MU + F::ROT + i);
        quat_matrix(q, Rm);
        UKFB_UNROLL
        for (int blk = 0; blk < 2; ++blk) {
            const int off = blk == 0 ? F::QB0 : F::QB1;
            double t[9]; /* t = rot * Q[blk] */
            UKFB_UNROLL
            for (int r = 0; r < 3; ++r) {
                UKFB_UNROLL
                for (int k = 0; k < 3; ++k) {
                    double s = 0.0;
                    UKFB_UNROLL
                    for (int l = 0; l < 3; ++l) s += Rm[r * 3 + l] * q_sym(Qp, off + l, off + k);
                    t[r * 3 + k] = s;
                }
            }
            UKFB_UNROLL
            for (int r = 0; r < 3; ++r) {
                UKFB_UNROLL
                for (int c = 0; c <= r; ++c) {
                    double s = 0.0;
                    UKFB_UNROLL
                    for (int k = 0; k < 3; ++k) s += t[r * 3 + k] * Rm[c * 3 + k];
                    sig[tri(off + r, off + c) * TILE] = scale * s;
                }
            }
        }
    } else if (F::KIND == 0) {
        UKFB_UNROLL
        for (int r = 0; r < 3; ++r) {
            UKFB_UNROLL
            for (int c = 0; c <= r; ++c) sig[tri(6 + r, 6 + c) * TILE] = 2.0 * acov[r * 3 + c];
        }
    }
}

/* ---- first half of ukfom update (App. A.4): innovation statistics, gain, sigma <- sigma - K S K^T, delta = K innov */
template <class F, int ST = TILE>
UKFB_D uint32_t update_first_half(double* sm, int lane, double* sig, int kind, const double* zm, const double* Rm, int r_ld,
                                  double gate_d2)
{
    typedef TSmem<F> TS;
    const bool rot = (F::KIND == 0) && kind == UKFB_MEAS_POSE_ORIENTATION;
    const int m = meas_dim(kind);
    uint32_t st = 0;

    /* mean of Z_p = h(X_p) (ukfom sigma_points_mean on the measurement space) */
    double zref[4];
    {
        double x[F::MU];
        sigma_point<F, false, ST>(sm, lane, 0, x);
        measure<F>(x, kind, zref);
    }
    {
        int it = 0;
        while (true) {
            double md[3] = {0.0, 0.0, 0.0};
            UKFB_NOUNROLL
            for (int p = 0; p < F::NS; ++p) {
                double x[F::MU], z[4], dz[3];
                sigma_point<F, false, ST>(sm, lane, p, x);
                measure<F>(x, kind, z);
#if UKFB_EUCLID_MEAS_DIRECT_MEAN
                if (!rot)
                    dz[0] = z[0], dz[1] = z[1], dz[2] = z[2];
                else
#endif
                    meas_boxminus(z, zref, rot, dz);
                md[0] += dz[0], md[1] += dz[1], md[2] += dz[2];
            }
            UKFB_UNROLL
            for (int c = 0; c < 3; ++c) md[c] = div_ns<F::NS>(md[c]);
#if UKFB_EUCLID_MEAS_DIRECT_MEAN
            if (!rot) {
                zref[0] = md[0], zref[1] = md[1], zref[2] = md[2];
                break;
            }
#endif
            const double n2 = md[0] * md[0] + md[1] * md[1] + md[2] * md[2];
            if (rot) {
                so3_boxplus(zref, md, 1.0);
            } else {
                zref[0] += md[0];
                zref[1] += md[1];
                zref[2] += md[2];
            }
            if (!(n2 > UKFB_MEAN_TOL * UKFB_MEAN_TOL)) break;
            if (++it >= UKFB_MEAN_MAX_IT) {
                st = UKFB_STATUS_MEAN_NO_CONVERGE;
                break;
            }
        }
    }

    /* S = 0.5 sum dz dz^T + R,  Sxz = 0.5 sum dx dz^T with dx = X_p [-] mu (the PRIOR mu) */
    double Sxz[F::N * 3], S[9];
    UKFB_UNROLL
    for (int i = 0; i < F::N * 3; ++i) Sxz[i] = 0.0;
    UKFB_UNROLL
    for (int i = 0; i < 9; ++i) S[i] = 0.0;
    UKFB_NOUNROLL
    for (int p = 0; p < F::NS; ++p) {
        double x[F::MU], z[4], dz[3], dx[F::N], mu[F::MU];
        sigma_point<F, false, ST>(sm, lane, p, x);
        measure<F>(x, kind, z);
        meas_boxminus(z, zref, rot, dz);
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) mu[i] = UKFB_TS(TS::OFF_MU + i);
        state_boxminus<F>(x, mu, dx);
        UKFB_UNROLL
        for (int i = 0; i < F::N; ++i) {
            UKFB_UNROLL
            for (int c = 0; c < 3; ++c) Sxz[i * 3 + c] = fma(dx[i], dz[c], Sxz[i * 3 + c]);
        }
        UKFB_UNROLL
        for (int a = 0; a < 3; ++a) {
            UKFB_UNROLL
            for (int c = 0; c < 3; ++c) S[a * 3 + c] = fma(dz[a], dz[c], S[a * 3 + c]);
        }
    }
    UKFB_UNROLL
    for (int i = 0; i < F::N * 3; ++i) Sxz[i] *= 0.5;
    UKFB_UNROLL
    for (int a = 0; a < 3; ++a) {
        UKFB_UNROLL
        for (int c = 0; c < 3; ++c) {
            const double rr = (a < m && c < m) ? Rm[a * r_ld + c] : (a == c ? 1.0 : 0.0); /* R padded with identity */
            S[a * 3 + c] = fma(0.5, S[a * 3 + c], rr);
        }
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
    /* innovation z [-] zbar */
    double innov[3];
    {
        double zin[4] = {0.0, 0.0, 0.0, 1.0};
        if (rot) {
            const double v[3] = {zm[0], zm[1], zm[2]};
            so3_exp(v, 1.0, zin); /* PoseUKF.cpp:135 */
        } else {
            UKFB_UNROLL
            for (int c = 0; c < 3; ++c) zin[c] = c < m ? zm[c] : 0.0;
        }
        meas_boxminus(zin, zref, rot, innov);
    }
    /* the accept functor: squared Mahalanobis distance against the gate (never taken with the reference's accept_any) */
    {
        double d2 = 0.0;
        UKFB_UNROLL
        for (int a = 0; a < 3; ++a) d2 += innov[a] * (Si[a * 3] * innov[0] + Si[a * 3 + 1] * innov[1] + Si[a * 3 + 2] * innov[2]);
        if (d2 > gate_d2) return st | UKFB_STATUS_MEAS_REJECTED;
    }
    /* K = Sxz S^-1 (in place of Sxz), KS = K S */
    double KS[F::N * 3];
    UKFB_UNROLL
    for (int i = 0; i < F::N; ++i) {
        double k3[3];
        UKFB_UNROLL
        for (int cc = 0; cc < 3; ++cc) {
            double s = 0.0;
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) s += Sxz[i * 3 + k] * Si[k * 3 + cc];
            k3[cc] = s;
        }
        double delta = 0.0;
        UKFB_UNROLL
        for (int c = 0; c < 3; ++c) {
            double ks = 0.0;
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) ks += k3[k] * S[k * 3 + c];
            KS[i * 3 + c] = ks;
            Sxz[i * 3 + c] = k3[c];
            delta += k3[c] * innov[c];
        }
        UKFB_TS(TS::OFF_DELTA + i) = delta;
    }
    /* sigma <- sigma - (K S) K^T, lower triangle, in the HBM record */
    UKFB_UNROLL
    for (int i = 0; i < F::N; ++i) {
        UKFB_UNROLL
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) s += KS[i * 3 + k] * Sxz[j * 3 + k];
            sig[tri(i, j) * TILE] -= s;
        }
    }
    return st;
}

/* ---- the kernel: one warp per block, one filter per lane ---------------------------------------------------- */
template <class F>
UKFB_GLOBAL void UKFB_LAUNCH_BOUNDS(TILE, 1) ukf_thread_kernel(const UKFB_GRID_CONSTANT StepParams p)
{
    typedef TSmem<F> TS;
    constexpr int ST = TILE;
    UKFB_SMEM_DECL
    double* sm = ukfb_smem;
    const int lane = threadIdx.x;
    const long long b = (long long)blockIdx.x * TILE + lane;
    const bool valid = b < p.B;
    const long long bb = valid ? b : p.B - 1; /* lanes past the end shadow the last filter and never store */
    double* rec = p.state + (long long)blockIdx.x * (TILE * F::REC) + lane; /* entry e at rec[e * TILE] */
    double* sig = rec + F::MU * TILE;

    UKFB_UNROLL
    for (int i = 0; i < F::MU; ++i) UKFB_TS(TS::OFF_MU + i) = rec[i * TILE];

    ModelArgs ma;
    ma.dt = 0.0;
    ma.has_acc = false;
    ma.neg_inv_tau_g = p.neg_inv_tau_g;
    ma.neg_inv_tau_a = p.neg_inv_tau_a;
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        ma.earth[i] = p.earth[i];
        ma.acc[i] = p.acc_mu[bb * 3 + i];
        ma.omega[i] = F::KIND == 1 ? p.gyro_mu[bb * 3 + i] : 0.0;
    }
    if (F::KIND == 1 && p.ori_params) { /* this filter's own constructor arguments (OrientationUKF.cpp:41-47) */
        ma.neg_inv_tau_g = p.ori_params[bb * 5], ma.neg_inv_tau_a = p.ori_params[bb * 5 + 1];
        ma.earth[0] = p.ori_params[bb * 5 + 2], ma.earth[1] = p.ori_params[bb * 5 + 3], ma.earth[2] = p.ori_params[bb * 5 + 4];
    }
    const double big = 1.79769313486231570e308;
    uint32_t status = 0;
    bool dirty_mu = false;
    int hist[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* hist[k]: state means that took k passes (7 = 7 or more) */

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
                    || (kind >= 0 && (F::KIND == 0 ? (kind == UKFB_MEAS_ORI_VELOCITY || kind > UKFB_EVENT_POSE_ACCELERATION)
                                                   : (kind < UKFB_MEAS_ORI_VELOCITY || kind == UKFB_EVENT_POSE_ACCELERATION)))) {
                    status |= UKFB_STATUS_BAD_EVENT;
                    idle = true;
                }
                if (idle) kind = -1;
                if (kind >= 0) Rm += kind * p.r_kind_stride;
                if (kind >= UKFB_EVENT_POSE_ACCELERATION) store = kind, kind = -1;
            }
            if (F::KIND == 1 && p.imu) { /* integrateMeasurement(RotationRate / Acceleration): check, store */
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
                if (kind >= 0) {
                    bool ok = true;
                    if (F::KIND == 1) { /* checkMeasurment: OrientationUKF only (OrientationUKF.cpp:67) */
                        const int m = meas_dim(kind);
                        const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
                        for (int a = 0; a < m; ++a) ok = ok && (fabs(zm[a]) <= big);
                        for (int a = 0; a < m; ++a)
                            for (int c = 0; c < m; ++c) ok = ok && (fabs(Rm[a * p.r_ld + c]) <= big);
                    }
                    if (ok)
                        do_upd = true;
                    else {
                        status |= UKFB_STATUS_NONFINITE_MEAS;
                        kind = -1;
                    }
                }
            }
        }

        /* ---- predict (ukfom predict, App. A.3) ------------------------------------------------------------- */
        int passes_a = 0, passes_b = 0;
        if (do_pred) {
            if (!cholesky_thread<F>(sig, sm, lane)) {
                status |= UKFB_STATUS_NOT_SPD;
                do_upd = false; /* every later factorisation of this covariance fails too */
            } else {
                if (F::KIND == 0)
                    ma.has_acc = (fabs(ma.acc[0]) <= big) && (fabs(ma.acc[1]) <= big) && (fabs(ma.acc[2]) <= big);
                store_noise<F>(sm, lane, sig, p.Q + b * p.q_stride, p.acc_cov ? p.acc_cov + b * 9 : nullptr, ma);
                status |= mean_and_cov<F, true>(sm, lane, sig, ma, &passes_a);
                UKFB_UNROLL
                for (int i = 0; i < F::MU; ++i) UKFB_TS(TS::OFF_MU + i) = UKFB_TS(TS::OFF_REF + i);
                dirty_mu = true;
            }
        }

        /* ---- storing events: the sample is kept for the next predict (after this slot's own predict) ----------------- */
        if (store >= 0) {
            const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
            if (F::KIND == 0) { /* PoseUKF.cpp:175-178: no finite check (NaN mu = "no acceleration") */
                UKFB_UNROLL
                for (int r = 0; r < 3; ++r) {
                    ma.acc[r] = zm[r];
                    UKFB_UNROLL
                    for (int c = 0; c < 3; ++c) p.acc_cov[b * 9 + r * 3 + c] = Rm[r * p.r_ld + c];
                }
            } else { /* OrientationUKF.cpp:53-63: checkMeasurment(mu, cov), then store */
                bool ok = true;
                for (int a = 0; a < 3; ++a) ok = ok && (fabs(zm[a]) <= big);
                for (int a = 0; a < 3; ++a)
                    for (int c = 0; c < 3; ++c) ok = ok && (fabs(Rm[a * p.r_ld + c]) <= big);
                if (!ok)
                    status |= UKFB_STATUS_NONFINITE_MEAS;
                else if (store == UKFB_EVENT_ORI_ROTATION_RATE)
                    ma.omega[0] = zm[0], ma.omega[1] = zm[1], ma.omega[2] = zm[2];
                else
                    ma.acc[0] = zm[0], ma.acc[1] = zm[1], ma.acc[2] = zm[2];
            }
        }

        /* ---- update (ukfom update + apply_delta, App. A.4) --------------------------------------------------- */
        if (do_upd) {
            if (!cholesky_thread<F>(sig, sm, lane)) {
                status |= UKFB_STATUS_NOT_SPD;
            } else {
                const uint32_t ust = update_first_half<F>(sm, lane, sig, kind, p.z + tick * p.z_kstride + b * p.z_stride, Rm, p.r_ld, p.gate_d2);
                status |= ust;
                /* the reference has already replaced sigma by sigma - K S K^T when MTK's assert fires inside
                 * apply_delta: on failure that matrix stays in the record, mu is left alone */
                if (ust & UKFB_STATUS_MEAS_REJECTED) {
                    /* gated out: nothing was modified */
                } else if (!cholesky_thread<F>(sig, sm, lane)) {
                    status |= UKFB_STATUS_NOT_SPD;
                } else {
                    status |= mean_and_cov<F, false>(sm, lane, sig, ma, &passes_b);
                    UKFB_UNROLL
                    for (int i = 0; i < F::MU; ++i) UKFB_TS(TS::OFF_MU + i) = UKFB_TS(TS::OFF_REF + i);
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

    /* ---- write back: the covariance is already in the record ------------------------------------------------------ */
    if (valid && dirty_mu) {
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) rec[i * TILE] = UKFB_TS(TS::OFF_MU + i);
    }
    if (valid && ((F::KIND == 1 && p.imu) || p.events)) {
        UKFB_UNROLL
        for (int i = 0; i < 3; ++i) {
            p.acc_mu[b * 3 + i] = ma.acc[i];
            if (F::KIND == 1) p.gyro_mu[b * 3 + i] = ma.omega[i];
        }
    }
    if (valid && status) p.status[b] |= status;
    if (p.hist && valid) {
        unsigned long long* hs = p.hist + (blockIdx.x % HIST_SLOTS) * 8;
        UKFB_UNROLL
        for (int k = 1; k < 8; ++k)
            if (hist[k]) atomicAdd(hs + k, (unsigned long long)hist[k]);
    }
}

#undef UKFB_TS

} /* namespace ukfb */

#endif /* UKFB_THREAD_CUH */
