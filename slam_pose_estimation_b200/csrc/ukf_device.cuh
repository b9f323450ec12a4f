/*
 * ukf_device.cuh -- device side of the batched unscented Kalman filter engine (sm_100a).
 *
 * What runs here is the arithmetic the reference delegates to ukfom::ukf<> / MTK
 *  plus the reference's own models and noise shaping:
 *   predictionStepImpl   PoseUKF.cpp:180-196, OrientationUKF.cpp:79-89
 *   process models       PoseUKF.cpp:75-97,   OrientationUKF.cpp:12-32
 *   measurement models   PoseUKF.cpp:7-69,    OrientationUKF.cpp:34-39
 *   time guards          UnscentedKalmanFilter.hpp:83-125
 *   ukf predict / update / apply_delta   (ukfom/ukf.hpp, App. A.2-A.4)
 *
 * Mapping: one warp owns a group of G filters.  Work that is
 * serial per filter (the Cholesky factorisations) runs LANE-PER-FILTER so the
 * sqrt / reciprocal chains of G filters share one warp instruction; work that is
 * parallel over the 2n+1 sigma points (boxplus, models, boxminus, the manifold
 * mean) runs LANE-PER-SIGMA-POINT, one filter at a time; the covariance
 * contractions run lane-per-2x2-tile of the lower triangle.  State and covariance
 * of the group live in shared memory between phases; HBM sees one coalesced read
 * and one coalesced write of each filter record per launch.
 */
#ifndef UKFB_DEVICE_CUH
#define UKFB_DEVICE_CUH

#include "so3.cuh"
#include "../../include/ukf_batch.h"

namespace ukfb {

/* ---- filter traits ---------------------------------------------------------- */

struct PoseF { /* PoseWithVelocity.hpp:18-23 */
    static constexpr int KIND = 0;
    static constexpr int N = UKFB_POSE_DOF;   /* tangent dimension */
    static constexpr int MU = UKFB_POSE_MU;   /* stored state size */
    static constexpr int NS = 2 * N + 1;      /* sigma points */
    static constexpr int LP = N * (N + 1) / 2;/* packed lower triangle */
    static constexpr int ROT = 3;             /* offset of the SO(3) block (tangent and mu) */
    static constexpr int REC = 96;            /* HBM record: mu padded to 16, then packed sigma, padded to 16 */
    static constexpr int DS = 15;             /* row stride of the deviation matrix (odd: conflict-free) */
    static constexpr int QB0 = 0, QB1 = 3;    /* rotated blocks of Q: position, orientation (PoseUKF.cpp:184-185) */
};

struct OriF { /* OrientationState.hpp:20-26 */
    static constexpr int KIND = 1;
    static constexpr int N = UKFB_ORI_DOF;
    static constexpr int MU = UKFB_ORI_MU;
    static constexpr int NS = 2 * N + 1;
    static constexpr int LP = N * (N + 1) / 2;
    static constexpr int ROT = 0;
    static constexpr int REC = 112;
    static constexpr int DS = 17;
    static constexpr int QB0 = 0, QB1 = 3;    /* orientation, velocity (OrientationUKF.cpp:84-85) */
};

constexpr int REC_MU_PAD = 16; /* sigma starts at this offset inside an HBM record */

template <class F>
UKFB_HD constexpr int mu_of(int t) { return t < F::ROT ? t : t + 1; } /* vector tangent index -> mu index */

/* packed lower-triangular index, i >= j */
UKFB_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

/* ---- shared-memory layout (doubles) ------------------------------------------- */
template <class F, int G>
struct Smem {
    /* per filter of the group */
    static constexpr int OFF_A = 0;                /* sigma, packed lower */
    static constexpr int OFF_B = F::LP;            /* Cholesky factor / updated sigma, packed lower */
    static constexpr int OFF_MU = 2 * F::LP;       /* mu */
    static constexpr int OFF_DELTA = OFF_MU + F::MU;
    static constexpr int OFF_IMU = OFF_DELTA + F::N; /* stored acceleration [0:3], rotation rate [3:6] */
    static constexpr int FREC_RAW = OFF_IMU + 6;
    static constexpr int FREC = FREC_RAW | 1;      /* odd stride: lane-per-filter accesses are conflict-free */
    /* per warp scratch */
    static constexpr int OFF_D = G * FREC;         /* deviations: 32 rows x DS ([dx | dz]) */
    static constexpr int OFF_SXZ = OFF_D + 32 * F::DS;
    static constexpr int OFF_KM = OFF_SXZ + 40;
    static constexpr int OFF_KS = OFF_KM + 40;
    static constexpr int OFF_SM = OFF_KS + 40;     /* S, 3x3 */
    static constexpr int OFF_MD = OFF_SM + 10;     /* mean delta broadcast */
    static constexpr int OFF_BC = OFF_MD + 16;     /* state broadcast */
    static constexpr int OFF_QR = OFF_BC + 16;     /* two rotated 3x3 blocks of Q */
    static constexpr int OFF_DT = OFF_QR + 18;     /* per filter dt */
    static constexpr int OFF_CTL = OFF_DT + G;     /* per filter: flags, kind (ints, 2 per double slot) */
    static constexpr int TOTAL_RAW = OFF_CTL + G + 4; /* + 8 ints: mean-pass histogram */
    static constexpr int TOTAL = (TOTAL_RAW + 1) & ~1;
};

/* control flags */
constexpr int CF_VALID = 1, CF_PRED = 2, CF_UPD = 4, CF_DIRTY = 8;

constexpr int HIST_SLOTS = 64;

struct StepParams {
    double* state;          /* B x REC */
    const double* Q;        /* packed lower, q_stride = 0 (broadcast) or LP */
    long long q_stride;
    long long B;
    uint32_t* status;       /* B */
    long long* t_last;      /* B, microseconds */
    unsigned long long* hist; /* HIST_SLOTS x 8 */
    /* predict */
    int do_predict;
    int time_mode;          /* 0: dt given, 1: sample timestamps given */
    const double* dt;
    long long dt_stride;
    const long long* ts;
    long long ts_stride;
    double min_dt, max_dt;
    double* acc_mu;         /* B x 3: POSE stored acceleration (NaN = none); ORIENTATION acceleration */
    const double* acc_cov;  /* B x 9: POSE only */
    double* gyro_mu;        /* B x 3: ORIENTATION only (both written back after an imu stream run) */
    double neg_inv_tau_g, neg_inv_tau_a; /* -1.0 / tau */
    double earth[3];
    /* update */
    int do_update;
    int kind;               /* uniform kind, or -2: per-filter kinds[] */
    const int8_t* kinds;
    const double* z;
    int z_stride;
    const double* R;
    long long r_stride;
    int r_ld;
    const uint8_t* mask;
    /* K consecutive ticks in one launch (state stays in shared memory between them):
     * element strides from tick k to tick k+1 of the per-tick streams */
    int K;
    long long dt_kstride, ts_kstride, z_kstride, r_kstride, kinds_kstride, mask_kstride;
    const int8_t* tick_kinds; /* K uniform kinds, one per tick (overrides `kind`), or null */
    const double* imu;        /* ORIENTATION: K x B x 6 (gyro xyz, acc xyz) stored before each predict, or null */
    long long imu_kstride;
};

UKFB_HD int meas_dim(int kind)
{
    switch (kind) {
        case 1: case 5: case 7: return 2;
        case 2: case 6: return 1;
        default: return 3;
    }
}

/* ---- manifold operations on a lane's state x[MU] --------------------------------- */

template <class F>
UKFB_D void state_boxplus(double* x, const double* d, double s)
{
    UKFB_UNROLL
    for (int t = 0; t < F::N; ++t) {
        if (t >= F::ROT && t < F::ROT + 3) continue;
        x[mu_of<F>(t)] += s * d[t];
    }
    so3_boxplus(x + F::ROT, d + F::ROT, s);
}

template <class F>
UKFB_D void state_boxminus(const double* x, const double* o, double* res)
{
    UKFB_UNROLL
    for (int t = 0; t < F::N; ++t) {
        if (t >= F::ROT && t < F::ROT + 3) continue;
        res[t] = x[mu_of<F>(t)] - o[mu_of<F>(t)];
    }
    so3_boxminus(x + F::ROT, o + F::ROT, res + F::ROT);
}

/* ---- lane-per-filter Cholesky, packed lower, src may alias dst ---------------------- */
/* LAPACK dpotf2('L') order: dot-product update, sqrt, scale by the reciprocal. */
template <class F>
UKFB_D bool cholesky_packed(const double* src, double* dst)
{
    UKFB_NOUNROLL
    for (int j = 0; j < F::N; ++j) {
        const int jj = tri(j, 0);
        double ajj = src[jj + j];
        for (int k = 0; k < j; ++k) ajj -= dst[jj + k] * dst[jj + k];
        if (!(ajj > 0.0) || !(ajj < 1.0e300)) return false;
        const double d = sqrt(ajj);
        dst[jj + j] = d;
        const double r = 1.0 / d;
        for (int i = j + 1; i < F::N; ++i) {
            const int ii = tri(i, 0);
            double s = src[ii + j];
            for (int k = 0; k < j; ++k) s -= dst[ii + k] * dst[jj + k];
            dst[ii + j] = s * r;
        }
    }
    return true;
}

/* ---- sigma points: X0 = mu + delta, X(2j+1) = mu + (delta + L[:,j]), X(2j+2) = mu + (delta - L[:,j]) */
template <class F>
UKFB_D void sigma_generate(const double* L, const double* mu, const double* delta, int lane, double* x)
{
    UKFB_UNROLL
    for (int i = 0; i < F::MU; ++i) x[i] = mu[i];
    const bool col = lane >= 1 && lane < F::NS;
    const int j = (lane - 1) >> 1;
    const bool plus = (lane & 1) != 0;
    double d[F::N];
    UKFB_UNROLL
    for (int i = 0; i < F::N; ++i) {
        const double l = (col && i >= j) ? L[tri(i, j)] : 0.0;
        const double dl = delta ? delta[i] : 0.0;
        d[i] = plus ? dl + l : dl - l;
    }
    state_boxplus<F>(x, d, 1.0);
}

/* ---- process models ------------------------------------------------------------- */

/* PoseUKF.cpp:75-83 / :88-97 */
UKFB_D void process_model_pose(double* x, double dt, bool has_acc, const double* acc)
{
    double* p = x;
    double* q = x + 3;
    double* v = x + 7;
    double* w = x + 10;
    if (has_acc) {
        v[0] += dt * acc[0];
        v[1] += dt * acc[1];
        v[2] += dt * acc[2];
    }
    double rv[3], rw[3];
    quat_rotate(q, v, rv);
    p[0] += dt * rv[0];
    p[1] += dt * rv[1];
    p[2] += dt * rv[2];
    quat_rotate(q, w, rw);
    so3_boxplus(q, rw, dt);
}

/* OrientationUKF.cpp:12-32 */
UKFB_D void process_model_ori(double* x, double dt, const double* acc, const double* omega, double neg_inv_tau_g,
                              double neg_inv_tau_a, const double* earth)
{
    double* q = x;
    double* v = x + 4;
    double* bg = x + 7;
    double* ba = x + 10;
    const double g = x[13];
    double wb[3] = {omega[0] - bg[0], omega[1] - bg[1], omega[2] - bg[2]};
    double av[3];
    quat_rotate(q, wb, av);
    av[0] -= earth[0];
    av[1] -= earth[1];
    av[2] -= earth[2];
    so3_boxplus(q, av, dt);

    double ab[3] = {acc[0] - ba[0], acc[1] - ba[1], acc[2] - ba[2]};
    double an[3];
    quat_rotate(q, ab, an);
    an[2] -= g;
    v[0] += dt * an[0];
    v[1] += dt * an[1];
    v[2] += dt * an[2];
    UKFB_UNROLL
    for (int i = 0; i < 3; ++i) {
        const double dg = neg_inv_tau_g * bg[i];
        bg[i] += dt * dg;
        const double da = neg_inv_tau_a * ba[i];
        ba[i] += dt * da;
    }
}

/* ---- measurement models (PoseUKF.cpp:7-69, OrientationUKF.cpp:34-39) --------------- */
/* z[0:3] (unused components 0) or a quaternion z[0:4] when kind == ORIENTATION. */
template <class F>
UKFB_D void measure(const double* x, int kind, double* z)
{
    z[0] = z[1] = z[2] = 0.0;
    z[3] = 1.0;
    if (F::KIND == 0) {
        switch (kind) {
            case 0: z[0] = x[0], z[1] = x[1], z[2] = x[2]; break;
            case 1: z[0] = x[0], z[1] = x[1]; break;
            case 2: z[0] = x[2]; break;
            case 3: z[0] = x[3], z[1] = x[4], z[2] = x[5], z[3] = x[6]; break;
            case 4: z[0] = x[7], z[1] = x[8], z[2] = x[9]; break;
            case 5: z[0] = x[7], z[1] = x[8]; break;
            case 6: z[0] = x[9]; break;
            case 7: z[0] = x[7], z[1] = x[12]; break;
            case 8: z[0] = x[10], z[1] = x[11], z[2] = x[12]; break;
            default: break;
        }
    } else {
        quat_inv_rotate(x, x + 4, z);
    }
}

UKFB_D void meas_boxminus(const double* z, const double* o, bool rot, double* res)
{
    if (rot) {
        so3_boxminus(z, o, res);
    } else {
        res[0] = z[0] - o[0];
        res[1] = z[1] - o[1];
        res[2] = z[2] - o[2];
    }
}

/* ---- warp context ----------------------------------------------------------------- */
template <class F>
struct Warp {
    double* D;
    double* SXZ;
    double* KM;
    double* KS;
    double* SM;
    double* MD;
    double* BC;
    double* QR;
    int* HP; /* mean-pass histogram of this warp, 8 ints */
    int lane;
};

/* ukfom sigma_points_mean on the state manifold: ref = X0; loop { md = mean(X_i [-] ref);
 * ref [+]= md } while (|md| > tol && ++i < max_it).  Every lane ends with the same ref. */
template <class F>
UKFB_D uint32_t manifold_mean(Warp<F>& w, const double* x, double* ref)
{
    const int lane = w.lane;
    if (lane == 0) {
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) w.BC[i] = x[i];
    }
    __syncwarp();
    UKFB_UNROLL
    for (int i = 0; i < F::MU; ++i) ref[i] = w.BC[i];
    uint32_t st = 0;
    int it = 0, passes = 0;
    while (true) {
        double d[F::N];
        state_boxminus<F>(x, ref, d);
        if (lane < F::NS) {
            UKFB_UNROLL
            for (int i = 0; i < F::N; ++i) w.D[lane * F::DS + i] = d[i];
        }
        __syncwarp();
        if (lane < F::N) {
            double a = 0.0;
            for (int p = 0; p < F::NS; ++p) a += w.D[p * F::DS + lane];
            w.MD[lane] = a / double(F::NS);
        }
        __syncwarp();
        double md[F::N];
        double n2 = 0.0;
        UKFB_UNROLL
        for (int i = 0; i < F::N; ++i) {
            md[i] = w.MD[i];
            n2 += md[i] * md[i];
        }
        state_boxplus<F>(ref, md, 1.0);
        ++passes;
        __syncwarp();
        if (!(sqrt(n2) > UKFB_MEAN_TOL)) break;
        if (++it >= UKFB_MEAN_MAX_IT) {
            st = UKFB_STATUS_MEAN_NO_CONVERGE;
            break;
        }
    }
    if (lane == 0) w.HP[passes < 7 ? passes : 7]++;
    return st;
}

/* covariance of the deviations in D[:, 0:N] (already written, synced): each lane owns one
 * 2x2 tile of the lower triangle;  out(i,j) = 0.5 * sum_p d_i d_j + noise(i,j). */
template <class F, class Noise>
UKFB_D void cov_store(Warp<F>& w, double* out, Noise noise)
{
    constexpr int NT1 = (F::N + 1) / 2;
    constexpr int NT = NT1 * (NT1 + 1) / 2;
    static_assert(NT <= 32, "one tile per lane");
    const int lane = w.lane;
    if (lane < NT) {
        int ta = 0;
        while ((ta + 1) * (ta + 2) / 2 <= lane) ++ta;
        const int tb = lane - ta * (ta + 1) / 2;
        const int i0 = 2 * ta, j0 = 2 * tb;
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
        for (int p = 0; p < F::NS; ++p) {
            const double* row = w.D + p * F::DS;
            const double a0 = row[i0], a1 = row[i0 + 1];
            const double b0 = row[j0], b1 = row[j0 + 1];
            c00 += a0 * b0;
            c01 += a0 * b1;
            c10 += a1 * b0;
            c11 += a1 * b1;
        }
        out[tri(i0, j0)] = 0.5 * c00 + noise(i0, j0);
        if (j0 + 1 <= i0) out[tri(i0, j0 + 1)] = 0.5 * c01 + noise(i0, j0 + 1);
        if (i0 + 1 < F::N) {
            out[tri(i0 + 1, j0)] = 0.5 * c10 + noise(i0 + 1, j0);
            out[tri(i0 + 1, j0 + 1)] = 0.5 * c11 + noise(i0 + 1, j0 + 1);
        }
    }
}

/* deviations of every sigma point from `ref` into D[:, 0:N] */
template <class F>
UKFB_D void write_deviations(Warp<F>& w, const double* x, const double* ref)
{
    double d[F::N];
    state_boxminus<F>(x, ref, d);
    if (w.lane < F::NS) {
        UKFB_UNROLL
        for (int i = 0; i < F::N; ++i) w.D[w.lane * F::DS + i] = d[i];
    }
}

/* symmetric lookup into packed-lower Q */
UKFB_D double q_sym(const double* Qp, int i, int j) { return i >= j ? UKFB_LDG(Qp + tri(i, j)) : UKFB_LDG(Qp + tri(j, i)); }

/* ---- predict of one filter by one warp (ukfom predict, App. A.3) --------------------- */
template <class F>
UKFB_D uint32_t sigma_predict(Warp<F>& w, const StepParams& p, long long b, double* A, double* Bs, double* mu,
                              const double* fimu, double dt)
{
    const int lane = w.lane;
    const double* Qp = p.Q + b * p.q_stride;

    /* acceleration branch of PoseUKF.cpp:188-193 */
    bool has_acc = false;
    const double acc[3] = {fimu[0], fimu[1], fimu[2]};
    const double omega[3] = {fimu[3], fimu[4], fimu[5]};
    if (F::KIND == 0)
        has_acc = (fabs(acc[0]) <= 1.79769313486231570e308) && (fabs(acc[1]) <= 1.79769313486231570e308) &&
                  (fabs(acc[2]) <= 1.79769313486231570e308);

    /* rotated blocks of Q: rot * Q[blk] * rot^T (PoseUKF.cpp:184-185, OrientationUKF.cpp:84-85) */
    if (!has_acc && lane < 18) {
        double Rm[9];
        quat_matrix(mu + F::ROT, Rm);
        const int off = lane < 9 ? F::QB0 : F::QB1;
        const int e = lane < 9 ? lane : lane - 9;
        const int r = e / 3, c = e % 3;
        double acc_rc = 0.0;
        UKFB_UNROLL
        for (int k = 0; k < 3; ++k) {
            double t = 0.0;
            UKFB_UNROLL
            for (int l = 0; l < 3; ++l) t += Rm[r * 3 + l] * q_sym(Qp, off + l, off + k);
            acc_rc += t * Rm[c * 3 + k];
        }
        w.QR[lane] = acc_rc;
    }

    double x[F::MU];
    sigma_generate<F>(Bs, mu, nullptr, lane, x);
    if (F::KIND == 0)
        process_model_pose(x, dt, has_acc, acc);
    else
        process_model_ori(x, dt, acc, omega, p.neg_inv_tau_g, p.neg_inv_tau_a, p.earth);

    double ref[F::MU];
    uint32_t st = manifold_mean<F>(w, x, ref);
    write_deviations<F>(w, x, ref);
    __syncwarp();

    const double scale = F::KIND == 0 ? dt : dt * dt; /* PoseUKF.cpp:186 vs OrientationUKF.cpp:86 */
    const double* QR = w.QR;
    const double* acov = p.acc_cov ? p.acc_cov + b * 9 : nullptr;
    cov_store<F>(w, A, [=](int i, int j) -> double {
        if (F::KIND == 0 && has_acc) {
            /* shadowing local of PoseUKF.cpp:190-191: unrotated, unscaled Q, velocity block = 2 acc.cov */
            if (i >= 6 && i < 9 && j >= 6 && j < 9) return 2.0 * UKFB_LDG(acov + (i - 6) * 3 + (j - 6));
            return q_sym(Qp, i, j);
        }
        if (i >= F::QB0 && i < F::QB0 + 3 && j >= F::QB0 && j < F::QB0 + 3)
            return scale * QR[(i - F::QB0) * 3 + (j - F::QB0)];
        if (i >= F::QB1 && i < F::QB1 + 3 && j >= F::QB1 && j < F::QB1 + 3)
            return scale * QR[9 + (i - F::QB1) * 3 + (j - F::QB1)];
        return scale * q_sym(Qp, i, j);
    });
    if (lane == 0) {
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) mu[i] = ref[i];
    }
    __syncwarp();
    return st;
}

/* ---- apply_delta of one filter (App. A.4): sigma points around mu [+] delta from the
 * factor in Bs, manifold mean, covariance (no additive noise) ------------------------- */
template <class F>
UKFB_D uint32_t sigma_apply_delta(Warp<F>& w, double* A, const double* Bs, double* mu, const double* delta)
{
    double x[F::MU];
    sigma_generate<F>(Bs, mu, delta, w.lane, x);
    double ref[F::MU];
    uint32_t st = manifold_mean<F>(w, x, ref);
    write_deviations<F>(w, x, ref);
    __syncwarp();
    cov_store<F>(w, A, [](int, int) -> double { return 0.0; });
    if (w.lane == 0) {
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) mu[i] = ref[i];
    }
    __syncwarp();
    return st;
}

/* ---- first half of update of one filter (App. A.4): innovation statistics, gain,
 * sigma <- sigma - K S K^T (in place), delta = K innov ----------------------- */
template <class F>
UKFB_D uint32_t sigma_update(Warp<F>& w, const StepParams& p, long long b, int tick, int kind, double* A,
                             const double* Bs,
                             const double* mu, double* delta)
{
    const int lane = w.lane;
    const bool rot = (F::KIND == 0) && kind == UKFB_MEAS_POSE_ORIENTATION;
    const int m = meas_dim(kind);
    uint32_t st = 0;

    double x[F::MU];
    sigma_generate<F>(Bs, mu, nullptr, lane, x);
    double z[4];
    measure<F>(x, kind, z);

    /* mean of Z */
    double zref[4];
    if (lane == 0) {
        w.BC[0] = z[0], w.BC[1] = z[1], w.BC[2] = z[2], w.BC[3] = z[3];
    }
    __syncwarp();
    zref[0] = w.BC[0], zref[1] = w.BC[1], zref[2] = w.BC[2], zref[3] = w.BC[3];
#if UKFB_EUCLID_MEAS_DIRECT_MEAN
    if (!rot) {
        if (lane < F::NS) {
            w.D[lane * F::DS + F::N + 0] = z[0];
            w.D[lane * F::DS + F::N + 1] = z[1];
            w.D[lane * F::DS + F::N + 2] = z[2];
        }
        __syncwarp();
        if (lane < 3) {
            double a = 0.0;
            for (int q = 0; q < F::NS; ++q) a += w.D[q * F::DS + F::N + lane];
            w.MD[lane] = a / double(F::NS);
        }
        __syncwarp();
        zref[0] = w.MD[0], zref[1] = w.MD[1], zref[2] = w.MD[2];
        __syncwarp();
    } else
#endif
    {
        int it = 0;
        while (true) {
            double dz[3];
            meas_boxminus(z, zref, rot, dz);
            if (lane < F::NS) {
                w.D[lane * F::DS + F::N + 0] = dz[0];
                w.D[lane * F::DS + F::N + 1] = dz[1];
                w.D[lane * F::DS + F::N + 2] = dz[2];
            }
            __syncwarp();
            if (lane < 3) {
                double a = 0.0;
                for (int q = 0; q < F::NS; ++q) a += w.D[q * F::DS + F::N + lane];
                w.MD[lane] = a / double(F::NS);
            }
            __syncwarp();
            const double md[3] = {w.MD[0], w.MD[1], w.MD[2]};
            const double n2 = md[0] * md[0] + md[1] * md[1] + md[2] * md[2];
            if (rot) {
                so3_boxplus(zref, md, 1.0);
            } else {
                zref[0] += md[0];
                zref[1] += md[1];
                zref[2] += md[2];
            }
            __syncwarp();
            if (!(sqrt(n2) > UKFB_MEAN_TOL)) break;
            if (++it >= UKFB_MEAN_MAX_IT) {
                st = UKFB_STATUS_MEAN_NO_CONVERGE;
                break;
            }
        }
    }

    /* deviations: dz = Z_i [-] zbar, dx = X_i [-] mu (the PRIOR mu, App. A.4) */
    {
        double dz[3];
        meas_boxminus(z, zref, rot, dz);
        double mur[F::MU];
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) mur[i] = mu[i];
        double dx[F::N];
        state_boxminus<F>(x, mur, dx);
        if (lane < F::NS) {
            UKFB_UNROLL
            for (int i = 0; i < F::N; ++i) w.D[lane * F::DS + i] = dx[i];
            w.D[lane * F::DS + F::N + 0] = dz[0];
            w.D[lane * F::DS + F::N + 1] = dz[1];
            w.D[lane * F::DS + F::N + 2] = dz[2];
        }
    }
    __syncwarp();

    /* S = 0.5 sum dz dz^T + R (R padded with identity to 3x3);  Sxz = 0.5 sum dx dz^T */
    const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
    const double* Rm = p.R + tick * p.r_kstride + b * p.r_stride;
    if (lane < 9) {
        const int a = lane / 3, c = lane % 3;
        double s = 0.0;
        for (int q = 0; q < F::NS; ++q) s += w.D[q * F::DS + F::N + a] * w.D[q * F::DS + F::N + c];
        const double r = (a < m && c < m) ? UKFB_LDG(Rm + a * p.r_ld + c) : (a == c ? 1.0 : 0.0);
        w.SM[lane] = 0.5 * s + r;
    }
    for (int e = lane; e < 3 * F::N; e += 32) {
        const int i = e / 3, c = e % 3;
        double s = 0.0;
        for (int q = 0; q < F::NS; ++q) s += w.D[q * F::DS + i] * w.D[q * F::DS + F::N + c];
        w.SXZ[e] = 0.5 * s;
    }
    __syncwarp();

    /* S^-1 by cofactors (Eigen fixed-size inverse), every lane redundantly */
    double S[9], Si[9];
    UKFB_UNROLL
    for (int i = 0; i < 9; ++i) S[i] = w.SM[i];
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
            const double v[3] = {UKFB_LDG(zm + 0), UKFB_LDG(zm + 1), UKFB_LDG(zm + 2)};
            so3_exp(v, 1.0, zin); /* PoseUKF.cpp:135 */
        } else {
            UKFB_UNROLL
            for (int c = 0; c < 3; ++c) zin[c] = c < m ? UKFB_LDG(zm + c) : 0.0;
        }
        meas_boxminus(zin, zref, rot, innov);
    }
    /* K = Sxz S^-1, KS = K S */
    for (int e = lane; e < 3 * F::N; e += 32) {
        const int i = e / 3, c = e % 3;
        double k3[3];
        UKFB_UNROLL
        for (int cc = 0; cc < 3; ++cc) {
            double s = 0.0;
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) s += w.SXZ[i * 3 + k] * Si[k * 3 + cc];
            k3[cc] = s;
        }
        double ks = 0.0;
        UKFB_UNROLL
        for (int k = 0; k < 3; ++k) ks += k3[k] * S[k * 3 + c];
        w.KM[e] = k3[c];
        w.KS[e] = ks;
    }
    __syncwarp();
    /* sigma' = sigma - (K S) K^T, lower triangle, in place (the factor in Bs is no longer needed) */
    for (int e = lane; e < F::LP; e += 32) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        double s = 0.0;
        UKFB_UNROLL
        for (int k = 0; k < 3; ++k) s += w.KS[i * 3 + k] * w.KM[j * 3 + k];
        A[e] = A[e] - s;
    }
    if (lane < F::N) {
        double s = 0.0;
        UKFB_UNROLL
        for (int k = 0; k < 3; ++k) s += w.KM[lane * 3 + k] * innov[k];
        delta[lane] = s;
    }
    __syncwarp();
    return st;
}

/* ---- the kernel ------------------------------------------------------------------- */
constexpr int WPB = 4; /* warps per block; warps are independent (no block-level sync) */

template <class F, int G>
UKFB_GLOBAL void UKFB_LAUNCH_BOUNDS(WPB * 32, 1) ukf_step_kernel(const StepParams p)
{
    typedef Smem<F, G> SM;
    UKFB_SMEM_DECL
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long first = ((long long)blockIdx.x * WPB + warp) * G;
    if (first >= p.B) return;
    const int cnt = (p.B - first) < G ? int(p.B - first) : G;

    double* wsm = ukfb_smem + warp * SM::TOTAL;
    Warp<F> w;
    w.D = wsm + SM::OFF_D;
    w.SXZ = wsm + SM::OFF_SXZ;
    w.KM = wsm + SM::OFF_KM;
    w.KS = wsm + SM::OFF_KS;
    w.SM = wsm + SM::OFF_SM;
    w.MD = wsm + SM::OFF_MD;
    w.BC = wsm + SM::OFF_BC;
    w.QR = wsm + SM::OFF_QR;
    w.lane = lane;
    double* cdt = wsm + SM::OFF_DT;
    int* cflag = reinterpret_cast<int*>(wsm + SM::OFF_CTL);
    int* ckind = cflag + G;
    w.HP = ckind + G;
    if (lane < 8) w.HP[lane] = 0;

    /* ---- load the group's records (coalesced), scatter into the per-filter layout */
    {
        const double* src = p.state + first * F::REC;
        for (int i = lane; i < cnt * F::REC; i += 32) {
            const int g = i / F::REC, k = i - g * F::REC;
            const double v = src[i];
            double* fr = wsm + g * SM::FREC;
            if (k < F::MU)
                fr[SM::OFF_MU + k] = v;
            else if (k >= REC_MU_PAD && k < REC_MU_PAD + F::LP)
                fr[SM::OFF_A + (k - REC_MU_PAD)] = v;
        }
        /* the stored IMU sample: acceleration (PoseUKF.cpp:175-178, OrientationUKF.cpp:59-63)
         * and rotation rate (OrientationUKF.cpp:53-57) */
        for (int i = lane; i < cnt * 6; i += 32) {
            const int g = i / 6, k = i - g * 6;
            double v = 0.0;
            if (k < 3)
                v = p.acc_mu[(first + g) * 3 + k];
            else if (F::KIND == 1)
                v = p.gyro_mu[(first + g) * 3 + (k - 3)];
            wsm[g * SM::FREC + SM::OFF_IMU + k] = v;
        }
    }
    uint32_t my_status = 0; /* lane g: status bits of filter g */
    bool dirty = false;     /* lane g: record of filter g changed */
    __syncwarp();

    UKFB_NOUNROLL
    for (int tick = 0; tick < p.K; ++tick) {
        /* ---- per-filter control: time guards (UnscentedKalmanFilter.hpp:83-125), masks, checks */
        if (lane < G) {
            int flags = 0, kind = -1;
            if (lane < cnt) {
                const long long b = first + lane;
                flags = CF_VALID;
                if (F::KIND == 1 && p.imu) { /* integrateMeasurement(RotationRate / Acceleration): check, store */
                    const double* s6 = p.imu + tick * p.imu_kstride + b * 6;
                    double* fimu = wsm + lane * SM::FREC + SM::OFF_IMU;
                    const double g0 = s6[0], g1 = s6[1], g2 = s6[2], a0 = s6[3], a1 = s6[4], a2 = s6[5];
                    const double big = 1.79769313486231570e308;
                    if (fabs(g0) <= big && fabs(g1) <= big && fabs(g2) <= big)
                        fimu[3] = g0, fimu[4] = g1, fimu[5] = g2;
                    else
                        my_status |= UKFB_STATUS_NONFINITE_MEAS;
                    if (fabs(a0) <= big && fabs(a1) <= big && fabs(a2) <= big)
                        fimu[0] = a0, fimu[1] = a1, fimu[2] = a2;
                    else
                        my_status |= UKFB_STATUS_NONFINITE_MEAS;
                }
                if (p.do_predict) {
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
                        if (dt < 0.0)
                            my_status |= UKFB_STATUS_NEG_DT;
                        else if (dt <= p.min_dt) {
                            /* delta time is zero or close to zero: no-op */
                        } else if (dt > p.max_dt)
                            my_status |= UKFB_STATUS_DT_TOO_LARGE;
                        else {
                            flags |= CF_PRED;
                            cdt[lane] = dt;
                        }
                    }
                }
                if (p.do_update) {
                    kind = p.tick_kinds ? int(p.tick_kinds[tick])
                                        : (p.kind == -2 ? int(p.kinds[tick * p.kinds_kstride + b]) : p.kind);
                    if (p.mask && !p.mask[tick * p.mask_kstride + b]) kind = -1;
                    if (kind >= 0) {
                        bool ok = true;
                        if (F::KIND == 1) { /* checkMeasurment, OrientationUKF.cpp:67 */
                            const int m = meas_dim(kind);
                            const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
                            const double* Rm = p.R + tick * p.r_kstride + b * p.r_stride;
                            for (int a = 0; a < m; ++a) ok = ok && (fabs(zm[a]) <= 1.79769313486231570e308);
                            for (int a = 0; a < m; ++a)
                                for (int c = 0; c < m; ++c) ok = ok && (fabs(Rm[a * p.r_ld + c]) <= 1.79769313486231570e308);
                        }
                        if (ok)
                            flags |= CF_UPD;
                        else {
                            my_status |= UKFB_STATUS_NONFINITE_MEAS;
                            kind = -1;
                        }
                    }
                }
            }
            cflag[lane] = flags;
            ckind[lane] = kind;
        }
        __syncwarp();

        /* ---- predict ------------------------------------------------------------------ */
        if (p.do_predict) {
            if (lane < G && (cflag[lane] & CF_PRED)) {
                double* fr = wsm + lane * SM::FREC;
                if (!cholesky_packed<F>(fr + SM::OFF_A, fr + SM::OFF_B)) {
                    my_status |= UKFB_STATUS_NOT_SPD;
                    cflag[lane] &= ~(CF_PRED | CF_UPD);
                }
            }
            __syncwarp();
            for (int g = 0; g < cnt; ++g) {
                if (!(cflag[g] & CF_PRED)) continue;
                double* fr = wsm + g * SM::FREC;
                const uint32_t st = sigma_predict<F>(w, p, first + g, fr + SM::OFF_A, fr + SM::OFF_B, fr + SM::OFF_MU,
                                                     fr + SM::OFF_IMU, cdt[g]);
                if (lane == g) {
                    my_status |= st;
                    dirty = true;
                }
            }
            __syncwarp();
        }

        /* ---- update --------------------------------------------------------------------- */
        if (p.do_update) {
            if (lane < G && (cflag[lane] & CF_UPD)) {
                double* fr = wsm + lane * SM::FREC;
                if (!cholesky_packed<F>(fr + SM::OFF_A, fr + SM::OFF_B)) {
                    my_status |= UKFB_STATUS_NOT_SPD;
                    cflag[lane] &= ~CF_UPD;
                }
            }
            __syncwarp();
            for (int g = 0; g < cnt; ++g) {
                if (!(cflag[g] & CF_UPD)) continue;
                double* fr = wsm + g * SM::FREC;
                const uint32_t st = sigma_update<F>(w, p, first + g, tick, ckind[g], fr + SM::OFF_A, fr + SM::OFF_B,
                                                    fr + SM::OFF_MU, fr + SM::OFF_DELTA);
                if (lane == g) my_status |= st;
            }
            __syncwarp();
            if (lane < G && (cflag[lane] & CF_UPD)) {
                double* fr = wsm + lane * SM::FREC;
                if (!cholesky_packed<F>(fr + SM::OFF_A, fr + SM::OFF_B)) {
                    /* the reference has already replaced sigma by sigma - K S K^T when MTK's
                     * assert fires inside apply_delta; keep that matrix, leave mu alone */
                    my_status |= UKFB_STATUS_NOT_SPD;
                    cflag[lane] &= ~CF_UPD;
                    dirty = true;
                }
            }
            __syncwarp();
            for (int g = 0; g < cnt; ++g) {
                if (!(cflag[g] & CF_UPD)) continue;
                double* fr = wsm + g * SM::FREC;
                const uint32_t st =
                    sigma_apply_delta<F>(w, fr + SM::OFF_A, fr + SM::OFF_B, fr + SM::OFF_MU, fr + SM::OFF_DELTA);
                if (lane == g) {
                    my_status |= st;
                    dirty = true;
                }
            }
            __syncwarp();
        }
    }

    /* ---- store dirty records (coalesced), stored IMU sample, status, histogram --------------- */
    if (lane < G) cflag[lane] = dirty ? CF_DIRTY : 0;
    __syncwarp();
    {
        double* dst = p.state + first * F::REC;
        for (int i = lane; i < cnt * F::REC; i += 32) {
            const int g = i / F::REC, k = i - g * F::REC;
            if (!(cflag[g] & CF_DIRTY)) continue;
            const double* fr = wsm + g * SM::FREC;
            if (k < F::MU)
                dst[i] = fr[SM::OFF_MU + k];
            else if (k >= REC_MU_PAD && k < REC_MU_PAD + F::LP)
                dst[i] = fr[SM::OFF_A + (k - REC_MU_PAD)];
        }
        if (F::KIND == 1 && p.imu) {
            for (int i = lane; i < cnt * 6; i += 32) {
                const int g = i / 6, k = i - g * 6;
                const double v = wsm[g * SM::FREC + SM::OFF_IMU + k];
                if (k < 3)
                    p.acc_mu[(first + g) * 3 + k] = v;
                else
                    p.gyro_mu[(first + g) * 3 + (k - 3)] = v;
            }
        }
    }
    if (lane < cnt && my_status) p.status[first + lane] |= my_status;
    __syncwarp();
    if (lane >= 1 && lane < 8 && p.hist && w.HP[lane])
        atomicAdd(p.hist + (blockIdx.x % HIST_SLOTS) * 8 + lane, (unsigned long long)w.HP[lane]);
}

} /* namespace ukfb */

#endif /* UKFB_DEVICE_CUH */
