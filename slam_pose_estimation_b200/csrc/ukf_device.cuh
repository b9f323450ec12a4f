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
 * Mapping.  One warp owns a group of G filters whose records stay in
 * shared memory for the whole launch (K ticks).  Three thread mappings alternate:
 *   - T = 32/G LANES PER FILTER for the Cholesky factorisations: each lane keeps its rows
 *     of the factor in registers, the pivot row travels by warp shuffle, nothing touches
 *     shared memory until the factor is complete (so a failed factorisation leaves the
 *     covariance intact, as the reference's early return does);
 *   - ONE LANE PER SIGMA POINT (25 / 27 of 32 lanes), one filter at a time, for boxplus,
 *     the process / measurement models, boxminus and the manifold mean;
 *   - the FP64 TENSOR-CORE path (mma.sync m8n8k4, SASS DMMA) for the contractions over the
 *     sigma points: the deviations are written once as a [component][point] matrix and the
 *     covariance, cross-covariance and innovation covariance come out of 14-21 DMMA tiles
 *     instead of ~100 shared-memory-bound FMA rounds.
 * HBM sees one coalesced read and one coalesced write of each filter record per launch
 * (plus one L2-resident spill of the predicted covariance inside an update).
 */
#ifndef UKFB_DEVICE_CUH
#define UKFB_DEVICE_CUH

#include "so3.cuh"
#include "../../include/ukf_batch.h"

namespace ukfb {

/* ---- filter traits ---------------------------------------------------------- */

struct PoseF { /* PoseWithVelocity.hpp:18-23 */
    static constexpr int KIND = 0;
    static constexpr int N = UKFB_POSE_DOF;    /* tangent dimension */
    static constexpr int MU = UKFB_POSE_MU;    /* stored state size */
    static constexpr int NS = 2 * N + 1;       /* sigma points */
    static constexpr int LP = N * (N + 1) / 2; /* packed lower triangle */
    static constexpr int ROT = 3;              /* offset of the SO(3) block (tangent and mu) */
    static constexpr int REC = MU + LP;        /* HBM record: mu, then packed sigma (91 doubles, odd) */
    static constexpr int QB0 = 0, QB1 = 3;     /* rotated blocks of Q: position, orientation (PoseUKF.cpp:184-185) */
};

struct OriF { /* OrientationState.hpp:20-26 */
    static constexpr int KIND = 1;
    static constexpr int N = UKFB_ORI_DOF;
    static constexpr int MU = UKFB_ORI_MU;
    static constexpr int NS = 2 * N + 1;
    static constexpr int LP = N * (N + 1) / 2;
    static constexpr int ROT = 0;
    static constexpr int REC = MU + LP;        /* 105 doubles, odd */
    static constexpr int QB0 = 0, QB1 = 3;     /* orientation, velocity (OrientationUKF.cpp:84-85) */
};

template <class F>
UKFB_HD constexpr int mu_of(int t) { return t < F::ROT ? t : t + 1; } /* vector tangent index -> mu index */

/* packed lower-triangular index, i >= j */
UKFB_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

/* ---- shared-memory layout (doubles) ------------------------------------------- */
constexpr int DT_ROWS = 16; /* deviation matrix: tangent components (<= 13) then measurement components (3) */
constexpr int DT_COLS = 28; /* sigma points padded to a multiple of the DMMA k = 4 */
constexpr int DT_LD = 36;   /* = 4 mod 16: the DMMA fragment loads hit 16 distinct 8-byte banks */

template <class F, int G>
struct Smem {
    /* per filter of the group; [mu | sigma] is a straight copy of the HBM record */
    static constexpr int OFF_MU = 0;
    static constexpr int OFF_SIG = F::MU;               /* sigma, packed lower -- or its Cholesky factor */
    static constexpr int OFF_DELTA = OFF_SIG + F::LP;   /* K * innovation */
    static constexpr int OFF_IMU = OFF_DELTA + F::N;    /* stored acceleration [0:3], rotation rate [3:6] */
    static constexpr int OFF_DT = OFF_IMU + 6;          /* this tick's delta time */
    static constexpr int OFF_Z = OFF_DT + 1;            /* this tick's measurement, zero padded to 3 */
    static constexpr int OFF_R = OFF_Z + 3;             /* its covariance, identity padded to 3x3 */
    static constexpr int FS = (OFF_R + 9) | 1;          /* odd stride: lane-per-filter accesses are conflict-free */
    /* per warp scratch */
    static constexpr int OFF_D = G * FS + ((G * FS) & 1); /* deviation matrix [DT_ROWS][DT_LD] */
    static constexpr int OFF_SXZ = OFF_D + DT_ROWS * DT_LD;
    static constexpr int OFF_KM = OFF_SXZ + 40;
    static constexpr int OFF_KS = OFF_KM + 40;
    static constexpr int OFF_SM = OFF_KS + 40;     /* S, 3x3 */
    static constexpr int OFF_MD = OFF_SM + 10;     /* mean delta broadcast */
    static constexpr int OFF_BC = OFF_MD + 16;     /* state broadcast */
    static constexpr int OFF_NT = OFF_BC + 16;     /* this predict's process noise, packed lower */
    static constexpr int OFF_CTL = OFF_NT + F::LP + (F::LP & 1); /* ints: flags[G], kind[G], status[G], mean-pass histogram[8], store[G] */
    static constexpr int TOTAL_RAW = OFF_CTL + (4 * G + 8 + 1) / 2;
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
    double* acc_cov;        /* B x 9: POSE only (written by UKFB_EVENT_POSE_ACCELERATION events) */
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
    /* event streams (ukfb_run_events): kinds[tick][b] is one UKFB_EVENT_* / UKFB_MEAS_* code per filter and slot, ts its
     * sample time; an idle slot touches nothing; kinds >= 10 only store their sample after the predict.  The covariance
     * of an event is at R + tick * r_kstride + b * r_stride + kind * r_kind_stride (per-sensor table or per event). */
    int events;
    long long r_kind_stride;
    /* the accept functor of ukfom::ukf::update: a measurement with innov^T S^-1 innov > gate_d2 is not integrated
     * (+inf = accept_any_mahalanobis_distance, the reference's choice at PoseUKF.cpp:116) */
    double gate_d2;
    /* ORIENTATION, one parameter set per filter: B x 5 (-1/tau_g, -1/tau_a, earth rotation xyz), or null */
    const double* ori_params;
    /* fast kernels: a starting warp asks L2 for the record of the tile `prefetch_tiles` ahead of its own -- the tile a
     * warp of the next wave will start on (0 = off; set by the launcher to the number of resident warps) -- one request
     * per `prefetch_bytes` of the record */
    long long prefetch_tiles;
    int prefetch_bytes;
    /* fast kernels: 1 = the broadcast Q is diagonal (the reference's own default, PoseUKF.cpp:103-107, and the usual
     * configuration): only its diagonal is loaded; 2 = moreover entries 0..2 are equal and entries 3..5 are equal, so the
     * two blocks the filters rotate into the navigation frame (R Q_blk R^T) are multiples of the identity and stay so */
    int q_diagonal;
    /* fast kernels: launches of one handle may overlap (the next launch starts in the slots this one's last wave leaves
     * empty: simt.cuh, pdl_launch_dependents).  tile_done[t] counts the lanes that have finished tile t, 32 per launch;
     * the warp that owns tile t in launch number `launch_seq` (1, 2, ...) first waits for 32 (launch_seq - 1).  Null: off. */
    unsigned long long* tile_done;
    unsigned long long launch_seq;
};

UKFB_HD int meas_dim(int kind)
{
    switch (kind) {
        case 1: case 5: case 7: return 2;
        case 2: case 6: return 1;
        default: return 3;
    }
}

/* does filter class F (F::KIND) have an integrateMeasurement overload that calls ukf->update for `kind`? */
template <class F>
UKFB_HD bool meas_kind_of_class(int kind)
{
    return F::KIND == 0 ? (kind >= 0 && kind <= UKFB_MEAS_POSE_ANGULAR_VELOCITY) : kind == UKFB_MEAS_ORI_VELOCITY;
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

/* ---- Cholesky, T lanes per filter, rows in registers -------------------------------------- */
/* LAPACK dpotf2('L') order per column: dot-product update of the pivot, sqrt, update of the
 * rows below scaled by the reciprocal of the pivot.  Lane t of a filter's T lanes owns rows
 * i = r*T + t.  `sig` (packed lower, shared memory) is read at entry and overwritten by the
 * factor only when every pivot was positive and finite; all 32 lanes must call (shuffles). */
template <class F, int T>
UKFB_D bool cholesky_rows(double* sig, int lane, bool active)
{
    constexpr int N = F::N, R = (N + T - 1) / T;
    const int t = lane & (T - 1);
    double a[R][N];
    UKFB_UNROLL
    for (int r = 0; r < R; ++r) {
        const int i = r * T + t;
        UKFB_UNROLL
        for (int k = 0; k < N; ++k) {
            if (k > r * T + T - 1) continue; /* beyond the longest row of this register row */
            const bool in = i < N && k <= i;
            a[r][k] = in ? sig[tri(in ? i : 0, in ? k : 0)] : 0.0;
        }
    }
    bool ok = true;
    UKFB_UNROLL
    for (int j = 0; j < N; ++j) {
        const int rj = j / T, src = (lane & ~(T - 1)) | (j % T);
        double pj[N];
        double ajj = warp_shfl(a[rj][j], src);
        UKFB_UNROLL
        for (int k = 0; k < j; ++k) {
            pj[k] = warp_shfl(a[rj][k], src);
            ajj -= pj[k] * pj[k];
        }
        if (!(ajj > 0.0) || !(ajj < 1.0e300)) {
            ok = false;
            ajj = 1.0; /* keep the arithmetic finite; nothing is written back */
        }
        double d, rinv;
        fast_sqrt_rsqrt(ajj, d, rinv);
        UKFB_UNROLL
        for (int r = 0; r < R; ++r) {
            if (r * T + T - 1 < j) continue; /* all rows of this register row are above the pivot */
            const int i = r * T + t;
            double s = a[r][j];
            if (r * T + T - 1 > j) { /* some lane's row is below the pivot */
                UKFB_UNROLL
                for (int k = 0; k < j; ++k) s -= a[r][k] * pj[k];
                s *= rinv;
            }
            a[r][j] = (i == j) ? d : s;
        }
    }
    if (ok && active) {
        UKFB_UNROLL
        for (int r = 0; r < R; ++r) {
            const int i = r * T + t;
            UKFB_UNROLL
            for (int k = 0; k < N; ++k) {
                if (k > r * T + T - 1) continue;
                if (i < N && k <= i) sig[tri(i, k)] = a[r][k];
            }
        }
    }
    return ok;
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
struct Warp {
    double* D;   /* deviation matrix D[c * DT_LD + p]: component c of sigma point p */
    double* SXZ;
    double* KM;
    double* KS;
    double* SM;
    double* MD;
    double* BC;
    double* NT;
    int* HP; /* mean-pass histogram of this warp, 8 ints */
    int lane;
};

/* Row sums of the deviation matrix over the sigma points: rows 0..15 by lanes (c = lane & 15), each row split
 * between the two half-warps (columns 0..15 and 16..27) and combined by one shuffle.  Within a half-warp lane c
 * visits the columns of each aligned group of four in the order s ^ (c >> 2), which puts the 16 lanes on 16
 * distinct 8-byte banks (row stride 36 = 4 mod 16).  The padding columns NS..27 hold zeros. */
UKFB_D double row_sum(const double* D, int lane)
{
    const int c = lane & 15, h = lane >> 4, k = c >> 2;
    const double* base = D + c * DT_LD + 16 * h;
    const double* b0 = base + (0 ^ k);
    const double* b1 = base + (1 ^ k);
    const double* b2 = base + (2 ^ k);
    const double* b3 = base + (3 ^ k);
    double a0 = 0.0, a1 = 0.0;
    UKFB_UNROLL
    for (int g4 = 0; g4 < 3; ++g4) {
        a0 += b0[4 * g4];
        a1 += b1[4 * g4];
        a0 += b2[4 * g4];
        a1 += b3[4 * g4];
    }
    if (h == 0) { /* the first half-warp has one more group: columns 12..15 */
        a0 += b0[12];
        a1 += b1[12];
        a0 += b2[12];
        a1 += b3[12];
    }
    const double a = a0 + a1;
    return a + warp_shfl(a, lane ^ 16);
}

/* x / NS as the reference's `mean_delta /= X.size()` computes it (a true division), through
 * one reciprocal multiply and a residual correction */
template <int NS>
UKFB_D double div_ns(double x)
{
    constexpr double rinv = 1.0 / double(NS);
    const double q = x * rinv;
    return fma(fma(-double(NS), q, x), rinv, q);
}

/* DMMA fragment of component tile I (rows 8I..8I+7 of D) and sigma points 4s..4s+3: serves as the
 * A operand (component x point) of tile I and as the B operand (point x component) of tile I. */
UKFB_D double frag(const double* D, int lane, int I, int s) { return D[(8 * I + (lane >> 2)) * DT_LD + 4 * s + (lane & 3)]; }

/* symmetric lookup into packed-lower Q */
UKFB_D double q_sym(const double* Qp, int i, int j) { return i >= j ? UKFB_LDG(Qp + tri(i, j)) : UKFB_LDG(Qp + tri(j, i)); }

/* ---- one sigma-point pass over one filter by one warp --------------------------------------------
 * MODE_PREDICT  ukfom predict (App. A.3): sigma points from the factor, process model, manifold mean,
 *               covariance + process noise.
 * MODE_UPDATE   first half of ukfom update (App. A.4): sigma points, measurement model, innovation statistics,
 *               gain, sigma <- sigma_prior - K S K^T, delta = K innov.
 * MODE_APPLY    apply_delta (App. A.4): sigma points around mu [+] delta, manifold mean, covariance.
 * `fr` is the filter's shared-memory record; its sigma slot holds the Cholesky factor at entry and the new
 * covariance at exit.  MODE is a compile-time constant: three specialised copies, no mode branches at run time. */
constexpr int MODE_PREDICT = 0, MODE_UPDATE = 1, MODE_APPLY = 2;

template <class F, int G, int MODE>
UKFB_D uint32_t sigma_pass(const Warp w, const StepParams* pp, const long long b, const int kind, double* fr,
                           const double* sigma_prior)
{
    constexpr int mode = MODE;
    typedef Smem<F, G> SM;
    constexpr int ZC = F::N; /* rows of D holding the measurement deviations */
    const StepParams& p = *pp;
    const int lane = w.lane;
    double* sig = fr + SM::OFF_SIG;
    double* mu = fr + SM::OFF_MU;
    double* delta = fr + SM::OFF_DELTA;
    uint32_t st = 0;

    /* ---- sigma points: X0 = mu + delta, X(2j+1) = mu + (delta + L[:,j]), X(2j+2) = mu + (delta - L[:,j]) */
    double x[F::MU];
    {
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) x[i] = mu[i];
        const bool col = lane >= 1 && lane < F::NS;
        const int j = col ? (lane - 1) >> 1 : 0;
        const double sgn = col ? ((lane & 1) ? 1.0 : -1.0) : 0.0;
        const double* Lc = sig + j; /* column j of the packed factor: element (i, j) at tri(i, 0) + j */
        double d[F::N];
        UKFB_UNROLL
        for (int i = 0; i < F::N; ++i) {
            const double l = (i >= j) ? Lc[tri(i, 0)] : 0.0; /* i < j never reads past row i */
            d[i] = sgn * l;
        }
        if (mode == MODE_APPLY) {
            UKFB_UNROLL
            for (int i = 0; i < F::N; ++i) d[i] += delta[i];
        }
        state_boxplus<F>(x, d, 1.0);
    }

    double dt = 0.0;
    bool has_acc = false;
    const double* Qp = p.Q + b * p.q_stride;
    if (mode == MODE_PREDICT) {
        /* ---- process model; acceleration branch of PoseUKF.cpp:188-193 */
        const double* fimu = fr + SM::OFF_IMU;
        dt = fr[SM::OFF_DT];
        const double acc[3] = {fimu[0], fimu[1], fimu[2]};
        if (F::KIND == 0)
            has_acc = (fabs(acc[0]) <= 1.79769313486231570e308) && (fabs(acc[1]) <= 1.79769313486231570e308) &&
                      (fabs(acc[2]) <= 1.79769313486231570e308);
        /* this step's process noise, packed lower, into the warp's noise table:
         *   no acceleration: scale * Q with the two rotated blocks rot * Q[blk] * rot^T, rot from the PRIOR
         *     orientation; scale = dt (PoseUKF.cpp:182-186) or dt^2 (OrientationUKF.cpp:81-86);
         *   acceleration (the shadowing local of PoseUKF.cpp:190-191): Q unrotated and unscaled, velocity
         *     block = 2 acc.cov */
        const double scale = has_acc ? 1.0 : (F::KIND == 0 ? dt : dt * dt);
        UKFB_UNROLL
        for (int q = 0; q < (F::LP + 31) / 32; ++q) {
            const int e = lane + 32 * q;
            if (e < F::LP) w.NT[e] = scale * UKFB_LDG(Qp + e);
        }
        __syncwarp();
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
            if (c <= r) w.NT[tri(off + r, off + c)] = scale * acc_rc;
        }
        if (F::KIND == 0 && has_acc && lane < 9) {
            const int r = lane / 3, c = lane % 3;
            if (c <= r) w.NT[tri(6 + r, 6 + c)] = 2.0 * p.acc_cov[b * 9 + r * 3 + c]; /* plain load: an event of this launch may have written it */
        }
        if (F::KIND == 0) {
            process_model_pose(x, dt, has_acc, acc);
        } else {
            const double omega[3] = {fimu[3], fimu[4], fimu[5]};
            if (p.ori_params) { /* this filter's own constructor arguments (OrientationUKF.cpp:41-47) */
                const double* op = p.ori_params + b * 5;
                const double earth[3] = {op[2], op[3], op[4]};
                process_model_ori(x, dt, acc, omega, op[0], op[1], earth);
            } else
                process_model_ori(x, dt, acc, omega, p.neg_inv_tau_g, p.neg_inv_tau_a, p.earth);
        }
    }

    if (mode == MODE_UPDATE) {
        const bool rot = (F::KIND == 0) && kind == UKFB_MEAS_POSE_ORIENTATION;
        /* prior covariance entries this lane will downdate (fragment-shaped ownership, see below); issued
         * early so that the L2 round trip hides behind the measurement statistics */
        const int fr_r = lane >> 2, fr_c = 2 * (lane & 3);
        double sp00[2], sp10[2], sp11[2];
        UKFB_UNROLL
        for (int e = 0; e < 2; ++e) {
            sp00[e] = (fr_c + e <= fr_r) ? UKFB_LDCG(sigma_prior + tri(fr_r, fr_c + e)) : 0.0;
            sp10[e] = (8 + fr_r < F::N) ? UKFB_LDCG(sigma_prior + tri(8 + fr_r, fr_c + e)) : 0.0;
            sp11[e] = (8 + fr_r < F::N && fr_c + e <= fr_r) ? UKFB_LDCG(sigma_prior + tri(8 + fr_r, 8 + fr_c + e)) : 0.0;
        }

        double z[4];
        measure<F>(x, kind, z);

        /* mean of Z (ukfom sigma_points_mean on the measurement space) */
        double zref[4];
        if (lane == 0) {
            w.BC[0] = z[0], w.BC[1] = z[1], w.BC[2] = z[2], w.BC[3] = z[3];
        }
        __syncwarp();
        zref[0] = w.BC[0], zref[1] = w.BC[1], zref[2] = w.BC[2], zref[3] = w.BC[3];
        {
            int it = 0;
            while (true) {
                double dz[3];
#if UKFB_EUCLID_MEAS_DIRECT_MEAN
                if (!rot)
                    dz[0] = z[0], dz[1] = z[1], dz[2] = z[2];
                else
#endif
                    meas_boxminus(z, zref, rot, dz);
                if (lane < F::NS) {
                    UKFB_UNROLL
                    for (int c = 0; c < 3; ++c) w.D[(ZC + c) * DT_LD + lane] = dz[c];
                }
                __syncwarp();
                const double rs = row_sum(w.D, lane);
                if (lane >= ZC && lane < ZC + 3) w.MD[lane - ZC] = div_ns<F::NS>(rs);
                __syncwarp();
                const double md[3] = {w.MD[0], w.MD[1], w.MD[2]};
                const double n2 = md[0] * md[0] + md[1] * md[1] + md[2] * md[2];
                __syncwarp();
#if UKFB_EUCLID_MEAS_DIRECT_MEAN
                if (!rot) {
                    zref[0] = md[0], zref[1] = md[1], zref[2] = md[2];
                    break;
                }
#endif
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
                for (int i = 0; i < F::N; ++i) w.D[i * DT_LD + lane] = dx[i];
                UKFB_UNROLL
                for (int c = 0; c < 3; ++c) w.D[(ZC + c) * DT_LD + lane] = dz[c];
            }
        }
        __syncwarp();

        /* S = 0.5 sum dz dz^T + R (R padded with identity to 3x3);  Sxz = 0.5 sum dx dz^T:
         * component tile 1 (rows 8..15 of D) holds dz; two DMMA tiles per 4 sigma points */
        {
            double c01[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
            UKFB_UNROLL
            for (int s = 0; s < DT_COLS / 4; ++s) {
                const double f0 = frag(w.D, lane, 0, s), f1 = frag(w.D, lane, 1, s);
                warp_dmma(c01[0], c01[1], f0, f1);
                warp_dmma(c11[0], c11[1], f1, f1);
            }
            const double* Rm = fr + SM::OFF_R;
            UKFB_UNROLL
            for (int e = 0; e < 2; ++e) {
                const int cz = 8 + fr_c + e - ZC; /* measurement component of this column */
                if (cz >= 0 && cz < 3) {
                    w.SXZ[fr_r * 3 + cz] = 0.5 * c01[e];
                    if (8 + fr_r < F::N) w.SXZ[(8 + fr_r) * 3 + cz] = 0.5 * c11[e];
                    const int az = 8 + fr_r - ZC;
                    if (az >= 0 && az < 3) w.SM[az * 3 + cz] = 0.5 * c11[e] + Rm[az * 3 + cz];
                }
            }
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
            const double* zm = fr + SM::OFF_Z;
            double zin[4] = {zm[0], zm[1], zm[2], 1.0};
            if (rot) {
                const double v[3] = {zm[0], zm[1], zm[2]};
                so3_exp(v, 1.0, zin); /* PoseUKF.cpp:135 */
            }
            meas_boxminus(zin, zref, rot, innov);
        }
        /* the accept functor (every lane holds the same innov and S^-1): a rejected measurement leaves the filter as it
         * was -- the shared-memory covariance, overwritten by the factor, comes back from the record */
        {
            double d2 = 0.0;
            UKFB_UNROLL
            for (int a = 0; a < 3; ++a) d2 += innov[a] * (Si[a * 3] * innov[0] + Si[a * 3 + 1] * innov[1] + Si[a * 3 + 2] * innov[2]);
            if (d2 > p.gate_d2) {
                for (int e = lane; e < F::LP; e += 32) sig[e] = UKFB_LDCG(sigma_prior + e);
                __syncwarp();
                return st | UKFB_STATUS_MEAS_REJECTED;
            }
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
        /* sigma <- sigma_prior - (K S) K^T, lower triangle; overwrites the factor, which is no longer needed */
        UKFB_UNROLL
        for (int e = 0; e < 2; ++e) {
            const int i0 = fr_r, i1 = 8 + fr_r, j0 = fr_c + e, j1 = 8 + fr_c + e;
            double s00 = 0.0, s10 = 0.0, s11 = 0.0;
            UKFB_UNROLL
            for (int k = 0; k < 3; ++k) {
                s00 += w.KS[i0 * 3 + k] * w.KM[j0 * 3 + k];
                if (i1 < F::N) s10 += w.KS[i1 * 3 + k] * w.KM[j0 * 3 + k];
                if (i1 < F::N && j1 < F::N) s11 += w.KS[i1 * 3 + k] * w.KM[j1 * 3 + k];
            }
            if (j0 <= i0) sig[tri(i0, j0)] = sp00[e] - s00;
            if (i1 < F::N) sig[tri(i1, j0)] = sp10[e] - s10;
            if (i1 < F::N && j0 <= i0) sig[tri(i1, j1)] = sp11[e] - s11;
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

    /* ---- ukfom sigma_points_mean on the state manifold: ref = X0; loop { md = mean(X_i [-] ref);
     * ref [+]= md } while (|md| > tol && ++i < max_it).  Every lane ends with the same ref. */
    double ref[F::MU];
    if (lane == 0) {
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) w.BC[i] = x[i];
    }
    __syncwarp();
    UKFB_UNROLL
    for (int i = 0; i < F::MU; ++i) ref[i] = w.BC[i];
    {
        int it = 0, passes = 0;
        bool converged = false;
        while (true) {
            /* deviations from the current reference; after convergence the same code produces the
             * deviations from the final mean for the covariance */
            double d[F::N];
            state_boxminus<F>(x, ref, d);
            if (lane < F::NS) {
                UKFB_UNROLL
                for (int i = 0; i < F::N; ++i) w.D[i * DT_LD + lane] = d[i];
            }
            __syncwarp();
            if (converged) break;
            const double rs = row_sum(w.D, lane);
            if (lane < F::N) w.MD[lane] = div_ns<F::NS>(rs);
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
            if (!(n2 > UKFB_MEAN_TOL * UKFB_MEAN_TOL)) converged = true; /* |md| <= tol */
            else if (++it >= UKFB_MEAN_MAX_IT) {
                st = UKFB_STATUS_MEAN_NO_CONVERGE;
                converged = true;
            }
        }
        if (lane == 0) w.HP[passes < 7 ? passes : 7]++;
    }

    /* ---- covariance of the deviations: out(i,j) = 0.5 * sum_p d_i d_j + noise(i,j), lower triangle,
     * three 8x8 tiles on the DMMA path */
    {
        const int r = lane >> 2, c = 2 * (lane & 3);
        double nz00[2] = {0.0, 0.0}, nz10[2] = {0.0, 0.0}, nz11[2] = {0.0, 0.0};
        if (mode == MODE_PREDICT) {
            UKFB_UNROLL
            for (int e = 0; e < 2; ++e) {
                if (c + e <= r) nz00[e] = w.NT[tri(r, c + e)];
                if (8 + r < F::N) nz10[e] = w.NT[tri(8 + r, c + e)];
                if (8 + r < F::N && c + e <= r) nz11[e] = w.NT[tri(8 + r, 8 + c + e)];
            }
        }
        double c00[2] = {0.0, 0.0}, c10[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
        UKFB_UNROLL
        for (int s = 0; s < DT_COLS / 4; ++s) {
            const double f0 = frag(w.D, lane, 0, s), f1 = frag(w.D, lane, 1, s);
            warp_dmma(c00[0], c00[1], f0, f0);
            warp_dmma(c10[0], c10[1], f1, f0);
            warp_dmma(c11[0], c11[1], f1, f1);
        }
        UKFB_UNROLL
        for (int e = 0; e < 2; ++e) {
            if (c + e <= r) sig[tri(r, c + e)] = fma(0.5, c00[e], nz00[e]);
            if (8 + r < F::N) sig[tri(8 + r, c + e)] = fma(0.5, c10[e], nz10[e]);
            if (8 + r < F::N && c + e <= r) sig[tri(8 + r, 8 + c + e)] = fma(0.5, c11[e], nz11[e]);
        }
    }
    if (lane == 0) {
        UKFB_UNROLL
        for (int i = 0; i < F::MU; ++i) mu[i] = ref[i];
    }
    __syncwarp();
    return st;
}

/* ---- the kernel ------------------------------------------------------------------- */
/* Cholesky phase for every filter of the group whose flag has `bit`: T = 32/G lanes each.  Out of line: one
 * copy of the unrolled factorisation serves the three call sites (few values are live across the calls). */
template <class F, int G>
UKFB_DNI void cholesky_phase(double* wsm, int* cflag, int* cstat, int lane, int cnt, int bit, int clear_bits, int dirty_on_fail)
{
    typedef Smem<F, G> SM;
    constexpr int T = 32 / G;
    const int g = lane / T;
    const bool active = g < cnt && (cflag[g < cnt ? g : 0] & bit);
    double* sig = wsm + (g < cnt ? g : 0) * SM::FS + SM::OFF_SIG;
    const bool ok = cholesky_rows<F, T>(sig, lane, active);
    __syncwarp();
    if (active && !ok && (lane & (T - 1)) == 0) {
        cstat[g] |= int(UKFB_STATUS_NOT_SPD);
        cflag[g] = (cflag[g] & ~clear_bits) | (dirty_on_fail ? CF_DIRTY : 0);
    }
    __syncwarp();
}

template <class F, int G, int WPB, int MINB>
UKFB_GLOBAL void UKFB_LAUNCH_BOUNDS(WPB * 32, MINB) ukf_step_kernel(const UKFB_GRID_CONSTANT StepParams p)
{
    typedef Smem<F, G> SM;
    UKFB_SMEM_DECL
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long first = ((long long)blockIdx.x * WPB + warp) * G;
    if (first >= p.B) return;
    const int cnt = (p.B - first) < G ? int(p.B - first) : G;

    double* wsm = ukfb_smem + warp * SM::TOTAL;
    Warp w;
    w.D = wsm + SM::OFF_D;
    w.SXZ = wsm + SM::OFF_SXZ;
    w.KM = wsm + SM::OFF_KM;
    w.KS = wsm + SM::OFF_KS;
    w.SM = wsm + SM::OFF_SM;
    w.MD = wsm + SM::OFF_MD;
    w.BC = wsm + SM::OFF_BC;
    w.NT = wsm + SM::OFF_NT;
    w.lane = lane;
    int* cflag = reinterpret_cast<int*>(wsm + SM::OFF_CTL);
    int* ckind = cflag + G;
    int* cstat = ckind + G;
    w.HP = cstat + G;
    int* cstore = w.HP + 8; /* a storing event (UKFB_EVENT_* >= 10) of this slot, applied after the predict */
    if (lane < 8) w.HP[lane] = 0;
    if (lane < G) cstat[lane] = 0, cflag[lane] = 0;
    for (int i = lane; i < DT_ROWS * DT_LD; i += 32) w.D[i] = 0.0; /* padding columns / rows stay zero */

    /* ---- load the group's records (coalesced) into the per-filter layout */
    double* rec = p.state + first * F::REC;
    {
        for (int i = lane; i < cnt * F::REC; i += 32) {
            const int g = i / F::REC, k = i - g * F::REC;
            wsm[g * SM::FS + k] = rec[i];
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
            wsm[g * SM::FS + SM::OFF_IMU + k] = v;
        }
    }
    __syncwarp();

    UKFB_NOUNROLL
    for (int tick = 0; tick < p.K; ++tick) {
        /* ---- per-filter control: time guards (UnscentedKalmanFilter.hpp:83-125), masks, checks */
        if (lane < G) {
            int flags = cflag[lane] & CF_DIRTY, kind = -1, st = 0, store = -1;
            if (lane < cnt) {
                const long long b = first + lane;
                double* fr = wsm + lane * SM::FS;
                flags |= CF_VALID;
                bool idle = false;
                const double* Rev = p.R + tick * p.r_kstride + b * p.r_stride;
                if (p.events) { /* one queued sample per filter and slot; UKFB_EVENT_IDLE: nothing happens */
                    kind = int(p.kinds[tick * p.kinds_kstride + b]);
                    idle = kind == UKFB_EVENT_IDLE;
                    if (kind >= UKFB_EVENT_KIND_COUNT || kind < UKFB_EVENT_IDLE
                        || (kind >= 0 && (F::KIND == 0 ? (kind == UKFB_MEAS_ORI_VELOCITY || kind > UKFB_EVENT_POSE_ACCELERATION)
                                                       : (kind < UKFB_MEAS_ORI_VELOCITY || kind == UKFB_EVENT_POSE_ACCELERATION)))) {
                        st |= UKFB_STATUS_BAD_EVENT;
                        idle = true;
                    }
                    if (idle) kind = -1;
                    if (kind >= 0) Rev += kind * p.r_kind_stride;
                    if (kind >= UKFB_EVENT_POSE_ACCELERATION) store = kind, kind = -1;
                }
                if (F::KIND == 1 && p.imu) { /* integrateMeasurement(RotationRate / Acceleration): check, store */
                    const double* s6 = p.imu + tick * p.imu_kstride + b * 6;
                    double* fimu = fr + SM::OFF_IMU;
                    const double g0 = s6[0], g1 = s6[1], g2 = s6[2], a0 = s6[3], a1 = s6[4], a2 = s6[5];
                    const double big = 1.79769313486231570e308;
                    if (fabs(g0) <= big && fabs(g1) <= big && fabs(g2) <= big)
                        fimu[3] = g0, fimu[4] = g1, fimu[5] = g2;
                    else
                        st |= UKFB_STATUS_NONFINITE_MEAS;
                    if (fabs(a0) <= big && fabs(a1) <= big && fabs(a2) <= big)
                        fimu[0] = a0, fimu[1] = a1, fimu[2] = a2;
                    else
                        st |= UKFB_STATUS_NONFINITE_MEAS;
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
                            st |= UKFB_STATUS_NEG_DT;
                            if (p.events) kind = -1, store = -1; /* the reference's callback leaves here (the throw): this sample is neither integrated nor stored */
                        } else if (dt <= p.min_dt) {
                            /* delta time is zero or close to zero: no-op */
                        } else if (dt > p.max_dt) {
                            st |= UKFB_STATUS_DT_TOO_LARGE;
                            if (p.events) kind = -1, store = -1;
                        }
                        else {
                            flags |= CF_PRED;
                            fr[SM::OFF_DT] = dt;
                        }
                    }
                }
                if (p.do_update && !idle) {
                    if (!p.events) {
                        kind = p.tick_kinds ? int(p.tick_kinds[tick])
                                            : (p.kind == -2 ? int(p.kinds[tick * p.kinds_kstride + b]) : p.kind);
                        if (p.kind == -2 && !p.tick_kinds && kind != UKFB_MEAS_NONE && !meas_kind_of_class<F>(kind)) {
                            st |= UKFB_STATUS_BAD_EVENT; /* per-filter kinds on the device: a kind of the other filter class is ignored */
                            kind = -1;
                        }
                        if (p.mask && !p.mask[tick * p.mask_kstride + b]) kind = -1;
                    }
                    if (kind >= 0 || store >= 0) {
                        /* stage the measurement: z zero padded to 3, R identity padded to 3x3 */
                        const int m = store >= 0 ? 3 : meas_dim(kind);
                        const double* zm = p.z + tick * p.z_kstride + b * p.z_stride;
                        const double* Rm = Rev;
                        bool ok = true;
                        UKFB_UNROLL
                        for (int a = 0; a < 3; ++a) {
                            const double zv = a < m ? zm[a] : 0.0;
                            ok = ok && (fabs(zv) <= 1.79769313486231570e308);
                            fr[SM::OFF_Z + a] = zv;
                            UKFB_UNROLL
                            for (int c = 0; c < 3; ++c) {
                                const double rv = (a < m && c < m) ? Rm[a * p.r_ld + c] : (a == c ? 1.0 : 0.0);
                                ok = ok && (fabs(rv) <= 1.79769313486231570e308);
                                fr[SM::OFF_R + a * 3 + c] = rv;
                            }
                        }
                        /* checkMeasurment: OrientationUKF only (OrientationUKF.cpp:55,61,67); PoseUKF never checks */
                        if (F::KIND == 0 || ok) {
                            if (kind >= 0) flags |= CF_UPD;
                        } else {
                            st |= UKFB_STATUS_NONFINITE_MEAS;
                            kind = -1, store = -1;
                        }
                    }
                }
            }
            cflag[lane] = flags;
            ckind[lane] = kind;
            cstat[lane] |= st;
            cstore[lane] = store;
        }
        __syncwarp();

        /* ---- predict ------------------------------------------------------------------ */
        if (p.do_predict) {
            cholesky_phase<F, G>(wsm, cflag, cstat, lane, cnt, CF_PRED, CF_PRED | CF_UPD, 0);
            for (int g = 0; g < cnt; ++g) {
                if (!(cflag[g] & CF_PRED)) continue;
                const uint32_t st = sigma_pass<F, G, MODE_PREDICT>(w, &p, first + g, -1, wsm + g * SM::FS, nullptr);
                if (lane == 0) {
                    cstat[g] |= int(st);
                    cflag[g] |= CF_DIRTY;
                }
            }
            __syncwarp();
        }

        /* ---- storing events: the sample is kept for the next predict (PoseUKF.cpp:175-178, OrientationUKF.cpp:53-63) */
        if (p.events) {
            if (lane < cnt && cstore[lane] >= 0) {
                double* fr = wsm + lane * SM::FS;
                double* fimu = fr + SM::OFF_IMU;
                const int off = cstore[lane] == UKFB_EVENT_ORI_ROTATION_RATE ? 3 : 0;
                fimu[off] = fr[SM::OFF_Z], fimu[off + 1] = fr[SM::OFF_Z + 1], fimu[off + 2] = fr[SM::OFF_Z + 2];
                if (F::KIND == 0)
                    for (int i = 0; i < 9; ++i) p.acc_cov[(first + lane) * 9 + i] = fr[SM::OFF_R + i];
            }
            __syncwarp();
        }

        /* ---- update --------------------------------------------------------------------- */
        if (p.do_update) {
            /* spill the prior covariance of the updating filters to their HBM records (L2 resident): the factor
             * is about to overwrite it in shared memory and sigma - K S K^T needs it back */
            for (int g = 0; g < cnt; ++g) {
                if (!(cflag[g] & CF_UPD)) continue;
                const double* sig = wsm + g * SM::FS + SM::OFF_SIG;
                double* dst = rec + g * F::REC + F::MU;
                for (int e = lane; e < F::LP; e += 32) dst[e] = sig[e];
            }
            cholesky_phase<F, G>(wsm, cflag, cstat, lane, cnt, CF_UPD, CF_UPD, 0);
            for (int g = 0; g < cnt; ++g) {
                if (!(cflag[g] & CF_UPD)) continue;
                const uint32_t st = sigma_pass<F, G, MODE_UPDATE>(w, &p, first + g, ckind[g], wsm + g * SM::FS, rec + g * F::REC + F::MU);
                if (lane == 0) {
                    cstat[g] |= int(st);
                    if (st & UKFB_STATUS_MEAS_REJECTED) cflag[g] &= ~CF_UPD; /* gated out: no apply_delta */
                }
            }
            __syncwarp();
            /* the reference has already replaced sigma by sigma - K S K^T when MTK's assert fires inside
             * apply_delta: on failure keep that matrix (dirty), leave mu alone */
            cholesky_phase<F, G>(wsm, cflag, cstat, lane, cnt, CF_UPD, CF_UPD, 1);
            for (int g = 0; g < cnt; ++g) {
                if (!(cflag[g] & CF_UPD)) continue;
                const uint32_t st = sigma_pass<F, G, MODE_APPLY>(w, &p, first + g, -1, wsm + g * SM::FS, nullptr);
                if (lane == 0) {
                    cstat[g] |= int(st);
                    cflag[g] |= CF_DIRTY;
                }
            }
            __syncwarp();
        }
    }

    /* ---- store dirty records (coalesced), stored IMU sample, status, histogram --------------- */
    {
        for (int i = lane; i < cnt * F::REC; i += 32) {
            const int g = i / F::REC, k = i - g * F::REC;
            /* a spilled-but-not-updated covariance equals the shared-memory copy, so clean records can be skipped */
            if (!(cflag[g] & CF_DIRTY)) continue;
            rec[i] = wsm[g * SM::FS + k];
        }
        if ((F::KIND == 1 && p.imu) || p.events) {
            for (int i = lane; i < cnt * 6; i += 32) {
                const int g = i / 6, k = i - g * 6;
                const double v = wsm[g * SM::FS + SM::OFF_IMU + k];
                if (k < 3)
                    p.acc_mu[(first + g) * 3 + k] = v;
                else if (F::KIND == 1)
                    p.gyro_mu[(first + g) * 3 + (k - 3)] = v;
            }
        }
    }
    if (lane < cnt && cstat[lane]) p.status[first + lane] |= uint32_t(cstat[lane]);
    __syncwarp();
    if (lane >= 1 && lane < 8 && p.hist && w.HP[lane])
        atomicAdd(p.hist + (blockIdx.x % HIST_SLOTS) * 8 + lane, (unsigned long long)w.HP[lane]);
}

} /* namespace ukfb */

#endif /* UKFB_DEVICE_CUH */
