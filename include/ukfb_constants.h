/*
 * ukfb_constants.h -- the compile-time conventions of the un-vendored `slam/mtk`
 * dependency (ukfom::ukf<>, MTK::SO3, MTK::vect) that the reference's hot path
 * runs on.  SURVEY.md Appendix A records them from the public upstream MTK
 * library; they cannot be verified in this container.  Every such convention is
 * ONE named constant here, included by the CUDA kernels, the C-ABI and the CPU
 * oracle alike, so the product and its checker can never disagree on it.
 *
 * Reference call sites these constants stand behind:
 *   UnscentedKalmanFilter.hpp:23-25,42,55-56   (ukfom::ukf / mtkwrap types)
 *   PoseUKF.cpp:80-81,93-95,135,192,195        (boxplus, SO3::exp, predict)
 *   OrientationUKF.cpp:19-29,38,69,88          (boxplus, inverse, update, predict)
 */
#ifndef UKFB_CONSTANTS_H
#define UKFB_CONSTANTS_H

/* SO(3) frame convention of MTK::SO3::boxplus / boxminus.
 *   1 (default): global frame, q <- exp(v*s) * q,  boxminus(o) = log(q * o^-1).
 *     This is the only convention under which the reference's call sites are
 *     kinematically right (PoseUKF.cpp:81 rotates the body rate into the nav frame
 *     before boxplus; PoseUKF.cpp:185 rotates the orientation block of Q likewise).
 *   0: upstream OpenSLAM MTK, body frame, q <- q * exp(v*s),  boxminus(o) = log(o^-1 * q).
 */
#ifndef UKFB_SO3_BOXPLUS_LEFT
#define UKFB_SO3_BOXPLUS_LEFT 1
#endif

/* ukfom::ukf::sigma_points_mean: loop `while (norm(mean_delta) > tol && ++i < max_it)`.
 * The tolerance is a recollection of the un-vendored slam/mtk like everything else in this file; it decides how many
 * passes a mean takes (one at 1e-5 on the benchmark workload).  tools/bench_mean_tol.py builds engine and oracle with a
 * tighter value to put the cost of a second pass on record (profiles/). */
#ifndef UKFB_MEAN_TOL
#define UKFB_MEAN_TOL 1e-5
#endif
#define UKFB_MEAN_MAX_IT 10000

/* MTK::tolerance<double>() -- the floor on ||q.vec|| inside SO3::log (mtkmath.hpp). */
#define UKFB_MTK_TOLERANCE 1e-11

/* cos_sinc_sqrt(x2): Taylor branch iff x2 < sqrt(sqrt(DBL_EPSILON)). */
#define UKFB_TAYLOR_N_BOUND 1.220703125e-4 /* = 2^-13 = DBL_EPSILON^(1/4), exact */

/* Mean of Euclidean (plain vector) measurement sigma points.
 *   0 (default): the same iterative loop as for manifolds.
 *   1: single pass  sum(Z_i)/N  (upstream's Euclidean special case).
 * The two differ by rounding only (<= a few ulp of |z|). */
#ifndef UKFB_EUCLID_MEAS_DIRECT_MEAN
#define UKFB_EUCLID_MEAS_DIRECT_MEAN 0
#endif

/* base::Time: int64 microseconds, 0 == "null" (UnscentedKalmanFilter.hpp:30,86). */
#define UKFB_US_PER_S 1000000.0

/* UnscentedKalmanFilter ctor defaults (UnscentedKalmanFilter.hpp:27-33). */
#define UKFB_DEFAULT_MIN_DT 1.0e-9
/* max_time_delta default is DBL_MAX */

/* GravitationalModel.hpp:16  EARTHW = 2*pi/86164 */
#define UKFB_EARTHW (6.283185307179586476925286766559 / 86164.0)

/* PoseUKF default process noise diagonal (PoseUKF.cpp:103-107). */
#define UKFB_POSE_Q_POSITION 0.01
#define UKFB_POSE_Q_ORIENTATION 0.001
#define UKFB_POSE_Q_VELOCITY 0.00001
#define UKFB_POSE_Q_ANGULAR_VELOCITY 0.00001

/* State layouts (storage order of mu; quaternion stored x,y,z,w like Eigen).
 *   PoseWithVelocity  (PoseWithVelocity.hpp:14-25):  p[0:3] q[3:7] v[7:10] w[10:13]
 *       tangent: pos[0:3] ori[3:6] vel[6:9] angvel[9:12]
 *   OrientationState  (OrientationState.hpp:15-26):  q[0:4] v[4:7] bg[7:10] ba[10:13] g[13]
 *       tangent: ori[0:3] vel[3:6] bias_gyro[6:9] bias_acc[9:12] gravity[12]
 */
#define UKFB_POSE_DOF 12
#define UKFB_POSE_MU 13
#define UKFB_ORI_DOF 13
#define UKFB_ORI_MU 14

#endif /* UKFB_CONSTANTS_H */
