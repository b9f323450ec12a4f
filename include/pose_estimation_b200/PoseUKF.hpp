/*
 * PoseUKF.hpp -- host-side mirror of pose_estimation::PoseUKF (reference
 * src/pose_with_velocity/PoseUKF.hpp:17-93, PoseUKF.cpp:99-196) over the C ABI.
 * State layout (PoseWithVelocity.hpp:18-23): position[0:3], orientation quaternion x,y,z,w [3:7],
 * velocity (body frame) [7:10], angular_velocity (body frame) [10:13]; covariance 12 x 12 in the tangent order
 * position, orientation, velocity, angular_velocity.
 */
#ifndef POSE_ESTIMATION_B200_POSE_UKF_HPP
#define POSE_ESTIMATION_B200_POSE_UKF_HPP

#include "Measurement.hpp"
#include "UnscentedKalmanFilter.hpp"

namespace pose_estimation_b200
{

class PoseUKF : public UnscentedKalmanFilter<UKFB_POSE, 12, 13>
{
public:
    UKFB_MEASUREMENT(PositionMeasurement, 3)
    UKFB_MEASUREMENT(XYMeasurement, 2)
    UKFB_MEASUREMENT(ZMeasurement, 1)
    UKFB_MEASUREMENT(OrientationMeasurement, 3)
    UKFB_MEASUREMENT(VelocityMeasurement, 3)
    UKFB_MEASUREMENT(XYVelocityMeasurement, 2)
    UKFB_MEASUREMENT(ZVelocityMeasurement, 1)
    UKFB_MEASUREMENT(XVelYawVelMeasurement, 2)
    UKFB_MEASUREMENT(AngularVelocityMeasurement, 3)
    UKFB_MEASUREMENT(AccelerationMeasurement, 3)

    /* PoseUKF.cpp:99-110: initializeFilter, default process noise, acceleration = NaN (both set by ukfb_create) */
    PoseUKF(const State& initial_state, const Covariance& state_cov, int device = 0) : UnscentedKalmanFilter(1, device)
    {
        initializeFilter(initial_state, state_cov);
    }
    /* `batch` filters, arrays of states and covariances */
    PoseUKF(int64_t batch, const State* initial_state, const Covariance* state_cov, int device = 0) : UnscentedKalmanFilter(batch, device)
    {
        initializeFilter(initial_state, state_cov);
    }

    /* `batch` filters split by filter index over the listed GPUs (one object, one host thread per device inside) */
    PoseUKF(int64_t batch, const State* initial_state, const Covariance* state_cov, const std::vector<int>& devices)
        : UnscentedKalmanFilter(batch, devices)
    {
        initializeFilter(initial_state, state_cov);
    }

    /* PoseUKF.cpp:112-173.  The measurement is applied to every filter of the batch (batch = 1: the
     * reference call); integrateMeasurements takes one measurement per filter. */
    void integrateMeasurement(const PositionMeasurement& m) { update(UKFB_MEAS_POSE_POSITION, &m, false); }
    void integrateMeasurement(const XYMeasurement& m) { update(UKFB_MEAS_POSE_XY, &m, false); }
    void integrateMeasurement(const ZMeasurement& m) { update(UKFB_MEAS_POSE_Z, &m, false); }
    void integrateMeasurement(const OrientationMeasurement& m) { update(UKFB_MEAS_POSE_ORIENTATION, &m, false); }
    void integrateMeasurement(const VelocityMeasurement& m) { update(UKFB_MEAS_POSE_VELOCITY, &m, false); }
    void integrateMeasurement(const XYVelocityMeasurement& m) { update(UKFB_MEAS_POSE_XY_VELOCITY, &m, false); }
    void integrateMeasurement(const ZVelocityMeasurement& m) { update(UKFB_MEAS_POSE_Z_VELOCITY, &m, false); }
    void integrateMeasurement(const XVelYawVelMeasurement& m) { update(UKFB_MEAS_POSE_XVEL_YAWVEL, &m, false); }
    void integrateMeasurement(const AngularVelocityMeasurement& m) { update(UKFB_MEAS_POSE_ANGULAR_VELOCITY, &m, false); }
    template <class M>
    void integrateMeasurements(int meas_kind, const M* per_filter) { update(meas_kind, per_filter, true); }

    /* PoseUKF.cpp:175-178: stored (unchecked) for the next prediction step */
    void integrateMeasurement(const AccelerationMeasurement& m)
    {
        std::vector<double> mu(size_t(batch_size) * 3);
        for (int64_t b = 0; b < batch_size; ++b)
            for (int i = 0; i < 3; ++i) mu[size_t(b) * 3 + i] = m.mu[i];
        check(ukfb_set_acceleration(h, mu.data(), m.cov, 0, nullptr));
    }
};

}  // namespace pose_estimation_b200

#endif
