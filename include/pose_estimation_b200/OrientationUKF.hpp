/*
 * OrientationUKF.hpp -- host-side mirror of pose_estimation::OrientationUKF (reference
 * src/orientation_estimator/OrientationUKF.hpp:20-59, OrientationUKF.cpp:41-89) over the C ABI.
 * State layout (OrientationState.hpp:20-26): orientation quaternion x,y,z,w [0:4], velocity (navigation frame)
 * [4:7], bias_gyro [7:10], bias_acc [10:13], gravity [13]; covariance 13 x 13 in the tangent order orientation,
 * velocity, bias_gyro, bias_acc, gravity.
 */
#ifndef POSE_ESTIMATION_B200_ORIENTATION_UKF_HPP
#define POSE_ESTIMATION_B200_ORIENTATION_UKF_HPP

#include "Measurement.hpp"
#include "UnscentedKalmanFilter.hpp"

namespace pose_estimation_b200
{

/* OrientationUKFConfig.hpp:24-34 */
struct LocationConfiguration {
    double latitude;  /* rad */
    double longitude; /* rad */
    double altitude;  /* m */
};

class OrientationUKF : public UnscentedKalmanFilter<UKFB_ORIENTATION, 13, 14>
{
public:
    UKFB_MEASUREMENT(RotationRate, 3)
    UKFB_MEASUREMENT(Acceleration, 3)
    UKFB_MEASUREMENT(VelocityMeasurement, 3)

    /* OrientationUKF.cpp:41-51 */
    OrientationUKF(const State& initial_state, const Covariance& state_cov, double gyro_bias_tau, double acc_bias_tau,
                   const LocationConfiguration& location, int device = 0)
        : UnscentedKalmanFilter(1, device)
    {
        check(ukfb_set_orientation_params(h, gyro_bias_tau, acc_bias_tau, location.latitude));
        initializeFilter(initial_state, state_cov); /* also stores rotation_rate = 0, acceleration = (0, 0, gravity) */
    }
    OrientationUKF(int64_t batch, const State* initial_state, const Covariance* state_cov, double gyro_bias_tau, double acc_bias_tau,
                   const LocationConfiguration& location, int device = 0)
        : UnscentedKalmanFilter(batch, device)
    {
        check(ukfb_set_orientation_params(h, gyro_bias_tau, acc_bias_tau, location.latitude));
        initializeFilter(initial_state, state_cov);
    }

    OrientationUKF(int64_t batch, const State* initial_state, const Covariance* state_cov, double gyro_bias_tau, double acc_bias_tau,
                   const LocationConfiguration& location, const std::vector<int>& devices)
        : UnscentedKalmanFilter(batch, devices)
    {
        check(ukfb_set_orientation_params(h, gyro_bias_tau, acc_bias_tau, location.latitude));
        initializeFilter(initial_state, state_cov);
    }

    /* OrientationUKF.cpp:53-57: finite check, then store */
    void integrateMeasurement(const RotationRate& m)
    {
        std::vector<double> mu = replicate(m.mu);
        check(ukfb_set_rotation_rate(h, mu.data(), m.cov, 0, nullptr));
        raise();
    }
    /* OrientationUKF.cpp:59-63 */
    void integrateMeasurement(const Acceleration& m)
    {
        std::vector<double> mu = replicate(m.mu);
        check(ukfb_set_acceleration(h, mu.data(), m.cov, 0, nullptr));
        raise();
    }
    /* OrientationUKF.cpp:65-72 */
    void integrateMeasurement(const VelocityMeasurement& m) { update(UKFB_MEAS_ORI_VELOCITY, &m, false); }
    void integrateMeasurements(const VelocityMeasurement* per_filter) { update(UKFB_MEAS_ORI_VELOCITY, per_filter, true); }

    /* OrientationUKF.cpp:74-77: unbiased rotation rate in the IMU frame, 3 doubles per filter */
    void getRotationRate(double* out) { check(ukfb_get_rotation_rate(h, out)); }

private:
    std::vector<double> replicate(const double* v3) const
    {
        std::vector<double> mu(size_t(batch_size) * 3);
        for (int64_t b = 0; b < batch_size; ++b)
            for (int i = 0; i < 3; ++i) mu[size_t(b) * 3 + i] = v3[i];
        return mu;
    }
};

}  // namespace pose_estimation_b200

#endif
