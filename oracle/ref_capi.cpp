/*
 * oracle/ref_capi.cpp -- TEST INFRASTRUCTURE.  The batch C interface of oracle_capi.cpp (orc_*) over the REFERENCE'S
 * OWN filter classes: pose_estimation::PoseUKF and pose_estimation::OrientationUKF, compiled unmodified from
 *     /root/reference/src/pose_with_velocity/PoseUKF.cpp
 *     /root/reference/src/orientation_estimator/OrientationUKF.cpp
 *     /root/reference/src/UnscentedKalmanFilter.hpp (+ PoseWithVelocity.hpp, OrientationState.hpp, Measurement.hpp,
 *                                                      GravitationalModel.hpp, OrientationUKFConfig.hpp)
 * against the stand-in headers of oracle/ref_shim (the reference's dependencies -- Eigen, Boost, base-types and the
 * un-vendored slam/mtk -- are absent from this container).  Built by oracle/ref_recipe.mk into oracle/_ref/libref.so;
 * tests/test_ref_pin.py checks the oracle against it.
 *
 * What is reference text in that library and what is not:
 *   REFERENCE TEXT  predictionStepFromSampleTime / predictionStep guards and the time latch, initializeFilter,
 *                   getCurrentState; PoseUKF: default process noise, NaN acceleration sentinel, the nine measurement
 *                   models, processModel / processModelWithAcceleration, predictionStepImpl incl. the shadowed
 *                   process_noise of PoseUKF.cpp:190; OrientationUKF: constructor (earth rotation), checkMeasurment
 *                   calls, processModel, velocityMeasurementModel, getRotationRate, predictionStepImpl (dt^2).
 *   RESTATEMENT     ukfom::ukf (predict / update / apply_delta), MTK::SO3 exp / log / boxplus / boxminus, MTK::vect,
 *                   MTK_BUILD_MANIFOLD, Eigen's quaternion and fixed-size matrix arithmetic: oracle/ref_shim on top of
 *                   oracle/ukf_oracle.hpp.  That layer stays unpinned by the reference.
 *
 * The adapters below only convert between the flat arrays of the batch interface and the reference's types, and reach
 * the reference's protected members (they derive from its classes); they add no filter arithmetic.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>

#include <pose_estimation/GravitationalModel.hpp>
#include <pose_estimation/orientation_estimator/OrientationUKF.hpp>
#include <pose_estimation/pose_with_velocity/PoseUKF.hpp>

#include "ukf_oracle.hpp"

namespace refad {

template <class V>
inline void put3(V& dst, const double* src)
{
    for (int i = 0; i < 3; ++i) dst[i] = src[i];
}
template <class M>
inline void fill_meas(M& m, const double* mu, const double* cov)
{
    const int d = M::Mu::RowsAtCompileTime;
    for (int i = 0; i < d; ++i) m.mu[i] = mu[i];
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) m.cov(i, j) = cov[i * d + j];
}

class RefPose : public pose_estimation::PoseUKF {
public:
    typedef pose_estimation::PoseUKF Ref;
    static State to_state(const orc::PoseState<double>& s)
    {
        State x;
        put3(x.position, s.position);
        x.orientation = pose_estimation::RotationType(MTK::SO3<double>(Eigen::Quaterniond(s.orientation)));
        put3(x.velocity, s.velocity);
        put3(x.angular_velocity, s.angular_velocity);
        return x;
    }
    static Covariance to_cov(const double* sigma)
    {
        Covariance c;
        for (int i = 0; i < 144; ++i) c[i] = sigma[i];
        return c;
    }
    RefPose(const orc::PoseState<double>& s, const double* sigma) : Ref(to_state(s), to_cov(sigma)) {}
    void initializeFilter(const orc::PoseState<double>& s, const double* sigma) { Ref::initializeFilter(to_state(s), to_cov(sigma)); }
    void predictionStepFromSampleTime(int64_t us) { Ref::predictionStepFromSampleTime(base::Time::fromMicroseconds(us)); }
    /* the nine overloads of PoseUKF.cpp:112-173, selected by the batch interface's kind */
    void integrateMeasurement(int kind, const double* z, const double* zc)
    {
        switch (kind) {
            case 0: { PositionMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            case 1: { XYMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            case 2: { ZMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            case 3: { OrientationMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            case 4: { VelocityMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            case 5: { XYVelocityMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            case 6: { ZVelocityMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            case 7: { XVelYawVelMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            case 8: { AngularVelocityMeasurement m; fill_meas(m, z, zc); Ref::integrateMeasurement(m); break; }
            default: throw std::invalid_argument("bad PoseUKF measurement kind");
        }
    }
    void setAcceleration(const double* mu, const double* cov)
    {
        AccelerationMeasurement m;
        fill_meas(m, mu, cov);
        Ref::integrateMeasurement(m);
    }
    /* protected members of the reference, for the accessors below */
    MTK_UKF& engine() { return *ukf; }
    const MTK_UKF& engine() const { return *ukf; }
    Covariance& q() { return process_noise_cov; }
};

class RefOri : public pose_estimation::OrientationUKF {
public:
    typedef pose_estimation::OrientationUKF Ref;
    static State to_state(const orc::OrientationState<double>& s)
    {
        State x;
        x.orientation = pose_estimation::RotationType(MTK::SO3<double>(Eigen::Quaterniond(s.orientation)));
        put3(x.velocity, s.velocity);
        put3(x.bias_gyro, s.bias_gyro);
        put3(x.bias_acc, s.bias_acc);
        x.gravity[0] = s.gravity[0];
        return x;
    }
    static Covariance to_cov(const double* sigma)
    {
        Covariance c;
        for (int i = 0; i < 169; ++i) c[i] = sigma[i];
        return c;
    }
    static pose_estimation::LocationConfiguration location(double latitude)
    {
        pose_estimation::LocationConfiguration l;
        l.latitude = latitude, l.longitude = 0.0, l.altitude = 0.0;
        return l;
    }
    RefOri(const orc::OrientationState<double>& s, const double* sigma, double tau_g, double tau_a, double latitude)
        : Ref(to_state(s), to_cov(sigma), tau_g, tau_a, location(latitude))
    {
    }
    void initializeFilter(const orc::OrientationState<double>& s, const double* sigma) { Ref::initializeFilter(to_state(s), to_cov(sigma)); }
    void predictionStepFromSampleTime(int64_t us) { Ref::predictionStepFromSampleTime(base::Time::fromMicroseconds(us)); }
    void setRotationRate(const double* mu, const double* cov)
    {
        RotationRate m;
        fill_meas(m, mu, cov);
        Ref::integrateMeasurement(m);
    }
    void setAcceleration(const double* mu, const double* cov)
    {
        Acceleration m;
        fill_meas(m, mu, cov);
        Ref::integrateMeasurement(m);
    }
    void integrateVelocity(const double* z, const double* zc)
    {
        VelocityMeasurement m;
        fill_meas(m, z, zc);
        Ref::integrateMeasurement(m);
    }
    void getRotationRate(double out[3])
    {
        const RotationRate::Mu r = Ref::getRotationRate();
        for (int i = 0; i < 3; ++i) out[i] = r[i];
    }
    /* the constructor arguments of OrientationUKF.cpp:41-47 changed after construction (the batch interface sets them per
     * filter): the same assignments the reference's constructor makes */
    void set_params(double tau_g, double tau_a, double latitude)
    {
        gyro_bias_tau = tau_g;
        acc_bias_tau = tau_a;
        earth_rotation = Eigen::Vector3d(pose_estimation::EARTHW * cos(latitude), 0., pose_estimation::EARTHW * sin(latitude));
    }
    MTK_UKF& engine() { return *ukf; }
    const MTK_UKF& engine() const { return *ukf; }
    Covariance& q() { return process_noise_cov; }
};

}  // namespace refad

#define ORC_CUSTOM_IMPL 1
typedef refad::RefPose PoseImpl;
typedef refad::RefOri OriImpl;

namespace ax {
inline void store_mu(const PoseImpl& f, double* mu)
{
    const auto& x = f.engine().mu();
    for (int i = 0; i < 3; ++i) mu[i] = x.position[i], mu[7 + i] = x.velocity[i], mu[10 + i] = x.angular_velocity[i];
    mu[3] = x.orientation.x(), mu[4] = x.orientation.y(), mu[5] = x.orientation.z(), mu[6] = x.orientation.w();
}
inline void store_mu(const OriImpl& f, double* mu)
{
    const auto& x = f.engine().mu();
    mu[0] = x.orientation.x(), mu[1] = x.orientation.y(), mu[2] = x.orientation.z(), mu[3] = x.orientation.w();
    for (int i = 0; i < 3; ++i) mu[4 + i] = x.velocity[i], mu[7 + i] = x.bias_gyro[i], mu[10 + i] = x.bias_acc[i];
    mu[13] = x.gravity[0];
}
template <class F> void copy_sigma(const F& f, double* sigma)
{
    const auto& s = f.engine().sigma();
    for (int i = 0; i < int(F::DOF) * int(F::DOF); ++i) sigma[i] = s[i];
}
template <class F> void set_q(F& f, int k, double v) { f.q()[k] = v; }
template <class F> void set_dt_bounds(F& f, double lo, double hi) { f.setMinTimeDelta(lo), f.setMaxTimeDelta(hi); }
template <class F> void set_gate(F& f, double d2) { f.engine().engine().accept_max_d2 = d2; }
template <class F> bool rejected(const F& f) { return f.engine().engine().last_update_rejected; }
template <class F> void set_last_time(F& f, int64_t t) { f.setLastMeasurementTime(base::Time::fromMicroseconds(t)); }
template <class F> int64_t last_time(const F& f) { return f.getLastMeasurementTime().microseconds; }
template <class F> uint32_t status(const F& f) { return f.engine().engine().status; }
template <class F> void clear_status(F& f) { f.engine().engine().status = 0; }
template <class F> uint64_t mean_iters(const F& f, int k) { return f.engine().engine().mean_iters[k]; }
inline void set_ori_params(OriImpl& f, double tau_g, double tau_a, double latitude) { f.set_params(tau_g, tau_a, latitude); }
}  // namespace ax

#include "oracle_capi.cpp"
