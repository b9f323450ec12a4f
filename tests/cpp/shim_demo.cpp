// shim_demo.cpp -- drives the C++ host shim (include/pose_estimation_b200) the way a caller of the reference
// classes would, and prints the resulting states for tests/test_cpp_shim.py to compare with the oracle.
// Build: g++ -std=c++17 -I include tests/cpp/shim_demo.cpp -L slam_pose_estimation_b200/lib -lukfb -Wl,-rpath,...
#include <cmath>
#include <cstdio>
#include <limits>

#include <pose_estimation_b200/BodyStateMeasurement.hpp>
#include <pose_estimation_b200/EventQueue.hpp>
#include <pose_estimation_b200/OrientationUKF.hpp>
#include <pose_estimation_b200/PoseUKF.hpp>

using namespace pose_estimation_b200;

template <class S>
static void print_state(const char* tag, const S& s, int n)
{
    printf("%s", tag);
    for (int i = 0; i < n; ++i) printf(" %.17g", s.v[i]);
    printf("\n");
}

int main()
{
    // ---- PoseUKF: predict from sample times, then one update of each kind ------------------------------
    PoseUKF::State x0 = {};
    x0.v[6] = 1.0;  // identity orientation
    x0.v[7] = 1.0;  // body velocity x
    x0.v[12] = 0.05;
    PoseUKF::Covariance p0 = {};
    const double d0[12] = {1, 1, 1, 0.01, 0.01, 0.01, 0.1, 0.1, 0.1, 0.01, 0.01, 0.01};
    for (int i = 0; i < 12; ++i) p0.v[i * 12 + i] = d0[i];
    PoseUKF pose(x0, p0);
    printf("pose_state_size %u initialized %d\n", pose.getStateSize(), int(pose.isInitialized()));
    pose.predictionStepFromSampleTime(int64_t(1000000));  // first call only latches
    pose.predictionStepFromSampleTime(int64_t(1010000));
    PoseUKF::PositionMeasurement pm;
    pm.mu[0] = 0.02, pm.mu[1] = -0.01, pm.mu[2] = 0.005;
    for (int i = 0; i < 3; ++i) pm.cov[i * 3 + i] = 0.25;
    pose.integrateMeasurement(pm);
    PoseUKF::XYMeasurement xy;
    xy.mu[0] = 0.01, xy.mu[1] = 0.0;
    pose.integrateMeasurement(xy);
    PoseUKF::ZMeasurement zm;
    zm.mu[0] = -0.02;
    pose.integrateMeasurement(zm);
    PoseUKF::OrientationMeasurement om;
    om.mu[2] = 0.001;
    for (int i = 0; i < 3; ++i) om.cov[i * 3 + i] = 1e-4;
    pose.integrateMeasurement(om);
    PoseUKF::VelocityMeasurement vm;
    vm.mu[0] = 1.01;
    for (int i = 0; i < 3; ++i) vm.cov[i * 3 + i] = 1e-4;
    pose.integrateMeasurement(vm);
    PoseUKF::XYVelocityMeasurement xyv;
    xyv.mu[0] = 0.99;
    pose.integrateMeasurement(xyv);
    PoseUKF::ZVelocityMeasurement zv;
    pose.integrateMeasurement(zv);
    PoseUKF::XVelYawVelMeasurement xw;
    xw.mu[0] = 1.0, xw.mu[1] = 0.05;
    pose.integrateMeasurement(xw);
    PoseUKF::AngularVelocityMeasurement wm;
    wm.mu[2] = 0.049;
    for (int i = 0; i < 3; ++i) wm.cov[i * 3 + i] = 1e-6;
    pose.integrateMeasurement(wm);
    PoseUKF::AccelerationMeasurement am;
    am.mu[0] = 0.1;
    for (int i = 0; i < 3; ++i) am.cov[i * 3 + i] = 1e-4;
    pose.integrateMeasurement(am);
    pose.predictionStep(0.01);
    PoseUKF::State xs;
    PoseUKF::Covariance ps;
    if (!pose.getCurrentState(xs, ps)) return 2;
    print_state("pose_mu", xs, 13);
    print_state("pose_sigma", ps, 144);
    printf("pose_last_time %lld\n", (long long)pose.getLastMeasurementTime());

    // ---- the reference's exceptions ----------------------------------------------------------------------
    int caught = 0;
    try {
        pose.predictionStep(-0.5);
    } catch (const std::runtime_error& e) {
        printf("caught: %s\n", e.what());
        ++caught;
    }
    pose.setMaxTimeDelta(1.0);
    try {
        pose.predictionStep(2.0);
    } catch (const std::runtime_error& e) {
        printf("caught: %s\n", e.what());
        ++caught;
    }
    pose.predictionStep(0.0);  // silently ignored (:114-118)

    // ---- OrientationUKF ------------------------------------------------------------------------------------
    OrientationUKF::State o0 = {};
    o0.v[3] = 1.0;
    o0.v[13] = 9.81;
    OrientationUKF::Covariance op = {};
    const double od[13] = {0.01, 0.01, 0.01, 0.01, 0.01, 0.01, 1e-6, 1e-6, 1e-6, 1e-4, 1e-4, 1e-4, 1e-4};
    for (int i = 0; i < 13; ++i) op.v[i * 13 + i] = od[i];
    LocationConfiguration loc = {0.92698121, 0.154595663, 0.0};
    OrientationUKF ori(o0, op, 3600.0, 3600.0, loc);
    OrientationUKF::Covariance oq = {};
    const double qd[13] = {1e-6, 1e-6, 1e-6, 1e-4, 1e-4, 1e-4, 1e-10, 1e-10, 1e-10, 1e-8, 1e-8, 1e-8, 1e-12};
    for (int i = 0; i < 13; ++i) oq.v[i * 13 + i] = qd[i];
    ori.setProcessNoiseCovariance(oq);
    OrientationUKF::RotationRate rr;
    rr.mu[2] = 0.05;
    OrientationUKF::Acceleration ac;
    ac.mu[2] = 9.81;
    for (int k = 0; k < 5; ++k) {
        ori.integrateMeasurement(rr);
        ori.integrateMeasurement(ac);
        ori.predictionStepFromSampleTime(int64_t(1000000 + 1000 * k));
    }
    OrientationUKF::VelocityMeasurement ov;
    ov.mu[0] = 0.01;
    for (int i = 0; i < 3; ++i) ov.cov[i * 3 + i] = 1e-4;
    ori.integrateMeasurement(ov);
    try {
        OrientationUKF::VelocityMeasurement bad;
        bad.mu[1] = std::numeric_limits<double>::quiet_NaN();
        ori.integrateMeasurement(bad);
    } catch (const std::runtime_error& e) {
        printf("caught: %s\n", e.what());
        ++caught;
    }
    OrientationUKF::State os;
    OrientationUKF::Covariance oss;
    if (!ori.getCurrentState(os, oss)) return 3;
    print_state("ori_mu", os, 14);
    print_state("ori_sigma", oss, 169);
    double rate[3];
    ori.getRotationRate(rate);
    printf("ori_rate %.17g %.17g %.17g\n", rate[0], rate[1], rate[2]);
    // ---- EventQueue: three PoseUKF filters with queues of different depth, one launch -------------------
    {
        const int B = 3;
        PoseUKF::State xs[B] = {x0, x0, x0};
        PoseUKF::Covariance ps[B] = {p0, p0, p0};
        PoseUKF batch(B, xs, ps);
        EventQueue q(batch.handle());
        for (int b = 0; b < B; ++b) {
            for (int k = 0; k < 5; ++k) {
                PoseUKF::AngularVelocityMeasurement w;
                w.mu[2] = 0.05 + 0.001 * b;
                for (int i = 0; i < 3; ++i) w.cov[i * 3 + i] = 1e-6;
                q.push(b, 1000000 + 1000 * k + 100 * b, UKFB_MEAS_POSE_ANGULAR_VELOCITY, w);
                if (k == 2 && b >= 1) {
                    PoseUKF::VelocityMeasurement v;
                    v.mu[0] = 1.0 + 0.01 * b;
                    for (int i = 0; i < 3; ++i) v.cov[i * 3 + i] = 1e-4;
                    q.push(b, 1002500, UKFB_MEAS_POSE_VELOCITY, v);
                }
                if (k == 3 && b == 2) {
                    PoseUKF::AccelerationMeasurement a;
                    a.mu[0] = 0.2;
                    for (int i = 0; i < 3; ++i) a.cov[i * 3 + i] = 1e-4;
                    q.push(b, 1003500, UKFB_EVENT_POSE_ACCELERATION, a);
                }
            }
        }
        printf("evq_depth %zu\n", q.depth());
        q.flush();
        PoseUKF::State es[B];
        PoseUKF::Covariance ec[B];
        if (!batch.getCurrentState(es, ec)) return 4;
        for (int b = 0; b < B; ++b) {
            print_state("evq_mu", es[b], 13);
            print_state("evq_sigma", ec[b], 144);
        }
    }
    // ---- BodyStateMeasurement: a RigidBodyState in, one predict, a RigidBodyState out ---------------------
    {
        RigidBodyState rbs = {};
        rbs.position[0] = 1.0, rbs.position[1] = 2.0, rbs.position[2] = -3.0;
        rbs.orientation[2] = std::sin(0.3), rbs.orientation[3] = std::cos(0.3);
        rbs.velocity[0] = 0.5, rbs.velocity[1] = 0.1;
        rbs.angular_velocity[2] = 0.02;
        for (int i = 0; i < 3; ++i)
            rbs.cov_position[i * 4] = 0.5, rbs.cov_orientation[i * 4] = 0.01, rbs.cov_velocity[i * 4] = 0.1, rbs.cov_angular_velocity[i * 4] = 0.01;
        rbs.cov_position[1] = rbs.cov_position[3] = 0.05;
        PoseUKF f(x0, p0);
        BodyStateMeasurement::fromRigidBodyState(&rbs, f);
        f.predictionStep(0.05);
        RigidBodyState out;
        if (!BodyStateMeasurement::toRigidBodyState(f, &out)) return 5;
        printf("rbs");
        for (int i = 0; i < UKFB_RBS_DOUBLES; ++i) printf(" %.17g", out.position[i]);
        printf("\n");
    }
    printf("caught_total %d\n", caught);
    return caught == 3 ? 0 : 1;
}
