// sharded_demo.cpp -- one C++ filter object whose filters are split by index over several GPUs (ukfb_create_sharded
// behind pose_estimation_b200::PoseUKF(batch, states, covs, devices)), driven the way a caller of the reference classes
// would: sample-time predicts, per-filter measurements, an EventQueue flush, getCurrentState as the final gather.
// The same calls go to a one-device object; the two must agree bit for bit (filters share nothing:
// UnscentedKalmanFilter.hpp:150-154).  A few filters are printed for tests/test_sharded.py to compare with the oracle.
// usage: sharded_demo <batch> <device>[,<device>...]      e.g.  sharded_demo 1000 0,1,2,3,4,5,6,7   or   0,0,0
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <pose_estimation_b200/EventQueue.hpp>
#include <pose_estimation_b200/PoseUKF.hpp>

using namespace pose_estimation_b200;

// deterministic per-filter values (no RNG state): tests/test_sharded.py restates this function
static double wobble(int64_t b, int k, int c) { return std::sin(0.37 * double(b) + 1.3 * double(k) + 0.71 * double(c)); }

static void drive(PoseUKF& f, int64_t B)
{
    f.predictionStepFromSampleTime(int64_t(1000000));  // first call only latches
    const size_t n = size_t(B);
    std::vector<int64_t> ts(n);
    for (int64_t b = 0; b < B; ++b) ts[size_t(b)] = 1010000 + 10 * (b % 7);  // per-filter sample times
    f.predictionStepFromSampleTime(ts.data());
    std::vector<PoseUKF::AngularVelocityMeasurement> w(n);
    std::vector<PoseUKF::XYMeasurement> xy(n);
    for (int64_t b = 0; b < B; ++b) {
        for (int i = 0; i < 3; ++i) {
            w[size_t(b)].mu[i] = (i == 2 ? 0.05 : 0.0) + 1e-3 * wobble(b, 0, i);
            w[size_t(b)].cov[i * 3 + i] = 1e-6 * (1.0 + double(b % 3));  // per-filter covariances
        }
        for (int i = 0; i < 2; ++i) {
            xy[size_t(b)].mu[i] = 0.01 * wobble(b, 1, i);
            xy[size_t(b)].cov[i * 2 + i] = 0.25;
        }
    }
    f.integrateMeasurements(UKFB_MEAS_POSE_ANGULAR_VELOCITY, w.data());
    f.integrateMeasurements(UKFB_MEAS_POSE_XY, xy.data());
    PoseUKF::VelocityMeasurement v;  // one measurement for the whole batch
    v.mu[0] = 1.01;
    for (int i = 0; i < 3; ++i) v.cov[i * 3 + i] = 1e-4;
    f.integrateMeasurement(v);
    f.predictionStep(0.01);
    // ragged per-filter queues through the event scheduler (slot-major arrays: the strided case for a shard)
    EventQueue q(f.handle());
    for (int64_t b = 0; b < B; ++b) {
        for (int k = 0; k < 3 + int(b % 3); ++k) {
            PoseUKF::AngularVelocityMeasurement wk;
            for (int i = 0; i < 3; ++i) {
                wk.mu[i] = (i == 2 ? 0.05 : 0.0) + 1e-3 * wobble(b, 2 + k, i);
                wk.cov[i * 3 + i] = 1e-6;
            }
            q.push(b, 1030000 + 1000 * k + 10 * (b % 5), UKFB_MEAS_POSE_ANGULAR_VELOCITY, wk);
        }
        if (b % 4 == 1) {
            PoseUKF::ZMeasurement z;
            z.mu[0] = 0.02 * wobble(b, 9, 0);
            q.push(b, 1040000, UKFB_MEAS_POSE_Z, z);
        }
    }
    q.flush();
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <batch> <device>[,<device>...]\n", argv[0]);
        return 64;
    }
    const int64_t B = atoll(argv[1]);
    std::vector<int> devices;
    for (char* tok = strtok(argv[2], ","); tok; tok = strtok(nullptr, ",")) devices.push_back(atoi(tok));
    const size_t n = size_t(B);
    std::vector<PoseUKF::State> x0(n);
    std::vector<PoseUKF::Covariance> p0(n);
    const double d0[12] = {1, 1, 1, 0.01, 0.01, 0.01, 0.1, 0.1, 0.1, 0.01, 0.01, 0.01};
    for (int64_t b = 0; b < B; ++b) {
        PoseUKF::State& x = x0[size_t(b)];
        memset(&x, 0, sizeof(x));
        memset(&p0[size_t(b)], 0, sizeof(PoseUKF::Covariance));
        const double yaw = 0.3 * wobble(b, 20, 0);
        x.v[0] = wobble(b, 21, 0), x.v[1] = wobble(b, 21, 1);
        x.v[5] = std::sin(0.5 * yaw), x.v[6] = std::cos(0.5 * yaw);
        x.v[7] = 1.0 + 0.1 * wobble(b, 22, 0);
        x.v[12] = 0.05;
        for (int i = 0; i < 12; ++i) p0[size_t(b)].v[i * 12 + i] = d0[i] * (1.0 + 0.2 * wobble(b, 23, i));
    }
    try {
        PoseUKF sharded(B, x0.data(), p0.data(), devices);
        PoseUKF single(B, x0.data(), p0.data(), devices[0]);
        printf("shards %d batch %lld\n", ukfb_shard_count(sharded.handle()), (long long)ukfb_batch(sharded.handle()));
        drive(sharded, B);
        drive(single, B);
        std::vector<PoseUKF::State> xs(n), xr(n);
        std::vector<PoseUKF::Covariance> ps(n), pr(n);
        if (!sharded.getCurrentState(xs.data(), ps.data()) || !single.getCurrentState(xr.data(), pr.data())) return 2;
        const bool same = memcmp(xs.data(), xr.data(), sizeof(PoseUKF::State) * size_t(B)) == 0 &&
                          memcmp(ps.data(), pr.data(), sizeof(PoseUKF::Covariance) * size_t(B)) == 0;
        printf("sharded_equals_single %d\n", int(same));
        std::vector<int64_t> tl(n);
        if (ukfb_get_last_time(sharded.handle(), tl.data()) != UKFB_OK) return 3;
        for (int64_t b = 0; b < B; b += (B > 16 ? B / 16 : 1)) {
            printf("filter %lld %lld", (long long)b, (long long)tl[size_t(b)]);
            for (int i = 0; i < 13; ++i) printf(" %.17g", xs[size_t(b)].v[i]);
            for (int i = 0; i < 144; ++i) printf(" %.17g", ps[size_t(b)].v[i]);
            printf("\n");
        }
        return same ? 0 : 1;
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 3;
    }
}
