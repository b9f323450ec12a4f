// feeders_demo.cpp -- exercises the host-side feeders (GeographicProjection, GravitationalModel,
// StreamAlignmentVerifier mirrors) and prints values for tests/test_host_feeders.py.  No GPU involved.
#include <algorithm>
#include <cstdio>
#include <sstream>

#include <pose_estimation_b200/GeographicProjection.hpp>
#include <pose_estimation_b200/GravitationalModel.hpp>
#include <pose_estimation_b200/OrientationUKFConfig.hpp>
#include <pose_estimation_b200/StreamAlignmentVerifier.hpp>

using namespace pose_estimation_b200;

int main()
{
    // the reference's own test sequence (test/test_coordinate_projection.cpp:8-54)
    const double latitude = 0.92698121, longitude = 0.154595663;
    GeographicProjection projection(latitude, longitude);
    double x, y, lat2, lon2;
    printf("ref_identity_ok %d\n", int(projection.worldToNav(latitude, longitude, x, y)));
    printf("ref_identity %.17g %.17g\n", x, y);
    printf("ref_inverse_ok %d\n", int(projection.navToWorld(x, y, lat2, lon2)));
    printf("ref_inverse %.17g %.17g\n", lat2, lon2);
    GeographicProjection projection2(latitude, longitude, -500., 1234.);
    projection2.worldToNav(latitude, longitude, x, y);
    printf("ref_offset %.17g %.17g\n", x, y);
    projection2.navToWorld(x, y, lat2, lon2);
    printf("ref_offset_inverse %.17g %.17g\n", lat2, lon2);
    double x2, y2;
    projection2.worldToNav(latitude + 0.1, longitude + 0.1, x2, y2);
    printf("ref_plus %.17g %.17g\n", x2, y2);
    projection2.navToWorld(x - 10000, y - 10000, lat2, lon2);
    printf("ref_minus %.17g %.17g\n", lat2, lon2);

    // a grid around the reference position: forward and back
    for (int i = -2; i <= 2; ++i)
        for (int j = -2; j <= 2; ++j) {
            const double la = latitude + 0.005 * i, lo = longitude + 0.008 * j;
            projection.worldToNav(la, lo, x, y);
            projection.navToWorld(x, y, lat2, lon2);
            printf("grid %.17g %.17g %.17g %.17g %.17g %.17g\n", la, lo, x, y, lat2, lon2);
        }
    double bx, by;
    printf("domain %d %d\n", int(projection.worldToNav(1.5707963267948966, longitude, bx, by)), int(projection.worldToNav(latitude, longitude + 2.0, bx, by)));

    // gravity
    const double lats[4] = {0.0, 0.5, 0.92698121, 1.5707963267948966};
    for (double la : lats)
        for (double alt : {0.0, 100.0, -50.0}) printf("gravity %.17g %.17g %.17g\n", la, alt, GravitationalModel::WGS_84(la, alt));
    printf("earthw %.17g\n", EARTHW);

    // stream alignment verifier
    std::ostringstream log;
    StreamAlignmentVerifier v;
    v.setLogStream(&log);
    StreamAlignerStatus st;
    st.streams.resize(4);
    const char* names[4] = {"imu", "dvl", "gps", "depth"};
    for (int i = 0; i < 4; ++i) st.streams[i].name = names[i];
    unsigned fail = 99, crit = 99;
    auto feed = [&](int64_t t, const size_t recv[4], const size_t drop[4]) {
        st.time_us = t;
        for (int i = 0; i < 4; ++i) {
            st.streams[i].samples_received = recv[i];
            st.streams[i].samples_dropped_buffer_full = drop[i] / 2;
            st.streams[i].samples_dropped_late_arriving = drop[i] - drop[i] / 2;
        }
        v.verifyStreamAlignerStatus(st, fail, crit);
        printf("verifier %lld %u %u\n", (long long)t, fail, crit);
    };
    const size_t r0[4] = {100, 10, 2, 0}, d0[4] = {0, 0, 0, 0};
    feed(1000000, r0, d0);   // 1 s after 0: interval (2 s) not elapsed, outputs untouched
    feed(2500000, r0, d0);   // first verification: every stream is only recorded
    const size_t r1[4] = {2100, 40, 4, 0}, d1[4] = {100, 20, 2, 0};
    feed(3000000, r1, d1);   // too early
    feed(5000000, r1, d1);   // imu 5 % ok, dvl 20/30 -> failure, gps 2 new samples -> too few, depth never seen
    const size_t r2[4] = {4100, 70, 20, 50}, d2[4] = {150, 50, 18, 0};
    feed(7500000, r2, d2);   // dvl 30/30 -> critical, gps 16/16 -> critical, depth first seen
    v.setDropRateWarningThreshold(0.01);
    const size_t r3[4] = {6100, 100, 40, 100}, d3[4] = {250, 50, 18, 0};
    feed(10000000, r3, d3);  // imu 100/2000 = 5 % > 1 % -> failure
    const std::string lines = log.str();
    printf("verifier_log_lines %zu\n", size_t(std::count(lines.begin(), lines.end(), '\n')));
    OrientationUKFConfig cfg = {};
    cfg.location.latitude = latitude;
    printf("config %.17g %zu\n", cfg.location.latitude, sizeof(cfg) / sizeof(double));
    return 0;
}
