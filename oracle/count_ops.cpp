/*
 * oracle/count_ops.cpp -- CPU ORACLE tooling.  TEST INFRASTRUCTURE ONLY.
 *
 * Runs the oracle with a counting scalar to MEASURE the algorithmic work per
 * filter-step that SURVEY.md section 8(d) / Appendix B estimate by hand
 * (add/sub/mul = 1 flop, nothing fused; sqrt, div, sin, cos, atan counted
 * separately as special-function evaluations).  DESIGN.md quotes this output.
 *
 *   ./count_ops            -> one JSON object on stdout
 */
#include <cstdio>

namespace cnt {
struct Counts {
    unsigned long long add = 0, mul = 0, div = 0, sqrt_ = 0, sin_ = 0, cos_ = 0, atan_ = 0, cmp = 0;
};
static Counts g;
}  // namespace cnt

struct C {
    double v;
    C() : v(0) {}
    C(double x) : v(x) {}
    explicit operator double() const { return v; }
};
inline C operator+(const C& a, const C& b) { cnt::g.add++; return C(a.v + b.v); }
inline C operator-(const C& a, const C& b) { cnt::g.add++; return C(a.v - b.v); }
inline C operator*(const C& a, const C& b) { cnt::g.mul++; return C(a.v * b.v); }
inline C operator/(const C& a, const C& b) { cnt::g.div++; return C(a.v / b.v); }
inline C operator-(const C& a) { return C(-a.v); }
inline bool operator<(const C& a, const C& b) { cnt::g.cmp++; return a.v < b.v; }
inline bool operator>(const C& a, const C& b) { cnt::g.cmp++; return a.v > b.v; }
inline bool operator>=(const C& a, const C& b) { cnt::g.cmp++; return a.v >= b.v; }
inline bool operator<=(const C& a, const C& b) { cnt::g.cmp++; return a.v <= b.v; }

#include <cmath>
inline C sqrt(const C& a) { cnt::g.sqrt_++; return C(std::sqrt(a.v)); }
inline C sin(const C& a) { cnt::g.sin_++; return C(std::sin(a.v)); }
inline C cos(const C& a) { cnt::g.cos_++; return C(std::cos(a.v)); }
inline C atan(const C& a) { cnt::g.atan_++; return C(std::atan(a.v)); }

#include "ukf_oracle.hpp"

using namespace orc;

static void report(const char* name, const cnt::Counts& a, const cnt::Counts& b, bool last = false)
{
    const unsigned long long flops = (b.add - a.add) + (b.mul - a.mul);
    const unsigned long long spec = (b.div - a.div) + (b.sqrt_ - a.sqrt_) + (b.sin_ - a.sin_) + (b.cos_ - a.cos_) +
                                    (b.atan_ - a.atan_);
    std::printf(
        "  \"%s\": {\"flops\": %llu, \"add\": %llu, \"mul\": %llu, \"specials\": %llu, \"div\": %llu, "
        "\"sqrt\": %llu, \"sin\": %llu, \"cos\": %llu, \"atan\": %llu}%s\n",
        name, flops, b.add - a.add, b.mul - a.mul, spec, b.div - a.div, b.sqrt_ - a.sqrt_, b.sin_ - a.sin_,
        b.cos_ - a.cos_, b.atan_ - a.atan_, last ? "" : ",");
}

int main()
{
    std::printf("{\n");
    {
        /* PoseUKF, config C3-like state */
        PoseState<C> s;
        const double mu[13] = {0.1, -0.2, 0.3, 0.01, 0.02, 0.03, 0.9993, 1, 0.05, -0.02, 0.01, -0.01, 0.05};
        s.load(mu);
        C sig[144];
        for (int i = 0; i < 144; ++i) sig[i] = C(0);
        const double d[12] = {1, 1, 1, .01, .01, .01, .1, .1, .1, .01, .01, .01};
        for (int i = 0; i < 12; ++i) sig[i * 12 + i] = C(d[i]);
        PoseFilter<C> f(s, sig);
        f.predictionStep(1e-3);
        const double z[3] = {0.01, -0.01, 0.05}, R[9] = {1e-6, 0, 0, 0, 1e-6, 0, 0, 0, 1e-6};
        f.integrateMeasurement(MEAS_POSE_ANGULAR_VELOCITY, z, R);
        /* steady state-ish: measure the second step */
        cnt::Counts a = cnt::g;
        f.predictionStep(1e-3);
        cnt::Counts b = cnt::g;
        f.integrateMeasurement(MEAS_POSE_ANGULAR_VELOCITY, z, R);
        cnt::Counts c = cnt::g;
        const double z2[2] = {0.1, 0.2}, R2[4] = {0.25, 0, 0, 0.25};
        f.integrateMeasurement(MEAS_POSE_XY, z2, R2);
        cnt::Counts d2 = cnt::g;
        const double z1[1] = {0.3}, R1[1] = {0.25};
        f.integrateMeasurement(MEAS_POSE_Z, z1, R1);
        cnt::Counts d1 = cnt::g;
        const double zo[3] = {0.02, 0.04, 0.06};
        f.integrateMeasurement(MEAS_POSE_ORIENTATION, zo, R);
        cnt::Counts d3 = cnt::g;
        report("pose_predict", a, b);
        report("pose_update_m3", b, c);
        report("pose_predict_plus_update_m3", a, c);
        report("pose_update_m2", c, d2);
        report("pose_update_m1", d2, d1);
        report("pose_update_so3", d1, d3);
        std::printf("  \"pose_mean_pass_hist\": [%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu],\n",
                    (unsigned long long)f.ukf.mean_iters[0], (unsigned long long)f.ukf.mean_iters[1],
                    (unsigned long long)f.ukf.mean_iters[2], (unsigned long long)f.ukf.mean_iters[3],
                    (unsigned long long)f.ukf.mean_iters[4], (unsigned long long)f.ukf.mean_iters[5],
                    (unsigned long long)f.ukf.mean_iters[6], (unsigned long long)f.ukf.mean_iters[7]);
    }
    {
        /* OrientationUKF, config C1 state */
        OrientationState<C> s;
        const double mu[14] = {0.01, 0.02, 0.03, 0.9993, 0.1, 0.0, -0.1, 1e-4, 0, 0, 1e-3, 0, 0, 9.81};
        s.load(mu);
        C sig[169];
        for (int i = 0; i < 169; ++i) sig[i] = C(0);
        const double d[13] = {.01, .01, .01, .01, .01, .01, 1e-6, 1e-6, 1e-6, 1e-4, 1e-4, 1e-4, 1e-4};
        for (int i = 0; i < 13; ++i) sig[i * 13 + i] = C(d[i]);
        OrientationFilter<C> f(s, sig, 3600., 3600., 0.92698121);
        const double qd[13] = {1e-6, 1e-6, 1e-6, 1e-4, 1e-4, 1e-4, 1e-10, 1e-10, 1e-10, 1e-8, 1e-8, 1e-8, 1e-12};
        for (int i = 0; i < 13; ++i) f.process_noise_cov[i * 13 + i] = C(qd[i]);
        const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        const double gyro[3] = {0.001, -0.002, 0.05}, acc[3] = {0.05, -0.02, 9.80};
        f.setRotationRate(gyro, I3);
        f.setAcceleration(acc, I3);
        f.predictionStep(1e-3);
        const double z[3] = {0.1, 0.0, -0.1}, R[9] = {1e-4, 0, 0, 0, 1e-4, 0, 0, 0, 1e-4};
        f.integrateVelocity(z, R);
        cnt::Counts a = cnt::g;
        f.predictionStep(1e-3);
        cnt::Counts b = cnt::g;
        f.integrateVelocity(z, R);
        cnt::Counts c = cnt::g;
        report("orientation_predict", a, b);
        report("orientation_update_velocity", b, c);
        report("orientation_predict_plus_update", a, c, true);
    }
    std::printf("}\n");
    return 0;
}
