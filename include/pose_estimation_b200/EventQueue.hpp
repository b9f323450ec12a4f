/*
 * EventQueue.hpp -- per-filter queues of time-stamped sensor samples in front of a batched filter object.
 *
 * The reference has no such class: its callers (Rock oroGen tasks, implied by manifest.xml:14 and
 * StreamAlignmentVerifier.hpp:7) get one aggregator::StreamAligner callback per sensor sample, in timestamp
 * order, and each callback does
 *     filter.predictionStepFromSampleTime(ts);  filter.integrateMeasurement(sample);
 * (UnscentedKalmanFilter.hpp:83-100, PoseUKF.cpp:112-178, OrientationUKF.cpp:53-72).  With many filters behind
 * one object that loop would cost one kernel launch per sample; this queue collects the callbacks of all filters
 * and flush() hands them to ukfb_run_events, which runs every queued sample of every filter -- in each filter's own
 * order -- in ONE launch.  The result is the same as making the calls one by one with a try / catch around each
 * callback: when predictionStep throws for a sample (negative or too large time step, UnscentedKalmanFilter.hpp:110-122)
 * that callback ends, i.e. the sample is neither integrated nor stored, and the next sample is served normally.
 *
 * Exceptions: like the filter classes, flush() turns the status bits of the launch into the reference's
 * exceptions (first offending condition; all other samples have been integrated).
 */
#ifndef POSE_ESTIMATION_B200_EVENT_QUEUE_HPP
#define POSE_ESTIMATION_B200_EVENT_QUEUE_HPP

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../ukf_batch.h"

namespace pose_estimation_b200
{

class EventQueue
{
public:
    /* `handle` is the filter object's handle(); the queue does not own it */
    explicit EventQueue(ukfb_handle* handle) : h(handle), batch(ukfb_batch(handle)), queues(size_t(batch)) {}

    /* one aggregator callback of filter `filter`: the sample time and the measurement.  `kind` is a UKFB_MEAS_* or
     * storing UKFB_EVENT_* kind of the filter class, or UKFB_MEAS_NONE to advance the time only.  mu: m values,
     * cov: m x m row-major (m = Dim of the measurement struct; 3 for the storing kinds). */
    void push(int64_t filter, int64_t sample_time_us, int kind, const double* mu, const double* cov)
    {
        if (filter < 0 || filter >= batch) throw std::out_of_range("EventQueue::push: filter index");
        Sample s;
        s.ts = sample_time_us;
        s.kind = int8_t(kind);
        const int m = kind >= UKFB_EVENT_POSE_ACCELERATION ? 3 : (kind >= 0 ? ukfb_meas_dim(kind) : 0);
        for (int i = 0; i < 3; ++i) s.mu[i] = 0.0;
        for (int i = 0; i < 9; ++i) s.cov[i] = (i % 4 == 0) ? 1.0 : 0.0;
        for (int i = 0; i < m; ++i) s.mu[i] = mu[i];
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) s.cov[i * 3 + j] = cov[i * m + j];
        queues[size_t(filter)].push_back(s);
    }
    template <class M>
    void push(int64_t filter, int64_t sample_time_us, int kind, const M& measurement)
    {
        push(filter, sample_time_us, kind, measurement.mu, measurement.cov);
    }

    /* the deepest queue = the number of slots the next flush() runs */
    size_t depth() const
    {
        size_t d = 0;
        for (const auto& q : queues) d = std::max(d, q.size());
        return d;
    }

    /* integrates every queued sample (one kernel launch) and empties the queues */
    void flush()
    {
        const size_t K = depth(), B = size_t(batch);
        if (!K) return;
        std::vector<int64_t> ts(K * B, 0);
        std::vector<int8_t> kinds(K * B, int8_t(UKFB_EVENT_IDLE));
        std::vector<double> mu3(K * B * 3, 0.0), cov(K * B * 9, 0.0);
        for (size_t b = 0; b < B; ++b) {
            for (size_t k = 0; k < queues[b].size(); ++k) {
                const Sample& s = queues[b][k];
                const size_t e = k * B + b;
                ts[e] = s.ts;
                kinds[e] = s.kind;
                std::copy(s.mu, s.mu + 3, mu3.begin() + e * 3);
                std::copy(s.cov, s.cov + 9, cov.begin() + e * 9);
            }
            queues[b].clear();
        }
        if (ukfb_run_events(h, int(K), ts.data(), kinds.data(), mu3.data(), cov.data(), 1) != UKFB_OK)
            throw std::logic_error(std::string("ukf_batch: ") + ukfb_last_error());
        int64_t n = 0;
        uint32_t bits = 0;
        if (ukfb_status_summary(h, &n, &bits) != UKFB_OK) throw std::logic_error(std::string("ukf_batch: ") + ukfb_last_error());
        if (!bits) return;
        ukfb_clear_status(h);
        if (bits & UKFB_STATUS_BAD_EVENT) throw std::logic_error("EventQueue: a queued sample kind has no integrateMeasurement overload in this filter class");
        if (bits & UKFB_STATUS_NEG_DT) throw std::runtime_error("Delta time is negative!");
        if (bits & UKFB_STATUS_DT_TOO_LARGE) throw std::runtime_error("Delta time is greater then the allowed maximum!");
        if (bits & UKFB_STATUS_NONFINITE_MEAS) throw std::runtime_error("Measurement or covariance contains non-finite values!");
        if (bits & UKFB_STATUS_NOT_SPD) throw std::runtime_error("ukfom: covariance is not positive definite");
        if (bits & UKFB_STATUS_MEAN_NO_CONVERGE) throw std::runtime_error("ukfom: sigma point mean did not converge");
    }

private:
    struct Sample {
        int64_t ts;
        int8_t kind;
        double mu[3], cov[9];
    };
    ukfb_handle* h;
    int64_t batch;
    std::vector<std::vector<Sample>> queues;
};

}  // namespace pose_estimation_b200

#endif
