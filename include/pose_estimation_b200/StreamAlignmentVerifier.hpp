/*
 * StreamAlignmentVerifier.hpp -- host-side mirror of pose_estimation::StreamAlignmentVerifier
 * (reference src/StreamAlignmentVerifier.hpp:12-38, StreamAlignmentVerifier.cpp:6-72): a periodic drop-rate check over
 * the status of the sample queues that feed the filters.  No filter arithmetic; stays on the host (north star).
 * aggregator::StreamAlignerStatus is not in this image: the two plain structs below carry the members the reference
 * reads.  Log lines go to std::cerr instead of base-logging.
 */
#ifndef POSE_ESTIMATION_B200_STREAM_ALIGNMENT_VERIFIER_HPP
#define POSE_ESTIMATION_B200_STREAM_ALIGNMENT_VERIFIER_HPP

#include <cstddef>
#include <cstdint>
#include <iostream>
#include <map>
#include <string>
#include <vector>

namespace pose_estimation_b200
{

struct StreamStatus { /* aggregator::StreamStatus: the members read at StreamAlignmentVerifier.cpp:28-36 */
    std::string name;
    size_t samples_received = 0;
    size_t samples_dropped_buffer_full = 0;
    size_t samples_dropped_late_arriving = 0;
    size_t samples_backward_in_time = 0;
};

struct StreamAlignerStatus { /* aggregator::StreamAlignerStatus: :19, :23 */
    int64_t time_us = 0;
    std::vector<StreamStatus> streams;
};

class StreamAlignmentVerifier
{
public:
    StreamAlignmentVerifier() /* StreamAlignmentVerifier.cpp:6-13 */
        : aligner_last_verified_us(0), verification_interval(2.0), drop_rate_warning(0.5), drop_rate_critical(1.0), min_new_samples(5), log(&std::cerr)
    {
    }
    virtual ~StreamAlignmentVerifier() {}

    /* StreamAlignmentVerifier.cpp:15-66: once per verification interval, per stream: the share of samples dropped since
     * the last check; >= critical counts as critical, > warning as failure; a stream is judged only with more than
     * min_new_samples new samples; a stream first seen (stored count 0) is only recorded.  The counters are left
     * untouched when the interval has not elapsed. */
    void verifyStreamAlignerStatus(const StreamAlignerStatus& status, unsigned& streams_with_alignment_failures,
                                   unsigned& streams_with_critical_alignment_failures)
    {
        if (double(status.time_us - aligner_last_verified_us) / 1e6 <= verification_interval) return;
        streams_with_alignment_failures = 0;
        streams_with_critical_alignment_failures = 0;
        for (const StreamStatus& s : status.streams) {
            if (aligner_samples_received[s.name] == 0) {
                aligner_samples_received[s.name] = s.samples_received;
                continue;
            }
            const size_t fresh = s.samples_received - aligner_samples_received[s.name];
            const size_t dropped = s.samples_dropped_buffer_full + s.samples_dropped_late_arriving + s.samples_backward_in_time;
            const size_t fresh_dropped = dropped - aligner_samples_dropped[s.name];
            if (fresh > min_new_samples) {
                const double rate = double(fresh_dropped) / double(fresh);
                if (rate >= drop_rate_critical) {
                    ++streams_with_critical_alignment_failures;
                    report("Critical transformation alignment failure in stream ", s.name, rate);
                } else if (rate > drop_rate_warning) {
                    ++streams_with_alignment_failures;
                    report("Transformation alignment failure in stream ", s.name, rate);
                }
            } else if (log)
                *log << "To few samples received to validate the drop rate in stream " << s.name << std::endl;
            aligner_samples_received[s.name] = s.samples_received;
            aligner_samples_dropped[s.name] = dropped;
        }
        aligner_last_verified_us = status.time_us;
    }
    void verifyStreamAlignerStatus(const StreamAlignerStatus& status, unsigned& streams_with_alignment_failures) /* :68-72 */
    {
        unsigned critical;
        verifyStreamAlignerStatus(status, streams_with_alignment_failures, critical);
    }

    void setVerificationInterval(double v) { verification_interval = v; } /* StreamAlignmentVerifier.hpp:21-26 */
    double getVerificationInterval() { return verification_interval; }
    void setDropRateWarningThreshold(double v) { drop_rate_warning = v; }
    double getDropRateWarningThreshold() { return drop_rate_warning; }
    void setDropRateCriticalThreshold(double v) { drop_rate_critical = v; }
    double getDropRateCriticalThreshold() { return drop_rate_critical; }
    void setLogStream(std::ostream* os) { log = os; } /* nullptr = quiet */

protected:
    void report(const char* what, const std::string& name, double rate)
    {
        if (log)
            *log << what << name << ". " << rate * 100.0 << "% of all samples were dropped in the last " << verification_interval
                 << " seconds." << std::endl;
    }

    std::map<std::string, size_t> aligner_samples_received;
    std::map<std::string, size_t> aligner_samples_dropped;
    int64_t aligner_last_verified_us;
    double verification_interval;
    double drop_rate_warning;
    double drop_rate_critical;
    unsigned min_new_samples;
    std::ostream* log;
};

}  // namespace pose_estimation_b200

#endif
