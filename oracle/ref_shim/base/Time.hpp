/* oracle/ref_shim -- TEST INFRASTRUCTURE: base::Time of Rock's base-types as far as UnscentedKalmanFilter.hpp uses it
 * (an int64 microsecond count; isNull() == 0; difference; toSeconds()) */
#ifndef REF_SHIM_BASE_TIME
#define REF_SHIM_BASE_TIME
#include <cstdint>
namespace base {
struct Time {
    int64_t microseconds;
    Time() : microseconds(0) {}
    static Time fromMicroseconds(int64_t us)
    {
        Time t;
        t.microseconds = us;
        return t;
    }
    bool isNull() const { return microseconds == 0; }
    double toSeconds() const { return static_cast<double>(microseconds) / 1000000.0; }
    Time operator-(const Time& o) const { return fromMicroseconds(microseconds - o.microseconds); }
};
}
#endif
