/* oracle/ref_shim -- TEST INFRASTRUCTURE: base::NaN<T>() (PoseUKF.cpp:109) */
#ifndef REF_SHIM_BASE_FLOAT
#define REF_SHIM_BASE_FLOAT
#include <limits>
namespace base {
template <class T>
inline T NaN() { return std::numeric_limits<T>::quiet_NaN(); }
}
#endif
