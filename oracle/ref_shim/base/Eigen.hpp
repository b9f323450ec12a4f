/* oracle/ref_shim -- TEST INFRASTRUCTURE: base::Vector3d (OrientationUKFConfig.hpp) */
#ifndef REF_SHIM_BASE_EIGEN
#define REF_SHIM_BASE_EIGEN
#include <Eigen/Core>
namespace base {
typedef Eigen::Matrix<double, 3, 1> Vector3d;
}
#endif
