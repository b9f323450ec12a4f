/* oracle/ref_shim -- TEST INFRASTRUCTURE: boost::shared_ptr as the reference uses it (reset, get, ->) */
#ifndef REF_SHIM_BOOST_SHARED_PTR
#define REF_SHIM_BOOST_SHARED_PTR
#include <memory>
namespace boost {
template <class T>
using shared_ptr = std::shared_ptr<T>;
}
#endif
