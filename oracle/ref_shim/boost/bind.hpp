/* oracle/ref_shim -- TEST INFRASTRUCTURE: boost::bind and the global placeholders _1 .. (boost/bind.hpp puts them in
 * the global namespace, which is how PoseUKF.cpp:114 writes `boost::bind(measurementPosition<WState>, _1)`) */
#ifndef REF_SHIM_BOOST_BIND
#define REF_SHIM_BOOST_BIND
#include <functional>
namespace boost {
using std::bind;
}
using namespace std::placeholders;
#endif
