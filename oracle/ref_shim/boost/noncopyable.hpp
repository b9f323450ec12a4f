/* oracle/ref_shim -- TEST INFRASTRUCTURE: boost::noncopyable */
#ifndef REF_SHIM_BOOST_NONCOPYABLE
#define REF_SHIM_BOOST_NONCOPYABLE
namespace boost {
class noncopyable {
protected:
    noncopyable() {}
    ~noncopyable() {}
    noncopyable(const noncopyable&) = delete;
    noncopyable& operator=(const noncopyable&) = delete;
};
}
#endif
