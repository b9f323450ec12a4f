/*
 * oracle/ref_shim/ukfom/mtkwrap.hpp -- TEST INFRASTRUCTURE.  ukfom::mtkwrap<M>: M with the operator forms of the
 * manifold operations (x + delta = boxplus, x - y = boxminus) and the typedefs ukfom::ukf wants.
 */
#ifndef REF_SHIM_UKFOM_MTKWRAP
#define REF_SHIM_UKFOM_MTKWRAP

#include <Eigen/Core>

namespace ukfom {

template <class M>
struct mtkwrap : public M {
    typedef mtkwrap<M> self;
    typedef typename M::scalar scalar_type;
    typedef typename M::scalar scalar;
    enum { DOF = M::DOF };
    typedef Eigen::Matrix<scalar_type, DOF, 1> vectorized_type;

    mtkwrap() : M() {}
    mtkwrap(const M& m) : M(m) {}

    self& operator+=(const vectorized_type& delta)
    {
        M::boxplus(delta.data());
        return *this;
    }
    self operator+(const vectorized_type& delta) const
    {
        self r(*this);
        r += delta;
        return r;
    }
    vectorized_type operator-(const self& other) const
    {
        vectorized_type r;
        M::boxminus(r.data(), other);
        return r;
    }
};

}  // namespace ukfom

#endif
