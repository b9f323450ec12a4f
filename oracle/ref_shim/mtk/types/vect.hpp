/*
 * oracle/ref_shim/mtk/types/vect.hpp -- TEST INFRASTRUCTURE.  MTK::vect<D, scalar>: R^D as a manifold
 * (boxplus: x += s v; boxminus: x - o), restated from the published MTK; `slam/mtk` itself is
 * not vendored by the reference and absent here.
 */
#ifndef REF_SHIM_MTK_VECT
#define REF_SHIM_MTK_VECT

#include <Eigen/Core>

namespace MTK {

template <int D, class S = double>
struct vect : public Eigen::Matrix<S, D, 1> {
    typedef Eigen::Matrix<S, D, 1> base;
    typedef S scalar;
    enum { DOF = D };
    vect() : base() {} /* MTK: vect(const base& src = base::Zero()) */
    vect(const base& src) : base(src) {}
    void boxplus(const S* vec, S scale = S(1))
    {
        for (int i = 0; i < D; ++i) (*this)[i] = (*this)[i] + scale * vec[i];
    }
    void boxplus(const base& vec, S scale = S(1)) { boxplus(vec.data(), scale); }
    void boxminus(S* res, const vect& other) const
    {
        for (int i = 0; i < D; ++i) res[i] = (*this)[i] - other[i];
    }
};

}  // namespace MTK

#endif
