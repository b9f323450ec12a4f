/*
 * oracle/ref_shim/mtk/src/SubManifold.hpp -- TEST INFRASTRUCTURE.  MTK::SubManifold<T, idx>: a member of a compound
 * manifold that knows its start index in the tangent vector; MTK::subblock / MTK::setDiagonal address the matching
 * diagonal block of a covariance through a pointer to that member (PoseUKF.cpp:104-107,184-185).
 */
#ifndef REF_SHIM_MTK_SUBMANIFOLD
#define REF_SHIM_MTK_SUBMANIFOLD

#include <Eigen/Core>

namespace MTK {

template <class T, int idx>
struct SubManifold : public T {
    enum { IDX = idx, DIM = T::DOF };
    SubManifold() : T() {}
    SubManifold(const T& t) : T(t) {}
    SubManifold& operator=(const T& t)
    {
        T::operator=(t);
        return *this;
    }
};

/* SubManifold<T, idx> spelled without a top-level comma (the manifold macro passes member declarations through
 * the preprocessor) */
template <int idx>
struct shim_at {
    template <class T>
    using sub = SubManifold<T, idx>;
};

template <class S, int N, class Base, class T, int idx>
Eigen::FixedBlock<S, N, N, T::DOF, T::DOF> subblock(Eigen::Matrix<S, N, N>& cov, SubManifold<T, idx> Base::*)
{
    return cov.template fixed_block<T::DOF, T::DOF>(idx, idx);
}

template <class S, int N, class Base, class T, int idx>
void setDiagonal(Eigen::Matrix<S, N, N>& cov, SubManifold<T, idx> Base::*, const S& val)
{
    for (int i = 0; i < T::DOF; ++i) cov(idx + i, idx + i) = val;
}

}  // namespace MTK

#endif
