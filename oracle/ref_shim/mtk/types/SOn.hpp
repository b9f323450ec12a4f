/*
 * oracle/ref_shim/mtk/types/SOn.hpp -- TEST INFRASTRUCTURE.  MTK::SO3<scalar>: a unit quaternion as a 3-DOF manifold,
 * restated from the published MTK: exp / log through cos_sinc_sqrt and the atan form, frame
 * convention per UKFB_SO3_BOXPLUS_LEFT (include/ukfb_constants.h).  The arithmetic is the CPU oracle's
 * (oracle/ukf_oracle.hpp so3_*): this layer is a RESTATEMENT, not reference text.
 */
#ifndef REF_SHIM_MTK_SON
#define REF_SHIM_MTK_SON

#include <Eigen/Geometry>

#include "vect.hpp"

namespace MTK {

template <class S = double>
struct SO3 : public Eigen::Quaternion<S> {
    typedef Eigen::Quaternion<S> base;
    typedef S scalar;
    typedef vect<3, S> vect_type;
    enum { DOF = 3 };
    SO3() : base() {}
    SO3(const base& q) : base(q) {}
    void boxplus(const S* vec, S scale = S(1)) { orc::so3_boxplus(this->raw(), vec, scale); }
    void boxplus(const Eigen::Matrix<S, 3, 1>& vec, S scale = S(1)) { boxplus(vec.data(), scale); }
    void boxminus(S* res, const SO3& other) const { orc::so3_boxminus(this->raw(), other.raw(), res); }
    static SO3 exp(const Eigen::Matrix<S, 3, 1>& vec, S scale = S(1)) { return SO3(base(orc::so3_exp(vec.data(), scale))); }
    static Eigen::Matrix<S, 3, 1> log(const SO3& q)
    {
        Eigen::Matrix<S, 3, 1> r;
        orc::so3_log(q.raw(), r.data());
        return r;
    }
};

}  // namespace MTK

#endif
