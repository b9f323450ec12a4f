/*
 * oracle/ref_shim/ukfom/ukf.hpp -- TEST INFRASTRUCTURE.  ukfom::ukf<state> with the interface the reference calls
 * (UnscentedKalmanFilter.hpp:23-25,42,55-56; PoseUKF.cpp:114-195; OrientationUKF.cpp:69,88): constructor (mu, sigma),
 * mu(), sigma(), predict(g, Q), update(z, h, R, accept), plus ukfom::id and ukfom::accept_any_mahalanobis_distance.
 *
 * The estimator behind it is the CPU oracle's engine (oracle/ukf_oracle.hpp orc::Ukf: sigma points, manifold mean,
 * covariances, gain, apply_delta as restated in SURVEY.md App. A.2-A.4), instantiated on the MTK-style state the
 * reference declares.  So in oracle/_ref the wrapper layers (time guards, process models, measurement models, process
 * noise shaping, the acceleration branch) are the REFERENCE'S OWN TEXT, compiled unmodified, while this layer and the
 * manifold primitives under mtk/ remain a restatement of the un-vendored `slam/mtk`.
 */
#ifndef REF_SHIM_UKFOM_UKF
#define REF_SHIM_UKFOM_UKF

#include <boost/bind.hpp>

#include <Eigen/Core>
#include <cstdint>
#include <limits>
#include <type_traits>

#include "../../ukf_oracle.hpp"

namespace ukfom {

template <class T>
const T& id(const T& x) { return x; }

template <class scalar>
bool accept_any_mahalanobis_distance(const scalar&) { return true; }

namespace detail {

/* plain vector measurements (Eigen vectors, mtkwrap<vect<M>>): the oracle's Euclidean measurement */
template <class T, int M>
std::integral_constant<int, M> vec_dim(const Eigen::Matrix<T, M, 1>*);
std::integral_constant<int, 0> vec_dim(...);

/* a manifold-valued measurement (RotationType): boxplus / boxminus of the type itself */
template <class Z>
struct ManifoldMeas {
    enum { DOF = Z::DOF, EUCLID = 0 };
    Z z;
    void boxplus(const double* d, const double& s = 1.0) { z.boxplus(d, s); }
    void boxminus(double* res, const ManifoldMeas& o) const { z.boxminus(res, o.z); }
};

}  // namespace detail

template <class state>
class ukf {
public:
    typedef typename state::scalar scalar_type;
    enum { n = state::DOF };
    typedef Eigen::Matrix<scalar_type, n, n> cov;

    ukf(const state& mu, const cov& sigma)
    {
        eng_.mu = mu;
        for (int i = 0; i < n * n; ++i) eng_.sigma[i] = sigma[i];
        sync();
    }
    const state& mu() const { return eng_.mu; }
    const cov& sigma() const { return sigma_; }

    template <class G>
    void predict(G g, const cov& Q)
    {
        eng_.predict([&](const state& x) -> state { return g(x); }, Q.data());
        sync();
    }

    template <class Z, class H, class RF, class Accept>
    void update(const Z& z, H h, RF R, Accept accept)
    {
        typedef decltype(h(eng_.mu)) HZ;
        constexpr int m_vec = decltype(detail::vec_dim(static_cast<const HZ*>(nullptr)))::value;
        const auto Rm = R();
        (void)accept; /* the reference passes accept_any (PoseUKF.cpp:116); orc::Ukf::accept_max_d2 models the slot */
        update_impl<Z, H, HZ>(z, h, Rm.data(), std::integral_constant<int, m_vec>());
        sync();
    }

    /* instrumentation of the engine, for the comparison with the oracle */
    orc::Ukf<double, state>& engine() { return eng_; }
    const orc::Ukf<double, state>& engine() const { return eng_; }

private:
    template <class Z, class H, class HZ, int M>
    void update_impl(const Z& z, H h, const double* R, std::integral_constant<int, M>)
    {
        typedef orc::EuclidMeas<double, M> E;
        E ze;
        const Eigen::Matrix<double, M, 1>& zv = z;
        for (int i = 0; i < M; ++i) ze.a[i] = zv[i];
        eng_.update(ze, [&](const state& x) {
            const Eigen::Matrix<double, M, 1> hv = h(x);
            E r;
            for (int i = 0; i < M; ++i) r.a[i] = hv[i];
            return r;
        }, R);
    }
    template <class Z, class H, class HZ>
    void update_impl(const Z& z, H h, const double* R, std::integral_constant<int, 0>)
    {
        typedef detail::ManifoldMeas<HZ> W;
        W zw;
        zw.z = z;
        eng_.update(zw, [&](const state& x) {
            W r;
            r.z = h(x);
            return r;
        }, R);
    }
    void sync()
    {
        for (int i = 0; i < n * n; ++i) sigma_[i] = eng_.sigma[i];
    }

    orc::Ukf<double, state> eng_;
    cov sigma_;
};

}  // namespace ukfom

#endif
