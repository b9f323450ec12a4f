/*
 * Measurement.hpp -- the reference's MEASUREMENT(NAME, DIM) macro (src/Measurement.hpp:6-16) without Eigen:
 * a POD {mu, cov} with mu = 0 and cov = identity defaults, double, cov row-major (it is symmetric, so the
 * reference's column-major Eigen storage reads the same).  With Eigen available a caller maps these arrays
 * with Eigen::Map<Eigen::Matrix<double, DIM, 1>>(m.mu) / Map<Matrix<double, DIM, DIM>>(m.cov).
 */
#ifndef POSE_ESTIMATION_B200_MEASUREMENT_HPP
#define POSE_ESTIMATION_B200_MEASUREMENT_HPP

#define UKFB_MEASUREMENT(NAME, DIM)                                      \
    struct NAME {                                                        \
        enum { Dim = DIM };                                              \
        double mu[DIM];                                                  \
        double cov[DIM * DIM];                                           \
        NAME()                                                           \
        {                                                                \
            for (int i = 0; i < DIM; ++i) mu[i] = 0.0;                   \
            for (int i = 0; i < DIM * DIM; ++i) cov[i] = 0.0;            \
            for (int i = 0; i < DIM; ++i) cov[i * DIM + i] = 1.0;        \
        }                                                                \
    };

#endif
