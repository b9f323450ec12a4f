/*
 * GravitationalModel.hpp -- host-side mirror of pose_estimation::GravitationalModel and the WGS-84 constants next to
 * it (reference src/GravitationalModel.hpp:10-16, 33-44).  Stays on the host (north star): it produces scalars that
 * callers put into the initial state (gravity) -- the filter path itself only uses EARTHW (OrientationUKF.cpp:47),
 * which the engine shares through ukfb_constants.h.
 */
#ifndef POSE_ESTIMATION_B200_GRAVITATIONAL_MODEL_HPP
#define POSE_ESTIMATION_B200_GRAVITATIONAL_MODEL_HPP

#include <cmath>

#include "../ukfb_constants.h"

namespace pose_estimation_b200
{

static const double EQUATORIAL_RADIUS = 6378137.0;  /* m                                   GravitationalModel.hpp:10 */
static const double ECC = 0.0818191908426;          /* first eccentricity                  :11 */
static const double GRAVITY = 9.79766542;           /* mean gravity, WGS-84, m/s^2         :12 */
static const double GRAVITY_SI = 9.80665;           /* standard gravity                    :13 */
static const double GWGS0 = 9.7803267714;           /* gravity at the equator              :14 */
static const double GWGS1 = 0.00193185138639;       /* gravity formula constant            :15 */
static const double EARTHW = UKFB_EARTHW;           /* earth angular velocity, 2 pi / 86164 rad/s   :16 */

class GravitationalModel
{
public:
    /* theoretical gravity on the WGS-84 ellipsoid (Somigliana), reduced to `altitude` with the free-air factor
     * (R / (R + h))^2, R = equatorial radius (:33-44).  latitude in rad, altitude in m. */
    static double WGS_84(double latitude, double altitude)
    {
        const double s2 = std::pow(std::sin(latitude), 2);
        const double g0 = GWGS0 * ((1 + GWGS1 * s2) / std::sqrt(1 - std::pow(ECC, 2) * s2));
        return g0 * std::pow(EQUATORIAL_RADIUS / (EQUATORIAL_RADIUS + altitude), 2);
    }
};

}  // namespace pose_estimation_b200

#endif
