/*
 * OrientationUKFConfig.hpp -- the plain configuration structs of the reference
 * (src/orientation_estimator/OrientationUKFConfig.hpp:9-49) with base::Vector3d as double[3].  Only
 * LocationConfiguration is consumed on the filter path (OrientationUKF.cpp:42,47); it is defined in OrientationUKF.hpp.
 */
#ifndef POSE_ESTIMATION_B200_ORIENTATION_UKF_CONFIG_HPP
#define POSE_ESTIMATION_B200_ORIENTATION_UKF_CONFIG_HPP

#include "OrientationUKF.hpp"

namespace pose_estimation_b200
{

struct InertialNoiseParameters { /* :9-22 */
    double randomwalk[3];       /* (m/s^2)/sqrt(Hz) for accelerometers, (rad/s)/sqrt(Hz) for gyros */
    double bias_offset[3];      /* initial bias value */
    double bias_instability[3]; /* m/s^2 or rad/s */
    double bias_tau;            /* seconds */
};

struct OrientationUKFConfig { /* :36-49 */
    InertialNoiseParameters acceleration;
    InertialNoiseParameters rotation_rate;
    LocationConfiguration location;
    double max_velocity[3]; /* m/s */
};

}  // namespace pose_estimation_b200

#endif
