/*
 * BodyStateMeasurement.hpp -- mirror of pose_estimation::BodyStateMeasurement
 * (reference src/pose_with_velocity/BodyStateMeasurement.hpp:12-41) for batched PoseUKF objects.
 *
 * The reference converts one base::samples::RigidBodyState to / from (PoseWithVelocity, 12 x 12 covariance) on the
 * host, and its callers then hand the result to initializeFilter or read it from getCurrentState.  base-types is
 * absent from this image, so RigidBodyState here is the plain struct below (the members the reference touches, in
 * the order of the ABI record); the conversion of a whole batch runs on the device next to the filter records
 * (ukfb_initialize_from_body_states / ukfb_get_body_states) -- there is no host implementation.
 */
#ifndef POSE_ESTIMATION_B200_BODY_STATE_MEASUREMENT_HPP
#define POSE_ESTIMATION_B200_BODY_STATE_MEASUREMENT_HPP

#include "PoseUKF.hpp"

namespace pose_estimation_b200
{

struct RigidBodyState {
    double position[3];
    double orientation[4]; /* x, y, z, w */
    double velocity[3];
    double angular_velocity[3];
    double cov_position[9], cov_orientation[9], cov_velocity[9], cov_angular_velocity[9]; /* 3 x 3 */
};
static_assert(sizeof(RigidBodyState) == UKFB_RBS_DOUBLES * sizeof(double), "RigidBodyState must match the ABI record");

struct BodyStateMeasurement {
    /* fromRigidBodyState (:14-26) for each of the filter's batch() body states, then initializeFilter: position,
     * orientation, velocity and angular velocity as they are; covariance = the four blocks, zero elsewhere */
    static void fromRigidBodyState(const RigidBodyState* body_states, PoseUKF& filter)
    {
        if (ukfb_initialize_from_body_states(filter.handle(), body_states->position) != UKFB_OK)
            throw std::logic_error(std::string("ukf_batch: ") + ukfb_last_error());
    }
    /* getCurrentState, then toRigidBodyState (:28-39): velocity rotated into the navigation frame, covariance blocks
     * copied unrotated.  @returns false if the filter has not been initialized */
    static bool toRigidBodyState(PoseUKF& filter, RigidBodyState* body_states)
    {
        const int rc = ukfb_get_body_states(filter.handle(), body_states->position);
        if (rc == UKFB_ERR_NOT_INITIALIZED) return false;
        if (rc != UKFB_OK) throw std::logic_error(std::string("ukf_batch: ") + ukfb_last_error());
        return true;
    }
};

}  // namespace pose_estimation_b200

#endif
