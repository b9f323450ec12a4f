"""slam_pose_estimation_b200 -- B200-native batched unscented Kalman filter engine.

The product is lib/libukfb.so (hand-written sm_100a CUDA behind the C ABI of
include/ukf_batch.h).  This package only holds its build recipe, its ctypes binding and
the synthetic stream generators; importing it does not need a GPU, using it does.
"""
from .batch import (  # noqa: F401
    MEAS_NONE, MEAS_ORI_VELOCITY, MEAS_POSE_ANGULAR_VELOCITY, MEAS_POSE_ORIENTATION, MEAS_POSE_POSITION,
    MEAS_POSE_VELOCITY, MEAS_POSE_XVEL_YAWVEL, MEAS_POSE_XY, MEAS_POSE_XY_VELOCITY, MEAS_POSE_Z,
    MEAS_POSE_Z_VELOCITY, ORIENTATION, POSE, UkfBatch, UkfbError,
)
