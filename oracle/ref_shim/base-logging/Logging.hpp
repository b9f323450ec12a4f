/* oracle/ref_shim -- TEST INFRASTRUCTURE: OrientationUKF.cpp includes base-logging but logs nothing */
#ifndef REF_SHIM_BASE_LOGGING
#define REF_SHIM_BASE_LOGGING
#endif
