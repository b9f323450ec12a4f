// geo_capi.cpp -- the host-side GeographicProjection mirror (include/pose_estimation_b200/GeographicProjection.hpp,
// reference src/GeographicProjection.cpp:29-44) behind two C functions, so that the Python tests can produce
// XYMeasurement coordinates the way the reference's callers do: GPS fix (lat, lon) -> worldToNav -> (x, y).
// TEST INFRASTRUCTURE: built on the fly by tests/test_events.py; the product does not link it.
#include <pose_estimation_b200/GeographicProjection.hpp>

using pose_estimation_b200::GeographicProjection;

extern "C" int geo_world_to_nav(double lat0, double lon0, long n, const double* lat, const double* lon, double* x, double* y)
{
    GeographicProjection p(lat0, lon0);
    int bad = 0;
    for (long i = 0; i < n; ++i) bad += !p.worldToNav(lat[i], lon[i], x[i], y[i]);
    return bad;
}

extern "C" int geo_nav_to_world(double lat0, double lon0, long n, const double* x, const double* y, double* lat, double* lon)
{
    GeographicProjection p(lat0, lon0);
    int bad = 0;
    for (long i = 0; i < n; ++i) bad += !p.navToWorld(x[i], y[i], lat[i], lon[i]);
    return bad;
}
