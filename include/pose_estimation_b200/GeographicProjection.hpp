/*
 * GeographicProjection.hpp -- host-side mirror of pose_estimation::GeographicProjection
 * (reference src/GeographicProjection.hpp:14-32, GeographicProjection.cpp:6-44): GPS latitude / longitude (rad, WGS-84)
 * <-> a local transverse-Mercator plane in NWU coordinates (x north, y west) around a reference position.  Stays on the
 * host (north star); its XY output is what callers put into PoseUKF::XYMeasurement (PoseUKF.cpp:119-124).
 *
 * The reference delegates to GDAL/OGR: SetTM(latitude, longitude, 0.9996, 0, 0) on WGS-84 (GeographicProjection.cpp:15),
 * Transform(lon, lat) -> (easting, northing), then x = offset.x + northing, y = offset.y - easting (:29-37).  GDAL is not
 * in this image, so the transverse Mercator itself is restated here: the Krueger series in the third flattening n to
 * n^6 (Karney, "Transverse Mercator with an accuracy of a few nanometers", J. Geodesy 2011 -- the same series PROJ's
 * tmerc evaluates), scale 0.9996, origin at the reference latitude on the central meridian.  Pinned by the reference's
 * own test values (test/test_coordinate_projection.cpp:11-53) and an independent meridian-arc / Snyder-series check in
 * tests/test_host_feeders.py.
 */
#ifndef POSE_ESTIMATION_B200_GEOGRAPHIC_PROJECTION_HPP
#define POSE_ESTIMATION_B200_GEOGRAPHIC_PROJECTION_HPP

#include <cmath>

namespace pose_estimation_b200
{

class GeographicProjection
{
public:
    /* reference position in rad; (x, y): nav-frame coordinates of that position (GeographicProjection.cpp:6) */
    GeographicProjection(double latitude, double longitude, double x = 0., double y = 0.)
        : lat0(latitude), lon0(longitude), off_x(x), off_y(y)
    {
        const double f = 1.0 / 298.257223563; /* WGS-84 */
        const double a = 6378137.0;
        k0 = 0.9996;
        e = std::sqrt(f * (2.0 - f));
        const double n = f / (2.0 - f), n2 = n * n, n3 = n2 * n, n4 = n3 * n, n5 = n4 * n, n6 = n5 * n;
        A = a / (1.0 + n) * (1.0 + n2 / 4.0 + n4 / 64.0 + n6 / 256.0);
        al[0] = n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800;
        al[1] = 13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360;
        al[2] = 61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440;
        al[3] = 49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600;
        al[4] = 34729 * n5 / 80640 - 3418889 * n6 / 1995840;
        al[5] = 212378941 * n6 / 319334400;
        be[0] = n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800;
        be[1] = n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720;
        be[2] = 17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720;
        be[3] = 4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600;
        be[4] = 4583 * n5 / 161280 - 108847 * n6 / 3991680;
        be[5] = 20648693 * n6 / 638668800;
        double east;
        north0 = 0.0;
        forward(lat0, 0.0, east, north0); /* northing of the origin latitude on the central meridian */
    }
    virtual ~GeographicProjection() {}
    GeographicProjection(const GeographicProjection&) = delete; /* boost::noncopyable, GeographicProjection.hpp:16 */
    GeographicProjection& operator=(const GeographicProjection&) = delete;

    /* GeographicProjection.cpp:29-37.  @returns false outside the projection's domain (as a failed OGR transform) */
    bool worldToNav(double latitude, double longitude, double& x, double& y)
    {
        double east, north;
        if (!(std::fabs(latitude) < 1.5707963267948966) || !std::isfinite(longitude)) return false;
        const double dl = std::remainder(longitude - lon0, 6.283185307179586476925286766559);
        if (std::fabs(dl) >= 1.5707963267948966) return false;
        forward(latitude, dl, east, north);
        x = off_x + (north - north0);
        y = off_y - east;
        return std::isfinite(x) && std::isfinite(y);
    }

    /* GeographicProjection.cpp:39-44 */
    bool navToWorld(double x, double y, double& latitude, double& longitude)
    {
        const double north = x - off_x, east = off_y - y;
        if (!std::isfinite(north) || !std::isfinite(east)) return false;
        double dl;
        inverse(east, north + north0, latitude, dl);
        longitude = lon0 + dl;
        return std::isfinite(latitude) && std::isfinite(longitude);
    }

protected:
    /* (latitude, longitude difference to the central meridian) -> (easting, northing from the equator) */
    void forward(double phi, double dl, double& east, double& north) const
    {
        const double tau = std::tan(phi);
        const double sigma = std::sinh(e * std::atanh(e * tau / std::sqrt(1.0 + tau * tau)));
        const double taup = tau * std::sqrt(1.0 + sigma * sigma) - sigma * std::sqrt(1.0 + tau * tau);
        const double xip = std::atan2(taup, std::cos(dl));
        const double etap = std::asinh(std::sin(dl) / std::sqrt(taup * taup + std::cos(dl) * std::cos(dl)));
        double xi = xip, eta = etap;
        for (int j = 1; j <= 6; ++j) {
            xi += al[j - 1] * std::sin(2 * j * xip) * std::cosh(2 * j * etap);
            eta += al[j - 1] * std::cos(2 * j * xip) * std::sinh(2 * j * etap);
        }
        east = k0 * A * eta;
        north = k0 * A * xi;
    }

    void inverse(double east, double north, double& phi, double& dl) const
    {
        const double xi = north / (k0 * A), eta = east / (k0 * A);
        double xip = xi, etap = eta;
        for (int j = 1; j <= 6; ++j) {
            xip -= be[j - 1] * std::sin(2 * j * xi) * std::cosh(2 * j * eta);
            etap -= be[j - 1] * std::cos(2 * j * xi) * std::sinh(2 * j * eta);
        }
        const double sh = std::sinh(etap), c = std::cos(xip);
        const double taup = std::sin(xip) / std::sqrt(sh * sh + c * c);
        dl = std::atan2(sh, c);
        /* tau from the conformal tau' by Newton (Karney eq. 19-21) */
        const double e2m = 1.0 - e * e;
        double tau = taup;
        for (int it = 0; it < 8; ++it) {
            const double sigma = std::sinh(e * std::atanh(e * tau / std::sqrt(1.0 + tau * tau)));
            const double tpi = tau * std::sqrt(1.0 + sigma * sigma) - sigma * std::sqrt(1.0 + tau * tau);
            const double d = (taup - tpi) / std::sqrt(1.0 + tpi * tpi) * (1.0 + e2m * tau * tau) / (e2m * std::sqrt(1.0 + tau * tau));
            tau += d;
            if (std::fabs(d) <= 1e-17 * (1.0 + std::fabs(tau))) break;
        }
        phi = std::atan(tau);
    }

    double lat0, lon0, off_x, off_y;
    double k0, e, A, north0;
    double al[6], be[6];
};

}  // namespace pose_estimation_b200

#endif
