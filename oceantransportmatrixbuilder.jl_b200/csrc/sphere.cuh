// sphere.cuh — degree-argument trigonometry and the haversine distance, device side.
// Restates Distances.jl v0.10 `haversine` (third-party, not under /root/reference; compat at
// /root/reference/Project.toml:14) on top of Julia-Base-style sind/cosd: exact quadrant
// reduction in degrees, x/180 then a double-double multiply by pi, fdlibm kernels.
#pragma once
#include "common.cuh"

struct DD {
    double hi, lo;
};
__device__ __forceinline__ DD mulpi_ext(double x) {
    const double m = 3.141592653589793, m_hi = 3.1415926218032837, m_lo = 3.178650954705639e-8;
    const double x_hi = __longlong_as_double(__double_as_longlong(x) & 0xfffffffff8000000ll);
    const double x_lo = x - x_hi;
    const double y_hi = m * x;
    const double y_lo = x_hi * m_lo + (x_lo * m_hi + ((x_hi * m_hi - y_hi) + x_lo * m_lo));
    return DD{y_hi, y_lo};
}
__device__ __forceinline__ DD deg2rad_ext(double x) { return mulpi_ext(x / 180.0); }
__device__ __forceinline__ double sin_kernel(DD y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = y.hi * y.hi, w = z * z;
    const double r = (S2 + z * (S3 + z * S4)) + z * w * (S5 + z * S6);
    const double v = z * y.hi;
    return y.hi - ((z * (0.5 * y.lo - v * r) - y.lo) - v * S1);
}
__device__ __forceinline__ double cos_kernel(DD y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = y.hi * y.hi, w = z * z;
    const double r = z * (C1 + z * (C2 + z * C3)) + w * w * (C4 + z * (C5 + z * C6));
    const double hz = 0.5 * z;
    const double ww = 1.0 - hz;
    return ww + (((1.0 - ww) - hz) + (z * r - y.hi * y.lo));
}
__device__ inline double sind_dev(double x) {
    if (isnan(x) || isinf(x)) return __longlong_as_double(0x7ff8000000000000ll);
    const double rx = copysign(fmod(x, 360.0), x);
    const double arx = fabs(rx);
    if (rx == 0.0) return rx;
    if (arx < 45.0) return sin_kernel(deg2rad_ext(rx));
    if (arx <= 135.0) return copysign(cos_kernel(deg2rad_ext(90.0 - arx)), rx);
    if (arx == 180.0) return copysign(0.0, rx);
    if (arx < 225.0) return sin_kernel(deg2rad_ext((180.0 - arx) * (rx > 0 ? 1.0 : -1.0)));
    if (arx <= 315.0) return -copysign(cos_kernel(deg2rad_ext(270.0 - arx)), rx);
    return sin_kernel(deg2rad_ext(rx - copysign(360.0, rx)));
}
__device__ inline double cosd_dev(double x) {
    if (isnan(x) || isinf(x)) return __longlong_as_double(0x7ff8000000000000ll);
    const double rx = fabs(fmod(x, 360.0));
    if (rx <= 45.0) return cos_kernel(deg2rad_ext(rx));
    if (rx < 135.0) return sin_kernel(deg2rad_ext(90.0 - rx));
    if (rx <= 225.0) return -cos_kernel(deg2rad_ext(180.0 - rx));
    if (rx < 315.0) return sin_kernel(deg2rad_ext(rx - 270.0));
    return cos_kernel(deg2rad_ext(360.0 - rx));
}
__device__ inline double haversine_dev(double lon1, double lat1, double lon2, double lat2) {
    const double dl = lon2 - lon1, dp = lat2 - lat1;
    const double s1 = sind_dev(dp / 2), s2 = sind_dev(dl / 2);
    const double a = s1 * s1 + cosd_dev(lat1) * cosd_dev(lat2) * (s2 * s2);
    return 2 * (6371000.0 * asin(jl_min(sqrt(a), 1.0)));
}

