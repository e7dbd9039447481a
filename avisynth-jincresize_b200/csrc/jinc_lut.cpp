// jinc_lut.cpp -- Jinc kernel evaluation and the radial weight LUT (host, FP64).
//
// Product code of libjinc_b200.so.  Replaces the reference's jinc_sqr / sample_sqr / Lut::InitLut
// (src/JincResize.cpp:200-275).  The LUT has 1024 entries and is built once per filter, so it stays on
// the host, in the same libm/libstdc++ arithmetic the reference uses: entry i samples the normalised squared
// radius t2 = i/1023 and holds  jinc(radius*t/blur) * jinc(z1*t)  -- the EWA kernel times a Jinc window
// stretched so that its first zero z1 lands on the support radius.  The device kernels only ever consume
// the table as float (Lut::GetFactor, :277-282).
#include <cmath>

#include "jinc_constants.h"
#include "jinc_internal.h"

namespace {

constexpr double kPi = 3.14159265358979323846;

// Truncated power series in x^2; more terms as x^2 approaches the next zero (:203-230).
inline double series(double x2, int terms)
{
    double s = 0.0;
    for (int k = terms; k-- > 0;)
        s = s * x2 + JINC_TAYLOR[k];
    return s;
}

struct Rational7 {
    double p[7], q[7];
    // P(z)/Q(z); for |z| > 1 the polynomials are evaluated in 1/z with reversed coefficients (:110-140).
    double operator()(double z) const
    {
        double a, b;
        if (z <= 1.0) {
            a = p[6];
            b = q[6];
            for (int i = 5; i >= 0; --i) {
                a = a * z + p[i];
                b = b * z + q[i];
            }
        } else {
            z = 1.0 / z;
            a = p[0];
            b = q[0];
            for (int i = 1; i < 7; ++i) {
                a = a * z + p[i];
                b = b * z + q[i];
            }
        }
        return a / b;
    }
};

// Hankel-type asymptotic form of J1 with Boost.Math's rational corrections (bessel_j1.hpp PC/QC, PS/QS;
// Boost Software License 1.0), which the reference uses for 52.57 <= x^2 < 68.07 (:148-198).
const Rational7 kCos = {{-4.4357578167941278571e+06, -9.9422465050776411957e+06, -6.6033732483649391093e+06,
                         -1.5235293511811373833e+06, -1.0982405543459346727e+05, -1.6116166443246101165e+03, 0.0},
                        {-4.4357578167941278568e+06, -9.9341243899345856590e+06, -6.5853394797230870728e+06,
                         -1.5118095066341608816e+06, -1.0726385991103820119e+05, -1.4550094401904961825e+03, 1.0}};
const Rational7 kSin = {{3.3220913409857223519e+04, 8.5145160675335701966e+04, 6.6178836581270835179e+04,
                         1.8494262873223866797e+04, 1.7063754290207680021e+03, 3.5265133846636032186e+01, 0.0},
                        {7.0871281941028743574e+05, 1.8194580422439972989e+06, 1.4194606696037208929e+06,
                         4.0029443582266975117e+05, 3.7890229745772202641e+04, 8.6383677696049909675e+02, 1.0}};

double jinc_large(double x2)
{
    const double y2 = kPi * kPi * x2;
    const double y = std::sqrt(y2);
    const double z = 64.0 / y2;
    const double sn = std::sin(y), cs = std::cos(y);
    return (std::sqrt(y / kPi) * 2.0 / y2) * (kCos(z) * (sn - cs) + (8.0 / y) * kSin(z) * (sn + cs));
}

double jinc_bessel(double x2)
{
    const double x = kPi * std::sqrt(x2);
    return 2.0 * std::cyl_bessel_j(1, x) / x; // :231-235,240-244
}

} // namespace

extern "C" double jinc_eval_sqr(double x2)
{
    if (x2 < 1.49)
        return series(x2, 16);
    if (x2 < 4.97)
        return series(x2, 21);
    if (x2 < 10.49)
        return series(x2, 26);
    if (x2 < 17.99)
        return series(x2, 31);
    if (x2 < 52.57)
        return jinc_bessel(x2);
    if (x2 < 68.07)
        return jinc_large(x2);
    return jinc_bessel(x2);
}

extern "C" double jinc_radius_for_tap(int tap)
{
    return (tap < 1 || tap > JINC_MAX_TAP) ? 0.0 : JINC_ZEROS[tap - 1];
}

void jinc_lut_build_host(double radius, double blur, double* lut)
{
    if (blur == 0.0)
        blur = 1.0;
    const double r2 = radius * radius, b2 = blur * blur;
    // kernel sample: argument scaled by 1/blur^2, cut off at the support radius (:247-256)
    auto cut = [r2](double x2, double scale2) {
        if (scale2 > 0.0)
            x2 /= scale2;
        return x2 < r2 ? jinc_eval_sqr(x2) : 0.0;
    };
    for (int i = 0; i < JINC_LUT_SAMPLES; ++i) {
        const double t2 = i / (JINC_LUT_SAMPLES - 1.0);
        lut[i] = cut(r2 * t2, b2) * cut(JINC_FIRST_ZERO_SQR * t2, 1.0);
    }
}

extern "C" int jinc_lut_build(double radius, double blur, double* lut)
{
    if (!lut || !(radius > 0.0))
        return jinc_fail(JINC_E_INVALID, "jinc_lut_build: radius must be positive and lut non-null");
    jinc_lut_build_host(radius, blur, lut);
    return JINC_OK;
}
