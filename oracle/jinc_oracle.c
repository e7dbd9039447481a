/*
 * jinc_oracle.c -- CPU restatement of the reference's EWA-Jinc hot path.  See jinc_oracle.h:
 * TEST INFRASTRUCTURE, parity pinned by execution of the unmodified reference (oracle/_ref).
 *
 * Build with -ffp-contract=off: the reference's table code is compiled without FMA
 * (CMakeLists.txt:57-61 gives -mfma only to the SIMD files), so every product and sum below is
 * rounded separately, in the precision the reference's expression has.
 */
#define _DEFAULT_SOURCE /* j1() */
#include "jinc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "jinc_constants.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ Jinc evaluation */

/* Horner sum of the first `terms` Taylor coefficients in x2 (src/JincResize.cpp:203-230). */
static double taylor(double x2, int terms)
{
    double acc = 0.0;
    for (int j = terms - 1; j >= 0; --j)
        acc = acc * x2 + JINC_TAYLOR[j];
    return acc;
}

/* Ratio of two degree-(n-1) polynomials, evaluated in z or in 1/z for |z|>1
 * (src/JincResize.cpp:110-140, after Boost.Math tools/rational.hpp). */
static double rational(const double* num, const double* den, double z, int n)
{
    double a, b;
    if (z <= 1.0) {
        a = num[n - 1];
        b = den[n - 1];
        for (int i = n - 2; i >= 0; --i) {
            a = a * z + num[i];
            b = b * z + den[i];
        }
    } else {
        z = 1.0 / z;
        a = num[0];
        b = den[0];
        for (int i = 1; i < n; ++i) {
            a = a * z + num[i];
            b = b * z + den[i];
        }
    }
    return a / b;
}

/* Large-argument J1 via Boost.Math's bessel_j1 asymptotic form (coefficients PC/QC/PS/QS of
 * boost/math/special_functions/detail/bessel_j1.hpp, Boost Software License 1.0), as used by
 * src/JincResize.cpp:148-198 for 52.57 <= x2 < 68.07. */
static double jinc_asymptotic(double x2)
{
    static const double PC[7] = {-4.4357578167941278571e+06, -9.9422465050776411957e+06, -6.6033732483649391093e+06,
                                 -1.5235293511811373833e+06, -1.0982405543459346727e+05, -1.6116166443246101165e+03,
                                 0.0};
    static const double QC[7] = {-4.4357578167941278568e+06, -9.9341243899345856590e+06, -6.5853394797230870728e+06,
                                 -1.5118095066341608816e+06, -1.0726385991103820119e+05, -1.4550094401904961825e+03,
                                 1.0};
    static const double PS[7] = {3.3220913409857223519e+04, 8.5145160675335701966e+04, 6.6178836581270835179e+04,
                                 1.8494262873223866797e+04, 1.7063754290207680021e+03, 3.5265133846636032186e+01,
                                 0.0};
    static const double QS[7] = {7.0871281941028743574e+05, 1.8194580422439972989e+06, 1.4194606696037208929e+06,
                                 4.0029443582266975117e+05, 3.7890229745772202641e+04, 8.6383677696049909675e+02,
                                 1.0};
    const double y2 = M_PI * M_PI * x2;
    const double y = sqrt(y2);
    const double w = 64.0 / y2;
    const double s = sin(y), c = cos(y);
    const double rc = rational(PC, QC, w, 7);
    const double rs = rational(PS, QS, w, 7);
    return (sqrt(y / M_PI) * 2.0 / y2) * (rc * (s - c) + (8.0 / y) * rs * (s + c));
}

double jo_jinc_sqr(double x2)
{
    if (x2 < 1.49)
        return taylor(x2, 16);
    if (x2 < 4.97)
        return taylor(x2, 21);
    if (x2 < 10.49)
        return taylor(x2, 26);
    if (x2 < 17.99)
        return taylor(x2, 31);
    if (x2 >= 52.57 && x2 < 68.07)
        return jinc_asymptotic(x2);
    /* src/JincResize.cpp:231-235,240-244 call std::cyl_bessel_j(1, x); POSIX j1() is the same function
     * from a different library -- the LUT is only ever consumed as float (Lut::GetFactor), where the
     * two agree (tests/test_oracle_vs_ref.py checks all 16 taps). */
    const double x = M_PI * sqrt(x2);
    return 2.0 * j1(x) / x;
}

double jo_radius_for_tap(int tap)
{
    return (tap >= 1 && tap <= JINC_MAX_TAP) ? JINC_ZEROS[tap - 1] : 0.0;
}

/* sample_sqr (src/JincResize.cpp:247-256) */
static double sample_sqr(double x2, double blur2, double radius2)
{
    if (blur2 > 0.0)
        x2 /= blur2;
    return x2 < radius2 ? jo_jinc_sqr(x2) : 0.0;
}

void jo_lut_init(double* lut, double radius, double blur)
{
    if (blur == 0.0) /* src/JincResize.cpp:772-774 */
        blur = 1.0;
    const double radius2 = radius * radius;
    const double blur2 = blur * blur;
    for (int i = 0; i < JO_LUT_SAMPLES; ++i) {
        const double t2 = i / (JO_LUT_SAMPLES - 1.0);
        lut[i] = sample_sqr(radius2 * t2, blur2, radius2) * sample_sqr(JINC_FIRST_ZERO_SQR * t2, 1.0, radius2);
    }
}

float jo_lut_factor(const double* lut, int index)
{
    return index >= JO_LUT_SAMPLES ? 0.f : (float)lut[index];
}

/* ------------------------------------------------------------------ coefficient table */

static float clampf(float v, float lo, float hi)
{
    v = v > hi ? hi : v; /* upper bound first, as avs/minmax.h does */
    return v < lo ? lo : v;
}

static double min_d(double a, double b) { return a < b ? a : b; }
static float max_f(float a, float b) { return a > b ? a : b; }
static int max_i(int a, int b) { return a > b ? a : b; }

void jo_table_free(jo_table* t)
{
    free(t->meta);
    free(t->factor);
    free(t->border);
    free(t->phase);
    memset(t, 0, sizeof(*t));
}

int jo_table_generate(const jo_table_params* p, const double* lut, jo_table* out)
{
    memset(out, 0, sizeof(*out));
    const int qx_n = p->quant_x, qy_n = p->quant_y;
    const int src_w = p->src_w, src_h = p->src_h, dst_w = p->dst_w, dst_h = p->dst_h;

    /* scalars: src/JincResize.cpp:349-364 */
    const double step_x = min_d((double)dst_w / p->crop_w, 1.0);
    const double step_y = min_d((double)dst_h / p->crop_h, 1.0);
    const float support_x = (float)(p->radius / step_x);
    const float support_y = (float)(p->radius / step_y);
    const float support = max_f(support_x, support_y);
    const int fs = max_i((int)ceil(support_x * 2.0), (int)ceil(support_y * 2.0));
    const float x0 = (float)(p->crop_left + (p->crop_w / dst_w - 1.0) / 2.0);
    const float dxpos = (float)(p->crop_w / dst_w);
    const float dypos = (float)(p->crop_h / dst_h);
    float xpos = x0;
    float ypos = (float)(p->crop_top + (p->crop_h - dst_h) / (double)(dst_h * (int64_t)2));
    const double radius2 = p->radius * p->radius;

    /* init_coeff_table: src/JincResize.cpp:286-306 */
    const int stride = (fs + 15) & ~15;
    const size_t block = (size_t)stride * fs;
    const size_t npix = (size_t)dst_w * dst_h;
    out->filter_size = fs;
    out->coeff_stride = stride;
    out->dst_w = dst_w;
    out->dst_h = dst_h;
    out->meta = (int32_t*)calloc(npix * 3, sizeof(int32_t));
    out->border = (uint8_t*)calloc(npix, 1);
    out->phase = (int32_t*)calloc(npix * 2, sizeof(int32_t));
    int32_t* seen = (int32_t*)calloc((size_t)qx_n * qy_n, sizeof(int32_t)); /* factor_map: offset+1 */
    size_t cap = block * 64, top = 0;
    float* arena = (float*)malloc(cap * sizeof(float));
    if (!out->meta || !out->border || !out->phase || !seen || !arena)
        goto oom;

    for (int y = 0; y < dst_h; ++y) {
        for (int x = 0; x < dst_w; ++x) {
            const size_t pix = (size_t)y * dst_w + x;
            int border = 0;

            /* window placement from the UNquantised position: src/JincResize.cpp:392-421 */
            int end_x = (int)(xpos + support);
            int end_y = (int)(ypos + support);
            if (end_x >= src_w) { end_x = src_w - 1; border = 1; }
            if (end_y >= src_h) { end_y = src_h - 1; border = 1; }
            int begin_x = end_x - fs + 1;
            int begin_y = end_y - fs + 1;
            if (begin_x < 0) { begin_x = 0; border = 1; }
            if (begin_y < 0) { begin_y = 0; border = 1; }
            out->meta[pix * 3 + 0] = begin_x;
            out->meta[pix * 3 + 1] = begin_y;
            out->border[pix] = (uint8_t)border;

            /* sub-pixel phase: src/JincResize.cpp:424-429 */
            const int qxi = (int)(xpos * qx_n);
            const int qyi = (int)(ypos * qy_n);
            const int qxv = qxi % qx_n;
            const int qyv = qyi % qy_n;
            const float qxpos = (float)qxi / qx_n;
            const float qypos = (float)qyi / qy_n;
            out->phase[pix * 2 + 0] = qxv;
            out->phase[pix * 2 + 1] = qyv;

            if (!border && seen[qyv * qx_n + qxv] != 0) { /* :431-435 */
                out->meta[pix * 3 + 2] = seen[qyv * qx_n + qxv] - 1;
            } else {
                if (!border) { /* :446-451: the weights' own window comes from the quantised position */
                    begin_x = (int)(qxpos + support) - fs + 1;
                    begin_y = (int)(qypos + support) - fs + 1;
                }
                if (top + block > cap) {
                    cap = cap * 2 > top + block ? cap * 2 : top + block;
                    float* grown = (float*)realloc(arena, cap * sizeof(float));
                    if (!grown)
                        goto oom;
                    arena = grown;
                }
                float* w = arena + top;
                memset(w, 0, block * sizeof(float));

                const float cx = clampf(border ? xpos : qxpos, 0.f, (float)(src_w - 1));
                const float cy = clampf(border ? ypos : qypos, 0.f, (float)(src_h - 1));
                float sum = 0.f;
                for (int ly = 0; ly < fs; ++ly) {
                    for (int lx = 0; lx < fs; ++lx) { /* :485-493 */
                        const double dx = (cx - (begin_x + lx)) * step_x; /* float minus int, then double */
                        const double dy = (cy - (begin_y + ly)) * step_y;
                        /* nearbyint under the default rounding mode == the reference's 1.5*2^52 trick */
                        const int idx = (int)llround((JO_LUT_SAMPLES - 1) * (dx * dx + dy * dy) / radius2 + 6755399441055744.0);
                        const float f = jo_lut_factor(lut, idx);
                        w[(size_t)ly * stride + lx] = f;
                        sum += f;
                    }
                }
                for (int ly = 0; ly < fs; ++ly) /* :505-514 */
                    for (int lx = 0; lx < fs; ++lx)
                        w[(size_t)ly * stride + lx] /= sum;

                if (!border) /* :517-518 */
                    seen[qyv * qx_n + qxv] = (int32_t)top + 1;
                out->meta[pix * 3 + 2] = (int32_t)top;
                top += block;
                out->n_blocks++;
            }
            xpos += dxpos; /* :524 */
        }
        ypos += dypos; /* :527-528 */
        xpos = x0;
    }
    free(seen);
    out->factor = arena;
    out->factor_len = top;
    return 0;

oom:
    free(seen);
    free(arena);
    jo_table_free(out);
    return -1;
}

/* ------------------------------------------------------------------ per-plane geometry */

int jo_plane_params(int src_w, int src_h, int target_w, int target_h, double src_left, double src_top,
                    double src_width_arg, double src_height_arg, int quant_x, int quant_y, int tap, int sub_w,
                    int sub_h, int cplace, jo_table_params out[2])
{
    /* src/JincResize.cpp:762-770 */
    double crop_w = src_width_arg, crop_h = src_height_arg;
    if (crop_w <= 0.0)
        crop_w = src_w - src_left + crop_w;
    if (crop_h <= 0.0)
        crop_h = src_h - src_top + crop_h;

    const double radius = jo_radius_for_tap(tap);
    jo_table_params luma = {quant_x, quant_y, src_w, src_h, target_w, target_h, radius, src_left, src_top, crop_w, crop_h};
    out[0] = luma;
    if (sub_w == 0 && sub_h == 0)
        return 1;

    /* src/JincResize.cpp:833-862: chroma shift uses the FULL source width/height, not the crop */
    const double div_w = (double)(1 << sub_w), div_h = (double)(1 << sub_h);
    const double left_uv = (cplace == 0 || cplace == 2)
                               ? (0.5 * (1.0 - (double)src_w / target_w) + src_left) / div_w
                               : src_left / div_w;
    const double top_uv = (cplace == 2) ? (0.5 * (1.0 - (double)src_h / target_h) + src_top) / div_h : src_top / div_h;
    jo_table_params chroma = {quant_x, quant_y, src_w >> sub_w, src_h >> sub_h, target_w >> sub_w, target_h >> sub_h,
                              radius, left_uv, top_uv, crop_w / div_w, crop_h / div_h};
    out[1] = chroma;
    return 2;
}

/* ------------------------------------------------------------------ resampling */

#define JO_RESIZE_BODY(T, STORE)                                                                        \
    const int fs = t->filter_size, cs = t->coeff_stride, w = t->dst_w;                                  \
    for (int y = y0; y < y1; ++y) {                                                                     \
        T* drow = dst + (ptrdiff_t)y * dst_stride;                                                      \
        for (int x = 0; x < w; ++x) {                                                                   \
            const int32_t* m = t->meta + ((size_t)y * w + x) * 3;                                       \
            const T* s = src + (ptrdiff_t)m[1] * src_stride + m[0];                                     \
            const float* c = t->factor + m[2];                                                          \
            float acc = 0.f;                                                                            \
            for (int ly = 0; ly < fs; ++ly) {                                                           \
                for (int lx = 0; lx < fs; ++lx)                                                         \
                    acc += s[lx] * c[lx]; /* separate multiply and add, row-major (:572-579) */         \
                c += cs;                                                                                \
                s += src_stride;                                                                        \
            }                                                                                           \
            STORE;                                                                                      \
        }                                                                                               \
    }

static void rows_u8(const jo_table* t, const uint8_t* src, ptrdiff_t src_stride, uint8_t* dst, ptrdiff_t dst_stride,
                    float peak, int y0, int y1)
{
    JO_RESIZE_BODY(uint8_t, drow[x] = (uint8_t)lrintf(clampf(acc, 0.f, peak)))
}

static void rows_u16(const jo_table* t, const uint16_t* src, ptrdiff_t src_stride, uint16_t* dst, ptrdiff_t dst_stride,
                     float peak, int y0, int y1)
{
    JO_RESIZE_BODY(uint16_t, drow[x] = (uint16_t)lrintf(clampf(acc, 0.f, peak)))
}

static void rows_f32(const jo_table* t, const float* src, ptrdiff_t src_stride, float* dst, ptrdiff_t dst_stride,
                     int y0, int y1)
{
    JO_RESIZE_BODY(float, drow[x] = acc)
}

void jo_resize_plane_u8(const jo_table* t, const uint8_t* src, ptrdiff_t src_stride, uint8_t* dst, ptrdiff_t dst_stride,
                        float peak)
{
    rows_u8(t, src, src_stride, dst, dst_stride, peak, 0, t->dst_h);
}

void jo_resize_plane_u16(const jo_table* t, const uint16_t* src, ptrdiff_t src_stride, uint16_t* dst,
                         ptrdiff_t dst_stride, float peak)
{
    rows_u16(t, src, src_stride, dst, dst_stride, peak, 0, t->dst_h);
}

void jo_resize_plane_f32(const jo_table* t, const float* src, ptrdiff_t src_stride, float* dst, ptrdiff_t dst_stride)
{
    rows_f32(t, src, src_stride, dst, dst_stride, 0, t->dst_h);
}

void jo_resize_rows(const jo_table* t, int sample_bytes, const void* src, ptrdiff_t src_stride, void* dst,
                    ptrdiff_t dst_stride, float peak, int y0, int y1)
{
    if (sample_bytes == 1)
        rows_u8(t, (const uint8_t*)src, src_stride, (uint8_t*)dst, dst_stride, peak, y0, y1);
    else if (sample_bytes == 2)
        rows_u16(t, (const uint16_t*)src, src_stride, (uint16_t*)dst, dst_stride, peak, y0, y1);
    else
        rows_f32(t, (const float*)src, src_stride, (float*)dst, dst_stride, y0, y1);
}
