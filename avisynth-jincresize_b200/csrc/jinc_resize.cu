// jinc_resize.cu -- EWA resampling kernels for sm_100a and their launcher.
//
// Replaces JincResize::resize_plane_c<T,thr,subsampled> (src/JincResize.cpp:536-601) and the three SIMD copies of
// its inner loop.  Every output sample is  sum_{ly,lx} src[start_y+ly][start_x+lx] * w[ly][lx]  over an fs x fs
// window, followed for integer formats by clamp to [0,peak] and round-half-even (:581-582); float is raw (:583-584).
//
// One launch covers ALL planes that share a coefficient table, for a whole BATCH of frames, interior and border
// together; blocks take one of two roles:
//
//   interior tile (exact 2x upscale, every "JincNNResize(2w,2h)" use)
//       The table has 2x2 phase classes.  A thread owns TX=4 source-aligned cells x 2 cell rows = 8x4 output samples in
//       16 float2 accumulators.  The source tile lives in shared memory as VERTICAL PAIRS {S[r][c], S[r+1][c]} so one
//       packed FFMA2 (fma.rn.f32x2, new on sm_100) updates the same phase of two cell rows with a single scalar
//       weight.  Weights arrive as kernel parameters (constant bank) and reach the FMA pipe through uniform registers
//       (LDCU -> FFMA2 R, R, UR, R): weights cost no shared-memory or register-file bandwidth.  Pair columns are
//       de-interleaved by (c & 3) so a warp's LDS.64 is bank-conflict free.
//   strip chunk (256 output samples of the border strips, or of the whole plane when the table has no fast path)
//       One thread per output sample, all planes of the table in one pass.  Samples whose window was clamped get the
//       reference's per-pixel weights on the fly: exact LUT index per tap, divided by the per-pixel normaliser that
//       the table build stored (:443-514); other samples gather their shared phase block from the L2-resident table.
#include <algorithm>
#include <cstring>
#include <type_traits>

#include "jinc_internal.h"
#include "jinc_weights.cuh"

namespace {

// ------------------------------------------------------------------------------------------ sample conversion

// clamp to [0, peak] and round half to even (lrintf) in one saturating convert; NaN -> 0
__device__ __forceinline__ uint32_t finish_u8(float v, float peak)
{
    uint32_t r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(fminf(v, peak)));
    return r;
}
__device__ __forceinline__ uint32_t finish_u16(float v, float peak)
{
    uint32_t r;
    asm("cvt.rni.sat.u16.f32 %0, %1;" : "=r"(r) : "f"(fminf(v, peak)));
    return r;
}

template <typename T>
__device__ __forceinline__ T finish(float v, float peak);
template <>
__device__ __forceinline__ float finish<float>(float v, float)
{
    return v;
}
template <>
__device__ __forceinline__ uint8_t finish<uint8_t>(float v, float peak)
{
    return (uint8_t)finish_u8(v, peak);
}
template <>
__device__ __forceinline__ uint16_t finish<uint16_t>(float v, float peak)
{
    return (uint16_t)finish_u16(v, peak);
}

// Integer samples become floats without the (quarter-rate) I2F: drop the bits into the mantissa of 2^23 and subtract
// 2^23 again -- exact for any value below 2^23.
template <typename T>
__device__ __forceinline__ float load_sample(const T* p)
{
    return __uint_as_float(0x4B000000u | (uint32_t)__ldg(p)) - 8388608.f;
}
template <>
__device__ __forceinline__ float load_sample<float>(const float* p)
{
    return __ldg(p);
}
template <typename T>
__device__ __forceinline__ float sample_to_float(T v)
{
    return __uint_as_float(0x4B000000u | (uint32_t)v) - 8388608.f;
}
template <>
__device__ __forceinline__ float sample_to_float<float>(float v)
{
    return v;
}

template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8], float peak);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8], float)
{
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<uint16_t>(uint16_t* p, const float (&v)[8], float peak)
{
    uint32_t q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        q[k] = finish_u16(v[2 * k], peak) | (finish_u16(v[2 * k + 1], peak) << 16);
    *reinterpret_cast<uint4*>(p) = make_uint4(q[0], q[1], q[2], q[3]);
}
template <>
__device__ __forceinline__ void store8<uint8_t>(uint8_t* p, const float (&v)[8], float peak)
{
    uint32_t q[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
        q[k] = finish_u8(v[4 * k], peak) | (finish_u8(v[4 * k + 1], peak) << 8) | (finish_u8(v[4 * k + 2], peak) << 16) |
               (finish_u8(v[4 * k + 3], peak) << 24);
    *reinterpret_cast<uint2*>(p) = make_uint2(q[0], q[1]);
}

// ------------------------------------------------------------------------------------------ shared argument blocks

struct Rect {
    int x0, y0, x1, y1;
};

// bits per component back from peak = (1 << bits) - 1
inline int t_bits_from_peak(float peak)
{
    int bits = 0;
    while (bits < 32 && (float)((1ll << bits) - 1) < peak)
        ++bits;
    return bits;
}

// planes of ONE frame that share the table being run (device pointers, pitches in elements)
struct PlanePtrs {
    const void* src[JINC_MAX_PLANES];
    void* dst[JINC_MAX_PLANES];
    long long src_pitch[JINC_MAX_PLANES];
    long long dst_pitch[JINC_MAX_PLANES];
};

// strips: up to four rectangles of output samples, each cut into patches of PW x PH outputs; one block per patch
struct StripArgs {
    const int32_t* start_x;
    const int32_t* start_y;
    const int32_t* rank_x;
    const int32_t* rank_y;
    const float* pos_x;
    const float* pos_y;
    const float* weights;
    const float* lut;
    const float* border_sum;
    const float* border_w; // resident per-pixel border weights [slot/32][tap][slot%32], or null
    const int32_t* border_block; // slot -> class block, or null
    const float* border_wb;      // [block][fs][fsp] class blocks, fsp = fs rounded up to 4
    BorderGeom bg;
    int fs, n_rank_x, src_w, src_h;
    double step_x, step_y, radius2, idx_scale;
    Rect rect[4];
    unsigned patch_begin[5];   // prefix sums of patches per rect
    unsigned patches_x[4];     // patches per row of patches
    int pw_log2[4];            // log2 of the patch width (patch height = outputs per block / width)
    unsigned blocks_per_plane; // = patch_begin[4]; the grid holds this many strip blocks per plane
    unsigned smem_floats;      // shared memory a strip block may use to stage its source footprint (0: none)
};

struct FrameSet {
    PlanePtrs one;           // used when frames == nullptr
    const PlanePtrs* frames; // device array [grid.y] for batched launches
    int n_planes;
    float peak;
};

__device__ __forceinline__ const PlanePtrs& frame_ptrs(const FrameSet& fs)
{
    return fs.frames ? fs.frames[blockIdx.y] : fs.one;
}

// Window rows as aligned 32-bit words (the plane base and pitch are 4-byte aligned), converted later: the words are
// funnel-shifted into place and every sample is dropped into the mantissa of 2^23 by one byte permute.
template <typename T, int FS>
struct RowWords {
    static constexpr int SB = (int)sizeof(T);
    static constexpr int NA = (FS * SB + 3) / 4;      // aligned words that hold the row
    static constexpr int NW = SB == 4 ? FS : NA + 1;  // words loaded (one more when the row starts inside a word)
    // rows whose loads are issued together (the strips are latency-bound: most sample rows miss L2), about 24 registers
    static constexpr int GROUP = (24 / NW) < 1 ? 1 : ((24 / NW) > FS ? FS : (24 / NW));
};

template <typename T, int FS>
__device__ __forceinline__ void load_row_words(const T* __restrict__ s, uint32_t (&w)[RowWords<T, FS>::NW])
{
    using R = RowWords<T, FS>;
    if constexpr (R::SB == 4) {
#pragma unroll
        for (int i = 0; i < FS; ++i)
            w[i] = __ldg(reinterpret_cast<const uint32_t*>(s) + i);
    } else {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(s);
        const uint32_t* __restrict__ p4 = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
#pragma unroll
        for (int j = 0; j < R::NA; ++j)
            w[j] = __ldg(p4 + j);
        w[R::NA] = ((unsigned)(addr & 3) + FS * R::SB > 4 * R::NA) ? __ldg(p4 + R::NA) : 0u; // never touch a word the row does not reach
    }
}

template <typename T, int FS>
__device__ __forceinline__ void row_words_to_float(const uint32_t (&w)[RowWords<T, FS>::NW], unsigned off, float (&v)[FS])
{
    using R = RowWords<T, FS>;
    if constexpr (R::SB == 4) {
#pragma unroll
        for (int i = 0; i < FS; ++i)
            v[i] = __uint_as_float(w[i]);
    } else {
        uint32_t al[R::NA];
#pragma unroll
        for (int j = 0; j < R::NA; ++j)
            al[j] = __funnelshift_r(w[j], w[j + 1], off * 8);
#pragma unroll
        for (int i = 0; i < FS; ++i) {
            uint32_t bits;
            if (R::SB == 1)
                bits = __byte_perm(al[i >> 2], 0x4B000000u, 0x7440 | (i & 3));
            else
                bits = __byte_perm(al[i >> 1], 0x4B000000u, (i & 1) ? 0x7432 : 0x7410);
            v[i] = __uint_as_float(bits) - 8388608.f;
        }
    }
}

// sum over an FS x FS window with a weight block whose rows are padded to 16 bytes
template <typename T, int FS>
__device__ __forceinline__ float dot_rows_vec(const T* __restrict__ s, int pitch, const float* __restrict__ w)
{
    using R = RowWords<T, FS>;
    constexpr int FSP = (FS + 3) & ~3;
    const unsigned off = (unsigned)(reinterpret_cast<uintptr_t>(s) & 3); // the pitch keeps it the same on every row
    float acc = 0.f;
#pragma unroll 1
    for (int ly0 = 0; ly0 < FS; ly0 += R::GROUP) {
        uint32_t words[R::GROUP][R::NW];
#pragma unroll
        for (int g = 0; g < R::GROUP; ++g) {
            const int ly = min(ly0 + g, FS - 1); // the last group may be short: re-read the last row, skipped below
            load_row_words<T, FS>(s + (long long)ly * pitch, words[g]);
        }
#pragma unroll
        for (int g = 0; g < R::GROUP; ++g) {
            if (FS % R::GROUP != 0 && ly0 + g >= FS)
                break;
            float wr[FSP], v[FS];
            const float4* __restrict__ w4 = reinterpret_cast<const float4*>(w + (ly0 + g) * FSP);
#pragma unroll
            for (int q = 0; q < FSP / 4; ++q) {
                const float4 t = __ldg(w4 + q);
                wr[4 * q] = t.x;
                wr[4 * q + 1] = t.y;
                wr[4 * q + 2] = t.z;
                wr[4 * q + 3] = t.w;
            }
            row_words_to_float<T, FS>(words[g], off, v);
#pragma unroll
            for (int lx = 0; lx < FS; ++lx)
                acc = fmaf(v[lx], wr[lx], acc);
        }
    }
    return acc;
}

constexpr int STRIP_THREADS = 256;

// One output sample of a strip for ONE plane, read straight from global memory.  Lean on purpose: 32-bit indexing,
// constant weight strides.  FSC > 0 fixes the window size at compile time (inner loops unroll).
template <typename T, int FSC>
__device__ __forceinline__ void strip_sample(const StripArgs& a, const FrameSet& fsx, int x, int y, int plane)
{
    const PlanePtrs& pp = frame_ptrs(fsx);
    const int fs = FSC > 0 ? FSC : a.fs;
    const int sx = a.start_x[x], sy = a.start_y[y];
    const int rx = a.rank_x[x], ry = a.rank_y[y];
    const int pitch = (int)pp.src_pitch[plane];
    const T* __restrict__ s = static_cast<const T*>(pp.src[plane]) + (long long)sy * pitch + sx;
    float acc = 0.f;

    const bool shared_block = rx >= 0 && ry >= 0;
    if (!shared_block && a.border_block) {
        // the block of this border pixel's class, rows padded to 16 bytes
        const int fsp = (fs + 3) & ~3;
        const float* __restrict__ w = a.border_wb + (size_t)a.border_block[jinc_border_slot(a.bg, x, y)] * (unsigned)(fs * fsp);
        if (FSC > 0 && ((reinterpret_cast<uintptr_t>(pp.src[plane]) | (uintptr_t)(pitch * (int)sizeof(T))) & 3) == 0) {
            acc = dot_rows_vec<T, (FSC > 0 ? FSC : 4)>(s, pitch, w);
        } else {
            for (int ly = 0; ly < fs; ++ly) {
#pragma unroll
                for (int lx = 0; lx < (FSC > 0 ? FSC : fs); ++lx)
                    acc = fmaf(load_sample(s + lx), __ldg(w + lx), acc);
                w += fsp;
                s += pitch;
            }
        }
    } else if (shared_block) {
        // shared phase block (:431-435), row-major fs x fs
        const float* __restrict__ w = a.weights + (unsigned)(ry * a.n_rank_x + rx) * (unsigned)(fs * fs);
        for (int ly = 0; ly < fs; ++ly) {
#pragma unroll
            for (int lx = 0; lx < (FSC > 0 ? FSC : fs); ++lx)
                acc = fmaf(load_sample(s + lx), __ldg(w + lx), acc);
            w += fs;
            s += pitch;
        }
    } else if (a.border_w) {
        // this border pixel's own resident block (:443-514), stored [slot / 32][tap][slot % 32]: neighbouring pixels
        // coalesce and the tap stride is the constant 32
        const long long slot = jinc_border_slot(a.bg, x, y);
        const float* __restrict__ w = a.border_w + (size_t)(slot >> 5) * (size_t)(fs * fs * 32) + (unsigned)(slot & 31);
        for (int ly = 0; ly < fs; ++ly) {
#pragma unroll
            for (int lx = 0; lx < (FSC > 0 ? FSC : fs); ++lx)
                acc = fmaf(load_sample(s + lx), __ldg(w + lx * 32), acc);
            w += fs * 32;
            s += pitch;
        }
    } else {
        // border weights did not fit the residency budget: rebuild them per sample from the UNquantised position
        // (:443-514), exact LUT index per tap, factor / divider
        const float px = a.pos_x[x], py = a.pos_y[y];
        const float sum = a.border_sum[jinc_border_slot(a.bg, x, y)];
        for (int ly = 0; ly < fs; ++ly) {
            const double dy2 = jinc_tap_dist2(py, a.src_h, sy + ly, a.step_y);
            for (int lx = 0; lx < fs; ++lx) {
                const double dx2 = jinc_tap_dist2(px, a.src_w, sx + lx, a.step_x);
                const float f = jinc_lut_weight(a.lut, __dadd_rn(dx2, dy2), a.radius2, a.idx_scale);
                acc = fmaf(load_sample(s + lx), __fdiv_rn(f, sum), acc);
            }
            s += pitch;
        }
    }
    static_cast<T*>(pp.dst[plane])[(long long)y * pp.dst_pitch[plane] + x] = finish<T>(acc, fsx.peak);
}

// What a strip sample needs besides its source window: gathered for all of a thread's samples before any is used, so
// the table loads of the SPT samples are in flight together.
struct StripMeta {
    int x, y;       // output sample (x < 0: none)
    int sx, sy;     // window origin
    const float* w; // weight block: [fs][wstride]
    int wstride;    // fs for a shared phase block, fs rounded up to 4 for a border class block; 0 = neither (slow kinds)
};

template <int FSC>
__device__ __forceinline__ StripMeta strip_meta(const StripArgs& a, int x, int y)
{
    const int fs = FSC > 0 ? FSC : a.fs;
    StripMeta m;
    m.x = x;
    m.y = y;
    m.sx = a.start_x[x];
    m.sy = a.start_y[y];
    const bool border = x < a.bg.bx0 || x >= a.bg.bx1 || y < a.bg.by0 || y >= a.bg.by1; // no table load needed to know
    if (!border) {
        m.w = a.weights + (unsigned)(a.rank_y[y] * a.n_rank_x + a.rank_x[x]) * (unsigned)(fs * fs);
        m.wstride = fs;
    } else if (a.border_block) {
        const int fsp = (fs + 3) & ~3;
        m.w = a.border_wb + (size_t)a.border_block[jinc_border_slot(a.bg, x, y)] * (unsigned)(fs * fsp);
        m.wstride = fsp;
    } else {
        m.w = nullptr;
        m.wstride = 0;
    }
    return m;
}

// One sample from a staged footprint: `tile` holds the source rectangle [sy_lo, ..) x [sx_lo, sx_lo + fw) as floats.
template <typename T, int FSC>
__device__ __forceinline__ void strip_sample_staged(const StripArgs& a, const FrameSet& fsx, const StripMeta& m, int plane,
                                                    const float* __restrict__ tile, int fw, int sx_lo, int sy_lo)
{
    const int fs = FSC > 0 ? FSC : a.fs;
    const float* __restrict__ s = tile + (m.sy - sy_lo) * fw + (m.sx - sx_lo);
    float acc = 0.f;
    if (m.wstride == fs) {
        const float* __restrict__ w = m.w;
        for (int ly = 0; ly < fs; ++ly) {
#pragma unroll
            for (int lx = 0; lx < (FSC > 0 ? FSC : fs); ++lx)
                acc = fmaf(s[lx], __ldg(w + lx), acc);
            w += fs;
            s += fw;
        }
    } else {
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(m.w); // rows padded to 16 bytes
        for (int ly = 0; ly < fs; ++ly) {
#pragma unroll
            for (int q = 0; q < (FSC > 0 ? (FSC + 3) / 4 : m.wstride / 4); ++q) {
                const float4 t = __ldg(w4 + q);
                const int lx = 4 * q;
                acc = fmaf(s[lx], t.x, acc);
                if (lx + 1 < fs)
                    acc = fmaf(s[lx + 1], t.y, acc);
                if (lx + 2 < fs)
                    acc = fmaf(s[lx + 2], t.z, acc);
                if (lx + 3 < fs)
                    acc = fmaf(s[lx + 3], t.w, acc);
            }
            w4 += m.wstride / 4;
            s += fw;
        }
    }
    const PlanePtrs& pp = frame_ptrs(fsx);
    static_cast<T*>(pp.dst[plane])[(long long)m.y * pp.dst_pitch[plane] + m.x] = finish<T>(acc, fsx.peak);
}

// SPT samples of one thread that share ONE class block (the usual case in the strips of the periodic geometries: a
// thread's samples lie in the same border row or column, a multiple of the phase period apart): the weights are loaded
// once and feed SPT independent accumulators.
template <typename T, int FSC, int SPT>
__device__ __forceinline__ void strip_samples_fused(const StripArgs& a, const FrameSet& fsx, const StripMeta (&m)[SPT], int plane,
                                                    const float* __restrict__ tile, int fw, int sx_lo, int sy_lo)
{
    const int fs = FSC > 0 ? FSC : a.fs;
    const float* __restrict__ s[SPT];
    float acc[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        s[k] = tile + (m[k].sy - sy_lo) * fw + (m[k].sx - sx_lo);
        acc[k] = 0.f;
    }
    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(m[0].w); // rows padded to 16 bytes
    const int wq = m[0].wstride / 4;
    for (int ly = 0; ly < fs; ++ly) {
#pragma unroll
        for (int q = 0; q < (FSC > 0 ? (FSC + 3) / 4 : wq); ++q) {
            const float4 t = __ldg(w4 + q);
            const int lx = 4 * q;
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
                acc[k] = fmaf(s[k][lx], t.x, acc[k]);
                if (lx + 1 < fs)
                    acc[k] = fmaf(s[k][lx + 1], t.y, acc[k]);
                if (lx + 2 < fs)
                    acc[k] = fmaf(s[k][lx + 2], t.z, acc[k]);
                if (lx + 3 < fs)
                    acc[k] = fmaf(s[k][lx + 3], t.w, acc[k]);
            }
        }
        w4 += wq;
#pragma unroll
        for (int k = 0; k < SPT; ++k)
            s[k] += fw;
    }
    const PlanePtrs& pp = frame_ptrs(fsx);
    T* __restrict__ dst = static_cast<T*>(pp.dst[plane]);
    const long long dp = pp.dst_pitch[plane];
#pragma unroll
    for (int k = 0; k < SPT; ++k)
        dst[(long long)m[k].y * dp + m[k].x] = finish<T>(acc[k], fsx.peak);
}

// Strip block `sb` of the grid (planes are the slow dimension): one patch of PW x PH outputs, SPT per thread.  The
// source rectangle the patch reads (window origins are monotonic along both axes) is staged into shared memory as
// floats when it fits, so the global loads are coalesced, converted once and every window row is read from shared
// memory; otherwise every sample reads global memory directly.
template <typename T, int FSC, int THREADS = STRIP_THREADS, int SPT = 1>
__device__ __forceinline__ void strip_block(const StripArgs& a, const FrameSet& fsx, unsigned sb, float* __restrict__ tile)
{
    const int fs = FSC > 0 ? FSC : a.fs;
    const unsigned plane = sb / a.blocks_per_plane;
    const unsigned pid = sb - plane * a.blocks_per_plane;
    const int r = (int)(pid >= a.patch_begin[1]) + (int)(pid >= a.patch_begin[2]) + (int)(pid >= a.patch_begin[3]);
    const unsigned lp = pid - a.patch_begin[r];
    const unsigned pyi = lp / a.patches_x[r], pxi = lp - pyi * a.patches_x[r];
    const int pwl = a.pw_log2[r];
    const int ox0 = a.rect[r].x0 + (int)(pxi << pwl), oy0 = a.rect[r].y0 + (int)pyi * ((THREADS * SPT) >> pwl);
    const int nx = min(1 << pwl, a.rect[r].x1 - ox0), ny = min((THREADS * SPT) >> pwl, a.rect[r].y1 - oy0);

    // footprint corners and the tables of this thread's samples: one round of loads
    const int sx_lo = a.start_x[ox0], sy_lo = a.start_y[oy0];
    const int fw = a.start_x[ox0 + nx - 1] + fs - sx_lo, fh = a.start_y[oy0 + ny - 1] + fs - sy_lo;
    StripMeta meta[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        const int o = (int)threadIdx.x + k * THREADS;
        const int lx = o & ((1 << pwl) - 1), ly = o >> pwl;
        if (lx < nx && ly < ny) {
            meta[k] = strip_meta<FSC>(a, ox0 + lx, oy0 + ly);
        } else {
            meta[k].x = -1;
            meta[k].wstride = 0;
        }
    }
    const unsigned n = (unsigned)(fw * fh);
    const bool staged = tile != nullptr && n <= a.smem_floats; // the same for the whole block
    if (staged) {
        const PlanePtrs& pp = frame_ptrs(fsx);
        const int pitch = (int)pp.src_pitch[plane];
        const T* __restrict__ src = static_cast<const T*>(pp.src[plane]) + (long long)sy_lo * pitch + sx_lo;
        const float inv_fw = 1.f / (float)fw;
        for (unsigned e0 = threadIdx.x; e0 < n; e0 += 4 * THREADS) {
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { // all four loads are issued before the first conversion
                const unsigned e = min(e0 + u * THREADS, n - 1);
                unsigned row = (unsigned)__float2int_rd(((float)e + 0.5f) * inv_fw);
                row -= (row * (unsigned)fw > e);
                row += ((row + 1) * (unsigned)fw <= e);
                v[u] = __ldg(src + (long long)row * pitch + (e - row * (unsigned)fw));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (e0 + u * THREADS < n)
                    tile[e0 + u * THREADS] = sample_to_float(v[u]);
        }
        __syncthreads();
    }
    if (SPT > 1 && staged) {
        bool same = true;
        const int fsp = (fs + 3) & ~3;
#pragma unroll
        for (int k = 0; k < SPT; ++k)
            same = same && meta[k].x >= 0 && meta[k].w == meta[0].w && meta[k].wstride == fsp;
        if (same) {
            strip_samples_fused<T, FSC, SPT>(a, fsx, meta, (int)plane, tile, fw, sx_lo, sy_lo);
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
        if (meta[k].x < 0)
            continue;
        if (staged && meta[k].wstride)
            strip_sample_staged<T, FSC>(a, fsx, meta[k], (int)plane, tile, fw, sx_lo, sy_lo);
        else
            strip_sample<T, FSC>(a, fsx, meta[k].x, meta[k].y, (int)plane);
    }
}

// Role of block b in a merged grid of `interior` tile blocks and `strips` strip blocks: the strip blocks are spread
// evenly through the grid (their latency-bound work then hides under the FMA-bound tiles sharing the SM) instead of
// trailing it.  Returns true for a strip block and its index in `id`, else the tile index.
__device__ __forceinline__ bool block_role(unsigned b, unsigned interior, unsigned strips, unsigned& id)
{
    const unsigned long long total = (unsigned long long)interior + strips;
    const unsigned s0 = (unsigned)((unsigned long long)b * strips / total);
    const unsigned s1 = (unsigned)((unsigned long long)(b + 1) * strips / total);
    id = s1 > s0 ? s0 : b - s0;
    return s1 > s0;
}

struct GeneralArgs {
    FrameSet fr;
    StripArgs st;
};

constexpr int GEN_SPT = 4;                  // outputs per thread of the general kernel
constexpr int GEN_MAX_PW = 64;              // its patches are 64 x 16 outputs
constexpr size_t GEN_SMEM = (size_t)96 << 10; // staging space (two blocks per SM)

template <typename T>
__global__ void __launch_bounds__(STRIP_THREADS) resample_strips(const __grid_constant__ GeneralArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    strip_block<T, 0, STRIP_THREADS, GEN_SPT>(a.st, a.fr, blockIdx.x, reinterpret_cast<float*>(smem_raw));
}

// weights of one output pixel exactly as the reference defines them (introspection for parity tests)
__global__ void pixel_weights_kernel(StripArgs a, int x, int y, float* out)
{
    const int fs = a.fs;
    const int rx = a.rank_x[x], ry = a.rank_y[y];
    if (rx >= 0 && ry >= 0) {
        const float* w = a.weights + (size_t)(ry * a.n_rank_x + rx) * fs * fs;
        for (int t = threadIdx.x; t < fs * fs; t += blockDim.x)
            out[t] = w[t];
        return;
    }
    if (a.border_block) {
        // what the resample kernels apply: the block of this border pixel's class
        const int fsp = (fs + 3) & ~3;
        const float* w = a.border_wb + (size_t)a.border_block[jinc_border_slot(a.bg, x, y)] * (unsigned)(fs * fsp);
        for (int t = threadIdx.x; t < fs * fs; t += blockDim.x)
            out[t] = w[(t / fs) * fsp + t % fs];
        return;
    }
    const int sx = a.start_x[x], sy = a.start_y[y];
    const float px = a.pos_x[x], py = a.pos_y[y];
    const float sum = a.border_sum[jinc_border_slot(a.bg, x, y)];
    for (int t = threadIdx.x; t < fs * fs; t += blockDim.x) {
        const int ly = t / fs, lx = t - ly * fs;
        const double dy2 = jinc_tap_dist2(py, a.src_h, sy + ly, a.step_y);
        const double dx2 = jinc_tap_dist2(px, a.src_w, sx + lx, a.step_x);
        out[t] = __fdiv_rn(jinc_lut_weight(a.lut, __dadd_rn(dx2, dy2), a.radius2, a.idx_scale), sum);
    }
}

void fill_strip_args(const jinc_table* t, StripArgs& a)
{
    memset(&a, 0, sizeof(a));
    a.start_x = t->ax[0].start;
    a.start_y = t->ax[1].start;
    a.rank_x = t->ax[0].rank;
    a.rank_y = t->ax[1].rank;
    a.pos_x = t->ax[0].pos;
    a.pos_y = t->ax[1].pos;
    a.weights = t->d_weights;
    a.lut = t->d_lut;
    a.border_sum = t->d_border_sum;
    a.border_w = t->d_border_w;
    a.border_block = t->d_border_block;
    a.border_wb = t->d_border_wb;
    a.bg = t->bgeom;
    a.fs = t->sc.fs;
    a.n_rank_x = t->ax[0].n_rank;
    a.src_w = t->sc.src_w;
    a.src_h = t->sc.src_h;
    a.step_x = t->sc.filt_step[0];
    a.step_y = t->sc.filt_step[1];
    a.radius2 = t->sc.radius2;
    a.idx_scale = t->sc.idx_scale;
}

// Cuts the rectangles into patches of `outputs` samples (one strip block each), at most `max_pw` wide; returns the
// number of strip blocks per plane.
long long set_strip_rects(StripArgs& a, const Rect* rects, int n_rects, int outputs, int max_pw, size_t smem_bytes)
{
    unsigned total = 0;
    int k = 0;
    max_pw = std::min(max_pw, outputs);
    for (int r = 0; r < n_rects; ++r) {
        const long long w = rects[r].x1 - rects[r].x0, h = rects[r].y1 - rects[r].y0;
        if (w <= 0 || h <= 0)
            continue;
        int pwl = 3; // patches are at least 8 wide
        while ((1 << pwl) < w && (1 << pwl) < max_pw)
            ++pwl;
        const long long pw = 1ll << pwl, ph = outputs / pw;
        a.rect[k] = rects[r];
        a.pw_log2[k] = pwl;
        a.patches_x[k] = (unsigned)((w + pw - 1) / pw);
        a.patch_begin[k] = total;
        total += a.patches_x[k] * (unsigned)((h + ph - 1) / ph);
        ++k;
    }
    for (int j = k; j < 4; ++j) {
        a.rect[j] = Rect{0, 0, 1, 1};
        a.pw_log2[j] = 3;
        a.patches_x[j] = 1;
        a.patch_begin[j] = total;
    }
    a.patch_begin[4] = total;
    a.blocks_per_plane = total;
    a.smem_floats = (unsigned)(smem_bytes / sizeof(float));
    return total;
}

// ------------------------------------------------------------------------------------------ exact-2x kernel

constexpr int UP_TX = 4;                     // cells per thread along x
constexpr int UP_WARPS = 8;
constexpr int UP_THREADS = UP_WARPS * 32;
constexpr int UP_CW = 32 * UP_TX;            // cells per tile row (128 -> 256 output samples)
constexpr int UP_RPW = 2;                    // cell-row pairs per warp
constexpr int UP_CH = 2 * UP_WARPS * UP_RPW; // cell rows per tile (32 -> 64 output rows)
constexpr int UP_STRIP_SPT = 4;              // strip role: outputs per thread (patches of 1024 outputs, up to 256 wide)
constexpr int UP_STRIP_MAX_PW = 1024;            // a thread's four samples share a border row: x, x + 256, ...

template <int FS>
struct UpGeom {
    static constexpr int FSP = (FS + 3) & ~3;       // weight row stride (16-byte rows)
    static constexpr int NSEG = UP_TX + 1 + FS - 1; // pair columns a thread reads per row (ox1 <= 1)
    static constexpr int NC = UP_CW + FS;           // pair columns per tile row (CW + ox1 + FS - 1)
    static constexpr int NCP = (NC + 3) & ~3;
    static constexpr int SUB = NCP / 4;             // columns are de-interleaved by (c & 3): 4 sub-rows of SUB
    static constexpr int NR = UP_CH + 1 + FS - 1;   // pair rows per tile (CH + oy1 + FS - 1)
    static constexpr size_t SMEM = (size_t)NR * NCP * sizeof(float2);
};

template <int FS>
struct alignas(16) UpWeights {
    float w[2][2][FS][UpGeom<FS>::FSP]; // [py][px][ly][lx]
};

struct UpArgs {
    FrameSet fr;
    StripArgs st;         // border strips around the interior (run by the blocks after the interior tiles)
    int src_w, src_h;
    int x0, y0, ncx;      // output origin of the periodic interior, cells per row
    int sx0, sy0;         // window origin of cell (0,0), phase (0,0)
    int cy_begin, cy_end; // cell rows to produce (row-band split)
    int tiles_x, tiles_per_plane, interior_blocks; // interior_blocks = tiles_per_plane * n_planes
    int strip_blocks;
};

template <typename T, int FS, int OX1, int OY1>
__global__ void __launch_bounds__(UP_THREADS, (FS >= 13 ? 2 : 3))
    resample_up2x(const __grid_constant__ UpArgs a, const __grid_constant__ UpWeights<FS> W)
{
    using G = UpGeom<FS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    unsigned role_id;
    if (block_role(blockIdx.x, a.interior_blocks, a.strip_blocks, role_id)) {
        // ---------------- strip role
        strip_block<T, FS, UP_THREADS, UP_STRIP_SPT>(a.st, a.fr, role_id, reinterpret_cast<float*>(smem_raw));
        return;
    }

    // -------------------- interior tile role
    float2* tile = reinterpret_cast<float2*>(smem_raw); // [NR][4][SUB] pairs {S[r][c], S[r+1][c]}
    const int plane = role_id / a.tiles_per_plane;
    const int tidx = role_id - plane * a.tiles_per_plane;
    const int tile_y = tidx / a.tiles_x, tile_x = tidx - tile_y * a.tiles_x;
    const PlanePtrs& pp = frame_ptrs(a.fr);
    const T* __restrict__ src = static_cast<const T*>(pp.src[plane]);
    T* __restrict__ dst = static_cast<T*>(pp.dst[plane]);
    const long long sp = pp.src_pitch[plane], dp = pp.dst_pitch[plane];

    const int cell_x0 = tile_x * UP_CW; // first cell of this tile
    const int cell_y0 = a.cy_begin + tile_y * UP_CH;
    const int tsx = a.sx0 + cell_x0, tsy = a.sy0 + cell_y0; // source coordinates of tile(0,0)

    // ---- stage the source tile.  A thread owns 4 consecutive columns and walks down a segment of rows, pairing each
    //      row with the one above it, so every source sample is loaded and converted once per segment.
    {
        constexpr int SEGS = UP_THREADS / G::SUB;       // row segments
        constexpr int ROWS = (G::NR + SEGS - 1) / SEGS; // pair rows per segment
        const int q = threadIdx.x % G::SUB, seg = threadIdx.x / G::SUB;
        if (seg < SEGS) {
            int xo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                xo[k] = min(max(tsx + 4 * q + k, 0), a.src_w - 1); // out-of-plane taps only feed discarded cells
            const int r0 = seg * ROWS, r1 = min(r0 + ROWS, G::NR);
            float prev[4], cur[4];
            {
                const T* row = src + (long long)min(max(tsy + r0, 0), a.src_h - 1) * sp;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    prev[k] = load_sample(row + xo[k]);
            }
            for (int r = r0; r < r1; ++r) {
                const T* row = src + (long long)min(max(tsy + r + 1, 0), a.src_h - 1) * sp;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    cur[k] = load_sample(row + xo[k]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tile[(r * 4 + k) * G::SUB + q] = make_float2(prev[k], cur[k]);
                    prev[k] = cur[k];
                }
            }
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

#pragma unroll 1
    for (int rp = warp; rp < UP_WARPS * UP_RPW; rp += UP_WARPS) {
        const int cy = cell_y0 + 2 * rp; // first cell row of the pair
        if (cy >= a.cy_end)
            break;
        float2 acc[2][2][UP_TX];
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px)
#pragma unroll
                for (int i = 0; i < UP_TX; ++i)
                    acc[py][px][i] = make_float2(0.f, 0.f);

        const float2* trow = tile + (size_t)(2 * rp) * G::NCP + lane;
#pragma unroll 1
        for (int rr = 0; rr < FS + OY1; ++rr) {
            float2 seg[G::NSEG];
#pragma unroll
            for (int m = 0; m < G::NSEG; ++m)
                seg[m] = trow[(m & 3) * G::SUB + (m >> 2)]; // column 4*lane + m
            trow += G::NCP;

            if (OY1 == 0 || rr < FS) { // phase row 0: ly = rr
#pragma unroll
                for (int lx = 0; lx < FS; ++lx) {
                    const float w0 = W.w[0][0][rr][lx], w1 = W.w[0][1][rr][lx];
#pragma unroll
                    for (int i = 0; i < UP_TX; ++i) {
                        acc[0][0][i] = __ffma2_rn(seg[i + lx], make_float2(w0, w0), acc[0][0][i]);
                        acc[0][1][i] = __ffma2_rn(seg[i + OX1 + lx], make_float2(w1, w1), acc[0][1][i]);
                    }
                }
            }
            const int ly1 = rr - OY1; // phase row 1
            if (OY1 == 0 || rr >= OY1) {
#pragma unroll
                for (int lx = 0; lx < FS; ++lx) {
                    const float w0 = W.w[1][0][ly1][lx], w1 = W.w[1][1][ly1][lx];
#pragma unroll
                    for (int i = 0; i < UP_TX; ++i) {
                        acc[1][0][i] = __ffma2_rn(seg[i + lx], make_float2(w0, w0), acc[1][0][i]);
                        acc[1][1][i] = __ffma2_rn(seg[i + OX1 + lx], make_float2(w1, w1), acc[1][1][i]);
                    }
                }
            }
        }

        // ---- epilogue: 4 output rows x 8 consecutive samples per thread
        const int cx = cell_x0 + UP_TX * lane;
        if (cx >= a.ncx)
            continue;
        const int ox = a.x0 + 2 * cx;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (cy + h < a.cy_end) {
#pragma unroll
                for (int py = 0; py < 2; ++py) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < UP_TX; ++i) {
                        v[2 * i] = h ? acc[py][0][i].y : acc[py][0][i].x;
                        v[2 * i + 1] = h ? acc[py][1][i].y : acc[py][1][i].x;
                    }
                    T* o = dst + (long long)(a.y0 + 2 * (cy + h) + py) * dp + ox;
                    if (cx + UP_TX <= a.ncx) {
                        store8<T>(o, v, a.fr.peak);
                    } else {
                        const int nk = 2 * (a.ncx - cx);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < nk)
                                o[k] = finish<T>(v[k], a.fr.peak);
                    }
                }
            }
        }
    }
}

template <typename T, int FS>
int launch_up2x_fs(const jinc_table* t, UpArgs& a, long long strip_blocks, int n_frames, cudaStream_t st)
{
    using G = UpGeom<FS>;
    const Up2xPlan& u = t->up2x;
    UpWeights<FS> w;
    memset(&w, 0, sizeof(w));
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            const float* blk = t->h_weights.data() + (size_t)u.wblock[py][px] * FS * FS;
            for (int ly = 0; ly < FS; ++ly)
                for (int lx = 0; lx < FS; ++lx)
                    w.w[py][px][ly][lx] = blk[ly * FS + lx];
        }
    auto kern = u.ox1 ? (u.oy1 ? resample_up2x<T, FS, 1, 1> : resample_up2x<T, FS, 1, 0>)
                      : (u.oy1 ? resample_up2x<T, FS, 0, 1> : resample_up2x<T, FS, 0, 0>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "cudaFuncSetAttribute(up2x smem %zu): %s", G::SMEM, cudaGetErrorString(e));
    a.strip_blocks = (int)strip_blocks;
    dim3 grid((unsigned)(a.interior_blocks + strip_blocks), n_frames, 1);
    kern<<<grid, UP_THREADS, G::SMEM, st>>>(a, w);
    e = cudaGetLastError();
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "resample_up2x launch failed: %s", cudaGetErrorString(e));
    return JINC_OK;
}

template <typename T>
int launch_up2x(const jinc_table* t, UpArgs& a, long long strip_blocks, int n_frames, cudaStream_t st)
{
    switch (t->sc.fs) {
    case 7: return launch_up2x_fs<T, 7>(t, a, strip_blocks, n_frames, st);   // tap 3  (Jinc36Resize)
    case 9: return launch_up2x_fs<T, 9>(t, a, strip_blocks, n_frames, st);   // tap 4  (Jinc64Resize)
    case 13: return launch_up2x_fs<T, 13>(t, a, strip_blocks, n_frames, st); // tap 6  (Jinc144Resize)
    case 17: return launch_up2x_fs<T, 17>(t, a, strip_blocks, n_frames, st); // tap 8  (Jinc256Resize)
    default: return 1;
    }
}

bool up2x_supported(int fs) { return fs == 7 || fs == 9 || fs == 13 || fs == 17; }

size_t up2x_smem_bytes(int fs)
{
    switch (fs) {
    case 7: return UpGeom<7>::SMEM;
    case 9: return UpGeom<9>::SMEM;
    case 13: return UpGeom<13>::SMEM;
    case 17: return UpGeom<17>::SMEM;
    default: return 0;
    }
}

// ------------------------------------------------------------------------------------------ integer-ratio downscale kernel
//
// Output (x,y) of the interior reads the FS x FS window at (sx0 + Q*x, sy0 + Q*y) with ONE weight block for every
// pixel (config 5: Q = 4, FS = 50, 2500 taps per sample).  All threads apply the same weight at the same time, so
// weights again come from the constant bank through uniform registers.  Three ideas shape the kernel:
//   * polyphase columns: lx = Q*m + p turns the x-sum into Q stride-1 convolutions over the de-interleaved sequences
//     S_p[j] = S[Q*j + p]; a thread that owns NX consecutive outputs reads a span of NX+MT-1 values per (row, p) and
//     uses each for up to NX outputs, and one weight fetch feeds NX FFMA2s;
//   * tap pairing: one packed FFMA2 multiplies the vertical sample pair {S[r][c], S[r+1][c]} with the weight pair
//     {w[ly][lx], w[ly+1][lx]} into the two halves of ONE output's accumulator (even-row and odd-row partial sums,
//     added in the epilogue).  Q is even, so every output row of the thread sees the same pairing and a staged pair
//     is reused for all NY output rows of the thread;
//   * raw sample pairs in shared memory for integer formats (two 16-bit samples per 32-bit word; u8 is widened while
//     staging), so a 64x32-output tile with its 302x174-sample footprint fits twice per SM.  The float value is made
//     after the shared-memory load.  For depths up to 15 bits that costs ONE byte permute per sample: staging stores
//     x << (15 - bits), and PRMT drops those 16 bits into mantissa bits [22:8] of 0x3F000000, i.e. f = 0.5 + x' / 65536
//     exactly.  The kernel accumulates sum(w * f) and the epilogue removes the 0.5 * sum(w) bias (host-computed per
//     accumulator half) and rescales by a power of two.  Precision matches a direct float sum of 15-bit samples (the
//     accumulator's ulp relative to one input LSB is the same).  16-bit samples use I2F instead (XU pipe).
// Columns are de-interleaved by c mod (Q*NX) so a warp's loads are bank-conflict free.
constexpr int DN_TW = 64;  // output columns per tile
constexpr int DN_TH = 32;  // output rows per tile (integer formats; float tiles are half as tall)

constexpr int DN_STRIP_SPT = 2;    // strip role: outputs per thread, patches at most 64 wide (the windows are wide)
constexpr int DN_STRIP_MAX_PW = 256;

enum { DN_CVT_I2F = 0, DN_CVT_PRMT = 1, DN_CVT_FLOAT = 2 };

template <typename T, int FS, int Q, int NX, int NY>
struct DownGeom {
    static constexpr bool IS_FLOAT = sizeof(T) == 4;
    using Word = typename std::conditional<IS_FLOAT, float2, uint32_t>::type; // {row 2k, row 2k+1}
    static constexpr int LX = DN_TW / NX;               // lanes along x
    static constexpr int LY = 32 / LX;                  // lanes along y
    static constexpr int TH = IS_FLOAT ? DN_TH / 2 : DN_TH;
    static constexpr int WARPS = TH / (LY * NY);
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int FSE = (FS + 1) & ~1;           // window rows rounded up to whole pairs
    static constexpr int NKW = FSE / 2;                 // weight row pairs
    static constexpr int MT = (FS + Q - 1) / Q;         // taps per polyphase component
    static constexpr int SPAN = NX + MT - 1;            // values a thread reads per (row pair, p)
    static constexpr int D = Q * NX;                    // column de-interleave modulus
    static constexpr int NCOL = Q * (DN_TW - 1) + Q * (MT - 1) + Q; // columns a tile row can be asked for
    static constexpr int SUB = (NCOL + D - 1) / D;
    static constexpr int NROWS = Q * (TH - 1) + FSE;
    static constexpr int NROWP = (NROWS + 1) / 2;       // row pairs per tile
    static constexpr int JSTEP = Q / 2;                 // row pairs between consecutive output rows
    static constexpr int NK = NKW + JSTEP * (NY - 1);   // row pairs a thread walks
    static constexpr int HALF = NY * JSTEP;             // row pairs between lane groups that differ in y
    static constexpr int rs_pad()
    {
        for (int pad = 0; pad < 32; ++pad) // lane group g lands on banks [g*LX, g*LX + LX)
            if (((D * SUB + pad) * HALF) % 32 == LX % 32)
                return pad;
        return 0;
    }
    static constexpr int RS = D * SUB + rs_pad();       // row-pair stride in words
    static constexpr size_t SMEM = (size_t)NROWP * RS * sizeof(Word);
    static_assert(Q % 2 == 0, "tap pairing needs an even ratio");
    static_assert(TH % (LY * NY) == 0 && WARPS >= 1, "tile rows must split evenly over the warps");
    static_assert(LX - 1 + (Q * (SPAN - 1) + Q - 1) / D < SUB, "span reaches past the tile row");
};

template <int FS, int Q>
struct alignas(16) DownWeights {
    static constexpr int MT = (FS + Q - 1) / Q;
    float2 w[((FS + 1) & ~1) / 2][Q][MT]; // [ly/2][p][m] = {w[ly][Q*m+p], w[ly+1][Q*m+p]}; entries outside the window are 0
};

struct DownArgs {
    FrameSet fr;
    StripArgs st;
    int src_w, src_h;
    int x0, y0, x1, y1;  // output rectangle produced by the tiles (y0..y1 already cut to the row band)
    int tsx0, tsy0;      // window origin of output (x0, y0)
    int tiles_x, tiles_per_plane, interior_blocks, strip_blocks;
    int pre_shift;       // PRMT conversion: samples are staged as x << pre_shift
    float bias_even, bias_odd, out_scale; // PRMT conversion: out = ((acc.x - bias_even) + (acc.y - bias_odd)) * out_scale
};

template <int CVT>
__device__ __forceinline__ float2 down_cvt(uint32_t w)
{
    if (CVT == DN_CVT_PRMT)
        return make_float2(__uint_as_float(__byte_perm(w, 0x3F000000u, 0x7104)), __uint_as_float(__byte_perm(w, 0x3F000000u, 0x7324)));
    return make_float2((float)(w & 0xffffu), (float)(w >> 16));
}
template <int CVT>
__device__ __forceinline__ float2 down_cvt(float2 w)
{
    return w;
}

template <typename T>
__device__ __forceinline__ void down_pack(uint32_t& out, const T* r0, const T* r1, int sh)
{
    out = ((uint32_t)__ldg(r0) << sh) | ((uint32_t)__ldg(r1) << (16 + sh));
}
__device__ __forceinline__ void down_pack(float2& out, const float* r0, const float* r1, int)
{
    out = make_float2(__ldg(r0), __ldg(r1));
}

template <typename T, int N>
__device__ __forceinline__ void store_run(T* p, const float (&v)[N], float peak)
{
    static_assert(N % 4 == 0, "runs are multiples of four samples");
#pragma unroll
    for (int q = 0; q < N; q += 4) {
        if (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(p + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        } else if (sizeof(T) == 2) {
            *reinterpret_cast<uint2*>(p + q) = make_uint2(finish_u16(v[q], peak) | (finish_u16(v[q + 1], peak) << 16),
                                                          finish_u16(v[q + 2], peak) | (finish_u16(v[q + 3], peak) << 16));
        } else {
            *reinterpret_cast<uint32_t*>(p + q) = finish_u8(v[q], peak) | (finish_u8(v[q + 1], peak) << 8) |
                                                  (finish_u8(v[q + 2], peak) << 16) | (finish_u8(v[q + 3], peak) << 24);
        }
    }
}

// one row pair of the thread's walk.  ALL: every output row of the thread is inside its window (no tests)
template <typename G, int FS, int Q, int NX, int NY, int CVT, bool ALL>
__device__ __forceinline__ void down_row_pair(const typename G::Word* __restrict__ trow, int k, const DownWeights<FS, Q>& W,
                                              float2 (&acc)[NY][NX])
{
#pragma unroll
    for (int p = 0; p < Q; ++p) {
        float2 s[G::SPAN];
#pragma unroll
        for (int m = 0; m < G::SPAN; ++m) {
            const int cc = Q * m + p;
            s[m] = down_cvt<CVT>(trow[(cc % G::D) * G::SUB + cc / G::D]);
        }
#pragma unroll
        for (int j = 0; j < NY; ++j) {
            const int kk = k - j * G::JSTEP; // weight row pair of output row j
            if (ALL || (kk >= 0 && kk < G::NKW)) {
#pragma unroll
                for (int m = 0; m < G::MT; ++m) {
                    if (Q * m + p < FS) {
                        const float2 w = W.w[kk][p][m];
#pragma unroll
                        for (int i = 0; i < NX; ++i)
                            acc[j][i] = __ffma2_rn(s[i + m], w, acc[j][i]);
                    }
                }
            }
        }
    }
}

template <typename T, int FS, int Q, int NX, int NY, int CVT>
__global__ void __launch_bounds__((DownGeom<T, FS, Q, NX, NY>::THREADS), 2)
    resample_down(const __grid_constant__ DownArgs a, const __grid_constant__ DownWeights<FS, Q> W)
{
    using G = DownGeom<T, FS, Q, NX, NY>;
    using Word = typename G::Word;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    unsigned role_id;
    if (block_role(blockIdx.x, a.interior_blocks, a.strip_blocks, role_id)) {
        strip_block<T, FS, G::THREADS, DN_STRIP_SPT>(a.st, a.fr, role_id, reinterpret_cast<float*>(smem_raw));
        return;
    }
    Word* tile = reinterpret_cast<Word*>(smem_raw); // [NROWP][D][SUB] (+pad): word (k, c) at k*RS + (c%D)*SUB + c/D
    const int plane = role_id / a.tiles_per_plane;
    const int tidx = role_id - plane * a.tiles_per_plane;
    const int tile_y = tidx / a.tiles_x, tile_x = tidx - tile_y * a.tiles_x;
    const PlanePtrs& pp = frame_ptrs(a.fr);
    const T* __restrict__ src = static_cast<const T*>(pp.src[plane]);
    T* __restrict__ dst = static_cast<T*>(pp.dst[plane]);
    const int sp = (int)pp.src_pitch[plane];
    const long long dp = pp.dst_pitch[plane];

    const int ox0 = a.x0 + tile_x * DN_TW, oy0 = a.y0 + tile_y * G::TH; // first output of the tile
    const int tsx = a.tsx0 + Q * (tile_x * DN_TW), tsy = a.tsy0 + Q * (tile_y * G::TH);

    // ---- stage the tile: a warp takes whole row pairs, a lane the columns lane + 32 q.  All loads of KU row pairs are
    //      issued before the first store so ~40 global loads per thread are in flight.
    {
        constexpr int NCOLS = G::D * G::SUB;
        constexpr int CQ = (NCOLS + 31) / 32;
        constexpr int KU = 2;
        const int sh = CVT == DN_CVT_PRMT ? a.pre_shift : 0;
        const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
        int gx[CQ];
#pragma unroll
        for (int q = 0; q < CQ; ++q)
            gx[q] = min(max(tsx + lane_ + 32 * q, 0), a.src_w - 1); // clamped taps only feed masked outputs or zero weights
        for (int k0 = warp_; k0 < G::NROWP; k0 += KU * G::WARPS) {
            Word wv[KU][CQ];
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const int k = min(k0 + u * G::WARPS, G::NROWP - 1);
                const T* r0 = src + (long long)min(max(tsy + 2 * k, 0), a.src_h - 1) * sp;
                const T* r1 = src + (long long)min(max(tsy + 2 * k + 1, 0), a.src_h - 1) * sp;
#pragma unroll
                for (int q = 0; q < CQ; ++q)
                    down_pack(wv[u][q], r0 + gx[q], r1 + gx[q], sh);
            }
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                const int k = k0 + u * G::WARPS;
                if (k < G::NROWP) {
#pragma unroll
                    for (int q = 0; q < CQ; ++q) {
                        const int c = lane_ + 32 * q;
                        if (NCOLS % 32 == 0 || c < NCOLS)
                            tile[k * G::RS + (c % G::D) * G::SUB + c / G::D] = wv[u][q];
                    }
                }
            }
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx_ = lane % G::LX, ly_ = lane / G::LX;
    const int row0 = (warp * G::LY + ly_) * NY; // first output row of this thread inside the tile
    const Word* __restrict__ trow = tile + (size_t)(row0 * G::JSTEP) * G::RS + lx_;

    float2 acc[NY][NX];
#pragma unroll
    for (int j = 0; j < NY; ++j)
#pragma unroll
        for (int i = 0; i < NX; ++i)
            acc[j][i] = make_float2(0.f, 0.f);

    constexpr int K_ALL0 = G::JSTEP * (NY - 1); // first row pair at which every output row is inside its window
    int k = 0;
#pragma unroll 1
    for (; k < K_ALL0; ++k, trow += G::RS)
        down_row_pair<G, FS, Q, NX, NY, CVT, false>(trow, k, W, acc);
#pragma unroll 1
    for (; k < G::NKW; ++k, trow += G::RS)
        down_row_pair<G, FS, Q, NX, NY, CVT, true>(trow, k, W, acc);
#pragma unroll 1
    for (; k < G::NK; ++k, trow += G::RS)
        down_row_pair<G, FS, Q, NX, NY, CVT, false>(trow, k, W, acc);

    // ---- epilogue: NY rows x NX consecutive samples
    const int ox = ox0 + NX * lx_;
    if (ox >= a.x1)
        return;
#pragma unroll
    for (int j = 0; j < NY; ++j) {
        const int oy = oy0 + row0 + j;
        if (oy >= a.y1)
            break;
        float v[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            if (CVT == DN_CVT_PRMT)
                v[i] = ((acc[j][i].x - a.bias_even) + (acc[j][i].y - a.bias_odd)) * a.out_scale;
            else
                v[i] = acc[j][i].x + acc[j][i].y;
        }
        T* o = dst + (long long)oy * dp + ox;
        if (ox + NX <= a.x1) {
            store_run<T, NX>(o, v, a.fr.peak);
        } else {
#pragma unroll
            for (int i = 0; i < NX; ++i)
                if (ox + i < a.x1)
                    o[i] = finish<T>(v[i], a.fr.peak);
        }
    }
}

template <typename T, int FS, int Q, int NX, int NY, int CVT>
int launch_down_cfg(DownArgs& a, const DownWeights<FS, Q>& w, long long strip_blocks_of, int n_frames, cudaStream_t st,
                    const Rect* rects, int n_rects)
{
    using G = DownGeom<T, FS, Q, NX, NY>;
    const long long strip_blocks =
        strip_blocks_of ? set_strip_rects(a.st, rects, n_rects, G::THREADS * DN_STRIP_SPT, DN_STRIP_MAX_PW, G::SMEM) * a.fr.n_planes : 0;
    a.tiles_x = (a.x1 - a.x0 + DN_TW - 1) / DN_TW;
    a.tiles_per_plane = a.tiles_x * ((a.y1 - a.y0 + G::TH - 1) / G::TH);
    if (a.interior_blocks)
        a.interior_blocks = a.tiles_per_plane * a.fr.n_planes;
    auto kern = resample_down<T, FS, Q, NX, NY, CVT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "cudaFuncSetAttribute(down smem %zu): %s", G::SMEM, cudaGetErrorString(e));
    if (a.interior_blocks + strip_blocks == 0)
        return 2;
    a.strip_blocks = (int)strip_blocks;
    dim3 grid((unsigned)(a.interior_blocks + strip_blocks), n_frames, 1);
    kern<<<grid, G::THREADS, G::SMEM, st>>>(a, w);
    e = cudaGetLastError();
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "resample_down launch failed: %s", cudaGetErrorString(e));
    return JINC_OK;
}

constexpr int DN_NX = 8, DN_NY = 2; // outputs per thread

template <typename T, int FS, int Q>
int launch_down_fs(const jinc_table* t, DownArgs& a, bool want_strips, int n_frames, cudaStream_t st, const Rect* rects, int n_rects)
{
    static_assert(sizeof(DownWeights<FS, Q>) + sizeof(DownArgs) < 32000, "kernel parameters exceed the 32 KB limit");
    const DownPlan& d = t->down;
    DownWeights<FS, Q> w;
    memset(&w, 0, sizeof(w));
    const float* blk = t->h_weights.data() + (size_t)d.wblock * FS * FS;
    double sum_even = 0.0, sum_odd = 0.0;
    for (int ly = 0; ly < FS; ++ly)
        for (int lx = 0; lx < FS; ++lx) {
            const float v = blk[ly * FS + lx];
            float2& e = w.w[ly >> 1][lx % Q][lx / Q];
            if (ly & 1) {
                e.y = v;
                sum_odd += v;
            } else {
                e.x = v;
                sum_even += v;
            }
        }
    if constexpr (sizeof(T) == 4) {
        return launch_down_cfg<T, FS, Q, DN_NX, DN_NY, DN_CVT_FLOAT>(a, w, want_strips, n_frames, st, rects, n_rects);
    } else {
        const int bits = t_bits_from_peak(a.fr.peak);
        if (bits <= 15) {
            // f = 0.5 + (x << pre_shift) / 65536  =>  sum(w f) = 0.5 sum(w) + sum(w x) * 2^(pre_shift - 16)
            a.pre_shift = 15 - bits;
            a.bias_even = (float)(0.5 * sum_even);
            a.bias_odd = (float)(0.5 * sum_odd);
            a.out_scale = (float)(1 << (16 - a.pre_shift));
            return launch_down_cfg<T, FS, Q, DN_NX, DN_NY, DN_CVT_PRMT>(a, w, want_strips, n_frames, st, rects, n_rects);
        }
        if constexpr (sizeof(T) == 2)
            return launch_down_cfg<T, FS, Q, DN_NX, DN_NY, DN_CVT_I2F>(a, w, want_strips, n_frames, st, rects, n_rects);
        return 1;
    }
}

// 0 launched, 2 nothing to do, 1 unsupported geometry, <0 error
template <typename T>
int launch_down(const jinc_table* t, DownArgs& a, bool want_strips, int n_frames, cudaStream_t st, const Rect* rects, int n_rects)
{
    const int key = t->down.qx * 1000 + t->sc.fs;
    switch (key) {
#define JINC_DOWN_CASE(Q_, FS_) \
    case Q_ * 1000 + FS_: return launch_down_fs<T, FS_, Q_>(t, a, want_strips, n_frames, st, rects, n_rects);
        JINC_DOWN_CASE(2, 13) // tap 3, 1/2
        JINC_DOWN_CASE(2, 17) // tap 4, 1/2
        JINC_DOWN_CASE(2, 25) // tap 6, 1/2
        JINC_DOWN_CASE(2, 33) // tap 8, 1/2
        JINC_DOWN_CASE(4, 26) // tap 3, 1/4
        JINC_DOWN_CASE(4, 34) // tap 4, 1/4
        JINC_DOWN_CASE(4, 50) // tap 6, 1/4
#undef JINC_DOWN_CASE
    default: return 1;
    }
}

bool down_supported(const jinc_table* t)
{
    if (!t->down.ok || t->down.qx != t->down.qy)
        return false;
    switch (t->down.qx * 1000 + t->sc.fs) {
    case 2013: case 2017: case 2025: case 2033: case 4026: case 4034: case 4050: return true;
    default: return false;
    }
}

// ------------------------------------------------------------------------------------------ launcher

template <typename T>
int launch_typed(const jinc_table* t, const FrameSet& fr, int n_frames, int y_begin, int y_end, cudaStream_t st, int* launches,
                 int parts)
{
    const int W = t->sc.dst_w;
    StripArgs sa;
    fill_strip_args(t, sa);
    Rect rects[4];
    int n_rects = 0;

    if (t->fast_path == JINC_PATH_UP2X && up2x_supported(t->sc.fs)) {
        const Up2xPlan& u = t->up2x;
        // cell rows whose 2 output rows lie inside [y_begin, y_end)
        const int cb = (std::max(y_begin, u.y0) - u.y0 + 1) / 2;
        const int ce = (std::min(y_end, u.y0 + 2 * u.ncy) - u.y0) / 2;
        if (ce > cb) {
            const int fy0 = u.y0 + 2 * cb, fy1 = u.y0 + 2 * ce;
            if (parts & JINC_PART_BORDER) {
                rects[n_rects++] = Rect{0, y_begin, W, fy0};  // top strip
                rects[n_rects++] = Rect{0, fy1, W, y_end};    // bottom strip
                rects[n_rects++] = Rect{0, fy0, t->ix0, fy1}; // left strip
                rects[n_rects++] = Rect{t->ix1, fy0, W, fy1}; // right strip
            }
            UpArgs a;
            memset(&a, 0, sizeof(a));
            a.fr = fr;
            a.st = sa;
            const long long strip_blocks =
                set_strip_rects(a.st, rects, n_rects, UP_THREADS * UP_STRIP_SPT, UP_STRIP_MAX_PW, up2x_smem_bytes(t->sc.fs)) * fr.n_planes;
            a.src_w = t->sc.src_w;
            a.src_h = t->sc.src_h;
            a.x0 = u.x0;
            a.y0 = u.y0;
            a.ncx = u.ncx;
            a.sx0 = u.sx0;
            a.sy0 = u.sy0;
            a.cy_begin = cb;
            a.cy_end = ce;
            a.tiles_x = (u.ncx + UP_CW - 1) / UP_CW;
            a.tiles_per_plane = a.tiles_x * ((ce - cb + UP_CH - 1) / UP_CH);
            a.interior_blocks = (parts & JINC_PART_INTERIOR) ? a.tiles_per_plane * fr.n_planes : 0;
            if (a.interior_blocks + strip_blocks == 0)
                return JINC_OK;
            const int rc = launch_up2x<T>(t, a, strip_blocks, n_frames, st);
            if (rc <= 0) {
                if (rc == 0)
                    ++*launches;
                return rc;
            }
        }
    }
    if (t->fast_path == JINC_PATH_DOWN_INT && down_supported(t)) {
        const DownPlan& d = t->down;
        const int fy0 = std::max(y_begin, d.y0), fy1 = std::min(y_end, d.y0 + d.ny);
        if (fy1 > fy0) {
            if (parts & JINC_PART_BORDER) {
                rects[n_rects++] = Rect{0, y_begin, W, fy0};
                rects[n_rects++] = Rect{0, fy1, W, y_end};
                rects[n_rects++] = Rect{0, fy0, t->ix0, fy1};
                rects[n_rects++] = Rect{t->ix1, fy0, W, fy1};
            }
            DownArgs a;
            memset(&a, 0, sizeof(a));
            a.fr = fr;
            a.st = sa;
            a.src_w = t->sc.src_w;
            a.src_h = t->sc.src_h;
            a.x0 = d.x0;
            a.x1 = d.x0 + d.nx;
            a.y0 = fy0;
            a.y1 = fy1;
            a.tsx0 = d.sx0;
            a.tsy0 = d.sy0 + d.qy * (fy0 - d.y0);
            a.interior_blocks = (parts & JINC_PART_INTERIOR) ? 1 : 0; // resolved to the tile count by the launcher
            const int rc = launch_down<T>(t, a, n_rects > 0, n_frames, st, rects, n_rects);
            if (rc != 1) {
                if (rc == 0)
                    ++*launches;
                return rc == 2 ? JINC_OK : rc;
            }
            n_rects = 0;
        }
    }
    if (!(parts & JINC_PART_BORDER))
        return JINC_OK;
    // no fast path for this geometry (or band): every sample goes through the strip role
    GeneralArgs ga;
    ga.fr = fr;
    ga.st = sa;
    rects[0] = Rect{0, y_begin, W, y_end};
    const long long blocks = set_strip_rects(ga.st, rects, 1, STRIP_THREADS * GEN_SPT, GEN_MAX_PW, GEN_SMEM) * fr.n_planes;
    if (blocks == 0)
        return JINC_OK;
    cudaError_t e = cudaFuncSetAttribute(resample_strips<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEN_SMEM);
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "cudaFuncSetAttribute(strips smem): %s", cudaGetErrorString(e));
    resample_strips<T><<<dim3((unsigned)blocks, n_frames), STRIP_THREADS, GEN_SMEM, st>>>(ga);
    e = cudaGetLastError();
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "resample_strips launch failed: %s", cudaGetErrorString(e));
    ++*launches;
    return JINC_OK;
}

int check_and_fill(PlanePtrs& pl, int sample_bytes, int n_planes, const void* const* d_src, const ptrdiff_t* src_pitch,
                   void* const* d_dst, const ptrdiff_t* dst_pitch)
{
    memset(&pl, 0, sizeof(pl));
    for (int i = 0; i < n_planes; ++i) {
        if (!d_src[i] || !d_dst[i])
            return jinc_fail(JINC_E_INVALID, "resize: null plane pointer");
        if (src_pitch[i] % sample_bytes || dst_pitch[i] % sample_bytes)
            return jinc_fail(JINC_E_INVALID, "resize: pitch must be a multiple of the sample size");
        if (reinterpret_cast<uintptr_t>(d_dst[i]) % 16 || dst_pitch[i] % 16)
            return jinc_fail(JINC_E_INVALID, "resize: device destination planes must be 16-byte aligned (base and pitch)");
        pl.src[i] = d_src[i];
        pl.dst[i] = d_dst[i];
        pl.src_pitch[i] = src_pitch[i] / sample_bytes;
        pl.dst_pitch[i] = dst_pitch[i] / sample_bytes;
    }
    return JINC_OK;
}

int dispatch(const jinc_table* t, int sample_bytes, const FrameSet& fr, int n_frames, int y_begin, int y_end, cudaStream_t stream,
             int* launches, int parts)
{
    y_begin = std::max(y_begin, 0);
    y_end = std::min(y_end, t->sc.dst_h);
    if (y_end <= y_begin || n_frames <= 0)
        return JINC_OK;
    int dummy = 0;
    if (!launches)
        launches = &dummy;
    switch (sample_bytes) {
    case 1: return launch_typed<uint8_t>(t, fr, n_frames, y_begin, y_end, stream, launches, parts);
    case 2: return launch_typed<uint16_t>(t, fr, n_frames, y_begin, y_end, stream, launches, parts);
    case 4: return launch_typed<float>(t, fr, n_frames, y_begin, y_end, stream, launches, parts);
    default: return jinc_fail(JINC_E_INVALID, "resize: sample_bytes must be 1, 2 or 4");
    }
}

} // namespace

size_t jinc_plane_ptrs_size() { return sizeof(PlanePtrs); }

int jinc_pack_plane_ptrs(void* out, int sample_bytes, int n_planes, const void* const* d_src, const ptrdiff_t* src_pitch,
                         void* const* d_dst, const ptrdiff_t* dst_pitch)
{
    if (n_planes < 1 || n_planes > JINC_MAX_PLANES)
        return jinc_fail(JINC_E_INVALID, "resize: n_planes must be 1..4");
    return check_and_fill(*static_cast<PlanePtrs*>(out), sample_bytes, n_planes, d_src, src_pitch, d_dst, dst_pitch);
}

int jinc_launch_resize_planes(jinc_ctx* ctx, const jinc_table* t, int sample_bytes, float peak, int n_planes,
                              const void* const* d_src, const ptrdiff_t* src_pitch, void* const* d_dst,
                              const ptrdiff_t* dst_pitch, int y_begin, int y_end, cudaStream_t stream, int* launches,
                              int parts)
{
    (void)ctx;
    FrameSet fr;
    memset(&fr, 0, sizeof(fr));
    if (int rc = jinc_pack_plane_ptrs(&fr.one, sample_bytes, n_planes, d_src, src_pitch, d_dst, dst_pitch))
        return rc;
    fr.frames = nullptr;
    fr.n_planes = n_planes;
    fr.peak = peak;
    return dispatch(t, sample_bytes, fr, 1, y_begin, y_end, stream, launches, parts);
}

int jinc_launch_resize_batch(jinc_ctx* ctx, const jinc_table* t, int sample_bytes, float peak, int n_planes,
                             const void* d_frame_ptrs, int n_frames, cudaStream_t stream, int* launches, int parts)
{
    (void)ctx;
    FrameSet fr;
    memset(&fr, 0, sizeof(fr));
    fr.frames = static_cast<const PlanePtrs*>(d_frame_ptrs);
    fr.n_planes = n_planes;
    fr.peak = peak;
    return dispatch(t, sample_bytes, fr, n_frames, 0, t->sc.dst_h, stream, launches, parts);
}

int jinc_launch_resize(jinc_ctx* ctx, const jinc_table* t, int sample_bytes, float peak, const void* d_src,
                       ptrdiff_t src_pitch, void* d_dst, ptrdiff_t dst_pitch, int y_begin, int y_end, cudaStream_t stream,
                       int* launches)
{
    return jinc_launch_resize_planes(ctx, t, sample_bytes, peak, 1, &d_src, &src_pitch, &d_dst, &dst_pitch, y_begin, y_end,
                                     stream, launches, JINC_PART_ALL);
}

int jinc_debug_pixel_weights(const jinc_table* t, int x, int y, float* out)
{
    JINC_CUDA(cudaSetDevice(t->ctx->device));
    StripArgs a;
    fill_strip_args(t, a);
    const size_t n = (size_t)t->sc.fs * t->sc.fs;
    float* d = nullptr;
    JINC_CUDA(cudaMalloc(&d, n * sizeof(float)));
    pixel_weights_kernel<<<1, 128, 0, t->ctx->stream>>>(a, x, y, d);
    cudaError_t e = cudaMemcpyAsync(out, d, n * sizeof(float), cudaMemcpyDeviceToHost, t->ctx->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(t->ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "pixel_weights: %s", cudaGetErrorString(e));
    return JINC_OK;
}

extern "C" int jinc_resize_plane_device(jinc_ctx* ctx, const jinc_table* t, int sample_bytes, float peak, const void* d_src,
                                        ptrdiff_t src_pitch, void* d_dst, ptrdiff_t dst_pitch, void* stream)
{
    if (!ctx || !t)
        return jinc_fail(JINC_E_INVALID, "jinc_resize_plane_device: null argument");
    JINC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return jinc_launch_resize(ctx, t, sample_bytes, peak, d_src, src_pitch, d_dst, dst_pitch, 0, t->sc.dst_h, st, nullptr);
}

extern "C" int jinc_table_launches_per_plane(const jinc_table* t)
{
    return t ? 1 : 0; // interior tiles and border strips of all planes sharing the table go out in one launch
}
