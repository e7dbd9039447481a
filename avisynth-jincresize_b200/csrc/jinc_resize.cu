// jinc_resize.cu -- EWA resampling kernels for sm_100a and their launcher.
//
// Replaces JincResize::resize_plane_c<T,thr,subsampled> (src/JincResize.cpp:536-601) and the three SIMD copies of
// its inner loop.  Every output sample is  sum_{ly,lx} src[start_y+ly][start_x+lx] * w[ly][lx]  over an fs x fs
// window, followed for integer formats by clamp to [0,peak] and round-half-even (:581-582); float is raw (:583-584).
//
// Kernels
//   resample_up2x     exact 2x upscale (all "JincNNResize(2w,2h)" uses).  The table has 2x2 phase classes; a thread
//                     owns TX=4 source-aligned cells x 2 cell rows = 8x4 output samples and keeps them in 16 float2
//                     accumulators.  Source rows live in shared memory as VERTICAL PAIRS {S[r][c], S[r+1][c]} so that
//                     one packed FFMA2 (fma.rn.f32x2, new on sm_100) updates the same phase of two cell rows with a
//                     single scalar weight.  Weights arrive as kernel parameters (constant bank) and are fed to the
//                     FMA pipe through uniform registers (LDCU.128 -> FFMA2 R, R, UR, R): no shared-memory or
//                     register-file traffic for weights at all.
//   resample_general  one thread per output sample; interior samples gather their phase block from the L2-resident
//                     table, border samples build their window weights on the fly from the LUT exactly as the
//                     reference does per border pixel (:443-514).  Runs the border strips around a fast-path interior
//                     and whole planes whose geometry has no fast path.
#include <algorithm>
#include <cstring>

#include "jinc_internal.h"
#include "jinc_weights.cuh"

namespace {

// ------------------------------------------------------------------------------------------ store helpers

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
    v = v > peak ? peak : v; // upper bound first (avs/minmax.h clamp)
    v = v < 0.f ? 0.f : v;
    return (uint8_t)__float2int_rn(v); // lrintf: round half to even
}
template <>
__device__ __forceinline__ uint16_t finish<uint16_t>(float v, float peak)
{
    v = v > peak ? peak : v;
    v = v < 0.f ? 0.f : v;
    return (uint16_t)__float2int_rn(v);
}

template <typename T>
__device__ __forceinline__ float load_sample(const T* p)
{
    return (float)__ldg(p);
}

// ------------------------------------------------------------------------------------------ general kernel

struct Rect {
    int x0, y0, x1, y1;
};

struct PlanePtrs {
    const void* src[JINC_MAX_PLANES];
    void* dst[JINC_MAX_PLANES];
    long long src_pitch[JINC_MAX_PLANES]; // in elements
    long long dst_pitch[JINC_MAX_PLANES];
};

struct GeneralArgs {
    PlanePtrs pl;
    const int32_t* start_x;
    const int32_t* start_y;
    const int32_t* rank_x;
    const int32_t* rank_y;
    const float* pos_x;
    const float* pos_y;
    const float* weights;
    const float* lut;
    int fs, n_rank_x, src_w, src_h;
    double step_x, step_y, radius2, idx_scale;
    float peak;
    Rect rect[4];
    int block_begin[5]; // prefix sum of 32x8 blocks per rect
    int blocks_x[4];
};

constexpr int GB_X = 32, GB_Y = 8;

template <typename T>
__global__ void __launch_bounds__(GB_X* GB_Y) resample_general(const __grid_constant__ GeneralArgs a)
{
    const int b = blockIdx.x;
    int r = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
        r += b >= a.block_begin[k];
    const int lb = b - a.block_begin[r];
    const int by = lb / a.blocks_x[r], bx = lb - by * a.blocks_x[r];
    const int x = a.rect[r].x0 + bx * GB_X + threadIdx.x;
    const int y = a.rect[r].y0 + by * GB_Y + threadIdx.y;
    if (x >= a.rect[r].x1 || y >= a.rect[r].y1)
        return;

    const int plane = blockIdx.y;
    const T* __restrict__ src = static_cast<const T*>(a.pl.src[plane]);
    T* __restrict__ dst = static_cast<T*>(a.pl.dst[plane]);
    const long long sp = a.pl.src_pitch[plane];
    const int fs = a.fs;
    const int sx = a.start_x[x], sy = a.start_y[y];
    const int rx = a.rank_x[x], ry = a.rank_y[y];
    const T* s = src + (long long)sy * sp + sx;
    float acc = 0.f;

    if (rx >= 0 && ry >= 0) {
        // interior: shared phase block (:431-435)
        const float* __restrict__ w = a.weights + (size_t)(ry * a.n_rank_x + rx) * fs * fs;
        for (int ly = 0; ly < fs; ++ly) {
            for (int lx = 0; lx < fs; ++lx)
                acc = fmaf(load_sample(s + lx), __ldg(w + lx), acc);
            w += fs;
            s += sp;
        }
    } else {
        // border: per-pixel weights from the UNquantised position and the clamped window (:443-514)
        const float px = a.pos_x[x], py = a.pos_y[y];
        float sum = 0.f;
        for (int ly = 0; ly < fs; ++ly) {
            const double dy2 = jinc_tap_dist2(py, a.src_h, sy + ly, a.step_y);
            for (int lx = 0; lx < fs; ++lx) {
                const double dx2 = jinc_tap_dist2(px, a.src_w, sx + lx, a.step_x);
                sum = __fadd_rn(sum, jinc_lut_weight(a.lut, __dadd_rn(dx2, dy2), a.radius2, a.idx_scale));
            }
        }
        for (int ly = 0; ly < fs; ++ly) {
            const double dy2 = jinc_tap_dist2(py, a.src_h, sy + ly, a.step_y);
            for (int lx = 0; lx < fs; ++lx) {
                const double dx2 = jinc_tap_dist2(px, a.src_w, sx + lx, a.step_x);
                const float f = jinc_lut_weight(a.lut, __dadd_rn(dx2, dy2), a.radius2, a.idx_scale);
                acc = fmaf(load_sample(s + lx), __fdiv_rn(f, sum), acc);
            }
            s += sp;
        }
    }
    dst[(long long)y * a.pl.dst_pitch[plane] + x] = finish<T>(acc, a.peak);
}

// weights of one output pixel exactly as resample_general applies them (introspection for parity tests)
__global__ void pixel_weights_kernel(GeneralArgs a, int x, int y, float* out)
{
    const int fs = a.fs;
    const int rx = a.rank_x[x], ry = a.rank_y[y];
    if (rx >= 0 && ry >= 0) {
        const float* w = a.weights + (size_t)(ry * a.n_rank_x + rx) * fs * fs;
        for (int t = threadIdx.x; t < fs * fs; t += blockDim.x)
            out[t] = w[t];
        return;
    }
    if (threadIdx.x != 0)
        return;
    const int sx = a.start_x[x], sy = a.start_y[y];
    const float px = a.pos_x[x], py = a.pos_y[y];
    float sum = 0.f;
    for (int ly = 0; ly < fs; ++ly) {
        const double dy2 = jinc_tap_dist2(py, a.src_h, sy + ly, a.step_y);
        for (int lx = 0; lx < fs; ++lx) {
            const double dx2 = jinc_tap_dist2(px, a.src_w, sx + lx, a.step_x);
            const float f = jinc_lut_weight(a.lut, __dadd_rn(dx2, dy2), a.radius2, a.idx_scale);
            out[ly * fs + lx] = f;
            sum = __fadd_rn(sum, f);
        }
    }
    for (int t = 0; t < fs * fs; ++t)
        out[t] = __fdiv_rn(out[t], sum);
}

void fill_general_args(const jinc_table* t, GeneralArgs& a, float peak)
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
    a.fs = t->sc.fs;
    a.n_rank_x = t->ax[0].n_rank;
    a.src_w = t->sc.src_w;
    a.src_h = t->sc.src_h;
    a.step_x = t->sc.filt_step[0];
    a.step_y = t->sc.filt_step[1];
    a.radius2 = t->sc.radius2;
    a.idx_scale = t->sc.idx_scale;
    a.peak = peak;
}

// ------------------------------------------------------------------------------------------ exact-2x kernel

constexpr int UP_TX = 4;                  // cells per thread along x
constexpr int UP_WARPS = 8;
constexpr int UP_THREADS = UP_WARPS * 32;
constexpr int UP_CW = 32 * UP_TX;         // cells per tile row (128 -> 256 output samples)
constexpr int UP_RPW = 2;                 // cell-row pairs per warp
constexpr int UP_CH = 2 * UP_WARPS * UP_RPW; // cell rows per tile (32 -> 64 output rows)

template <int FS>
struct UpGeom {
    static constexpr int FSP = (FS + 3) & ~3;         // weight row stride (16-byte rows for LDCU.128)
    static constexpr int NSEG = UP_TX + 1 + FS - 1;   // pair columns a thread reads per row (ox1 <= 1)
    static constexpr int NC = UP_CW + FS;             // pair columns per tile row (CW + ox1 + FS - 1)
    static constexpr int NCP = (NC + 3) & ~3;
    static constexpr int SUB = NCP / 4;               // columns are de-interleaved by (c & 3): 4 sub-rows of SUB
    static constexpr int NR = UP_CH + 1 + FS - 1;     // pair rows per tile (CH + oy1 + FS - 1)
    static constexpr size_t SMEM = (size_t)NR * NCP * sizeof(float2);
};

template <int FS>
struct alignas(16) UpWeights {
    float w[2][2][FS][UpGeom<FS>::FSP]; // [py][px][ly][lx]
};

struct UpArgs {
    PlanePtrs pl;
    int src_w, src_h;
    int x0, y0, ncx, ncy; // output origin of the periodic interior, cells per axis
    int sx0, sy0;         // window origin of cell (0,0), phase (0,0)
    int oy1;              // window-origin offset of phase row 1 (0 or 1)
    int cy_begin, cy_end; // cell rows to produce (row-band split)
    float peak;
};

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
        q[k] = (uint32_t)finish<uint16_t>(v[2 * k], peak) | ((uint32_t)finish<uint16_t>(v[2 * k + 1], peak) << 16);
    *reinterpret_cast<uint4*>(p) = make_uint4(q[0], q[1], q[2], q[3]);
}
template <>
__device__ __forceinline__ void store8<uint8_t>(uint8_t* p, const float (&v)[8], float peak)
{
    uint32_t q[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
        q[k] = (uint32_t)finish<uint8_t>(v[4 * k], peak) | ((uint32_t)finish<uint8_t>(v[4 * k + 1], peak) << 8) |
               ((uint32_t)finish<uint8_t>(v[4 * k + 2], peak) << 16) | ((uint32_t)finish<uint8_t>(v[4 * k + 3], peak) << 24);
    *reinterpret_cast<uint2*>(p) = make_uint2(q[0], q[1]);
}

template <typename T, int FS, int OX1>
__global__ void __launch_bounds__(UP_THREADS, 2)
    resample_up2x(const __grid_constant__ UpArgs a, const __grid_constant__ UpWeights<FS> W)
{
    using G = UpGeom<FS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tile = reinterpret_cast<float2*>(smem_raw); // [NR][4][SUB] pairs {S[r][c], S[r+1][c]}

    const int plane = blockIdx.z;
    const T* __restrict__ src = static_cast<const T*>(a.pl.src[plane]);
    T* __restrict__ dst = static_cast<T*>(a.pl.dst[plane]);
    const long long sp = a.pl.src_pitch[plane], dp = a.pl.dst_pitch[plane];

    const int cell_x0 = blockIdx.x * UP_CW;               // first cell of this tile
    const int cell_y0 = a.cy_begin + blockIdx.y * UP_CH;
    const int tsx = a.sx0 + cell_x0, tsy = a.sy0 + cell_y0; // source coordinates of tile(0,0)

    // ---- stage the source tile: each item = one pair row x 4 consecutive columns
    for (int it = threadIdx.x; it < G::NR * G::SUB; it += UP_THREADS) {
        const int r = it / G::SUB, q = it - r * G::SUB;
        const int y0 = min(max(tsy + r, 0), a.src_h - 1), y1 = min(max(tsy + r + 1, 0), a.src_h - 1);
        const T* row0 = src + (long long)y0 * sp;
        const T* row1 = src + (long long)y1 * sp;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = min(max(tsx + 4 * q + k, 0), a.src_w - 1); // out-of-plane taps only feed discarded cells
            tile[(r * 4 + k) * G::SUB + q] = make_float2(load_sample(row0 + x), load_sample(row1 + x));
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int oy1 = a.oy1;

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
        for (int rr = 0; rr < FS + oy1; ++rr) {
            float2 seg[G::NSEG];
#pragma unroll
            for (int m = 0; m < G::NSEG; ++m)
                seg[m] = trow[(m & 3) * G::SUB + (m >> 2)]; // column 4*lane + m
            trow += G::NCP;

            if (rr < FS) { // phase row 0: ly = rr
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
            const int ly1 = rr - oy1; // phase row 1
            if (ly1 >= 0) {
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
            if (cy + h >= a.cy_end)
                break;
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
                    store8<T>(o, v, a.peak);
                } else {
                    for (int k = 0; k < 2 * (a.ncx - cx); ++k)
                        o[k] = finish<T>(v[k], a.peak);
                }
            }
        }
    }
}

template <typename T, int FS>
int launch_up2x_fs(const jinc_table* t, const UpArgs& a, int n_planes, cudaStream_t st)
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
    const int rows = a.cy_end - a.cy_begin;
    dim3 grid((u.ncx + UP_CW - 1) / UP_CW, (rows + UP_CH - 1) / UP_CH, n_planes);
    auto kern = u.ox1 ? resample_up2x<T, FS, 1> : resample_up2x<T, FS, 0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "cudaFuncSetAttribute(up2x smem %zu): %s", G::SMEM, cudaGetErrorString(e));
    kern<<<grid, UP_THREADS, G::SMEM, st>>>(a, w);
    e = cudaGetLastError();
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "resample_up2x launch failed: %s", cudaGetErrorString(e));
    return JINC_OK;
}

template <typename T>
int launch_up2x(const jinc_table* t, const UpArgs& a, int n_planes, cudaStream_t st)
{
    switch (t->sc.fs) {
    case 7: return launch_up2x_fs<T, 7>(t, a, n_planes, st);   // tap 3  (Jinc36Resize)
    case 9: return launch_up2x_fs<T, 9>(t, a, n_planes, st);   // tap 4  (Jinc64Resize)
    case 13: return launch_up2x_fs<T, 13>(t, a, n_planes, st); // tap 6  (Jinc144Resize)
    case 17: return launch_up2x_fs<T, 17>(t, a, n_planes, st); // tap 8  (Jinc256Resize)
    default: return 1; // no specialisation: caller falls back to the general kernel
    }
}

bool up2x_supported(int fs) { return fs == 7 || fs == 9 || fs == 13 || fs == 17; }

// ------------------------------------------------------------------------------------------ launcher

template <typename T>
int launch_general(const jinc_table* t, GeneralArgs& a, const Rect* rects, int n_rects, int n_planes, cudaStream_t st,
                   int* launches)
{
    int total = 0, k = 0;
    for (int r = 0; r < n_rects; ++r) {
        const int w = rects[r].x1 - rects[r].x0, h = rects[r].y1 - rects[r].y0;
        if (w <= 0 || h <= 0)
            continue;
        a.rect[k] = rects[r];
        a.blocks_x[k] = (w + GB_X - 1) / GB_X;
        a.block_begin[k] = total;
        total += a.blocks_x[k] * ((h + GB_Y - 1) / GB_Y);
        ++k;
    }
    for (int j = k; j < 5; ++j)
        a.block_begin[j] = total; // unused rects never match
    if (total == 0)
        return JINC_OK;
    for (int j = k; j < 4; ++j) {
        a.rect[j] = Rect{0, 0, 0, 0};
        a.blocks_x[j] = 1;
    }
    // block_begin[j] for j>=k equals `total`, so the rect search in the kernel stops at the last real rect
    resample_general<T><<<dim3(total, n_planes), dim3(GB_X, GB_Y), 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return jinc_fail(JINC_E_CUDA, "resample_general launch failed: %s", cudaGetErrorString(e));
    ++*launches;
    return JINC_OK;
}

template <typename T>
int launch_typed(jinc_ctx* ctx, const jinc_table* t, float peak, const PlanePtrs& pl, int n_planes, int y_begin, int y_end,
                 cudaStream_t st, int* launches, int parts)
{
    (void)ctx;
    GeneralArgs ga;
    fill_general_args(t, ga, peak);
    ga.pl = pl;
    const int W = t->sc.dst_w;
    Rect rects[4];
    int n_rects = 0;

    bool fast_done = false;
    int fy0 = 0, fy1 = 0; // output rows covered by the fast path
    if (t->fast_path == JINC_PATH_UP2X && up2x_supported(t->sc.fs)) {
        const Up2xPlan& u = t->up2x;
        // cell rows whose 2 output rows lie inside [y_begin, y_end); bands are cut on cell-pair boundaries
        int cb = (std::max(y_begin, u.y0) - u.y0 + 1) / 2;
        int ce = (std::min(y_end, u.y0 + 2 * u.ncy) - u.y0) / 2;
        if (ce > cb && !(parts & JINC_PART_INTERIOR)) {
            // interior deliberately skipped: still only the strips around it belong to the border part
            fast_done = true;
            fy0 = u.y0 + 2 * cb;
            fy1 = u.y0 + 2 * ce;
        } else if (ce > cb) {
            UpArgs a;
            memset(&a, 0, sizeof(a));
            a.pl = pl;
            a.src_w = t->sc.src_w;
            a.src_h = t->sc.src_h;
            a.x0 = u.x0;
            a.y0 = u.y0;
            a.ncx = u.ncx;
            a.ncy = u.ncy;
            a.sx0 = u.sx0;
            a.sy0 = u.sy0;
            a.oy1 = u.oy1;
            a.cy_begin = cb;
            a.cy_end = ce;
            a.peak = peak;
            const int rc = launch_up2x<T>(t, a, n_planes, st);
            if (rc < 0)
                return rc;
            if (rc == 0) {
                ++*launches;
                fast_done = true;
                fy0 = u.y0 + 2 * cb;
                fy1 = u.y0 + 2 * ce;
            }
        }
    }
    if (fast_done) {
        rects[n_rects++] = Rect{0, y_begin, W, fy0};            // top strip
        rects[n_rects++] = Rect{0, fy1, W, y_end};              // bottom strip
        rects[n_rects++] = Rect{0, fy0, t->ix0, fy1};           // left strip
        rects[n_rects++] = Rect{t->ix1, fy0, W, fy1};           // right strip
    } else {
        rects[n_rects++] = Rect{0, y_begin, W, y_end};
    }
    if (!(parts & JINC_PART_BORDER))
        return JINC_OK;
    return launch_general<T>(t, ga, rects, n_rects, n_planes, st, launches);
}

} // namespace

int jinc_launch_resize_planes(jinc_ctx* ctx, const jinc_table* t, int sample_bytes, float peak, int n_planes,
                              const void* const* d_src, const ptrdiff_t* src_pitch, void* const* d_dst,
                              const ptrdiff_t* dst_pitch, int y_begin, int y_end, cudaStream_t stream, int* launches,
                              int parts)
{
    if (n_planes < 1 || n_planes > JINC_MAX_PLANES)
        return jinc_fail(JINC_E_INVALID, "resize: n_planes must be 1..4");
    PlanePtrs pl;
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
    y_begin = std::max(y_begin, 0);
    y_end = std::min(y_end, t->sc.dst_h);
    if (y_end <= y_begin)
        return JINC_OK;
    int dummy = 0;
    if (!launches)
        launches = &dummy;
    switch (sample_bytes) {
    case 1: return launch_typed<uint8_t>(ctx, t, peak, pl, n_planes, y_begin, y_end, stream, launches, parts);
    case 2: return launch_typed<uint16_t>(ctx, t, peak, pl, n_planes, y_begin, y_end, stream, launches, parts);
    case 4: return launch_typed<float>(ctx, t, peak, pl, n_planes, y_begin, y_end, stream, launches, parts);
    default: return jinc_fail(JINC_E_INVALID, "resize: sample_bytes must be 1, 2 or 4");
    }
}

int jinc_launch_resize(jinc_ctx* ctx, const jinc_table* t, int sample_bytes, float peak, const void* d_src,
                       ptrdiff_t src_pitch, void* d_dst, ptrdiff_t dst_pitch, int y_begin, int y_end, cudaStream_t stream,
                       int* launches)
{
    return jinc_launch_resize_planes(ctx, t, sample_bytes, peak, 1, &d_src, &src_pitch, &d_dst, &dst_pitch, y_begin, y_end,
                                     stream, launches);
}

int jinc_debug_pixel_weights(const jinc_table* t, int x, int y, float* out)
{
    JINC_CUDA(cudaSetDevice(t->ctx->device));
    GeneralArgs a;
    fill_general_args(t, a, 0.f);
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
    if (!t)
        return 0;
    return (t->fast_path == JINC_PATH_UP2X && up2x_supported(t->sc.fs)) ? 2 : 1;
}
