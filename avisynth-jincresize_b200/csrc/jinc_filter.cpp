// jinc_filter.cpp -- device contexts, filter instances and the host<->device frame pipeline.
//
// Replaces the per-plane geometry derivation of Create_JincResize (src/JincResize.cpp:762-866), the
// process_frame step of JincResize_GetFrame (:615) and free_JincResize (:632-647).
//
// Frame pipeline: every GPU of the filter owns `slots_per_device` in-flight slots, each with its own CUDA
// stream, device source/destination buffers and pinned staging buffers laid out identically to the device
// buffers (so one cudaMemcpyAsync moves all planes).  A frame takes a slot (GPUs round-robin), is staged,
// copied H2D, resampled and copied D2H on that slot's stream; concurrent callers (AviSynth Prefetch threads,
// or jinc_filter_submit) therefore overlap staging, PCIe transfers in both directions and kernels across
// slots and GPUs.  Frames are independent, so there is no inter-GPU traffic and no collective.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>

#include "jinc_internal.h"

// ================================================================ contexts

extern "C" int jinc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int jinc_ctx_create(int device, jinc_ctx** out)
{
    if (!out)
        return jinc_fail(JINC_E_INVALID, "jinc_ctx_create: null output");
    *out = nullptr;
    const int n = jinc_device_count();
    if (n == 0)
        return jinc_fail(JINC_E_CUDA, "JincResize: no CUDA device available (this build has no CPU fallback)");
    if (device < 0 || device >= n)
        return jinc_fail(JINC_E_INVALID, "jinc_ctx_create: device %d out of range (0..%d)", device, n - 1);
    JINC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    JINC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return jinc_fail(JINC_E_UNSUPPORTED, "JincResize: device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
    auto* c = new jinc_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return jinc_fail(JINC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    *out = c;
    return JINC_OK;
}

extern "C" void jinc_ctx_destroy(jinc_ctx* ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int jinc_ctx_device(const jinc_ctx* ctx) { return ctx ? ctx->device : -1; }

// ================================================================ filter

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct PlaneLayout {
    int table = 0;          // which table (0 luma / shared, 1 subsampled chroma)
    int src_w = 0, src_h = 0, dst_w = 0, dst_h = 0;
    size_t src_off = 0, dst_off = 0; // byte offsets inside the slot buffers
    size_t src_pitch = 0, dst_pitch = 0;
};

struct Slot {
    int dev_index = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    unsigned char* d_src = nullptr;
    unsigned char* d_dst = nullptr;
    unsigned char* h_src = nullptr; // pinned
    unsigned char* h_dst = nullptr; // pinned
    bool busy = false;
    int64_t ticket = -1;
    jinc_frame pending{}; // destination of an outstanding submit
    bool dst_direct = false;
};

struct DeviceState {
    jinc_ctx* ctx = nullptr;
    jinc_table* tables[2] = {nullptr, nullptr};
    // ring of small device buffers holding the plane-pointer records of batched launches
    static constexpr int kRing = 8;
    unsigned char* batch_ptrs[kRing] = {};
    size_t batch_cap[kRing] = {};
    int batch_next = 0;
};

// rows of one plane between host and device: ONE linear transfer when both sides are tightly packed (the DMA engine
// then moves a single extent instead of a descriptor per row), else a 2-D copy
cudaError_t copy_plane_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, int rows,
                             cudaMemcpyKind kind, cudaStream_t st)
{
    if (dst_pitch == row_bytes && src_pitch == row_bytes)
        return cudaMemcpyAsync(dst, src, row_bytes * static_cast<size_t>(rows), kind, st);
    return cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, static_cast<size_t>(rows), kind, st);
}

bool is_pinned_host(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

} // namespace

struct jinc_filter {
    jinc_filter_params p{};
    float peak = 0.f;
    int n_tables = 1;
    jinc_table_params tparams[2]{};
    PlaneLayout planes[JINC_MAX_PLANES];
    size_t src_bytes = 0, dst_bytes = 0;
    std::vector<DeviceState> devs;
    std::vector<std::unique_ptr<Slot>> slots;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int64_t> next_ticket{0};
    std::atomic<int64_t> rr{0};
    std::atomic<int64_t> launches{0};
};

namespace {

// Create_JincResize geometry (:762-770 crop, :833-862 chroma shift).  cplace only matters for subsampled chroma.
void derive_table_params(const jinc_filter_params& p, jinc_table_params out[2], int* n_tables)
{
    double crop_w = p.src_width, crop_h = p.src_height;
    if (crop_w <= 0.0)
        crop_w = p.src_w - p.src_left + crop_w;
    if (crop_h <= 0.0)
        crop_h = p.src_h - p.src_top + crop_h;
    const double radius = jinc_radius_for_tap(p.tap);

    jinc_table_params& y = out[0];
    y.quant_x = p.quant_x;
    y.quant_y = p.quant_y;
    y.src_w = p.src_w;
    y.src_h = p.src_h;
    y.dst_w = p.target_w;
    y.dst_h = p.target_h;
    y.radius = radius;
    y.blur = p.blur;
    y.crop_left = p.src_left;
    y.crop_top = p.src_top;
    y.crop_w = crop_w;
    y.crop_h = crop_h;
    *n_tables = 1;
    if (p.n_planes < 2 || (p.sub_w == 0 && p.sub_h == 0))
        return;

    // The half-sample shift of co-sited chroma uses the FULL source/target widths, not the crop window (:838-841).
    const double div_w = static_cast<double>(1 << p.sub_w), div_h = static_cast<double>(1 << p.sub_h);
    const bool left_sited = p.cplace == JINC_CPLACE_MPEG2 || p.cplace == JINC_CPLACE_TOPLEFT;
    const bool top_sited = p.cplace == JINC_CPLACE_TOPLEFT;
    jinc_table_params& c = out[1];
    c = y;
    c.src_w = p.src_w >> p.sub_w;
    c.src_h = p.src_h >> p.sub_h;
    c.dst_w = p.target_w >> p.sub_w;
    c.dst_h = p.target_h >> p.sub_h;
    c.crop_left = left_sited ? (0.5 * (1.0 - static_cast<double>(p.src_w) / p.target_w) + p.src_left) / div_w
                             : p.src_left / div_w;
    c.crop_top = top_sited ? (0.5 * (1.0 - static_cast<double>(p.src_h) / p.target_h) + p.src_top) / div_h
                           : p.src_top / div_h;
    c.crop_w = crop_w / div_w;
    c.crop_h = crop_h / div_h;
    *n_tables = 2;
}

void release_slot(jinc_filter* f, Slot* s)
{
    {
        std::lock_guard<std::mutex> lk(f->mu);
        s->busy = false;
        s->ticket = -1;
    }
    f->cv.notify_all();
}

// Take a free slot, preferring GPU (n mod G) so consecutive frames spread over all GPUs.
Slot* acquire_slot(jinc_filter* f)
{
    const int nd = static_cast<int>(f->devs.size());
    const int want = static_cast<int>(f->rr.fetch_add(1) % nd);
    std::unique_lock<std::mutex> lk(f->mu);
    for (;;) {
        Slot* any = nullptr;
        for (auto& s : f->slots) {
            if (s->busy)
                continue;
            if (s->dev_index == want) {
                s->busy = true;
                return s.get();
            }
            if (!any)
                any = s.get();
        }
        if (any) {
            any->busy = true;
            return any;
        }
        f->cv.wait(lk);
    }
}

void copy_rows(unsigned char* dst, size_t dst_pitch, const unsigned char* src, ptrdiff_t src_pitch, size_t row_bytes, int rows)
{
    if (static_cast<ptrdiff_t>(dst_pitch) == src_pitch && dst_pitch == row_bytes) {
        memcpy(dst, src, row_bytes * rows);
        return;
    }
    for (int y = 0; y < rows; ++y)
        memcpy(dst + static_cast<size_t>(y) * dst_pitch, src + static_cast<ptrdiff_t>(y) * src_pitch, row_bytes);
}

// Enqueue H2D + kernels + D2H for output rows [y0,y1) (luma rows) of `frame` on slot s.
int enqueue_frame(jinc_filter* f, Slot* s, const jinc_frame* frame, int y0_luma, int y1_luma, bool whole)
{
    DeviceState& d = f->devs[s->dev_index];
    JINC_CUDA(cudaSetDevice(d.ctx->device));
    const int sb = f->p.sample_bytes;

    // ---- source planes -> device
    bool all_pinned = true;
    for (int i = 0; i < f->p.n_planes; ++i)
        all_pinned = all_pinned && is_pinned_host(frame->src[i]);
    // per plane: the source rows this band's windows reach
    int sy0[JINC_MAX_PLANES], sy1[JINC_MAX_PLANES], oy0[JINC_MAX_PLANES], oy1[JINC_MAX_PLANES];
    for (int i = 0; i < f->p.n_planes; ++i) {
        const PlaneLayout& pl = f->planes[i];
        const jinc_table* t = d.tables[pl.table];
        const int shift = (pl.table == 1) ? f->p.sub_h : 0;
        oy0[i] = whole ? 0 : (y0_luma >> shift);
        oy1[i] = whole ? pl.dst_h : std::min(pl.dst_h, (y1_luma + (1 << shift) - 1) >> shift);
        if (whole) {
            sy0[i] = 0;
            sy1[i] = pl.src_h;
        } else {
            int lo = pl.src_h, hi = 0;
            for (int y = oy0[i]; y < oy1[i]; ++y) {
                lo = std::min(lo, t->h_start[1][y]);
                hi = std::max(hi, t->h_start[1][y] + t->sc.fs);
            }
            sy0[i] = std::max(lo, 0);
            sy1[i] = std::min(hi, pl.src_h);
        }
    }
    if (whole && !all_pinned) {
        // pageable caller memory: stage into the slot's pinned mirror, then ONE copy for all planes
        for (int i = 0; i < f->p.n_planes; ++i) {
            const PlaneLayout& pl = f->planes[i];
            copy_rows(s->h_src + pl.src_off, pl.src_pitch, static_cast<const unsigned char*>(frame->src[i]),
                      frame->src_pitch[i], static_cast<size_t>(pl.src_w) * sb, pl.src_h);
        }
        JINC_CUDA(cudaMemcpyAsync(s->d_src, s->h_src, f->src_bytes, cudaMemcpyHostToDevice, s->stream));
    } else {
        for (int i = 0; i < f->p.n_planes; ++i) {
            const PlaneLayout& pl = f->planes[i];
            const int rows = sy1[i] - sy0[i];
            if (rows <= 0)
                continue;
            const unsigned char* hsrc = static_cast<const unsigned char*>(frame->src[i]) + static_cast<ptrdiff_t>(sy0[i]) * frame->src_pitch[i];
            size_t hpitch = static_cast<size_t>(frame->src_pitch[i]);
            if (!all_pinned) {
                unsigned char* stage = s->h_src + pl.src_off + static_cast<size_t>(sy0[i]) * pl.src_pitch;
                copy_rows(stage, pl.src_pitch, hsrc, frame->src_pitch[i], static_cast<size_t>(pl.src_w) * sb, rows);
                hsrc = stage;
                hpitch = pl.src_pitch;
            }
            JINC_CUDA(copy_plane_async(s->d_src + pl.src_off + static_cast<size_t>(sy0[i]) * pl.src_pitch, pl.src_pitch, hsrc, hpitch,
                                       static_cast<size_t>(pl.src_w) * sb, rows, cudaMemcpyHostToDevice, s->stream));
        }
    }

    // ---- kernels: planes that share a table go out in one launch
    for (int k = 0; k < f->n_tables; ++k) {
        const void* src[JINC_MAX_PLANES];
        void* dst[JINC_MAX_PLANES];
        ptrdiff_t sp[JINC_MAX_PLANES], dp[JINC_MAX_PLANES];
        int n = 0, yb = 0, ye = 0;
        for (int i = 0; i < f->p.n_planes; ++i) {
            const PlaneLayout& pl = f->planes[i];
            if (pl.table != k)
                continue;
            src[n] = s->d_src + pl.src_off;
            dst[n] = s->d_dst + pl.dst_off;
            sp[n] = static_cast<ptrdiff_t>(pl.src_pitch);
            dp[n] = static_cast<ptrdiff_t>(pl.dst_pitch);
            yb = oy0[i];
            ye = oy1[i];
            ++n;
        }
        if (n == 0)
            continue;
        int launched = 0;
        const int rc = jinc_launch_resize_planes(d.ctx, d.tables[k], sb, f->peak, n, src, sp, dst, dp, yb, ye, s->stream, &launched);
        f->launches.fetch_add(launched);
        if (rc != JINC_OK)
            return rc;
    }

    // ---- destination planes -> host
    bool dst_pinned = true;
    for (int i = 0; i < f->p.n_planes; ++i)
        dst_pinned = dst_pinned && is_pinned_host(frame->dst[i]);
    s->dst_direct = dst_pinned;
    if (whole && !dst_pinned) {
        JINC_CUDA(cudaMemcpyAsync(s->h_dst, s->d_dst, f->dst_bytes, cudaMemcpyDeviceToHost, s->stream));
    } else {
        for (int i = 0; i < f->p.n_planes; ++i) {
            const PlaneLayout& pl = f->planes[i];
            const int rows = oy1[i] - oy0[i];
            if (rows <= 0)
                continue;
            unsigned char* hdst;
            size_t hpitch;
            if (dst_pinned) {
                hdst = static_cast<unsigned char*>(frame->dst[i]) + static_cast<ptrdiff_t>(oy0[i]) * frame->dst_pitch[i];
                hpitch = static_cast<size_t>(frame->dst_pitch[i]);
            } else {
                hdst = s->h_dst + pl.dst_off + static_cast<size_t>(oy0[i]) * pl.dst_pitch;
                hpitch = pl.dst_pitch;
            }
            JINC_CUDA(copy_plane_async(hdst, hpitch, s->d_dst + pl.dst_off + static_cast<size_t>(oy0[i]) * pl.dst_pitch, pl.dst_pitch,
                                       static_cast<size_t>(pl.dst_w) * sb, rows, cudaMemcpyDeviceToHost, s->stream));
        }
    }
    JINC_CUDA(cudaEventRecord(s->done, s->stream));
    s->pending = *frame;
    return JINC_OK;
}

// Wait for the slot's work and move staged output rows [y0,y1) (luma) to the caller's planes.
int finish_frame(jinc_filter* f, Slot* s, int y0_luma, int y1_luma, bool whole)
{
    DeviceState& d = f->devs[s->dev_index];
    JINC_CUDA(cudaSetDevice(d.ctx->device));
    JINC_CUDA(cudaEventSynchronize(s->done));
    if (s->dst_direct)
        return JINC_OK;
    const int sb = f->p.sample_bytes;
    for (int i = 0; i < f->p.n_planes; ++i) {
        const PlaneLayout& pl = f->planes[i];
        const int shift = (pl.table == 1) ? f->p.sub_h : 0;
        const int a = whole ? 0 : (y0_luma >> shift);
        const int b = whole ? pl.dst_h : std::min(pl.dst_h, (y1_luma + (1 << shift) - 1) >> shift);
        if (b <= a)
            continue;
        copy_rows(static_cast<unsigned char*>(s->pending.dst[i]) + static_cast<ptrdiff_t>(a) * s->pending.dst_pitch[i],
                  static_cast<size_t>(s->pending.dst_pitch[i]), s->h_dst + pl.dst_off + static_cast<size_t>(a) * pl.dst_pitch,
                  static_cast<ptrdiff_t>(pl.dst_pitch), static_cast<size_t>(pl.dst_w) * sb, b - a);
    }
    return JINC_OK;
}

int check_frame(const jinc_filter* f, const jinc_frame* frame)
{
    if (!f || !frame)
        return jinc_fail(JINC_E_INVALID, "jinc_filter: null argument");
    for (int i = 0; i < f->p.n_planes; ++i)
        if (!frame->src[i] || !frame->dst[i])
            return jinc_fail(JINC_E_INVALID, "jinc_filter: plane %d has a null pointer", i);
    return JINC_OK;
}

} // namespace

extern "C" void jinc_filter_destroy(jinc_filter* f)
{
    if (!f)
        return;
    for (auto& s : f->slots) {
        cudaSetDevice(f->devs[s->dev_index].ctx->device);
        if (s->stream)
            cudaStreamSynchronize(s->stream);
        cudaFree(s->d_src);
        cudaFree(s->d_dst);
        cudaFreeHost(s->h_src);
        cudaFreeHost(s->h_dst);
        if (s->done)
            cudaEventDestroy(s->done);
        if (s->stream)
            cudaStreamDestroy(s->stream);
    }
    for (DeviceState& d : f->devs) {
        if (d.ctx)
            cudaSetDevice(d.ctx->device);
        for (unsigned char* b : d.batch_ptrs)
            cudaFree(b);
        jinc_table_destroy(d.tables[0]);
        jinc_table_destroy(d.tables[1]);
        jinc_ctx_destroy(d.ctx);
    }
    delete f;
}

extern "C" int jinc_filter_create(const jinc_filter_params* p, jinc_filter** out)
{
    if (!p || !out)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_create: null argument");
    *out = nullptr;
    if (p->n_planes < 1 || p->n_planes > JINC_MAX_PLANES)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_create: n_planes must be 1..4");
    if (p->sample_bytes != 1 && p->sample_bytes != 2 && p->sample_bytes != 4)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_create: sample_bytes must be 1, 2 or 4");
    if (p->tap < 1 || p->tap > 16)
        return jinc_fail(JINC_E_INVALID, "JincResize: tap must be between 1..16.");
    if (p->src_w < 1 || p->src_h < 1 || p->target_w < 1 || p->target_h < 1)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_create: clip dimensions must be positive");
    if (p->sub_w < 0 || p->sub_w > 2 || p->sub_h < 0 || p->sub_h > 2)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_create: bad chroma subsampling");

    int visible = jinc_device_count();
    if (visible == 0)
        return jinc_fail(JINC_E_CUDA, "JincResize: no CUDA device available (this build has no CPU fallback)");
    std::vector<int> dev_ids;
    if (p->n_devices <= 0) {
        for (int i = 0; i < visible && i < JINC_MAX_DEVICES; ++i)
            dev_ids.push_back(i);
    } else {
        for (int i = 0; i < p->n_devices && i < JINC_MAX_DEVICES; ++i)
            dev_ids.push_back(p->devices[i]);
    }

    std::unique_ptr<jinc_filter, void (*)(jinc_filter*)> f(new jinc_filter(), jinc_filter_destroy);
    f->p = *p;
    f->peak = (p->sample_bytes == 4) ? 0.f : static_cast<float>((1 << p->bits) - 1); // :793
    derive_table_params(*p, f->tparams, &f->n_tables);

    // plane layout shared by device and pinned buffers
    size_t so = 0, dof = 0;
    for (int i = 0; i < p->n_planes; ++i) {
        PlaneLayout& pl = f->planes[i];
        pl.table = (f->n_tables == 2 && (i == 1 || i == 2)) ? 1 : 0; // :552-558: alpha (3) uses the luma table
        const jinc_table_params& tp = f->tparams[pl.table];
        pl.src_w = tp.src_w;
        pl.src_h = tp.src_h;
        pl.dst_w = tp.dst_w;
        pl.dst_h = tp.dst_h;
        // 64-byte pitches: vector loads/stores stay aligned, and for the usual widths the pitch equals the row size, so a
        // plane whose host rows are tightly packed moves as ONE linear DMA transfer instead of a row-by-row 2-D copy
        pl.src_pitch = align_up(static_cast<size_t>(pl.src_w) * p->sample_bytes, 64);
        pl.dst_pitch = align_up(static_cast<size_t>(pl.dst_w) * p->sample_bytes, 64);
        pl.src_off = so;
        pl.dst_off = dof;
        so += align_up(pl.src_pitch * pl.src_h, 256);
        dof += align_up(pl.dst_pitch * pl.dst_h, 256);
    }
    f->src_bytes = so;
    f->dst_bytes = dof;

    // frames in flight per GPU: a slot is held from the staging copy of the source until the output has been copied out,
    // mostly host memcpy time for pageable callers, so small frames get more slots (3..8, about 512 MB per GPU)
    int spd = p->slots_per_device;
    if (spd <= 0) {
        const size_t per_slot = std::max<size_t>(so + dof, 1);
        spd = static_cast<int>(std::min<size_t>(8, std::max<size_t>(3, (static_cast<size_t>(512) << 20) / per_slot)));
    }
    for (size_t di = 0; di < dev_ids.size(); ++di) {
        DeviceState d;
        int rc = jinc_ctx_create(dev_ids[di], &d.ctx);
        if (rc != JINC_OK)
            return rc;
        f->devs.push_back(d);
        for (int k = 0; k < f->n_tables; ++k) {
            rc = jinc_table_create(f->devs.back().ctx, &f->tparams[k], &f->devs.back().tables[k]);
            if (rc != JINC_OK)
                return rc;
        }
        for (int sidx = 0; sidx < spd; ++sidx) {
            auto s = std::make_unique<Slot>();
            s->dev_index = static_cast<int>(di);
            JINC_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
            JINC_CUDA(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming));
            if (cudaMalloc(reinterpret_cast<void**>(&s->d_src), so) != cudaSuccess ||
                cudaMalloc(reinterpret_cast<void**>(&s->d_dst), dof) != cudaSuccess ||
                cudaHostAlloc(reinterpret_cast<void**>(&s->h_src), so, cudaHostAllocPortable) != cudaSuccess ||
                cudaHostAlloc(reinterpret_cast<void**>(&s->h_dst), dof, cudaHostAllocPortable) != cudaSuccess) {
                const char* msg = cudaGetErrorString(cudaGetLastError());
                f->slots.push_back(std::move(s));
                return jinc_fail(JINC_E_NOMEM, "JincResize: failed to allocate frame buffers (%zu + %zu bytes): %s", so, dof, msg);
            }
            f->slots.push_back(std::move(s));
        }
    }
    *out = f.release();
    return JINC_OK;
}

extern "C" const jinc_table* jinc_filter_table(const jinc_filter* f, int k)
{
    if (!f || k < 0 || k >= f->n_tables || f->devs.empty())
        return nullptr;
    return f->devs[0].tables[k];
}

extern "C" int jinc_filter_num_tables(const jinc_filter* f) { return f ? f->n_tables : 0; }
extern "C" int jinc_filter_num_devices(const jinc_filter* f) { return f ? static_cast<int>(f->devs.size()) : 0; }
extern "C" int64_t jinc_filter_kernel_launches(const jinc_filter* f) { return f ? f->launches.load() : 0; }

extern "C" int jinc_filter_process(jinc_filter* f, const jinc_frame* frame)
{
    if (int rc = check_frame(f, frame))
        return rc;
    Slot* s = acquire_slot(f);
    int rc = enqueue_frame(f, s, frame, 0, f->p.target_h, true);
    if (rc == JINC_OK)
        rc = finish_frame(f, s, 0, f->p.target_h, true);
    else
        cudaStreamSynchronize(s->stream);
    release_slot(f, s);
    return rc;
}

extern "C" int jinc_filter_process_device(jinc_filter* f, int device_index, const jinc_frame* frame, int table_mask, int parts,
                                          void* stream)
{
    if (int rc = check_frame(f, frame))
        return rc;
    if (device_index < 0 || device_index >= static_cast<int>(f->devs.size()))
        return jinc_fail(JINC_E_INVALID, "jinc_filter_process_device: bad device index %d", device_index);
    DeviceState& d = f->devs[device_index];
    JINC_CUDA(cudaSetDevice(d.ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : d.ctx->stream;
    for (int k = 0; k < f->n_tables; ++k) {
        if (!(table_mask & (1 << k)))
            continue;
        const void* src[JINC_MAX_PLANES];
        void* dst[JINC_MAX_PLANES];
        ptrdiff_t sp[JINC_MAX_PLANES], dp[JINC_MAX_PLANES];
        int n = 0;
        for (int i = 0; i < f->p.n_planes; ++i) {
            if (f->planes[i].table != k)
                continue;
            src[n] = frame->src[i];
            dst[n] = frame->dst[i];
            sp[n] = frame->src_pitch[i];
            dp[n] = frame->dst_pitch[i];
            ++n;
        }
        if (n == 0)
            continue;
        int launched = 0;
        const int rc = jinc_launch_resize_planes(d.ctx, d.tables[k], f->p.sample_bytes, f->peak, n, src, sp, dst, dp, 0,
                                                 d.tables[k]->sc.dst_h, st, &launched, parts);
        f->launches.fetch_add(launched);
        if (rc != JINC_OK)
            return rc;
    }
    return JINC_OK;
}

extern "C" int jinc_filter_process_device_batch(jinc_filter* f, int device_index, const jinc_frame* frames, int n_frames,
                                                int table_mask, int parts, void* stream)
{
    if (!f || !frames || n_frames < 1)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_process_device_batch: bad argument");
    if (device_index < 0 || device_index >= static_cast<int>(f->devs.size()))
        return jinc_fail(JINC_E_INVALID, "jinc_filter_process_device_batch: bad device index %d", device_index);
    DeviceState& d = f->devs[device_index];
    JINC_CUDA(cudaSetDevice(d.ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : d.ctx->stream;
    const size_t rec = jinc_plane_ptrs_size();
    for (int k = 0; k < f->n_tables; ++k) {
        if (!(table_mask & (1 << k)))
            continue;
        std::vector<unsigned char> host(rec * n_frames);
        int n = 0;
        for (int fi = 0; fi < n_frames; ++fi) {
            const void* src[JINC_MAX_PLANES];
            void* dst[JINC_MAX_PLANES];
            ptrdiff_t sp[JINC_MAX_PLANES], dp[JINC_MAX_PLANES];
            n = 0;
            for (int i = 0; i < f->p.n_planes; ++i) {
                if (f->planes[i].table != k)
                    continue;
                src[n] = frames[fi].src[i];
                dst[n] = frames[fi].dst[i];
                sp[n] = frames[fi].src_pitch[i];
                dp[n] = frames[fi].dst_pitch[i];
                ++n;
            }
            if (n == 0)
                break;
            if (int rc = jinc_pack_plane_ptrs(host.data() + rec * fi, f->p.sample_bytes, n, src, sp, dst, dp))
                return rc;
        }
        if (n == 0)
            continue;
        unsigned char* dbuf;
        {
            std::lock_guard<std::mutex> lk(f->mu);
            const int slot = d.batch_next;
            d.batch_next = (d.batch_next + 1) % DeviceState::kRing;
            if (d.batch_cap[slot] < host.size()) {
                cudaFree(d.batch_ptrs[slot]);
                d.batch_ptrs[slot] = nullptr;
                d.batch_cap[slot] = 0;
                if (cudaMalloc(reinterpret_cast<void**>(&d.batch_ptrs[slot]), host.size()) != cudaSuccess)
                    return jinc_fail(JINC_E_NOMEM, "jinc_filter_process_device_batch: cudaMalloc(%zu) failed", host.size());
                d.batch_cap[slot] = host.size();
            }
            dbuf = d.batch_ptrs[slot];
        }
        // pageable source: staged by the runtime before the call returns, ordered before the kernel on `st`
        JINC_CUDA(cudaMemcpyAsync(dbuf, host.data(), host.size(), cudaMemcpyHostToDevice, st));
        int launched = 0;
        const int rc = jinc_launch_resize_batch(d.ctx, d.tables[k], f->p.sample_bytes, f->peak, n, dbuf, n_frames, st, &launched, parts);
        f->launches.fetch_add(launched);
        if (rc != JINC_OK)
            return rc;
    }
    return JINC_OK;
}

extern "C" int jinc_filter_submit(jinc_filter* f, const jinc_frame* frame, int64_t* ticket)
{
    if (int rc = check_frame(f, frame))
        return rc;
    if (!ticket)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_submit: null ticket");
    Slot* s = acquire_slot(f);
    const int rc = enqueue_frame(f, s, frame, 0, f->p.target_h, true);
    if (rc != JINC_OK) {
        cudaStreamSynchronize(s->stream);
        release_slot(f, s);
        return rc;
    }
    s->ticket = f->next_ticket.fetch_add(1);
    *ticket = s->ticket;
    return JINC_OK;
}

extern "C" int jinc_filter_wait(jinc_filter* f, int64_t ticket)
{
    if (!f)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_wait: null filter");
    Slot* s = nullptr;
    {
        std::lock_guard<std::mutex> lk(f->mu);
        for (auto& c : f->slots)
            if (c->busy && c->ticket == ticket)
                s = c.get();
    }
    if (!s)
        return jinc_fail(JINC_E_INVALID, "jinc_filter_wait: unknown ticket %lld", static_cast<long long>(ticket));
    const int rc = finish_frame(f, s, 0, f->p.target_h, true);
    release_slot(f, s);
    return rc;
}

extern "C" int jinc_plan_frame_owner(int64_t frame, int n_parts)
{
    if (n_parts < 1 || frame < 0)
        return jinc_fail(JINC_E_INVALID, "jinc_plan_frame_owner: bad argument");
    return static_cast<int>(frame % n_parts);
}

extern "C" int jinc_plan_row_bands(int target_h, int n_parts, int32_t* y_begin, int32_t* y_end)
{
    if (target_h < 1 || n_parts < 1 || !y_begin || !y_end)
        return jinc_fail(JINC_E_INVALID, "jinc_plan_row_bands: bad argument");
    const int band = static_cast<int>(align_up(static_cast<size_t>((target_h + n_parts - 1) / n_parts), 16));
    for (int i = 0; i < n_parts; ++i) {
        y_begin[i] = std::min(target_h, i * band);
        y_end[i] = std::min(target_h, (i + 1) * band);
    }
    return JINC_OK;
}

extern "C" int jinc_filter_process_split(jinc_filter* f, const jinc_frame* frame)
{
    if (int rc = check_frame(f, frame))
        return rc;
    const int nd = static_cast<int>(f->devs.size());
    if (nd == 1)
        return jinc_filter_process(f, frame);
    // one slot per GPU, bands cut on multiples of 16 luma rows (whole cell pairs for luma and subsampled chroma)
    std::vector<Slot*> held(nd, nullptr);
    {
        std::unique_lock<std::mutex> lk(f->mu);
        for (;;) {
            bool ok = true;
            for (int di = 0; di < nd && ok; ++di) {
                held[di] = nullptr;
                for (auto& s : f->slots)
                    if (!s->busy && s->dev_index == di) {
                        held[di] = s.get();
                        break;
                    }
                ok = held[di] != nullptr;
            }
            if (ok)
                break;
            f->cv.wait(lk);
        }
        for (Slot* s : held)
            s->busy = true;
    }
    std::vector<int32_t> yb(nd), ye(nd);
    int rc = jinc_plan_row_bands(f->p.target_h, nd, yb.data(), ye.data());
    std::vector<std::pair<int, int>> ranges(nd);
    for (int di = 0; di < nd; ++di) {
        ranges[di] = {yb[di], ye[di]};
        if (ye[di] > yb[di] && rc == JINC_OK)
            rc = enqueue_frame(f, held[di], frame, yb[di], ye[di], false);
    }
    for (int di = 0; di < nd; ++di) {
        if (ranges[di].second > ranges[di].first) {
            if (rc == JINC_OK)
                rc = finish_frame(f, held[di], ranges[di].first, ranges[di].second, false);
            else
                cudaStreamSynchronize(held[di]->stream);
        }
        release_slot(f, held[di]);
    }
    return rc;
}
