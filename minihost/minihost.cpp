/*
 * minihost.cpp -- in-process stand-in for the AviSynth+ frame server.
 *
 * Implements the avs_* C API subset declared in include/avisynth_c.h plus the
 * mh_* driver surface of minihost.h.  It owns clips, frames (64-byte aligned
 * planes with padded pitches and tail slack, as AviSynth+ allocates them),
 * frame properties, the script-function registry and avs_invoke's mapping of
 * named arguments onto a function's parameter string.
 *
 * Test/bench infrastructure only: nothing here is on the resampling path.
 */
#include "minihost.h"

#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <variant>
#include <vector>

namespace {

std::atomic<long> g_live_frames{0};
std::atomic<long> g_live_clips{0};

using PropValue = std::variant<int64_t, double>;
struct PropMap {
    std::map<std::string, PropValue> kv;
};

struct Function {
    std::string name;
    std::string params;
    AVS_ApplyFunc apply;
    void* user_data;
};

struct Param {
    std::string name; /* empty for positional-only */
    char type;
    bool optional;
};

std::vector<Param> parse_params(const std::string& s)
{
    std::vector<Param> out;
    size_t i = 0;
    while (i < s.size()) {
        Param p{"", 0, false};
        if (s[i] == '[') {
            size_t j = s.find(']', i);
            p.name = s.substr(i + 1, j - i - 1);
            p.optional = true;
            i = j + 1;
        }
        p.type = s[i++];
        if (i < s.size() && (s[i] == '*' || s[i] == '+'))
            ++i; /* array qualifiers are not needed by any function we host */
        out.push_back(p);
    }
    return out;
}

bool type_ok(char want, const AVS_Value& v)
{
    switch (want) {
    case 'c': return avs_is_clip(v);
    case 'i': return avs_is_int(v);
    case 'f': return avs_is_float(v);
    case 's': return avs_is_string(v);
    case 'b': return avs_is_bool(v);
    case '.': return true;
    default: return false;
    }
}

int align_up(int v, int a) { return (v + a - 1) / a * a; }

} // namespace

struct AVS_ScriptEnvironment {
    std::mutex mu;
    std::map<std::string, Function> functions;
    std::deque<std::string> strings; /* stable storage for avs_save_string / errors */
    std::string last_error;
    int interface_version = AVISYNTH_INTERFACE_VERSION;
    int interface_bugfix = AVISYNTHPLUS_INTERFACE_BUGFIX_VERSION;
    int cpu_flags = 0;
    std::vector<void*> plugins;

    const char* save(const std::string& s)
    {
        std::lock_guard<std::mutex> lk(mu);
        strings.push_back(s);
        return strings.back().c_str();
    }
};

struct AVS_Clip {
    std::atomic<long> refcount{1};
    AVS_ScriptEnvironment* env = nullptr;
    bool is_filter = false;
    /* source clip */
    AVS_VideoInfo vi{};
    std::vector<AVS_VideoFrame*> stored;
    /* C filter */
    AVS_FilterInfo fi{};
};

/* ---------------------------------------------------------------- video info */

static int sample_bits_of(int pixel_type)
{
    switch (pixel_type & AVS_CS_SAMPLE_BITS_MASK) {
    case AVS_CS_SAMPLE_BITS_8: return 8;
    case AVS_CS_SAMPLE_BITS_10: return 10;
    case AVS_CS_SAMPLE_BITS_12: return 12;
    case AVS_CS_SAMPLE_BITS_14: return 14;
    case AVS_CS_SAMPLE_BITS_16: return 16;
    case AVS_CS_SAMPLE_BITS_32: return 32;
    default: return 8;
    }
}

static bool is_planar_yuv_family(const AVS_VideoInfo* p)
{
    return avs_is_planar(p) && (p->pixel_type & (AVS_CS_YUV | AVS_CS_YUVA));
}

static int generic_of(const AVS_VideoInfo* p)
{
    return p->pixel_type & AVS_CS_PLANAR_MASK & ~AVS_CS_SAMPLE_BITS_MASK;
}

AVSC_API(int, avs_is_y)(const AVS_VideoInfo* p)
{
    return generic_of(p) == (AVS_CS_GENERIC_Y & AVS_CS_PLANAR_FILTER);
}

AVSC_API(int, avs_is_420)(const AVS_VideoInfo* p)
{
    const int g = generic_of(p);
    return g == (AVS_CS_GENERIC_YUV420 & AVS_CS_PLANAR_FILTER) || g == (AVS_CS_GENERIC_YUVA420 & AVS_CS_PLANAR_FILTER);
}

AVSC_API(int, avs_is_422)(const AVS_VideoInfo* p)
{
    const int g = generic_of(p);
    return g == (AVS_CS_GENERIC_YUV422 & AVS_CS_PLANAR_FILTER) || g == (AVS_CS_GENERIC_YUVA422 & AVS_CS_PLANAR_FILTER);
}

AVSC_API(int, avs_is_444)(const AVS_VideoInfo* p)
{
    const int g = generic_of(p);
    return g == (AVS_CS_GENERIC_YUV444 & AVS_CS_PLANAR_FILTER) || g == (AVS_CS_GENERIC_YUVA444 & AVS_CS_PLANAR_FILTER);
}

AVSC_API(int, avs_is_yv411)(const AVS_VideoInfo* p)
{
    return (p->pixel_type & AVS_CS_PLANAR_MASK) == (AVS_CS_YV411 & AVS_CS_PLANAR_FILTER);
}

AVSC_API(int, avs_is_planar_rgb)(const AVS_VideoInfo* p)
{
    return generic_of(p) == (AVS_CS_GENERIC_RGBP & AVS_CS_PLANAR_FILTER);
}

AVSC_API(int, avs_is_planar_rgba)(const AVS_VideoInfo* p)
{
    return generic_of(p) == (AVS_CS_GENERIC_RGBAP & AVS_CS_PLANAR_FILTER);
}

AVSC_API(int, avs_num_components)(const AVS_VideoInfo* p)
{
    if (avs_is_y(p))
        return 1;
    if (avs_is_planar_rgba(p) || avs_is_yuva(p))
        return 4;
    if (!avs_is_planar(p)) /* packed formats */
        return (p->pixel_type & AVS_CS_RGBA_TYPE) ? 4 : 3;
    return 3;
}

AVSC_API(int, avs_bits_per_component)(const AVS_VideoInfo* p)
{
    return sample_bits_of(p->pixel_type);
}

AVSC_API(int, avs_component_size)(const AVS_VideoInfo* p)
{
    const int b = sample_bits_of(p->pixel_type);
    return b == 8 ? 1 : (b == 32 ? 4 : 2);
}

static int sub_shift(int code)
{
    switch (code) {
    case 3: return 0; /* _1 */
    case 0: return 1; /* _2 */
    case 1: return 2; /* _4 */
    default: return 0;
    }
}

AVSC_API(int, avs_get_plane_width_subsampling)(const AVS_VideoInfo* p, int plane)
{
    if (plane == AVS_PLANAR_U || plane == AVS_PLANAR_V) {
        if (!is_planar_yuv_family(p) || avs_is_y(p))
            return 0;
        return sub_shift((p->pixel_type >> AVS_CS_SHIFT_SUB_WIDTH) & 7);
    }
    return 0;
}

AVSC_API(int, avs_get_plane_height_subsampling)(const AVS_VideoInfo* p, int plane)
{
    if (plane == AVS_PLANAR_U || plane == AVS_PLANAR_V) {
        if (!is_planar_yuv_family(p) || avs_is_y(p))
            return 0;
        return sub_shift((p->pixel_type >> AVS_CS_SHIFT_SUB_HEIGHT) & 7);
    }
    return 0;
}

/* ---------------------------------------------------------------- frames */

/* Frame buffers are recycled through a small pool keyed by size, as AviSynth+'s frame cache does: a filter's output
   frame normally reuses memory that is already mapped (fresh pages are zeroed once, when first allocated). */
static std::mutex g_pool_mutex;
static std::unordered_map<size_t, std::vector<void*>> g_pool;
static const size_t kPoolPerSize = 64;

static void* pool_get(size_t total)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        auto it = g_pool.find(total);
        if (it != g_pool.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            return p;
        }
    }
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, total) != 0)
        return nullptr;
    memset(mem, 0, total);
    return mem;
}

static void pool_put(void* p, size_t total)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        auto& v = g_pool[total];
        if (v.size() < kPoolPerSize) {
            v.push_back(p);
            return;
        }
    }
    free(p);
}

static AVS_VideoFrame* alloc_frame(const AVS_VideoInfo* vi, const AVS_VideoFrame* prop_src)
{
    const int cs = avs_component_size(vi);
    const int ncomp = avs_is_planar(vi) ? avs_num_components(vi) : 1;
    const bool rgb = avs_is_rgb(vi);
    int w[4] = {0, 0, 0, 0}, h[4] = {0, 0, 0, 0};
    w[0] = vi->width;
    h[0] = vi->height;
    if (ncomp >= 3) {
        const int sw = rgb ? 0 : avs_get_plane_width_subsampling(vi, AVS_PLANAR_U);
        const int sh = rgb ? 0 : avs_get_plane_height_subsampling(vi, AVS_PLANAR_U);
        w[1] = w[2] = vi->width >> sw;
        h[1] = h[2] = vi->height >> sh;
    }
    if (ncomp == 4) {
        w[3] = vi->width;
        h[3] = vi->height;
    }
    int pitch[4], offset[4];
    size_t total = 0;
    for (int i = 0; i < 4; ++i) {
        pitch[i] = w[i] ? align_up(w[i] * cs, 64) : 0;
        offset[i] = (int)total;
        total += (size_t)pitch[i] * h[i];
        total = (total + 63) & ~(size_t)63;
    }
    total += 256; /* tail slack: SIMD readers may run past the last row */

    auto* vfb = new AVS_VideoFrameBuffer();
    void* mem = pool_get(total);
    if (!mem) {
        delete vfb;
        return nullptr;
    }
    vfb->data = static_cast<BYTE*>(mem);
    vfb->data_size = (int)total;
    vfb->sequence_number = 0;
    vfb->refcount = 1;
    vfb->device_data = nullptr;

    auto* f = new AVS_VideoFrame();
    f->refcount = 1;
    f->vfb = vfb;
    f->offset = offset[0];
    f->pitch = pitch[0];
    f->row_size = w[0] * cs;
    f->height = h[0];
    f->offsetU = offset[1];
    f->offsetV = offset[2];
    f->pitchUV = pitch[1];
    f->row_sizeUV = w[1] * cs;
    f->heightUV = h[1];
    f->offsetA = offset[3];
    f->pitchA = pitch[3];
    f->row_sizeA = w[3] * cs;
    auto* props = new PropMap();
    if (prop_src && prop_src->properties)
        *props = *static_cast<const PropMap*>(prop_src->properties);
    f->properties = props;
    g_live_frames.fetch_add(1);
    return f;
}

static void addref_frame(AVS_VideoFrame* f) { __atomic_add_fetch(&f->refcount, 1, __ATOMIC_SEQ_CST); }

AVSC_API(void, avs_release_video_frame)(AVS_VideoFrame* f)
{
    if (!f)
        return;
    if (__atomic_sub_fetch(&f->refcount, 1, __ATOMIC_SEQ_CST) == 0) {
        pool_put(f->vfb->data, (size_t)f->vfb->data_size);
        delete f->vfb;
        delete static_cast<PropMap*>(f->properties);
        delete f;
        g_live_frames.fetch_sub(1);
    }
}

AVSC_API(AVS_VideoFrame*, avs_copy_video_frame)(AVS_VideoFrame* f)
{
    addref_frame(f);
    return f;
}

AVSC_API(int, avs_is_writable)(const AVS_VideoFrame* p) { return p->refcount == 1 && p->vfb->refcount == 1; }

AVSC_API(int, avs_get_pitch_p)(const AVS_VideoFrame* p, int plane)
{
    switch (plane) {
    case AVS_PLANAR_U: case AVS_PLANAR_V: case AVS_PLANAR_B: case AVS_PLANAR_R: return p->pitchUV;
    case AVS_PLANAR_A: return p->pitchA;
    default: return p->pitch;
    }
}

AVSC_API(int, avs_get_row_size_p)(const AVS_VideoFrame* p, int plane)
{
    switch (plane) {
    case AVS_PLANAR_U: case AVS_PLANAR_V: case AVS_PLANAR_B: case AVS_PLANAR_R: return p->pitchUV ? p->row_sizeUV : 0;
    case AVS_PLANAR_A: return p->pitchA ? p->row_sizeA : 0;
    default: return p->row_size;
    }
}

AVSC_API(int, avs_get_height_p)(const AVS_VideoFrame* p, int plane)
{
    switch (plane) {
    case AVS_PLANAR_U: case AVS_PLANAR_V: case AVS_PLANAR_B: case AVS_PLANAR_R: return p->pitchUV ? p->heightUV : 0;
    case AVS_PLANAR_A: return p->pitchA ? p->height : 0;
    default: return p->height;
    }
}

static BYTE* plane_ptr(const AVS_VideoFrame* p, int plane)
{
    switch (plane) {
    case AVS_PLANAR_U: case AVS_PLANAR_B: return p->vfb->data + p->offsetU;
    case AVS_PLANAR_V: case AVS_PLANAR_R: return p->vfb->data + p->offsetV;
    case AVS_PLANAR_A: return p->vfb->data + p->offsetA;
    default: return p->vfb->data + p->offset; /* Y, G */
    }
}

AVSC_API(const BYTE*, avs_get_read_ptr_p)(const AVS_VideoFrame* p, int plane) { return plane_ptr(p, plane); }
AVSC_API(BYTE*, avs_get_write_ptr_p)(const AVS_VideoFrame* p, int plane) { return plane_ptr(p, plane); }

AVSC_API(AVS_VideoFrame*, avs_new_video_frame_a)(AVS_ScriptEnvironment*, const AVS_VideoInfo* vi, int)
{
    return alloc_frame(vi, nullptr);
}

AVSC_API(AVS_VideoFrame*, avs_new_video_frame_p)(AVS_ScriptEnvironment*, const AVS_VideoInfo* vi, const AVS_VideoFrame* prop_src)
{
    return alloc_frame(vi, prop_src);
}

AVSC_API(int, avs_make_writable)(AVS_ScriptEnvironment*, AVS_VideoFrame** pvf) { return avs_is_writable(*pvf); }

/* ---------------------------------------------------------------- properties */

AVSC_API(const AVS_Map*, avs_get_frame_props_ro)(AVS_ScriptEnvironment*, const AVS_VideoFrame* frame)
{
    return reinterpret_cast<const AVS_Map*>(frame->properties);
}

AVSC_API(AVS_Map*, avs_get_frame_props_rw)(AVS_ScriptEnvironment*, AVS_VideoFrame* frame)
{
    return reinterpret_cast<AVS_Map*>(frame->properties);
}

AVSC_API(int, avs_prop_num_keys)(AVS_ScriptEnvironment*, const AVS_Map* map)
{
    return (int)reinterpret_cast<const PropMap*>(map)->kv.size();
}

AVSC_API(char, avs_prop_get_type)(AVS_ScriptEnvironment*, const AVS_Map* map, const char* key)
{
    const auto& kv = reinterpret_cast<const PropMap*>(map)->kv;
    auto it = kv.find(key);
    if (it == kv.end())
        return 'u';
    return std::holds_alternative<int64_t>(it->second) ? 'i' : 'f';
}

AVSC_API(int64_t, avs_prop_get_int)(AVS_ScriptEnvironment*, const AVS_Map* map, const char* key, int index, int* error)
{
    const auto& kv = reinterpret_cast<const PropMap*>(map)->kv;
    auto it = kv.find(key);
    if (it == kv.end() || index != 0 || !std::holds_alternative<int64_t>(it->second)) {
        if (error)
            *error = 1;
        return 0;
    }
    if (error)
        *error = 0;
    return std::get<int64_t>(it->second);
}

AVSC_API(double, avs_prop_get_float)(AVS_ScriptEnvironment*, const AVS_Map* map, const char* key, int index, int* error)
{
    const auto& kv = reinterpret_cast<const PropMap*>(map)->kv;
    auto it = kv.find(key);
    if (it == kv.end() || index != 0 || !std::holds_alternative<double>(it->second)) {
        if (error)
            *error = 1;
        return 0;
    }
    if (error)
        *error = 0;
    return std::get<double>(it->second);
}

AVSC_API(int, avs_prop_set_int)(AVS_ScriptEnvironment*, AVS_Map* map, const char* key, int64_t i, int)
{
    reinterpret_cast<PropMap*>(map)->kv[key] = i;
    return 0;
}

AVSC_API(int, avs_prop_set_float)(AVS_ScriptEnvironment*, AVS_Map* map, const char* key, double d, int)
{
    reinterpret_cast<PropMap*>(map)->kv[key] = d;
    return 0;
}

AVSC_API(int, avs_prop_delete_key)(AVS_ScriptEnvironment*, AVS_Map* map, const char* key)
{
    return (int)reinterpret_cast<PropMap*>(map)->kv.erase(key);
}

/* ---------------------------------------------------------------- clips */

AVSC_API(AVS_Clip*, avs_copy_clip)(AVS_Clip* c)
{
    c->refcount.fetch_add(1);
    return c;
}

AVSC_API(void, avs_release_clip)(AVS_Clip* c)
{
    if (!c)
        return;
    if (c->refcount.fetch_sub(1) == 1) {
        if (c->is_filter) {
            if (c->fi.free_filter)
                c->fi.free_filter(&c->fi);
            avs_release_clip(c->fi.child);
        } else {
            for (AVS_VideoFrame* f : c->stored)
                avs_release_video_frame(f);
        }
        delete c;
        g_live_clips.fetch_sub(1);
    }
}

AVSC_API(const char*, avs_clip_get_error)(AVS_Clip* c) { return c->is_filter ? c->fi.error : nullptr; }

AVSC_API(const AVS_VideoInfo*, avs_get_video_info)(AVS_Clip* c) { return c->is_filter ? &c->fi.vi : &c->vi; }

AVSC_API(int, avs_get_version)(AVS_Clip* c) { return c->env->interface_version; }

AVSC_API(AVS_VideoFrame*, avs_get_frame)(AVS_Clip* c, int n)
{
    if (c->is_filter) {
        if (c->fi.get_frame)
            return c->fi.get_frame(&c->fi, n);
        return avs_get_frame(c->fi.child, n); /* not wired up yet: behave as a pass-through */
    }
    if (c->stored.empty())
        return nullptr;
    if (n < 0)
        n = 0;
    if (n >= c->vi.num_frames)
        n = c->vi.num_frames - 1;
    AVS_VideoFrame* f = c->stored[(size_t)n % c->stored.size()];
    addref_frame(f);
    return f;
}

AVSC_API(void, avs_set_to_clip)(AVS_Value* v, AVS_Clip* c)
{
    v->type = 'c';
    v->array_size = 0;
    v->d.clip = avs_copy_clip(c);
}

AVSC_API(AVS_Clip*, avs_take_clip)(AVS_Value v, AVS_ScriptEnvironment*)
{
    if (!avs_is_clip(v))
        return nullptr;
    return avs_copy_clip(static_cast<AVS_Clip*>(v.d.clip));
}

AVSC_API(void, avs_release_value)(AVS_Value v)
{
    if (avs_is_clip(v))
        avs_release_clip(static_cast<AVS_Clip*>(v.d.clip));
}

AVSC_API(void, avs_copy_value)(AVS_Value* dest, AVS_Value src)
{
    *dest = src;
    if (avs_is_clip(src))
        avs_copy_clip(static_cast<AVS_Clip*>(src.d.clip));
}

AVSC_API(AVS_Clip*, avs_new_c_filter)(AVS_ScriptEnvironment* e, AVS_FilterInfo** fi, AVS_Value child, int)
{
    auto* c = new AVS_Clip();
    g_live_clips.fetch_add(1);
    c->env = e;
    c->is_filter = true;
    c->fi.child = avs_take_clip(child, e);
    if (c->fi.child)
        c->fi.vi = *avs_get_video_info(c->fi.child);
    c->fi.env = e;
    *fi = &c->fi;
    return c;
}

/* ---------------------------------------------------------------- environment */

AVSC_API(const char*, avs_get_error)(AVS_ScriptEnvironment* e) { return e->last_error.empty() ? nullptr : e->last_error.c_str(); }

AVSC_API(int, avs_get_cpu_flags)(AVS_ScriptEnvironment* e) { return e->cpu_flags; }

/* 0 means "the host implements at least `version`" */
AVSC_API(int, avs_check_version)(AVS_ScriptEnvironment* e, int version) { return e->interface_version >= version ? 0 : -1; }

AVSC_API(size_t, avs_get_env_property)(AVS_ScriptEnvironment* e, int prop)
{
    switch (prop) {
    case AVS_AEP_INTERFACE_VERSION: return (size_t)e->interface_version;
    case AVS_AEP_INTERFACE_BUGFIX: return (size_t)e->interface_bugfix;
    case AVS_AEP_LOGICAL_CPUS: case AVS_AEP_PHYSICAL_CPUS: return std::thread::hardware_concurrency();
    default: return 0;
    }
}

AVSC_API(char*, avs_save_string)(AVS_ScriptEnvironment* e, const char* s, int length)
{
    return const_cast<char*>(e->save(length >= 0 ? std::string(s, (size_t)length) : std::string(s)));
}

AVSC_API(int, avs_add_function)(AVS_ScriptEnvironment* e, const char* name, const char* params, AVS_ApplyFunc apply, void* user_data)
{
    std::lock_guard<std::mutex> lk(e->mu);
    e->functions[name] = Function{name, params, apply, user_data};
    return 0;
}

AVSC_API(int, avs_function_exists)(AVS_ScriptEnvironment* e, const char* name)
{
    std::lock_guard<std::mutex> lk(e->mu);
    return e->functions.count(name) ? 1 : 0;
}

AVSC_API(AVS_Value, avs_invoke)(AVS_ScriptEnvironment* e, const char* name, AVS_Value args, const char** arg_names)
{
    Function fn;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        auto it = e->functions.find(name);
        if (it == e->functions.end())
            return avs_new_value_error(e->save(std::string("Script error: There is no function named '") + name + "'."));
        fn = it->second;
    }
    const std::vector<Param> params = parse_params(fn.params);
    std::vector<AVS_Value> slots(params.size(), avs_void);
    const int n = avs_array_size(args);
    size_t next_pos = 0;
    for (int i = 0; i < n; ++i) {
        const AVS_Value v = avs_array_elt(args, i);
        const char* an = arg_names ? arg_names[i] : nullptr;
        if (!an) {
            if (next_pos >= params.size() || !type_ok(params[next_pos].type, v))
                return avs_new_value_error(e->save(std::string("Script error: Invalid arguments to function '") + name + "'."));
            slots[next_pos++] = v;
        } else {
            size_t k = 0;
            for (; k < params.size(); ++k)
                if (params[k].name == an)
                    break;
            if (k == params.size())
                return avs_new_value_error(e->save(std::string("Script error: ") + name + " does not have a named argument \"" + an + "\"."));
            if (!type_ok(params[k].type, v))
                return avs_new_value_error(e->save(std::string("Script error: the named argument \"") + an + "\" to " + name + " had the wrong type."));
            slots[k] = v;
        }
    }
    for (size_t k = 0; k < params.size(); ++k)
        if (!params[k].optional && !avs_defined(slots[k]))
            return avs_new_value_error(e->save(std::string("Script error: Invalid arguments to function '") + name + "'."));
    return fn.apply(e, avs_new_value_array(slots.data(), (int)slots.size()), fn.user_data);
}

/* ================================================================ driver API */

static int detect_cpu_flags()
{
    int f = 0;
#if defined(__x86_64__) || defined(__i386__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("sse2")) f |= AVS_CPUF_SSE2;
    if (__builtin_cpu_supports("sse3")) f |= AVS_CPUF_SSE3;
    if (__builtin_cpu_supports("ssse3")) f |= AVS_CPUF_SSSE3;
    if (__builtin_cpu_supports("sse4.1")) f |= AVS_CPUF_SSE4_1;
    if (__builtin_cpu_supports("sse4.2")) f |= AVS_CPUF_SSE4_2;
    if (__builtin_cpu_supports("avx")) f |= AVS_CPUF_AVX;
    if (__builtin_cpu_supports("avx2")) f |= AVS_CPUF_AVX2;
    if (__builtin_cpu_supports("fma")) f |= AVS_CPUF_FMA3;
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") &&
        __builtin_cpu_supports("avx512vl"))
        f |= AVS_CPUF_AVX512F;
#endif
    return f;
}

extern "C" {

AVS_ScriptEnvironment* mh_env_create(void)
{
    auto* e = new AVS_ScriptEnvironment();
    e->cpu_flags = detect_cpu_flags();
    return e;
}

void mh_env_destroy(AVS_ScriptEnvironment* env)
{
    /* plugins stay mapped: their code may still be referenced by live clips */
    delete env;
}

void mh_env_set_interface(AVS_ScriptEnvironment* env, int version, int bugfix)
{
    env->interface_version = version;
    env->interface_bugfix = bugfix;
}

void mh_env_set_cpu_flags(AVS_ScriptEnvironment* env, int flags) { env->cpu_flags = flags; }

const char* mh_last_error(AVS_ScriptEnvironment* env) { return env->last_error.c_str(); }

const char* mh_load_plugin(AVS_ScriptEnvironment* env, const char* path)
{
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        env->last_error = std::string("dlopen failed: ") + dlerror();
        return nullptr;
    }
    using init_fn = const char*(AVSC_CC*)(AVS_ScriptEnvironment*);
    auto init = reinterpret_cast<init_fn>(dlsym(h, "avisynth_c_plugin_init"));
    if (!init) {
        env->last_error = std::string("no avisynth_c_plugin_init in ") + path;
        dlclose(h);
        return nullptr;
    }
    env->plugins.push_back(h);
    return init(env);
}

const char* mh_function_params(AVS_ScriptEnvironment* env, const char* name)
{
    std::lock_guard<std::mutex> lk(env->mu);
    auto it = env->functions.find(name);
    return it == env->functions.end() ? nullptr : it->second.params.c_str();
}

AVS_Clip* mh_source_create(AVS_ScriptEnvironment* env, int width, int height, int pixel_type, int num_frames, int distinct)
{
    auto* c = new AVS_Clip();
    g_live_clips.fetch_add(1);
    c->env = env;
    c->vi.width = width;
    c->vi.height = height;
    c->vi.fps_numerator = 24;
    c->vi.fps_denominator = 1;
    c->vi.num_frames = num_frames;
    c->vi.pixel_type = pixel_type;
    if (distinct < 1)
        distinct = 1;
    for (int k = 0; k < distinct; ++k) {
        AVS_VideoFrame* f = alloc_frame(&c->vi, nullptr);
        if (!f) {
            avs_release_clip(c);
            env->last_error = "mh_source_create: out of memory";
            return nullptr;
        }
        c->stored.push_back(f);
    }
    return c;
}

int mh_source_fill_plane(AVS_Clip* clip, int k, int plane, const void* src, ptrdiff_t src_pitch)
{
    if (clip->is_filter || k < 0 || (size_t)k >= clip->stored.size())
        return -1;
    AVS_VideoFrame* f = clip->stored[(size_t)k];
    const int rs = avs_get_row_size_p(f, plane), h = avs_get_height_p(f, plane), p = avs_get_pitch_p(f, plane);
    BYTE* d = avs_get_write_ptr_p(f, plane);
    for (int y = 0; y < h; ++y)
        memcpy(d + (size_t)y * p, static_cast<const BYTE*>(src) + (ptrdiff_t)y * src_pitch, (size_t)rs);
    return 0;
}

void mh_source_set_prop_int(AVS_Clip* clip, const char* key, int64_t value)
{
    for (AVS_VideoFrame* f : clip->stored)
        static_cast<PropMap*>(f->properties)->kv[key] = value;
}

void mh_source_clear_prop(AVS_Clip* clip, const char* key)
{
    for (AVS_VideoFrame* f : clip->stored)
        static_cast<PropMap*>(f->properties)->kv.erase(key);
}

struct mh_args {
    std::vector<AVS_Value> values;
    std::vector<const char*> names;
    std::deque<std::string> strings;
    bool any_named = false;
};

mh_args* mh_args_create(void) { return new mh_args(); }

void mh_args_destroy(mh_args* a)
{
    for (AVS_Value& v : a->values)
        avs_release_value(v);
    delete a;
}

static void push(mh_args* a, AVS_Value v, const char* name)
{
    a->values.push_back(v);
    if (name && *name) {
        a->strings.emplace_back(name);
        a->names.push_back(a->strings.back().c_str());
        a->any_named = true;
    } else {
        a->names.push_back(nullptr);
    }
}

void mh_args_add_clip(mh_args* a, AVS_Clip* clip, const char* name) { push(a, avs_new_value_clip(clip), name); }
void mh_args_add_int(mh_args* a, int v, const char* name) { push(a, avs_new_value_int(v), name); }
void mh_args_add_float(mh_args* a, float v, const char* name) { push(a, avs_new_value_float(v), name); }

void mh_args_add_string(mh_args* a, const char* s, const char* name)
{
    a->strings.emplace_back(s);
    push(a, avs_new_value_string(a->strings.back().c_str()), name);
}

AVS_Clip* mh_invoke_clip(AVS_ScriptEnvironment* env, const char* name, mh_args* a)
{
    env->last_error.clear();
    AVS_Value r = avs_invoke(env, name, avs_new_value_array(a->values.data(), (int)a->values.size()), a->names.data());
    if (avs_is_error(r)) {
        env->last_error = avs_as_error(r) ? avs_as_error(r) : "unknown error";
        return nullptr;
    }
    if (!avs_is_clip(r)) {
        env->last_error = "function did not return a clip";
        return nullptr;
    }
    return static_cast<AVS_Clip*>(r.d.clip); /* the value's reference passes to the caller */
}

AVS_FilterInfo* mh_clip_filter_info(AVS_Clip* clip) { return clip->is_filter ? &clip->fi : nullptr; }

int mh_clip_mt_mode(AVS_Clip* clip)
{
    if (!clip->is_filter || !clip->fi.set_cache_hints)
        return 0;
    return clip->fi.set_cache_hints(&clip->fi, AVS_CACHE_GET_MTMODE, 0);
}

int mh_frame_prop_int(AVS_ScriptEnvironment*, const AVS_VideoFrame* f, const char* key, int64_t* out)
{
    const auto& kv = static_cast<const PropMap*>(f->properties)->kv;
    auto it = kv.find(key);
    if (it == kv.end() || !std::holds_alternative<int64_t>(it->second))
        return 0;
    *out = std::get<int64_t>(it->second);
    return 1;
}

long mh_live_frames(void) { return g_live_frames.load(); }
long mh_live_clips(void) { return g_live_clips.load(); }

double mh_pull_frames(AVS_Clip* clip, int first, int count, int threads)
{
    if (threads < 1)
        threads = 1;
    std::atomic<int> next{0};
    std::atomic<int> failed{0};
    auto worker = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= count)
                return;
            AVS_VideoFrame* f = avs_get_frame(clip, first + i);
            if (!f)
                failed.fetch_add(1);
            else
                avs_release_video_frame(f);
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t)
            pool.emplace_back(worker);
        for (auto& t : pool)
            t.join();
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (failed.load())
        return -1.0;
    return std::chrono::duration<double>(t1 - t0).count();
}

} /* extern "C" */
