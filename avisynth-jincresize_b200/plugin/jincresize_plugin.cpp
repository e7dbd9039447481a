// jincresize_plugin.cpp -- AviSynth+ C-API plugin: the drop-in replacement for the reference's
// libjincresize (src/JincResize.cpp:603-1111), with every pixel computed on B200 GPUs through the C ABI of
// include/jinc_b200.h.  No CUDA headers here and no CPU resampling path: if the CUDA library cannot run, filter
// construction returns an AviSynth error value.
//
// Kept identical to the reference at the script boundary:
//   * the five script functions and their parameter strings (src/JincResize.cpp:1044-1108);
//   * argument defaults, validation order and error texts (:689-789);
//   * alias functions forward only the optional arguments that were given, by name, plus tap (:1007-1040);
//   * cplace defaults from frame 0's _ChromaLocation (:725-742);
//   * output frames inherit the source frame's properties; _ChromaLocation is written for 4:2:0/4:2:2/4:1:1 (:613-625).
//   * opt's validation, including the CPU-feature errors of opt=1/2/3 (:747-756), although opt selects nothing here;
//   * the output _ChromaLocation is what the reference writes: always 2 for 4:2:0 / 4:2:2 / 4:1:1 clips, because the
//     reference never stores the parsed cplace in its instance (src/JincResize.h:41, src/JincResize.cpp:617-625, 715).
//     JINCRESIZE_B200_CHROMALOC=actual writes the cplace actually used instead (what the reference's README documents).
// Deliberate differences (DESIGN.md "Boundary"):
//   * threads / opt / initial_capacity / initial_factor are accepted and validated but select nothing: there is one
//     GPU path;
//   * one instance serves all Prefetch threads (MT_NICE_FILTER): concurrent get_frame calls take different in-flight
//     slots of the GPU pipeline, which is what overlaps copies and kernels.  The reference asks for
//     MT_MULTI_INSTANCE (:649-652), i.e. one private table set per thread; JINCRESIZE_B200_MTMODE=2 reports that mode
//     instead, and is cheap here because instances created with identical arguments share ONE GPU filter (tables,
//     slots, streams) through a process-wide reference-counted cache.
// Environment: JINCRESIZE_B200_DEVICES ("0,1,..." or "all"; default: device 0), JINCRESIZE_B200_SLOTS,
// JINCRESIZE_B200_HOSTREG=0 (never page-lock the host's frame buffers), JINCRESIZE_B200_BANDS=n (cut every frame into n row
// bands over the filter's GPUs: latency of single large frames), JINCRESIZE_B200_CHROMALOC, JINCRESIZE_B200_MTMODE.
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "avisynth_c.h"
#include "jinc_b200.h"

namespace {

struct Instance {
    jinc_filter* filter = nullptr;
    std::string key; // entry of the shared-filter cache
    int bands = 0;   // > 0: every frame is cut into this many row bands (JINCRESIZE_B200_BANDS)
    int n_planes = 0;
    bool rgb = false;
    bool writes_chromaloc = false;
    int chromaloc = 0;
};

// Filters shared between instances created with identical parameters (every Prefetch thread of an MT_MULTI_INSTANCE
// host, or the same call twice in a script): jinc_filter is thread-safe, so one set of tables and slots serves all.
struct Shared {
    jinc_filter* filter = nullptr;
    int refs = 0;
};
std::mutex g_cache_mu;
std::map<std::string, Shared> g_cache;

jinc_filter* cache_acquire(const jinc_filter_params& p, std::string* key, std::string* err)
{
    key->assign(reinterpret_cast<const char*>(&p), sizeof(p));
    std::lock_guard<std::mutex> lk(g_cache_mu); // construction is serialised: a second thread finds the first one's filter
    Shared& e = g_cache[*key];
    if (!e.filter) {
        if (jinc_filter_create(&p, &e.filter) != JINC_OK) {
            *err = jinc_last_error();
            g_cache.erase(*key);
            return nullptr;
        }
    }
    ++e.refs;
    return e.filter;
}

void cache_release(const std::string& key)
{
    jinc_filter* dead = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_cache.find(key);
        if (it == g_cache.end())
            return;
        if (--it->second.refs == 0) {
            dead = it->second.filter;
            g_cache.erase(it);
        }
    }
    jinc_filter_destroy(dead);
}

// positions inside the JincResize argument array (src/JincResize.cpp:656-674)
enum Arg {
    A_CLIP, A_TARGET_W, A_TARGET_H, A_SRC_LEFT, A_SRC_TOP, A_SRC_WIDTH, A_SRC_HEIGHT, A_QUANT_X, A_QUANT_Y, A_TAP, A_BLUR,
    A_CPLACE, A_THREADS, A_OPT, A_INITIAL_CAPACITY, A_INITIAL_FACTOR
};

const int kPlanesYUV[4] = {AVS_PLANAR_Y, AVS_PLANAR_U, AVS_PLANAR_V, AVS_PLANAR_A};
const int kPlanesRGB[4] = {AVS_PLANAR_G, AVS_PLANAR_B, AVS_PLANAR_R, AVS_PLANAR_A};

AVS_VideoFrame* AVSC_CC get_frame(AVS_FilterInfo* fi, int n)
{
    auto* inst = static_cast<Instance*>(fi->user_data);
    AVS_VideoFrame* src = avs_get_frame(fi->child, n);
    if (!src)
        return nullptr;
    AVS_VideoFrame* dst = avs_new_video_frame_p(fi->env, &fi->vi, src); // properties copied from src

    jinc_frame fr;
    memset(&fr, 0, sizeof(fr));
    const int* ids = inst->rgb ? kPlanesRGB : kPlanesYUV;
    for (int i = 0; i < inst->n_planes; ++i) {
        fr.src[i] = avs_get_read_ptr_p(src, ids[i]);
        fr.src_pitch[i] = avs_get_pitch_p(src, ids[i]);
        fr.dst[i] = avs_get_write_ptr_p(dst, ids[i]);
        fr.dst_pitch[i] = avs_get_pitch_p(dst, ids[i]);
    }
    // whole frames by default (frames overlap each other across Prefetch threads); row bands overlap the transfers and
    // kernels of ONE frame, and spread it over the filter's GPUs, which is what a host without Prefetch wants
    const int rc = inst->bands > 0 ? jinc_filter_process_bands(inst->filter, &fr, inst->bands) : jinc_filter_process(inst->filter, &fr);
    if (rc != JINC_OK) {
        // the text lives in the environment's string heap: concurrent failing callers never share storage
        const std::string msg = std::string("JincResize: ") + jinc_last_error();
        fi->error = avs_save_string(fi->env, msg.c_str(), -1);
        avs_release_video_frame(src);
        avs_release_video_frame(dst);
        return nullptr;
    }
    if (inst->writes_chromaloc)
        avs_prop_set_int(fi->env, avs_get_frame_props_rw(fi->env, dst), "_ChromaLocation", inst->chromaloc, 0);
    avs_release_video_frame(src);
    return dst;
}

void AVSC_CC free_filter(AVS_FilterInfo* fi)
{
    auto* inst = static_cast<Instance*>(fi->user_data);
    if (inst) {
        cache_release(inst->key);
        delete inst;
    }
    fi->user_data = nullptr;
}

int AVSC_CC set_cache_hints(AVS_FilterInfo*, int cachehints, int)
{
    if (cachehints != AVS_CACHE_GET_MTMODE)
        return 0;
    const char* m = getenv("JINCRESIZE_B200_MTMODE");
    return (m && *m == '2') ? AVS_MT_MULTI_INSTANCE : AVS_MT_NICE_FILTER;
}

AVS_Value fail(AVS_Clip* clip, const char* msg)
{
    avs_release_clip(clip);
    return avs_new_value_error(msg);
}

AVS_Value AVSC_CC create_jincresize(AVS_ScriptEnvironment* env, AVS_Value args, void*)
{
    AVS_FilterInfo* fi = nullptr;
    AVS_Clip* clip = avs_new_c_filter(env, &fi, avs_array_elt(args, A_CLIP), 1);
    AVS_VideoInfo* vi = &fi->vi;
    auto arg = [&](int i) { return avs_array_elt(args, i); };
    auto given = [&](int i) { return avs_defined(avs_array_elt(args, i)) != 0; };

    // host interface gate: v10+, or v9 with bug-fix level >= 2 (:689-698)
    const char* too_old = "JincResize: AviSynth+ version must be r3688 or later.";
    if (avs_check_version(env, 9) != 0)
        return fail(clip, too_old);
    if (avs_check_version(env, 10) != 0 && avs_get_env_property(env, AVS_AEP_INTERFACE_BUGFIX) < 2)
        return fail(clip, too_old);

    if (!avs_is_planar(vi))
        return fail(clip, "JincResize: clip must be in planar format.");

    const int tap = given(A_TAP) ? avs_as_int(arg(A_TAP)) : 3;
    if (tap < 1 || tap > 16)
        return fail(clip, "JincResize: tap must be between 1..16.");
    const int quant_x = given(A_QUANT_X) ? avs_as_int(arg(A_QUANT_X)) : 256;
    if (quant_x < 1 || quant_x > 256)
        return fail(clip, "JincResize: quant_x must be between 1..256.");
    const int quant_y = given(A_QUANT_Y) ? avs_as_int(arg(A_QUANT_Y)) : 256;
    if (quant_y < 1 || quant_y > 256)
        return fail(clip, "JincResize: quant_y must be between 1..256.");

    std::string cplace = given(A_CPLACE) ? avs_as_string(arg(A_CPLACE)) : "";
    if (!cplace.empty()) {
        std::transform(cplace.begin(), cplace.end(), cplace.begin(), [](unsigned char c) { return (char)std::tolower(c); });
        if (cplace != "mpeg2" && cplace != "mpeg1" && cplace != "topleft")
            return fail(clip, "JincResize: cplace must be MPEG2, MPEG1 or topleft.");
    } else {
        // default from the first frame's _ChromaLocation, else MPEG2 (:725-742)
        cplace = "mpeg2";
        AVS_VideoFrame* frame0 = avs_get_frame(clip, 0);
        if (frame0) {
            const AVS_Map* props = avs_get_frame_props_ro(env, frame0);
            long long loc = -1;
            bool has = false;
            if (props && avs_prop_get_type(env, props, "_ChromaLocation") == 'i') {
                loc = avs_prop_get_int(env, props, "_ChromaLocation", 0, nullptr);
                has = true;
            }
            avs_release_video_frame(frame0); // (the reference leaks this frame)
            if (has) {
                if (loc == 0)
                    cplace = "mpeg2";
                else if (loc == 1)
                    cplace = "mpeg1";
                else if (loc == 2)
                    cplace = "topleft";
                else
                    return fail(clip, "JincResize: invalid _ChromaLocation");
            }
        }
    }
    if (cplace == "topleft" && !avs_is_420(vi))
        return fail(clip, "JincResize: topleft must be used only for 4:2:0 chroma subsampling.");

    const int opt = given(A_OPT) ? avs_as_int(arg(A_OPT)) : -1;
    const int cpu_flags = avs_get_cpu_flags(env);
    if (opt > 3)
        return fail(clip, "JincResize: opt higher than 3 is not allowed.");
    if (opt == 3 && !(cpu_flags & AVS_CPUF_AVX512F))
        return fail(clip, "JincResize: opt=3 requires AVX-512F.");
    if (opt == 2 && !(cpu_flags & AVS_CPUF_AVX2))
        return fail(clip, "JincResize: opt=2 requires AVX2.");
    if (opt == 1 && !(cpu_flags & AVS_CPUF_SSE4_1))
        return fail(clip, "JincResize: opt=1 requires SSE4.1.");
    const int threads = given(A_THREADS) ? avs_as_int(arg(A_THREADS)) : 0;
    if (threads < 0 || threads > 1)
        return fail(clip, "JincResize: threads must be either 0 or 1.");

    // script floats are 32-bit: avs_as_float widens a float (:762-770)
    const double src_left = given(A_SRC_LEFT) ? avs_as_float(arg(A_SRC_LEFT)) : 0.0;
    const double src_top = given(A_SRC_TOP) ? avs_as_float(arg(A_SRC_TOP)) : 0.0;
    const double src_width = given(A_SRC_WIDTH) ? avs_as_float(arg(A_SRC_WIDTH)) : static_cast<double>(vi->width);
    const double src_height = given(A_SRC_HEIGHT) ? avs_as_float(arg(A_SRC_HEIGHT)) : static_cast<double>(vi->height);
    const double blur = given(A_BLUR) ? avs_as_float(arg(A_BLUR)) : 0.0; // 0 => 1.0 inside the library (:772-774)

    const int target_w = avs_as_int(arg(A_TARGET_W));
    const int target_h = avs_as_int(arg(A_TARGET_H));

    const double initial_factor = given(A_INITIAL_FACTOR) ? avs_as_float(arg(A_INITIAL_FACTOR)) : 1.5;
    if (initial_factor < 1.0)
        return fail(clip, "JincResize: initial_factor must be eqaul to or greater than 1.0."); // sic, as the reference
    const int initial_capacity = given(A_INITIAL_CAPACITY) ? avs_as_int(arg(A_INITIAL_CAPACITY))
                                                           : std::max(target_w * target_h, vi->width * vi->height);
    if (initial_capacity <= 0)
        return fail(clip, "JincResize: initial_capacity must be greater than 0.");

    jinc_filter_params p;
    memset(&p, 0, sizeof(p));
    p.src_w = vi->width;
    p.src_h = vi->height;
    p.target_w = target_w;
    p.target_h = target_h;
    p.src_left = src_left;
    p.src_top = src_top;
    p.src_width = src_width;
    p.src_height = src_height;
    p.quant_x = quant_x;
    p.quant_y = quant_y;
    p.tap = tap;
    p.blur = blur;
    p.cplace = cplace == "mpeg2" ? JINC_CPLACE_MPEG2 : (cplace == "mpeg1" ? JINC_CPLACE_MPEG1 : JINC_CPLACE_TOPLEFT);
    p.n_planes = avs_num_components(vi);
    p.sample_bytes = avs_component_size(vi);
    p.bits = avs_bits_per_component(vi);
    const bool one_table = p.n_planes == 1 || avs_is_444(vi) || avs_is_rgb(vi); // :824-827
    p.sub_w = one_table ? 0 : avs_get_plane_width_subsampling(vi, AVS_PLANAR_U);
    p.sub_h = one_table ? 0 : avs_get_plane_height_subsampling(vi, AVS_PLANAR_U);
    // GPUs: device 0 unless JINCRESIZE_B200_DEVICES names others ("0,1,..." or "all"): every GPU of a filter gets its own
    // tables, streams and pinned frame slots, which a script with several JincResize calls should not pay eightfold
    const char* devs = getenv("JINCRESIZE_B200_DEVICES");
    if (!devs || !*devs) {
        p.devices[p.n_devices++] = 0;
    } else if (strcmp(devs, "all") != 0) {
        const char* s = devs;
        while (*s && p.n_devices < JINC_MAX_DEVICES) {
            char* end = nullptr;
            const long v = strtol(s, &end, 10);
            if (end == s)
                break;
            p.devices[p.n_devices++] = static_cast<int>(v);
            s = (*end == ',') ? end + 1 : end;
        }
    }
    if (const char* slots = getenv("JINCRESIZE_B200_SLOTS"))
        p.slots_per_device = atoi(slots);
    // AviSynth+ frame buffers: the padding inside a plane's pitch belongs to the frame; recycled buffers may be page-locked
    p.flags = JINC_FILTER_DST_PADDING_WRITABLE | JINC_FILTER_HOST_REGISTER;
    if (const char* hr = getenv("JINCRESIZE_B200_HOSTREG"))
        if (*hr == '0')
            p.flags &= ~JINC_FILTER_HOST_REGISTER;

    const bool subsampled_family = avs_is_420(vi) || avs_is_422(vi) || avs_is_yv411(vi);

    std::string key, msg;
    jinc_filter* filter = cache_acquire(p, &key, &msg);
    if (!filter) {
        if (msg.rfind("JincResize:", 0) != 0)
            msg = "JincResize: " + msg;
        return fail(clip, avs_save_string(env, msg.c_str(), -1));
    }

    auto* inst = new Instance();
    inst->filter = filter;
    inst->key = key;
    if (const char* b = getenv("JINCRESIZE_B200_BANDS"))
        inst->bands = std::max(0, std::min(atoi(b), 256));
    inst->n_planes = p.n_planes;
    inst->rgb = avs_is_rgb(vi) != 0;
    inst->writes_chromaloc = subsampled_family;
    // the reference always writes 2 (see the header comment); "actual" writes the cplace that was used
    const char* cl = getenv("JINCRESIZE_B200_CHROMALOC");
    const char* old = getenv("JINCRESIZE_B200_COMPAT_CHROMALOC"); // round-1 name of the switch: 0 = actual
    const bool actual = (cl && strcmp(cl, "actual") == 0) || (!cl && old && *old == '0');
    inst->chromaloc = actual ? p.cplace : 2;

    vi->width = target_w;
    vi->height = target_h;
    fi->user_data = inst;
    fi->get_frame = get_frame;
    fi->set_cache_hints = set_cache_hints;
    fi->free_filter = free_filter;

    AVS_Value v = avs_new_value_clip(clip);
    avs_release_clip(clip);
    return v;
}

// JincNNResize(clip, w, h, ...) == JincResize(clip, w, h, ..., tap=N): forward what was given, by name (:1007-1040)
template <int TAP>
AVS_Value AVSC_CC create_alias(AVS_ScriptEnvironment* env, AVS_Value args, void*)
{
    static const char* const kNames[8] = {"src_left", "src_top", "src_width", "src_height", "quant_x", "quant_y", "cplace", "threads"};
    AVS_Value values[12];
    const char* names[12];
    int n = 0;
    for (int i = 0; i < 3; ++i) {
        values[n] = avs_array_elt(args, i);
        names[n++] = nullptr;
    }
    for (int k = 0; k < 8; ++k) {
        const AVS_Value v = avs_array_elt(args, 3 + k);
        if (avs_defined(v)) {
            values[n] = v;
            names[n++] = kNames[k];
        }
    }
    values[n] = avs_new_value_int(TAP);
    names[n++] = "tap";
    return avs_invoke(env, "JincResize", avs_new_value_array(values, n), names);
}

#define JINC_COMMON_PARAMS "cii[src_left]f[src_top]f[src_width]f[src_height]f[quant_x]i[quant_y]i"

} // namespace

const char* AVSC_CC avisynth_c_plugin_init(AVS_ScriptEnvironment* env)
{
    avs_add_function(env, "JincResize",
                     JINC_COMMON_PARAMS "[tap]i[blur]f[cplace]s[threads]i[opt]i[initial_capacity]i[initial_factor]f",
                     create_jincresize, nullptr);
    avs_add_function(env, "Jinc36Resize", JINC_COMMON_PARAMS "[cplace]s[threads]i", create_alias<3>, nullptr);
    avs_add_function(env, "Jinc64Resize", JINC_COMMON_PARAMS "[cplace]s[threads]i", create_alias<4>, nullptr);
    avs_add_function(env, "Jinc144Resize", JINC_COMMON_PARAMS "[cplace]s[threads]i", create_alias<6>, nullptr);
    avs_add_function(env, "Jinc256Resize", JINC_COMMON_PARAMS "[cplace]s[threads]i", create_alias<8>, nullptr);
    return "JincResize";
}
