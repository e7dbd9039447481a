/*
 * avisynth_c.h -- clean-room subset of the AviSynth+ C API (interface v9/v10).
 *
 * The genuine AviSynth+ SDK is not available offline, so this header restates,
 * from public API knowledge, exactly the part of the C interface that an EWA
 * resizer plugin touches (SURVEY.md section 8b lists the symbols).  Struct
 * layouts, constant values and calling conventions follow AviSynth+ 3.7 so a
 * plugin compiled against this header has the same shape as one compiled
 * against the real SDK; all avs_* entry points are resolved at load time from
 * the host (here: libavs_minihost.so, minihost/minihost.cpp).
 *
 * Both the unmodified reference translation units (oracle/_ref) and the
 * B200 plugin are compiled against THIS header, so the two sides of every
 * parity test see the same host.
 */
#ifndef MINIHOST_AVISYNTH_C_H
#define MINIHOST_AVISYNTH_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#  define AVSC_EXTERN_C extern "C"
#else
#  define AVSC_EXTERN_C
#endif

#if defined(_WIN32) && !defined(_WIN64)
#  define AVSC_CC __stdcall
#else
#  define AVSC_CC
#endif

#if defined(_MSC_VER)
#  define AVS_FORCEINLINE __forceinline
#else
#  define AVS_FORCEINLINE inline __attribute__((always_inline))
#endif

#define AVSC_INLINE static inline
#define AVSC_API(ret, name) AVSC_EXTERN_C __attribute__((visibility("default"))) ret AVSC_CC name
#define AVSC_EXPORT AVSC_EXTERN_C __attribute__((visibility("default")))

typedef unsigned char BYTE;

/* ------------------------------------------------------------------ constants */

enum { AVISYNTH_INTERFACE_VERSION = 10, AVISYNTHPLUS_INTERFACE_BUGFIX_VERSION = 0 };

/* plane selectors */
enum {
    AVS_PLANAR_Y = 1 << 0,
    AVS_PLANAR_U = 1 << 1,
    AVS_PLANAR_V = 1 << 2,
    AVS_PLANAR_ALIGNED = 1 << 3,
    AVS_PLANAR_A = 1 << 4,
    AVS_PLANAR_R = 1 << 5,
    AVS_PLANAR_G = 1 << 6,
    AVS_PLANAR_B = 1 << 7
};

/* pixel_type bit fields */
enum {
    AVS_CS_YUVA = 1 << 27,
    AVS_CS_BGR = 1 << 28,
    AVS_CS_YUV = 1 << 29,
    AVS_CS_INTERLEAVED = 1 << 30,
    AVS_CS_PLANAR = (int)(1u << 31),

    AVS_CS_SHIFT_SUB_WIDTH = 0,
    AVS_CS_SHIFT_SUB_HEIGHT = 8,
    AVS_CS_SHIFT_SAMPLE_BITS = 16,

    AVS_CS_SUB_WIDTH_MASK = 7 << AVS_CS_SHIFT_SUB_WIDTH,
    AVS_CS_SUB_WIDTH_1 = 3 << AVS_CS_SHIFT_SUB_WIDTH, /* 4:4:4 */
    AVS_CS_SUB_WIDTH_2 = 0 << AVS_CS_SHIFT_SUB_WIDTH, /* 4:2:x */
    AVS_CS_SUB_WIDTH_4 = 1 << AVS_CS_SHIFT_SUB_WIDTH, /* 4:1:1 */

    AVS_CS_VPLANEFIRST = 1 << 3,
    AVS_CS_UPLANEFIRST = 1 << 4,

    AVS_CS_SUB_HEIGHT_MASK = 7 << AVS_CS_SHIFT_SUB_HEIGHT,
    AVS_CS_SUB_HEIGHT_1 = 3 << AVS_CS_SHIFT_SUB_HEIGHT,
    AVS_CS_SUB_HEIGHT_2 = 0 << AVS_CS_SHIFT_SUB_HEIGHT,
    AVS_CS_SUB_HEIGHT_4 = 1 << AVS_CS_SHIFT_SUB_HEIGHT,

    AVS_CS_SAMPLE_BITS_MASK = 7 << AVS_CS_SHIFT_SAMPLE_BITS,
    AVS_CS_SAMPLE_BITS_8 = 0 << AVS_CS_SHIFT_SAMPLE_BITS,
    AVS_CS_SAMPLE_BITS_10 = 5 << AVS_CS_SHIFT_SAMPLE_BITS,
    AVS_CS_SAMPLE_BITS_12 = 6 << AVS_CS_SHIFT_SAMPLE_BITS,
    AVS_CS_SAMPLE_BITS_14 = 7 << AVS_CS_SHIFT_SAMPLE_BITS,
    AVS_CS_SAMPLE_BITS_16 = 1 << AVS_CS_SHIFT_SAMPLE_BITS,
    AVS_CS_SAMPLE_BITS_32 = 2 << AVS_CS_SHIFT_SAMPLE_BITS,

    AVS_CS_PLANAR_MASK = AVS_CS_PLANAR | AVS_CS_INTERLEAVED | AVS_CS_YUV | AVS_CS_BGR | AVS_CS_YUVA |
                         AVS_CS_SAMPLE_BITS_MASK | AVS_CS_SUB_HEIGHT_MASK | AVS_CS_SUB_WIDTH_MASK,
    AVS_CS_PLANAR_FILTER = ~(AVS_CS_VPLANEFIRST | AVS_CS_UPLANEFIRST),

    AVS_CS_RGB_TYPE = 1 << 0,
    AVS_CS_RGBA_TYPE = 1 << 1,

    AVS_CS_GENERIC_YUV420 = AVS_CS_PLANAR | AVS_CS_YUV | AVS_CS_VPLANEFIRST | AVS_CS_SUB_HEIGHT_2 | AVS_CS_SUB_WIDTH_2,
    AVS_CS_GENERIC_YUV422 = AVS_CS_PLANAR | AVS_CS_YUV | AVS_CS_VPLANEFIRST | AVS_CS_SUB_HEIGHT_1 | AVS_CS_SUB_WIDTH_2,
    AVS_CS_GENERIC_YUV444 = AVS_CS_PLANAR | AVS_CS_YUV | AVS_CS_VPLANEFIRST | AVS_CS_SUB_HEIGHT_1 | AVS_CS_SUB_WIDTH_1,
    AVS_CS_GENERIC_Y = AVS_CS_PLANAR | AVS_CS_INTERLEAVED | AVS_CS_YUV,
    AVS_CS_GENERIC_RGBP = AVS_CS_PLANAR | AVS_CS_BGR | AVS_CS_RGB_TYPE,
    AVS_CS_GENERIC_RGBAP = AVS_CS_PLANAR | AVS_CS_BGR | AVS_CS_RGBA_TYPE,
    AVS_CS_GENERIC_YUVA420 = AVS_CS_PLANAR | AVS_CS_YUVA | AVS_CS_VPLANEFIRST | AVS_CS_SUB_HEIGHT_2 | AVS_CS_SUB_WIDTH_2,
    AVS_CS_GENERIC_YUVA422 = AVS_CS_PLANAR | AVS_CS_YUVA | AVS_CS_VPLANEFIRST | AVS_CS_SUB_HEIGHT_1 | AVS_CS_SUB_WIDTH_2,
    AVS_CS_GENERIC_YUVA444 = AVS_CS_PLANAR | AVS_CS_YUVA | AVS_CS_VPLANEFIRST | AVS_CS_SUB_HEIGHT_1 | AVS_CS_SUB_WIDTH_1,

    /* a few named formats */
    AVS_CS_BGR24 = AVS_CS_RGB_TYPE | AVS_CS_BGR | AVS_CS_INTERLEAVED,
    AVS_CS_BGR32 = AVS_CS_RGBA_TYPE | AVS_CS_BGR | AVS_CS_INTERLEAVED,
    AVS_CS_YUY2 = 1 << 2 | AVS_CS_YUV | AVS_CS_INTERLEAVED,
    AVS_CS_YV24 = AVS_CS_GENERIC_YUV444 | AVS_CS_SAMPLE_BITS_8,
    AVS_CS_YV16 = AVS_CS_GENERIC_YUV422 | AVS_CS_SAMPLE_BITS_8,
    AVS_CS_YV12 = AVS_CS_GENERIC_YUV420 | AVS_CS_SAMPLE_BITS_8,
    AVS_CS_YV411 = AVS_CS_PLANAR | AVS_CS_YUV | AVS_CS_VPLANEFIRST | AVS_CS_SUB_HEIGHT_1 | AVS_CS_SUB_WIDTH_4,
    AVS_CS_Y8 = AVS_CS_GENERIC_Y | AVS_CS_SAMPLE_BITS_8
};

/* cache hints */
enum { AVS_CACHE_GET_MTMODE = 509 };
enum { AVS_MT_NICE_FILTER = 1, AVS_MT_MULTI_INSTANCE = 2, AVS_MT_SERIALIZED = 3 };

/* avs_get_env_property selectors */
enum {
    AVS_AEP_PHYSICAL_CPUS = 1,
    AVS_AEP_LOGICAL_CPUS = 2,
    AVS_AEP_THREADPOOL_THREADS = 3,
    AVS_AEP_FILTERCHAIN_THREADS = 4,
    AVS_AEP_THREAD_ID = 5,
    AVS_AEP_VERSION = 6,
    AVS_AEP_HOST_SYSTEM_ENDIANNESS = 7,
    AVS_AEP_INTERFACE_VERSION = 8,
    AVS_AEP_INTERFACE_BUGFIX = 9
};

/* cpu flags */
enum {
    AVS_CPUF_SSE2 = 0x20,
    AVS_CPUF_SSE3 = 0x100,
    AVS_CPUF_SSSE3 = 0x200,
    AVS_CPUF_SSE4_1 = 0x400,
    AVS_CPUF_SSE4_2 = 0x800,
    AVS_CPUF_AVX = 0x1000,
    AVS_CPUF_AVX2 = 0x2000,
    AVS_CPUF_FMA3 = 0x4000,
    AVS_CPUF_AVX512F = 0x10000
};

/* ------------------------------------------------------------------ types */

typedef struct AVS_Clip AVS_Clip;
typedef struct AVS_ScriptEnvironment AVS_ScriptEnvironment;
typedef struct AVS_Map AVS_Map;

typedef struct AVS_VideoInfo {
    int width, height;
    unsigned fps_numerator, fps_denominator;
    int num_frames;
    int pixel_type;
    int audio_samples_per_second;
    int sample_type;
    int64_t num_audio_samples;
    int nchannels;
    int image_type;
} AVS_VideoInfo;

typedef struct AVS_VideoFrameBuffer {
    BYTE* data;
    int data_size;
    volatile long sequence_number;
    volatile long refcount;
    void* device_data;
} AVS_VideoFrameBuffer;

typedef struct AVS_VideoFrame {
    volatile long refcount;
    AVS_VideoFrameBuffer* vfb;
    int offset;
    int pitch, row_size, height;
    int offsetU, offsetV;
    int pitchUV;
    int row_sizeUV, heightUV;
    int offsetA;
    int pitchA, row_sizeA;
    void* properties;
} AVS_VideoFrame;

typedef struct AVS_Value AVS_Value;
struct AVS_Value {
    short type; /* 'a'rray 'c'lip 'b'ool 'i'nt 'f'loat 's'tring 'v'oid 'e'rror */
    short array_size;
    union {
        void* clip;
        char boolean;
        int integer;
        float floating_pt;
        const char* string;
        const AVS_Value* array;
    } d;
};

typedef struct AVS_FilterInfo AVS_FilterInfo;
struct AVS_FilterInfo {
    AVS_Clip* child;
    AVS_VideoInfo vi;
    AVS_ScriptEnvironment* env;
    AVS_VideoFrame*(AVSC_CC* get_frame)(AVS_FilterInfo*, int n);
    int(AVSC_CC* get_parity)(AVS_FilterInfo*, int n);
    int(AVSC_CC* get_audio)(AVS_FilterInfo*, void* buf, int64_t start, int64_t count);
    int(AVSC_CC* set_cache_hints)(AVS_FilterInfo*, int cachehints, int frame_range);
    void(AVSC_CC* free_filter)(AVS_FilterInfo*);
    const char* error;
    void* user_data;
};

typedef AVS_Value(AVSC_CC* AVS_ApplyFunc)(AVS_ScriptEnvironment*, AVS_Value args, void* user_data);

/* ------------------------------------------------------------------ AVS_Value helpers */

static const AVS_Value avs_void = {'v', 0, {0}};

AVSC_INLINE int avs_defined(AVS_Value v) { return v.type != 'v'; }
AVSC_INLINE int avs_is_clip(AVS_Value v) { return v.type == 'c'; }
AVSC_INLINE int avs_is_bool(AVS_Value v) { return v.type == 'b'; }
AVSC_INLINE int avs_is_int(AVS_Value v) { return v.type == 'i'; }
AVSC_INLINE int avs_is_float(AVS_Value v) { return v.type == 'f' || v.type == 'i'; }
AVSC_INLINE int avs_is_string(AVS_Value v) { return v.type == 's'; }
AVSC_INLINE int avs_is_array(AVS_Value v) { return v.type == 'a'; }
AVSC_INLINE int avs_is_error(AVS_Value v) { return v.type == 'e'; }

AVSC_INLINE int avs_as_bool(AVS_Value v) { return v.d.boolean; }
AVSC_INLINE int avs_as_int(AVS_Value v) { return v.d.integer; }
AVSC_INLINE const char* avs_as_string(AVS_Value v) { return avs_is_error(v) || avs_is_string(v) ? v.d.string : 0; }
/* script floats are 32-bit: the value widens from float, never from a double */
AVSC_INLINE double avs_as_float(AVS_Value v) { return avs_is_int(v) ? v.d.integer : v.d.floating_pt; }
AVSC_INLINE const char* avs_as_error(AVS_Value v) { return avs_is_error(v) ? v.d.string : 0; }
AVSC_INLINE const AVS_Value* avs_as_array(AVS_Value v) { return v.d.array; }
AVSC_INLINE int avs_array_size(AVS_Value v) { return avs_is_array(v) ? v.array_size : 1; }
AVSC_INLINE AVS_Value avs_array_elt(AVS_Value v, int index) { return avs_is_array(v) ? v.d.array[index] : v; }

AVSC_INLINE AVS_Value avs_new_value_bool(int v0) { AVS_Value v = {'b', 0, {0}}; v.d.boolean = v0 ? 1 : 0; return v; }
AVSC_INLINE AVS_Value avs_new_value_int(int v0) { AVS_Value v = {'i', 0, {0}}; v.d.integer = v0; return v; }
AVSC_INLINE AVS_Value avs_new_value_string(const char* v0) { AVS_Value v = {'s', 0, {0}}; v.d.string = v0; return v; }
AVSC_INLINE AVS_Value avs_new_value_float(float v0) { AVS_Value v = {'f', 0, {0}}; v.d.floating_pt = v0; return v; }
AVSC_INLINE AVS_Value avs_new_value_error(const char* v0) { AVS_Value v = {'e', 0, {0}}; v.d.string = v0; return v; }
AVSC_INLINE AVS_Value avs_new_value_array(AVS_Value* v0, int size) { AVS_Value v = {'a', 0, {0}}; v.d.array = v0; v.array_size = (short)size; return v; }

/* ------------------------------------------------------------------ AVS_VideoInfo helpers */

AVSC_INLINE int avs_is_rgb(const AVS_VideoInfo* p) { return !!(p->pixel_type & AVS_CS_BGR); }
AVSC_INLINE int avs_is_yuv(const AVS_VideoInfo* p) { return !!(p->pixel_type & AVS_CS_YUV); }
AVSC_INLINE int avs_is_yuva(const AVS_VideoInfo* p) { return !!(p->pixel_type & AVS_CS_YUVA); }
AVSC_INLINE int avs_is_planar(const AVS_VideoInfo* p) { return !!(p->pixel_type & AVS_CS_PLANAR); }

AVSC_API(int, avs_is_420)(const AVS_VideoInfo* p);
AVSC_API(int, avs_is_422)(const AVS_VideoInfo* p);
AVSC_API(int, avs_is_444)(const AVS_VideoInfo* p);
AVSC_API(int, avs_is_yv411)(const AVS_VideoInfo* p);
AVSC_API(int, avs_is_y)(const AVS_VideoInfo* p);
AVSC_API(int, avs_is_planar_rgb)(const AVS_VideoInfo* p);
AVSC_API(int, avs_is_planar_rgba)(const AVS_VideoInfo* p);
AVSC_API(int, avs_num_components)(const AVS_VideoInfo* p);
AVSC_API(int, avs_component_size)(const AVS_VideoInfo* p);
AVSC_API(int, avs_bits_per_component)(const AVS_VideoInfo* p);
AVSC_API(int, avs_get_plane_width_subsampling)(const AVS_VideoInfo* p, int plane);
AVSC_API(int, avs_get_plane_height_subsampling)(const AVS_VideoInfo* p, int plane);

/* ------------------------------------------------------------------ frame access */

AVSC_API(int, avs_get_pitch_p)(const AVS_VideoFrame* p, int plane);
AVSC_API(int, avs_get_row_size_p)(const AVS_VideoFrame* p, int plane);
AVSC_API(int, avs_get_height_p)(const AVS_VideoFrame* p, int plane);
AVSC_API(const BYTE*, avs_get_read_ptr_p)(const AVS_VideoFrame* p, int plane);
AVSC_API(BYTE*, avs_get_write_ptr_p)(const AVS_VideoFrame* p, int plane);
AVSC_API(int, avs_is_writable)(const AVS_VideoFrame* p);
AVSC_API(void, avs_release_video_frame)(AVS_VideoFrame*);
AVSC_API(AVS_VideoFrame*, avs_copy_video_frame)(AVS_VideoFrame*);

/* ------------------------------------------------------------------ clips */

AVSC_API(void, avs_release_clip)(AVS_Clip*);
AVSC_API(AVS_Clip*, avs_copy_clip)(AVS_Clip*);
AVSC_API(const char*, avs_clip_get_error)(AVS_Clip*);
AVSC_API(const AVS_VideoInfo*, avs_get_video_info)(AVS_Clip*);
AVSC_API(int, avs_get_version)(AVS_Clip*);
AVSC_API(AVS_VideoFrame*, avs_get_frame)(AVS_Clip*, int n);
AVSC_API(AVS_Clip*, avs_take_clip)(AVS_Value, AVS_ScriptEnvironment*);
AVSC_API(void, avs_set_to_clip)(AVS_Value*, AVS_Clip*);
AVSC_API(void, avs_release_value)(AVS_Value);
AVSC_API(void, avs_copy_value)(AVS_Value* dest, AVS_Value src);
AVSC_API(AVS_Clip*, avs_new_c_filter)(AVS_ScriptEnvironment* e, AVS_FilterInfo** fi, AVS_Value child, int store_child);

AVSC_INLINE AVS_Value avs_new_value_clip(AVS_Clip* v0)
{
    AVS_Value v;
    avs_set_to_clip(&v, v0);
    return v;
}

/* ------------------------------------------------------------------ environment */

AVSC_API(const char*, avs_get_error)(AVS_ScriptEnvironment*);
AVSC_API(int, avs_get_cpu_flags)(AVS_ScriptEnvironment*);
AVSC_API(int, avs_check_version)(AVS_ScriptEnvironment*, int version);
AVSC_API(size_t, avs_get_env_property)(AVS_ScriptEnvironment*, int prop);
AVSC_API(char*, avs_save_string)(AVS_ScriptEnvironment*, const char* s, int length);
AVSC_API(int, avs_add_function)(AVS_ScriptEnvironment*, const char* name, const char* params, AVS_ApplyFunc apply, void* user_data);
AVSC_API(int, avs_function_exists)(AVS_ScriptEnvironment*, const char* name);
AVSC_API(AVS_Value, avs_invoke)(AVS_ScriptEnvironment*, const char* name, AVS_Value args, const char** arg_names);
AVSC_API(AVS_VideoFrame*, avs_new_video_frame_a)(AVS_ScriptEnvironment*, const AVS_VideoInfo* vi, int align);
AVSC_API(AVS_VideoFrame*, avs_new_video_frame_p)(AVS_ScriptEnvironment*, const AVS_VideoInfo* vi, const AVS_VideoFrame* prop_src);
AVSC_API(int, avs_make_writable)(AVS_ScriptEnvironment*, AVS_VideoFrame** pvf);

AVSC_INLINE AVS_VideoFrame* avs_new_video_frame(AVS_ScriptEnvironment* env, const AVS_VideoInfo* vi)
{
    return avs_new_video_frame_a(env, vi, 64);
}

/* ------------------------------------------------------------------ frame properties (v8+) */

AVSC_API(const AVS_Map*, avs_get_frame_props_ro)(AVS_ScriptEnvironment*, const AVS_VideoFrame* frame);
AVSC_API(AVS_Map*, avs_get_frame_props_rw)(AVS_ScriptEnvironment*, AVS_VideoFrame* frame);
AVSC_API(int, avs_prop_num_keys)(AVS_ScriptEnvironment*, const AVS_Map* map);
AVSC_API(char, avs_prop_get_type)(AVS_ScriptEnvironment*, const AVS_Map* map, const char* key);
AVSC_API(int64_t, avs_prop_get_int)(AVS_ScriptEnvironment*, const AVS_Map* map, const char* key, int index, int* error);
AVSC_API(double, avs_prop_get_float)(AVS_ScriptEnvironment*, const AVS_Map* map, const char* key, int index, int* error);
AVSC_API(int, avs_prop_set_int)(AVS_ScriptEnvironment*, AVS_Map* map, const char* key, int64_t i, int append);
AVSC_API(int, avs_prop_set_float)(AVS_ScriptEnvironment*, AVS_Map* map, const char* key, double d, int append);
AVSC_API(int, avs_prop_delete_key)(AVS_ScriptEnvironment*, AVS_Map* map, const char* key);

/* ------------------------------------------------------------------ plugin entry */

/* every C plugin exports:  const char* AVSC_CC avisynth_c_plugin_init(AVS_ScriptEnvironment* env); */
AVSC_EXPORT const char* AVSC_CC avisynth_c_plugin_init(AVS_ScriptEnvironment* env);

#endif /* MINIHOST_AVISYNTH_C_H */
