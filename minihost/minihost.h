/*
 * minihost.h -- driver API of the in-process AviSynth+ C-API stand-in.
 *
 * libavs_minihost.so exports two things:
 *   1. the avs_* entry points declared in include/avisynth_c.h, so that C plugins
 *      (the unmodified reference build under oracle/_ref and the B200 plugin)
 *      can be dlopen()ed and driven exactly the way AviSynth+ drives them;
 *   2. the mh_* functions below -- a plain-C driver surface for tests and the
 *      bench harness (called from Python through ctypes, or from C++).
 *
 * This is test/bench infrastructure, not product code.
 */
#ifndef MINIHOST_H
#define MINIHOST_H

#include "avisynth_c.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MH_API __attribute__((visibility("default")))

/* --- environment ---------------------------------------------------------- */
MH_API AVS_ScriptEnvironment* mh_env_create(void);
MH_API void mh_env_destroy(AVS_ScriptEnvironment* env);
/* what avs_check_version / AVS_AEP_INTERFACE_BUGFIX report (default 10 / 0) */
MH_API void mh_env_set_interface(AVS_ScriptEnvironment* env, int version, int bugfix);
/* what avs_get_cpu_flags reports (default: detected from the host CPU) */
MH_API void mh_env_set_cpu_flags(AVS_ScriptEnvironment* env, int flags);
MH_API const char* mh_last_error(AVS_ScriptEnvironment* env);
/* dlopen(path) + avisynth_c_plugin_init(env); returns the plugin description or NULL */
MH_API const char* mh_load_plugin(AVS_ScriptEnvironment* env, const char* path);
/* parameter string a function was registered with, or NULL */
MH_API const char* mh_function_params(AVS_ScriptEnvironment* env, const char* name);

/* --- source clips ---------------------------------------------------------- */
/* A source clip serves `num_frames` frames that cycle over `distinct` stored
 * frames (frame n -> stored[n % distinct]); planes start zero-filled. */
MH_API AVS_Clip* mh_source_create(AVS_ScriptEnvironment* env, int width, int height, int pixel_type,
                                  int num_frames, int distinct);
/* copy rows into plane `plane` of stored frame `k` (src_pitch in bytes) */
MH_API int mh_source_fill_plane(AVS_Clip* clip, int k, int plane, const void* src, ptrdiff_t src_pitch);
/* set / clear an integer frame property on every stored frame */
MH_API void mh_source_set_prop_int(AVS_Clip* clip, const char* key, int64_t value);
MH_API void mh_source_clear_prop(AVS_Clip* clip, const char* key);

/* --- argument lists + invoke ----------------------------------------------- */
typedef struct mh_args mh_args;
MH_API mh_args* mh_args_create(void);
MH_API void mh_args_destroy(mh_args* a);
MH_API void mh_args_add_clip(mh_args* a, AVS_Clip* clip, const char* name);
MH_API void mh_args_add_int(mh_args* a, int v, const char* name);
MH_API void mh_args_add_float(mh_args* a, float v, const char* name);
MH_API void mh_args_add_string(mh_args* a, const char* s, const char* name);
/* Calls script function `name`; returns the resulting clip (caller releases with
 * avs_release_clip) or NULL, with the error text in mh_last_error(). */
MH_API AVS_Clip* mh_invoke_clip(AVS_ScriptEnvironment* env, const char* name, mh_args* a);

/* --- introspection ---------------------------------------------------------- */
/* AVS_FilterInfo of a clip created with avs_new_c_filter (NULL for source clips) */
MH_API AVS_FilterInfo* mh_clip_filter_info(AVS_Clip* clip);
MH_API int mh_clip_mt_mode(AVS_Clip* clip);
/* returns 1 and stores the value when the frame carries integer property `key` */
MH_API int mh_frame_prop_int(AVS_ScriptEnvironment* env, const AVS_VideoFrame* f, const char* key, int64_t* out);
/* number of live frames / clips (leak checks) */
MH_API long mh_live_frames(void);
MH_API long mh_live_clips(void);

/* --- timing ------------------------------------------------------------------ */
/* Pull frames [first, first+count) through avs_get_frame with `threads` host
 * threads (frame-parallel, what Prefetch(threads) does) and release them.
 * Returns wall seconds, or a negative value on error. */
MH_API double mh_pull_frames(AVS_Clip* clip, int first, int count, int threads);

#ifdef __cplusplus
}
#endif
#endif
