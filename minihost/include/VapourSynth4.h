/*
 * VapourSynth4.h -- CLEAN-ROOM restatement of the subset of the VapourSynth API 4 that a video filter plugin uses.
 *
 * VapourSynth's own header is not on this machine and cannot be fetched (no network).  This file restates, from the
 * published API documentation, the types, enumerations and function tables (names, argument lists and ORDER of the
 * entries) that the VapourSynth front-end in avisynth-jincresize_b200/vapoursynth/ needs, so that the plugin can be
 * compiled and driven by the in-process stand-in host (minihost/vs_minihost.cpp) in tests.
 *
 * It is TEST/BUILD INFRASTRUCTURE.  A plugin that is to be loaded by a real VapourSynth must be compiled against the
 * real VapourSynth4.h (make vs VSINC=/path/to/vapoursynth/include): the layout of struct VSAPI below follows the
 * documented order but has not been compared with the genuine header.
 */
#ifndef VAPOURSYNTH4_H
#define VAPOURSYNTH4_H

#include <stddef.h>
#include <stdint.h>

#define VS_MAKE_VERSION(major, minor) (((major) << 16) | (minor))
#define VAPOURSYNTH_API_MAJOR 4
#define VAPOURSYNTH_API_MINOR 0
#define VAPOURSYNTH_API_VERSION VS_MAKE_VERSION(VAPOURSYNTH_API_MAJOR, VAPOURSYNTH_API_MINOR)

#ifdef __cplusplus
#define VS_EXTERN_C extern "C"
#define VS_NOEXCEPT noexcept
#else
#define VS_EXTERN_C
#define VS_NOEXCEPT
#endif
#define VS_CC
#define VS_EXTERNAL_API(ret) VS_EXTERN_C __attribute__((visibility("default"))) ret VS_CC

typedef struct VSFrame VSFrame;
typedef struct VSNode VSNode;
typedef struct VSCore VSCore;
typedef struct VSPlugin VSPlugin;
typedef struct VSPluginFunction VSPluginFunction;
typedef struct VSFunction VSFunction;
typedef struct VSMap VSMap;
typedef struct VSLogHandle VSLogHandle;
typedef struct VSFrameContext VSFrameContext;
typedef struct VSPLUGINAPI VSPLUGINAPI;
typedef struct VSAPI VSAPI;

typedef enum VSColorFamily { cfUndefined = 0, cfGray = 1, cfRGB = 2, cfYUV = 3 } VSColorFamily;
typedef enum VSSampleType { stInteger = 0, stFloat = 1 } VSSampleType;
typedef enum VSFilterMode { fmParallel = 0, fmParallelRequests = 1, fmUnordered = 2, fmFrameState = 3 } VSFilterMode;
typedef enum VSMediaType { mtVideo = 1, mtAudio = 2 } VSMediaType;
typedef enum VSActivationReason { arError = -1, arInitial = 0, arAllFramesReady = 1 } VSActivationReason;
typedef enum VSMapPropertyError { peSuccess = 0, peUnset = 1, peType = 2, peIndex = 4, peError = 3 } VSMapPropertyError;
typedef enum VSMapAppendMode { maReplace = 0, maAppend = 1 } VSMapAppendMode;
typedef enum VSRequestPattern { rpGeneral = 0, rpNoFrameReuse = 1, rpStrictSpatial = 2 } VSRequestPattern;
typedef enum VSPropertyType { ptUnset = 0, ptInt = 1, ptFloat = 2, ptData = 3, ptFunction = 4, ptVideoNode = 5, ptAudioNode = 6,
                              ptVideoFrame = 7, ptAudioFrame = 8 } VSPropertyType;
typedef enum VSDataTypeHint { dtUnknown = -1, dtBinary = 0, dtUtf8 = 1 } VSDataTypeHint;

typedef struct VSVideoFormat {
    int colorFamily;
    int sampleType;
    int bitsPerSample;
    int bytesPerSample;
    int subSamplingW; /* log2 */
    int subSamplingH;
    int numPlanes;
} VSVideoFormat;

typedef struct VSVideoInfo {
    VSVideoFormat format;
    int64_t fpsNum;
    int64_t fpsDen;
    int width;
    int height;
    int numFrames;
} VSVideoInfo;

typedef struct VSAudioFormat VSAudioFormat;
typedef struct VSAudioInfo VSAudioInfo;
typedef struct VSCoreInfo VSCoreInfo;

typedef struct VSFilterDependency {
    VSNode* source;
    int requestPattern; /* VSRequestPattern */
} VSFilterDependency;

typedef const VSFrame*(VS_CC* VSFilterGetFrame)(int n, int activationReason, void* instanceData, void** frameData,
                                                 VSFrameContext* frameCtx, VSCore* core, const VSAPI* vsapi);
typedef void(VS_CC* VSFilterFree)(void* instanceData, VSCore* core, const VSAPI* vsapi);
typedef void(VS_CC* VSPublicFunction)(const VSMap* in, VSMap* out, void* userData, VSCore* core, const VSAPI* vsapi);
typedef void(VS_CC* VSFreeFunctionData)(void* userData);
typedef void(VS_CC* VSFrameDoneCallback)(void* userData, const VSFrame* f, int n, VSNode* node, const char* errorMsg);
typedef void(VS_CC* VSInitPlugin)(VSPlugin* plugin, const VSPLUGINAPI* vspapi);

struct VSPLUGINAPI {
    int(VS_CC* getAPIVersion)(void) VS_NOEXCEPT;
    int(VS_CC* configPlugin)(const char* identifier, const char* pluginNamespace, const char* name, int pluginVersion, int apiVersion,
                             int flags, VSPlugin* plugin) VS_NOEXCEPT;
    int(VS_CC* registerFunction)(const char* name, const char* args, const char* returnType, VSPublicFunction argsFunc, void* functionData,
                                 VSPlugin* plugin) VS_NOEXCEPT;
};

struct VSAPI {
    /* video and audio filters, nodes */
    void(VS_CC* createVideoFilter)(VSMap* out, const char* name, const VSVideoInfo* vi, VSFilterGetFrame getFrame, VSFilterFree free,
                                   int filterMode, const VSFilterDependency* dependencies, int numDeps, void* instanceData,
                                   VSCore* core) VS_NOEXCEPT;
    VSNode*(VS_CC* createVideoFilter2)(const char* name, const VSVideoInfo* vi, VSFilterGetFrame getFrame, VSFilterFree free, int filterMode,
                                       const VSFilterDependency* dependencies, int numDeps, void* instanceData, VSCore* core) VS_NOEXCEPT;
    void(VS_CC* createAudioFilter)(VSMap* out, const char* name, const VSAudioInfo* ai, VSFilterGetFrame getFrame, VSFilterFree free,
                                   int filterMode, const VSFilterDependency* dependencies, int numDeps, void* instanceData,
                                   VSCore* core) VS_NOEXCEPT;
    VSNode*(VS_CC* createAudioFilter2)(const char* name, const VSAudioInfo* ai, VSFilterGetFrame getFrame, VSFilterFree free, int filterMode,
                                       const VSFilterDependency* dependencies, int numDeps, void* instanceData, VSCore* core) VS_NOEXCEPT;
    int(VS_CC* setLinearFilter)(VSNode* node) VS_NOEXCEPT;
    void(VS_CC* setCacheMode)(VSNode* node, int mode) VS_NOEXCEPT;
    void(VS_CC* setCacheOptions)(VSNode* node, int fixedSize, int maxSize, int maxHistorySize) VS_NOEXCEPT;
    void(VS_CC* freeNode)(VSNode* node) VS_NOEXCEPT;
    VSNode*(VS_CC* addNodeRef)(VSNode* node) VS_NOEXCEPT;
    int(VS_CC* getNodeType)(VSNode* node) VS_NOEXCEPT;
    const VSVideoInfo*(VS_CC* getVideoInfo)(VSNode* node) VS_NOEXCEPT;
    const VSAudioInfo*(VS_CC* getAudioInfo)(VSNode* node) VS_NOEXCEPT;

    /* frames */
    VSFrame*(VS_CC* newVideoFrame)(const VSVideoFormat* format, int width, int height, const VSFrame* propSrc, VSCore* core) VS_NOEXCEPT;
    VSFrame*(VS_CC* newVideoFrame2)(const VSVideoFormat* format, int width, int height, const VSFrame** planeSrc, const int* planes,
                                    const VSFrame* propSrc, VSCore* core) VS_NOEXCEPT;
    VSFrame*(VS_CC* newAudioFrame)(const VSAudioFormat* format, int numSamples, const VSFrame* propSrc, VSCore* core) VS_NOEXCEPT;
    VSFrame*(VS_CC* newAudioFrame2)(const VSAudioFormat* format, int numSamples, const VSFrame** channelSrc, const int* channels,
                                    const VSFrame* propSrc, VSCore* core) VS_NOEXCEPT;
    void(VS_CC* freeFrame)(const VSFrame* f) VS_NOEXCEPT;
    const VSFrame*(VS_CC* addFrameRef)(const VSFrame* f) VS_NOEXCEPT;
    VSFrame*(VS_CC* copyFrame)(const VSFrame* f, VSCore* core) VS_NOEXCEPT;
    const VSMap*(VS_CC* getFramePropertiesRO)(const VSFrame* f) VS_NOEXCEPT;
    VSMap*(VS_CC* getFramePropertiesRW)(VSFrame* f) VS_NOEXCEPT;
    ptrdiff_t(VS_CC* getStride)(const VSFrame* f, int plane) VS_NOEXCEPT;
    const uint8_t*(VS_CC* getReadPtr)(const VSFrame* f, int plane) VS_NOEXCEPT;
    uint8_t*(VS_CC* getWritePtr)(VSFrame* f, int plane) VS_NOEXCEPT;
    const VSVideoFormat*(VS_CC* getVideoFrameFormat)(const VSFrame* f) VS_NOEXCEPT;
    const VSAudioFormat*(VS_CC* getAudioFrameFormat)(const VSFrame* f) VS_NOEXCEPT;
    int(VS_CC* getFrameType)(const VSFrame* f) VS_NOEXCEPT;
    int(VS_CC* getFrameWidth)(const VSFrame* f, int plane) VS_NOEXCEPT;
    int(VS_CC* getFrameHeight)(const VSFrame* f, int plane) VS_NOEXCEPT;
    int(VS_CC* getFrameLength)(const VSFrame* f) VS_NOEXCEPT;

    /* formats */
    int(VS_CC* getVideoFormatName)(const VSVideoFormat* format, char* buffer) VS_NOEXCEPT;
    int(VS_CC* getAudioFormatName)(const VSAudioFormat* format, char* buffer) VS_NOEXCEPT;
    int(VS_CC* queryVideoFormat)(VSVideoFormat* format, int colorFamily, int sampleType, int bitsPerSample, int subSamplingW,
                                 int subSamplingH, VSCore* core) VS_NOEXCEPT;
    int(VS_CC* queryAudioFormat)(VSAudioFormat* format, int sampleType, int bitsPerSample, uint64_t channelLayout, VSCore* core) VS_NOEXCEPT;
    uint32_t(VS_CC* queryVideoFormatID)(int colorFamily, int sampleType, int bitsPerSample, int subSamplingW, int subSamplingH,
                                        VSCore* core) VS_NOEXCEPT;
    int(VS_CC* getVideoFormatByID)(VSVideoFormat* format, uint32_t id, VSCore* core) VS_NOEXCEPT;

    /* frame requests */
    const VSFrame*(VS_CC* getFrame)(int n, VSNode* node, char* errorMsg, int bufSize) VS_NOEXCEPT;
    void(VS_CC* getFrameAsync)(int n, VSNode* node, VSFrameDoneCallback callback, void* userData) VS_NOEXCEPT;
    const VSFrame*(VS_CC* getFrameFilter)(int n, VSNode* node, VSFrameContext* frameCtx) VS_NOEXCEPT;
    void(VS_CC* requestFrameFilter)(int n, VSNode* node, VSFrameContext* frameCtx) VS_NOEXCEPT;
    void(VS_CC* releaseFrameEarly)(VSNode* node, int n, VSFrameContext* frameCtx) VS_NOEXCEPT;
    void(VS_CC* cacheFrame)(const VSFrame* frame, int n, VSFrameContext* frameCtx) VS_NOEXCEPT;
    void(VS_CC* setFilterError)(const char* errorMessage, VSFrameContext* frameCtx) VS_NOEXCEPT;

    /* external functions */
    VSFunction*(VS_CC* createFunction)(VSPublicFunction func, void* userData, VSFreeFunctionData free, VSCore* core) VS_NOEXCEPT;
    void(VS_CC* freeFunction)(VSFunction* f) VS_NOEXCEPT;
    VSFunction*(VS_CC* addFunctionRef)(VSFunction* f) VS_NOEXCEPT;
    void(VS_CC* callFunction)(VSFunction* func, const VSMap* in, VSMap* out) VS_NOEXCEPT;

    /* maps */
    VSMap*(VS_CC* createMap)(void) VS_NOEXCEPT;
    void(VS_CC* freeMap)(VSMap* map) VS_NOEXCEPT;
    void(VS_CC* clearMap)(VSMap* map) VS_NOEXCEPT;
    void(VS_CC* copyMap)(const VSMap* src, VSMap* dst) VS_NOEXCEPT;
    void(VS_CC* mapSetError)(VSMap* map, const char* errorMessage) VS_NOEXCEPT;
    const char*(VS_CC* mapGetError)(const VSMap* map) VS_NOEXCEPT;
    int(VS_CC* mapNumKeys)(const VSMap* map) VS_NOEXCEPT;
    const char*(VS_CC* mapGetKey)(const VSMap* map, int index) VS_NOEXCEPT;
    int(VS_CC* mapDeleteKey)(VSMap* map, const char* key) VS_NOEXCEPT;
    int(VS_CC* mapNumElements)(const VSMap* map, const char* key) VS_NOEXCEPT;
    int(VS_CC* mapGetType)(const VSMap* map, const char* key) VS_NOEXCEPT;
    int(VS_CC* mapSetEmpty)(VSMap* map, const char* key, int type) VS_NOEXCEPT;
    int64_t(VS_CC* mapGetInt)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    int(VS_CC* mapGetIntSaturated)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    const int64_t*(VS_CC* mapGetIntArray)(const VSMap* map, const char* key, int* error) VS_NOEXCEPT;
    int(VS_CC* mapSetInt)(VSMap* map, const char* key, int64_t i, int append) VS_NOEXCEPT;
    int(VS_CC* mapSetIntArray)(VSMap* map, const char* key, const int64_t* i, int size) VS_NOEXCEPT;
    double(VS_CC* mapGetFloat)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    float(VS_CC* mapGetFloatSaturated)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    const double*(VS_CC* mapGetFloatArray)(const VSMap* map, const char* key, int* error) VS_NOEXCEPT;
    int(VS_CC* mapSetFloat)(VSMap* map, const char* key, double d, int append) VS_NOEXCEPT;
    int(VS_CC* mapSetFloatArray)(VSMap* map, const char* key, const double* d, int size) VS_NOEXCEPT;
    const char*(VS_CC* mapGetData)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    int(VS_CC* mapGetDataSize)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    int(VS_CC* mapGetDataTypeHint)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    int(VS_CC* mapSetData)(VSMap* map, const char* key, const char* data, int size, int type, int append) VS_NOEXCEPT;
    VSNode*(VS_CC* mapGetNode)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    int(VS_CC* mapSetNode)(VSMap* map, const char* key, VSNode* node, int append) VS_NOEXCEPT;
    int(VS_CC* mapConsumeNode)(VSMap* map, const char* key, VSNode* node, int append) VS_NOEXCEPT;
    const VSFrame*(VS_CC* mapGetFrame)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    int(VS_CC* mapSetFrame)(VSMap* map, const char* key, const VSFrame* f, int append) VS_NOEXCEPT;
    int(VS_CC* mapConsumeFrame)(VSMap* map, const char* key, const VSFrame* f, int append) VS_NOEXCEPT;
    VSFunction*(VS_CC* mapGetFunction)(const VSMap* map, const char* key, int index, int* error) VS_NOEXCEPT;
    int(VS_CC* mapSetFunction)(VSMap* map, const char* key, VSFunction* func, int append) VS_NOEXCEPT;
    int(VS_CC* mapConsumeFunction)(VSMap* map, const char* key, VSFunction* func, int append) VS_NOEXCEPT;
    /* (plugin lookup, core and logging entries follow in the real header; a filter plugin does not call them) */
};

#endif /* VAPOURSYNTH4_H */
