// fma_peak.cu -- FP32 FMA-pipe microbenchmark for B200 (sm_100a).
//
// Measures the sustained FFMA rate the resample kernels can be held against (the "measured FP32 peak"
// used as roofline denominator next to the nominal 148 SM x 128 lanes x 2 x f_clk), for the operand
// shapes the kernels use:
//   rrr   : FFMA Rd, Ra, Rb, Rd       all-register operands
//   rcr   : FFMA Rd, Ra, c[0][imm], Rd  weight taken straight from the constant bank (kernel parameter)
//   rur   : FFMA Rd, Ra, URb, Rd      weight in a uniform register (ULDC with a runtime-uniform index)
//   ffma2 : FFMA2 (fma.rn.f32x2), packed two-lane FMA new on sm_100, register pairs
//   ffma2_ur: FFMA2 Rd.xy, Ra.xy, URb(scalar broadcast), Rd.xy
// Prints one JSON line per variant:  {"variant":..., "tflops":..., "fma_per_clk_per_sm":..., "sm_mhz":...}
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));         \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int NACC = 16;
constexpr int NW = 64;
struct Weights {
    float w[NW];
};

__global__ void __launch_bounds__(256) k_rrr(float* out, int iters, float a, float b)
{
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j)
        acc[j] = threadIdx.x * 1e-3f + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < NACC; ++j)
                acc[j] = fmaf(acc[j], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j)
        s += acc[j];
    if (s == 12345.678f)
        out[0] = s;
}

// acc[j] += v[j] * w[k]: weight from constant bank with immediate offsets (fully unrolled over k)
__global__ void __launch_bounds__(256) k_rcr(float* out, int iters, const __grid_constant__ Weights W)
{
    float acc[NACC], v[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
        acc[j] = 0.f;
        v[j] = threadIdx.x * 1e-3f + j;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int j = 0; j < NACC; ++j)
                acc[j] = fmaf(v[j], W.w[k * 8 + (j & 7)], acc[j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j)
        s += acc[j];
    if (s == 12345.678f)
        out[0] = s;
}

// same, but the weight index is a runtime-uniform loop variable (rolled loop): ULDC + uniform-register operand
__global__ void __launch_bounds__(256) k_rur(float* out, int iters, const __grid_constant__ Weights W)
{
    float acc[NACC], v[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
        acc[j] = 0.f;
        v[j] = threadIdx.x * 1e-3f + j;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int k = 0; k < NW; k += 4) {
            const float w0 = W.w[k], w1 = W.w[k + 1], w2 = W.w[k + 2], w3 = W.w[k + 3];
#pragma unroll
            for (int j = 0; j < NACC; j += 4) {
                acc[j + 0] = fmaf(v[j + 0], w0, acc[j + 0]);
                acc[j + 1] = fmaf(v[j + 1], w1, acc[j + 1]);
                acc[j + 2] = fmaf(v[j + 2], w2, acc[j + 2]);
                acc[j + 3] = fmaf(v[j + 3], w3, acc[j + 3]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j)
        s += acc[j];
    if (s == 12345.678f)
        out[0] = s;
}

// packed FFMA2, all operands register pairs
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b)
{
    float2 acc[NACC / 2];
    const float2 va = make_float2(a, a * 1.0001f), vb = make_float2(b, b * 0.5f);
#pragma unroll
    for (int j = 0; j < NACC / 2; ++j)
        acc[j] = make_float2(threadIdx.x * 1e-3f + j, j * 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < NACC / 2; ++j)
                acc[j] = __ffma2_rn(acc[j], va, vb);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC / 2; ++j)
        s += acc[j].x + acc[j].y;
    if (s == 12345.678f)
        out[0] = s;
}

// packed FFMA2 in the shape the resampler uses: acc.xy += v.xy * w, w a uniform scalar from the constant bank
__global__ void __launch_bounds__(256) k_ffma2u(float* out, int iters, const __grid_constant__ Weights W)
{
    float2 acc[NACC / 2], v[NACC / 2];
#pragma unroll
    for (int j = 0; j < NACC / 2; ++j) {
        acc[j] = make_float2(0.f, 0.f);
        v[j] = make_float2(threadIdx.x * 1e-3f + j, j * 0.25f);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int k = 0; k < NW; k += 4) {
            const float w0 = W.w[k], w1 = W.w[k + 1], w2 = W.w[k + 2], w3 = W.w[k + 3];
#pragma unroll
            for (int j = 0; j < NACC / 2; j += 4) {
                acc[j + 0] = __ffma2_rn(v[j + 0], make_float2(w0, w0), acc[j + 0]);
                acc[j + 1] = __ffma2_rn(v[j + 1], make_float2(w1, w1), acc[j + 1]);
                acc[j + 2] = __ffma2_rn(v[j + 2], make_float2(w2, w2), acc[j + 2]);
                acc[j + 3] = __ffma2_rn(v[j + 3], make_float2(w3, w3), acc[j + 3]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NACC / 2; ++j)
        s += acc[j].x + acc[j].y;
    if (s == 12345.678f)
        out[0] = s;
}

int main(int argc, char** argv)
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    float* out;
    CK(cudaMalloc(&out, 4));
    Weights W;
    for (int i = 0; i < NW; ++i)
        W.w[i] = 1.0f / (i + 3);
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    struct V {
        const char* name;
        double fma_per_thread;
    };
    for (int v = 0; v < 5; ++v) {
        const char* names[5] = {"rrr", "rcr", "rur", "ffma2", "ffma2_ur"};
        double per_thread = 0;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            switch (v) {
            case 0: k_rrr<<<blocks, threads>>>(out, iters, 0.999f, 0.001f); per_thread = (double)iters * 8 * NACC; break;
            case 1: k_rcr<<<blocks, threads>>>(out, iters, W); per_thread = (double)iters * 8 * NACC; break;
            case 2: k_rur<<<blocks, threads>>>(out, iters / 2, W); per_thread = (double)(iters / 2) * NW / 4 * NACC; break;
            case 3: k_ffma2<<<blocks, threads>>>(out, iters, 0.999f, 0.001f); per_thread = (double)iters * 8 * NACC; break;
            case 4: k_ffma2u<<<blocks, threads>>>(out, iters / 2, W); per_thread = (double)(iters / 2) * NW / 4 * NACC; break;
            }
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best)
                best = ms;
        }
        const double fmas = per_thread * blocks * threads;
        const double tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
        printf("{\"variant\":\"%s\",\"ms\":%.3f,\"tflops\":%.2f,\"fma_per_clk_per_sm_at_max_clock\":%.1f,\"sms\":%d,\"max_sm_mhz\":%d}\n",
               names[v], best, tflops, fmas / (best * 1e-3) / sms / (clk_khz * 1e3), sms, clk_khz / 1000);
    }
    return 0;
}
