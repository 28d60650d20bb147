// probe.cu — FP32 pipe microbenchmark used by bench.py to MEASURE the roofline denominator of the
// brute-force kernel on the box it runs on (MEASURED_PEAKS.json holds HBM and bf16 tensor peaks
// only).  Independent FFMA chains, 8 per thread, no memory traffic in the timed loop.
#include "common.cuh"
#include "ptx.cuh"

namespace pcuda {

template <bool PACKED>
__global__ void __launch_bounds__(256) fma_probe(float *out, int iters, float a, float b) {
    if (PACKED) {
        float2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f);
        const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ptx::fma2(v[k], aa, bb);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k].x + v[k].y;
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = threadIdx.x * 1e-3f + k;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = fmaf(v[k], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += v[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

}  // namespace pcuda

using namespace pcuda;

extern "C" int pcuda_probe_fp32(pcuda_ctx *ctx, int packed, int iters, int repeats,
                                double *tflops_out, float *ms_out) {
    if (!ctx || iters <= 0 || repeats <= 0) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "bad probe arguments");
    DeviceGuard guard(ctx->device);
    const int blocks = ctx->sm_count * 8, threads = 256;
    PCUDA_CUDA_TRY(ctx, ctx->d_misc.ensure((size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    PCUDA_CUDA_TRY(ctx, cudaEventCreate(&e0));
    PCUDA_CUDA_TRY(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < repeats + 1; ++r) {
        cudaEventRecord(e0, ctx->stream);
        if (packed)
            fma_probe<true><<<blocks, threads, 0, ctx->stream>>>(ctx->d_misc.as<float>(), iters, 0.999f, 1e-3f);
        else
            fma_probe<false><<<blocks, threads, 0, ctx->stream>>>(ctx->d_misc.as<float>(), iters, 0.999f, 1e-3f);
        cudaEventRecord(e1, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return fail(ctx, PCUDA_ERR_CUDA, "probe kernel: %s", cudaGetErrorString(e));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;  // first run is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * threads;
    if (tflops_out) *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return PCUDA_OK;
}
