// probe.cu — FP32 pipe microbenchmark used by bench.py to MEASURE the roofline denominator of the
// brute-force kernel on the box it runs on (MEASURED_PEAKS.json holds HBM and bf16 tensor peaks
// only).  Independent FFMA chains, 8 per thread, no memory traffic in the timed loop.
#include "common.cuh"
#include "ptx.cuh"

namespace pcuda {

// MODE 0: scalar FFMA, 16 chains.  1: FFMA2 with three distinct register operands, 8 chains.
// 2: FFMA2 v = v * a + v (two distinct registers).  3: FMUL2.  4: FADD2 with a scalar-broadcast
// operand (the form the pair kernel uses).  5: 1 FFMA2 : 2 FFMA interleaved.  6: 2 FFMA2 : 2 FFMA.
// 7: FFMA2 + MUFU.RSQ + FADD every 3rd.  8-11: 12 FFMA2 with 2 MUFU / 2 FMNMX / nothing / both.  All count 2 flop per lane-FMA (FMUL / FADD as one slot).
template <int MODE>
__global__ void __launch_bounds__(256) fma_probe(float *out, int iters, float a, float b) {
    float2 v[8];
    float w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f + 1.f);
        w[k] = threadIdx.x * 1e-3f + 0.25f * k;
    }
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) {
                v[k].x = fmaf(v[k].x, a, b);
                v[k].y = fmaf(v[k].y, a, b);
            } else if (MODE == 1) {
                v[k] = ptx::fma2(v[k], aa, bb);
            } else if (MODE == 2) {
                v[k] = ptx::fma2(v[k], aa, v[k]);
            } else if (MODE == 3) {
                v[k] = ptx::mul2(v[k], aa);
            } else if (MODE == 4) {
                v[k] = ptx::add2(v[k], ptx::splat(b));
            } else if (MODE == 5) {
                if (k < 4) v[k] = ptx::fma2(v[k], aa, bb);
                w[k] = fmaf(w[k], a, b);
            } else if (MODE == 6) {
                v[k] = ptx::fma2(v[k], aa, bb);
                w[k] = fmaf(w[k], a, b);
            } else if (MODE == 7) {
                v[k] = ptx::fma2(v[k], aa, bb);
                if (k % 3 == 0) w[k] = ptx::rsqrt_approx(w[k] + 1.5f);
            } else if (MODE == 8) {   // 12 FFMA2 (2 regs) : 2 MUFU, the pair kernel's ratio
                v[k] = ptx::fma2(v[k], aa, v[k]);
                if (k < 4) v[k] = ptx::fma2(v[k], aa, v[k]);
                if (k < 2) w[k] = ptx::rsqrt_approx(w[k]);
            } else if (MODE == 9) {   // 12 FFMA2 (2 regs) : 2 FMNMX
                v[k] = ptx::fma2(v[k], aa, v[k]);
                if (k < 4) v[k] = ptx::fma2(v[k], aa, v[k]);
                if (k < 2) w[k] = fmaxf(w[k], b + k);
            } else if (MODE == 10) {  // 12 FFMA2 (2 regs) alone
                v[k] = ptx::fma2(v[k], aa, v[k]);
                if (k < 4) v[k] = ptx::fma2(v[k], aa, v[k]);
            } else {                  // 12 FFMA2 (2 regs) : 2 MUFU : 2 FMNMX
                v[k] = ptx::fma2(v[k], aa, v[k]);
                if (k < 4) v[k] = ptx::fma2(v[k], aa, v[k]);
                if (k < 2) w[k] = ptx::rsqrt_approx(fmaxf(w[k], b + k));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k].x + v[k].y + w[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
        float *o = ctx->d_misc.as<float>();
        switch (packed) {
            case 0: fma_probe<0><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 1e-3f); break;
            case 1: fma_probe<1><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 1e-3f); break;
            case 2: fma_probe<2><<<blocks, threads, 0, ctx->stream>>>(o, iters, -1e-6f, 1e-3f); break;
            case 3: fma_probe<3><<<blocks, threads, 0, ctx->stream>>>(o, iters, 1.0f, 1e-3f); break;
            case 4: fma_probe<4><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 1e-3f); break;
            case 5: fma_probe<5><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 1e-3f); break;
            case 6: fma_probe<6><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 1e-3f); break;
            case 7: fma_probe<7><<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999f, 1e-3f); break;
            case 8: fma_probe<8><<<blocks, threads, 0, ctx->stream>>>(o, iters, -1e-6f, 1e-3f); break;
            case 9: fma_probe<9><<<blocks, threads, 0, ctx->stream>>>(o, iters, -1e-6f, 1e-3f); break;
            case 10: fma_probe<10><<<blocks, threads, 0, ctx->stream>>>(o, iters, -1e-6f, 1e-3f); break;
            default: fma_probe<11><<<blocks, threads, 0, ctx->stream>>>(o, iters, -1e-6f, 1e-3f); break;
        }
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
    // lane-slots per thread per iteration: 16 (modes 0-4), 4*2+8 = 16 (mode 5), 16+8 = 24 (mode 6),
    // 16 (mode 7; the MUFU is not counted)
    const double slots = packed == 6 || packed >= 8 ? 24.0 : 16.0;
    const double flops = 2.0 * slots * (double)iters * (double)blocks * threads;
    if (tflops_out) *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return PCUDA_OK;
}
