// ptx.cuh — thin inline-PTX wrappers for the sm_100a features the kernels use:
// mbarrier producer/consumer pipeline, 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP),
// approximate reciprocal square root (MUFU.RSQ) and packed FP32 pairs (FFMA2 / FADD2 / FMUL2).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace pcuda {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// Makes freshly initialised mbarriers visible to the async (TMA) proxy.
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Orders prior generic-proxy shared-memory writes before later async-proxy accesses.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PCUDA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PCUDA_DONE;\n"
        "bra PCUDA_WAIT;\n"
        "PCUDA_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk tensor-memory-accelerator copy global -> shared, completion signalled on `bar`
// (complete_tx::bytes).  bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// MUFU.RSQ: max relative error 2^-22.9 (PTX ISA, rsqrt.approx.f32), denormals flushed.
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One 32-bit read-only load that the compiler cannot merge with its neighbours into a vector load.
__device__ __forceinline__ float ldg_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// The SM this thread runs on.
__device__ __forceinline__ unsigned smid() {
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}

// Packed FP32 pairs (sm_100+: fma.rn.f32x2 etc.).  One issue slot does two lanes of work, which
// is what lets MUFU / LDS / loop overhead hide under the FMA pipe in the pair kernel.
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 splat(float s) { return make_float2(s, s); }

// x == +0.0 ? 1.0 : x, with integer instructions (a DSETP would take an FP64-pipe slot, and the
// FP64 pipe is the limiter of the f64 kernels).  x is a sum of squares: never -0.0.
__device__ __forceinline__ double one_if_zero(double x) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    return __hiloint2double((hi | lo) == 0 ? 0x3ff00000 : hi, lo);
}

// mu * x^(-3/2) in 7 FP64 operations (CUDA's rsqrt() + three multiplies take 8 plus a range
// check): y0 = MUFU.RSQ64H(x) carries ~20 bits; with e = 1 - x*y0^2 (|e| < 2^-19),
//     x^(-3/2) = y0^3 (1 - e)^(-3/2) = y0^3 (1 + e (3/2 + 15/8 e)) + O(e^3),   35/16 e^3 < 2^-56,
// so the result is good to a few ulp — far inside the 1e-12 parity bound.  x must be a normal
// positive number: the kernel's r2 is one (coincident pairs are handled before the call); x = 0
// (unchecked, eps = 0) gives NaN like the reference's 0 * inf.
__device__ __forceinline__ double mu_rcbrt2(double x, double mu) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double t = y0 * y0;
    const double e = fma(-x, t, 1.0);
    const double q = fma(1.875, e, 1.5);
    const double um = (t * y0) * mu;
    return fma(um, e * q, um);
}

}  // namespace ptx
}  // namespace pcuda
