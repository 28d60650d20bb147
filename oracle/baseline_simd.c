/*
 * baseline_simd.c — CPU timing baseline: restatement of particular's parallel::BruteForceSimd<8>.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline and --impl reference legs, tests/).
 * Never linked or called from the product path.
 *
 * Restates (paths relative to /root/reference/particular/src):
 *   AoS -> SoA lanes, zero-padded tail     lib.rs:398-401, gravity/impls/mod.rs:50-81
 *   SIMD pair term                         gravity/impls/mod.rs:224-241
 *       d = src - at; n = d.d; ns = n + eps*eps; mag = m * recip_sqrt(ns*ns*ns);
 *       CHECKED => d * blend(n != 0, mag, 0)
 *   per-target splat + fold + reduce_add   sequential.rs:402-417, gravity/impls/mod.rs:35-46
 *   rayon par_iter over affected           parallel.rs:266-287, 110-137  (OpenMP stands in)
 * `wide::f32x8::recip_sqrt` is the hardware approximate rsqrt (AVX: vrsqrtps), so this path is a
 * TIMING baseline, not a parity oracle: its per-pair relative error is ~2^-12.
 * Rust does not contract a*b+c, hence -ffp-contract=off and no FMA intrinsics.
 *
 * Build: gcc -O3 -mavx2 -ffp-contract=off -fopenmp (oracle/Makefile).
 */
#include <immintrin.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1; the baseline legs ask for all host cores explicitly. */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_baseline_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline float hsum8(__m256 v) { /* wide's reduce_add: horizontal sum of the 8 lanes */
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    __m128 s = _mm_add_ps(lo, hi);
    s = _mm_add_ps(s, _mm_movehl_ps(s, s));
    s = _mm_add_ss(s, _mm_shuffle_ps(s, s, 1));
    return _mm_cvtss_f32(s);
}

/* affected: na x 3 positions; affecting: nb x 4 {x,y,z,mu}; out: na x 3.
 * Returns 0, or -1 on allocation failure. */
int oracle_bruteforce_simd8_parallel_f32x3(const float *affected, size_t na,
                                           const float *affecting, size_t nb, float eps,
                                           int checked, float *out) {
    size_t nchunk = (nb + 7) / 8;
    float *soa = (float *)aligned_alloc(32, (nchunk ? nchunk : 1) * 32 * sizeof(float));
    if (!soa) return -1;
    memset(soa, 0, (nchunk ? nchunk : 1) * 32 * sizeof(float)); /* zero-mass padding at origin */
    for (size_t j = 0; j < nb; ++j) {
        size_t c = j / 8, l = j % 8;
        soa[c * 32 + 0 + l] = affecting[j * 4 + 0];
        soa[c * 32 + 8 + l] = affecting[j * 4 + 1];
        soa[c * 32 + 16 + l] = affecting[j * 4 + 2];
        soa[c * 32 + 24 + l] = affecting[j * 4 + 3];
    }
    const __m256 eps2 = _mm256_set1_ps(eps * eps);
    const __m256 zero = _mm256_setzero_ps();
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)na; ++i) {
        const __m256 px = _mm256_set1_ps(affected[i * 3 + 0]);
        const __m256 py = _mm256_set1_ps(affected[i * 3 + 1]);
        const __m256 pz = _mm256_set1_ps(affected[i * 3 + 2]);
        __m256 ax = zero, ay = zero, az = zero;
        for (size_t c = 0; c < nchunk; ++c) {
            const float *s = soa + c * 32;
            __m256 dx = _mm256_sub_ps(_mm256_load_ps(s), px);
            __m256 dy = _mm256_sub_ps(_mm256_load_ps(s + 8), py);
            __m256 dz = _mm256_sub_ps(_mm256_load_ps(s + 16), pz);
            __m256 n = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(dx, dx), _mm256_mul_ps(dy, dy)),
                                     _mm256_mul_ps(dz, dz));
            __m256 ns = _mm256_add_ps(n, eps2);
            __m256 mag = _mm256_mul_ps(_mm256_load_ps(s + 24),
                                       _mm256_rsqrt_ps(_mm256_mul_ps(_mm256_mul_ps(ns, ns), ns)));
            if (checked) mag = _mm256_and_ps(mag, _mm256_cmp_ps(n, zero, _CMP_NEQ_UQ));
            ax = _mm256_add_ps(ax, _mm256_mul_ps(dx, mag));
            ay = _mm256_add_ps(ay, _mm256_mul_ps(dy, mag));
            az = _mm256_add_ps(az, _mm256_mul_ps(dz, mag));
        }
        out[i * 3 + 0] = hsum8(ax);
        out[i * 3 + 1] = hsum8(ay);
        out[i * 3 + 2] = hsum8(az);
    }
    free(soa);
    return 0;
}
