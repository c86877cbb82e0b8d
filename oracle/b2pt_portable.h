/* oracle/b2pt_portable.h — TEST INFRASTRUCTURE (oracle side only).
 *
 * Portable, bit-reproducible definitions shared by the two CPU checkers
 * (oracle/pt_oracle.c = restatement, oracle/ref_harness.cpp = the real
 * reference sources): the Philox4x32-10 sample streams that replace the
 * reference's non-deterministic std::mt19937 (global.hpp:42-53) and a
 * sin/cos evaluated in IEEE double with a fixed operation order, which the
 * harness interposes over libm's sinf/cosf so that CPU and GPU produce the
 * same bits for the two places where the reference feeds sin/cos back into
 * the path (Renderer.cpp:58-60, Material.hpp:114-119).
 *
 * The product (csrc/) carries its own device-side copy of both definitions;
 * nothing under the package includes this file.
 */
#ifndef B2PT_PORTABLE_H
#define B2PT_PORTABLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants) -- */
static inline void b2pt_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout (DESIGN.md "sample streams"): one stream per (pixel, sample);
 * the three wavelength paths of a sample read the SAME stream, the camera
 * draws use stream tag 1.  Draw number `dim` of a stream is word dim&3 of the
 * block with counter (pixel, sample, dim>>2, tag). */
#define B2PT_STREAM_PATH 0u
#define B2PT_STREAM_CAMERA 1u

static inline uint32_t b2pt_stream_word(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel,
                                        uint32_t sample, uint32_t tag, uint32_t dim) {
    uint32_t ctr[4] = {pixel, sample, dim >> 2, tag}, key[2] = {seed_lo, seed_hi}, out[4];
    b2pt_philox4x32_10(ctr, key, out);
    return out[dim & 3u];
}
/* The uniform libstdc++'s uniform_real_distribution<float>(0,1) returns for a 32-bit engine word (generate_canonical<float, 24>):
 * float(word) / 2^32 with the int-to-float conversion's round-to-nearest, nextafter(1, 0) when that reaches 1. */
static inline float b2pt_u01(uint32_t word) {
    float u = (float)word * 2.3283064365386963e-10f;
    return u < 1.f ? u : 0.99999994f;
}

/* ---- portable sin/cos ------------------------------------------------------
 * double-precision evaluation, Cody-Waite reduction by pi/2 and the fdlibm
 * kernel polynomials, every operation written out (compile without FMA
 * contraction).  Valid for |x| < ~1e5; the callers pass 2*pi*u, u in [0,1). */
static inline void b2pt_sincos_d(double x, double *s_out, double *c_out) {
    const double two_over_pi = 6.36619772367581382433e-01;
    const double pio2_hi = 1.57079632673412561417e+00; /* 33 bits of pi/2 */
    const double pio2_lo = 6.07710050650619224932e-11;
    double t = x * two_over_pi;
    double kd = (t >= 0.0) ? (double)(long long)(t + 0.5) : -(double)(long long)(0.5 - t);
    long long k = (long long)kd;
    double r = (x - kd * pio2_hi) - kd * pio2_lo;
    double z = r * r;
    /* fdlibm __kernel_sin / __kernel_cos coefficients */
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double ps = S6;
    ps = ps * z + S5; ps = ps * z + S4; ps = ps * z + S3; ps = ps * z + S2; ps = ps * z + S1;
    double sr = r + (r * z) * ps;
    double pc = C6;
    pc = pc * z + C5; pc = pc * z + C4; pc = pc * z + C3; pc = pc * z + C2; pc = pc * z + C1;
    double cr = (1.0 - 0.5 * z) + (z * z) * pc;
    switch ((int)(k & 3)) {
    case 0: *s_out = sr; *c_out = cr; break;
    case 1: *s_out = cr; *c_out = -sr; break;
    case 2: *s_out = -sr; *c_out = -cr; break;
    default: *s_out = -cr; *c_out = sr; break;
    }
}
static inline float b2pt_sinf(float x) { double s, c; b2pt_sincos_d((double)x, &s, &c); return (float)s; }
static inline float b2pt_cosf(float x) { double s, c; b2pt_sincos_d((double)x, &s, &c); return (float)c; }

#ifdef __cplusplus
}
#endif
#endif
