// synth.cuh -- synthetic embedding rows generated on the device.
//
// Spec (DESIGN.md "Synthetic inputs"): Philox4x32-10, counter = (row_lo, row_hi, quad, stream),
// key = (seed_lo, seed_hi); each call yields four u32 -> two Box-Muller pairs -> four normals for
// columns 4*quad .. 4*quad+3.  u1 = (r + 0.5) * 2^-32, angle = 2*pi*r' * 2^-32.  log / sin / cos are
// evaluated with + - * / only, in a fixed order, and this file is compiled with --fmad=false, so a
// CPU that follows the same spec without FMA contraction regenerates the same bits.
#pragma once
#include <stdint.h>

#define SYNTH_STREAM_NOISE 0u
#define SYNTH_STREAM_CENTRE 1u
#define SYNTH_STREAM_ASSIGN 2u
#define SYNTH_STREAM_SHIFT 3u

__device__ __forceinline__ void synth_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                             uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ln(u), u a normal positive double: u = m * 2^e with m in [sqrt(1/2), sqrt(2)), atanh series.
__device__ __forceinline__ double synth_log(double u) {
    uint64_t bits = (uint64_t)__double_as_longlong(u);
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    double m = __longlong_as_double((long long)((bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
    if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
    double s = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0));
    double s2 = __dmul_rn(s, s);
    double p = 1.0 / 29.0;
#pragma unroll
    for (int d = 27; d >= 1; d -= 2) p = __dadd_rn(__dmul_rn(p, s2), 1.0 / (double)d);
    return __dadd_rn(__dmul_rn(__dmul_rn(2.0, s), p), __dmul_rn((double)e, 0.6931471805599453));
}

// sin, cos of 2*pi*r*2^-32: octant q = r >> 29, Taylor on phi = frac * pi/4, angle addition.
__device__ __forceinline__ void synth_sincos(uint32_t r, double* sn, double* cs) {
    uint32_t q = r >> 29;
    double f = __dmul_rn((double)(r & 0x1FFFFFFFu), 1.0 / 536870912.0);
    double a = __dmul_rn(f, 0.7853981633974483);
    double a2 = __dmul_rn(a, a);
    double ps = -1.0 / 121645100408832000.0;
    ps = __dadd_rn(__dmul_rn(ps, a2), 1.0 / 355687428096000.0);
    ps = __dadd_rn(__dmul_rn(ps, a2), -(1.0 / 1307674368000.0));
    ps = __dadd_rn(__dmul_rn(ps, a2), 1.0 / 6227020800.0);
    ps = __dadd_rn(__dmul_rn(ps, a2), -(1.0 / 39916800.0));
    ps = __dadd_rn(__dmul_rn(ps, a2), 1.0 / 362880.0);
    ps = __dadd_rn(__dmul_rn(ps, a2), -(1.0 / 5040.0));
    ps = __dadd_rn(__dmul_rn(ps, a2), 1.0 / 120.0);
    ps = __dadd_rn(__dmul_rn(ps, a2), -(1.0 / 6.0));
    ps = __dadd_rn(__dmul_rn(ps, a2), 1.0);
    double s0 = __dmul_rn(a, ps);
    double pc = 1.0 / 6402373705728000.0;
    pc = __dadd_rn(__dmul_rn(pc, a2), -(1.0 / 20922789888000.0));
    pc = __dadd_rn(__dmul_rn(pc, a2), 1.0 / 87178291200.0);
    pc = __dadd_rn(__dmul_rn(pc, a2), -(1.0 / 479001600.0));
    pc = __dadd_rn(__dmul_rn(pc, a2), 1.0 / 3628800.0);
    pc = __dadd_rn(__dmul_rn(pc, a2), -(1.0 / 40320.0));
    pc = __dadd_rn(__dmul_rn(pc, a2), 1.0 / 720.0);
    pc = __dadd_rn(__dmul_rn(pc, a2), -(1.0 / 24.0));
    pc = __dadd_rn(__dmul_rn(pc, a2), 0.5);
    double c0 = __dadd_rn(1.0, -__dmul_rn(a2, pc));
    const double H = 0.7071067811865476;
    // sin(q*pi/4), cos(q*pi/4) up to the factor H on odd octants
    double sq = (q == 0 || q == 4) ? 0.0 : (q < 4 ? 1.0 : -1.0);
    double cq = (q == 2 || q == 6) ? 0.0 : ((q < 2 || q == 7) ? 1.0 : -1.0);
    if (q & 1u) { sq = __dmul_rn(sq, H); cq = __dmul_rn(cq, H); }
    *sn = __dadd_rn(__dmul_rn(sq, c0), __dmul_rn(cq, s0));
    *cs = __dadd_rn(__dmul_rn(cq, c0), -__dmul_rn(sq, s0));
}

__device__ __forceinline__ void synth_normal4(uint64_t seed, uint32_t stream, uint64_t row, uint32_t quad, double z[4]) {
    uint32_t r[4];
    synth_philox((uint32_t)row, (uint32_t)(row >> 32), quad, stream, (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double u1 = __dmul_rn(__dadd_rn((double)r[2 * h], 0.5), 1.0 / 4294967296.0);
        double rad = __dsqrt_rn(__dmul_rn(-2.0, synth_log(u1)));
        double sn, cs;
        synth_sincos(r[2 * h + 1], &sn, &cs);
        z[2 * h] = __dmul_rn(rad, cs);
        z[2 * h + 1] = __dmul_rn(rad, sn);
    }
}

__device__ __forceinline__ void synth_row_quad(int kind, uint64_t seed, uint64_t row, uint32_t quad, uint32_t n_centres,
                                               double noise, double v[4]) {
    double z[4];
    synth_normal4(seed, SYNTH_STREAM_NOISE, row, quad, z);
    if (kind == 1) {
        uint32_t r[4];
        synth_philox((uint32_t)row, (uint32_t)(row >> 32), 0u, SYNTH_STREAM_ASSIGN, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        uint64_t centre = r[0] % n_centres;
        double c[4];
        synth_normal4(seed, SYNTH_STREAM_CENTRE, centre, quad, c);
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = __dadd_rn(c[t], __dmul_rn(noise, z[t]));
    } else if (kind == 2) {
        double s[4];
        synth_normal4(seed, SYNTH_STREAM_SHIFT, row, 0u, s);
        double shift = __dmul_rn(noise, s[0]);
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = __dadd_rn(z[t], shift);
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = z[t];
    }
}
