/*
 * oracle/exact_math.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Transcendentals of the CPU oracle, built from IEEE-754 binary32 primitives
 * only (mul, add, fma, round-to-nearest-integer, integer exponent insert), so
 * that a second implementation of the same recipe on another IEEE machine
 * produces identical bits.  The CUDA product path has its OWN implementation
 * of this recipe (multiposenet_b200/csrc/mpn_math.cuh); tests/ compare the two
 * bit for bit.  Nothing outside tests/, bench.py's cpu_baseline leg and
 * __graft_entry__.smoke() may include this file.
 *
 * Why: the reference evaluates tf.exp / tf.sigmoid / tf.nn.softmax with
 * TensorFlow 1.15's Eigen kernels (detector/utils/box_utils.py:132-133,
 * detector/retinanet.py:73, create_pb.py:74,117).  Those kernels are a
 * third-party dependency that is not under /root/reference and cannot be
 * installed here, so last-ulp agreement with them is unverifiable ("parity
 * unpinned").  What the tests CAN demand is that every keep / suppress / argmax
 * decision is taken on identical bits by oracle and device, which requires one
 * recipe for exp.  The recipe below is the classic Cephes expf
 * (Cody-Waite reduction by ln2 in two parts + degree-5 polynomial), < 1 ulp
 * from the true value over the range used; tests/test_oracle_kat.py checks it
 * against libm to 2e-7 relative.
 *
 * Compile with -ffp-contract=off (the only fused operations are the explicit
 * fmaf calls).
 */
#ifndef MPN_ORACLE_EXACT_MATH_H
#define MPN_ORACLE_EXACT_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float orc_bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t orc_float_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* exp(x), recipe "mpn-exp-v1":
 *   x < -87      -> +0            (no denormal results, so FTZ settings cannot matter)
 *   x >  88      -> +inf
 *   NaN          -> NaN
 *   k = rint(x * log2(e)); r = fma(k, -ln2_hi, x); r = fma(k, -ln2_lo, r)
 *   p = Horner(deg 5 in r) ; e = fma(p, r*r, r) + 1 ; result = e * 2^k
 */
static inline float orc_expf(float x)
{
    if (x != x) return x;
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) return INFINITY;
    float k = rintf(x * 1.44269504088896341f);
    float r = fmaf(k, -0.693359375f, x);
    r = fmaf(k, 2.12194440e-4f, r);
    float z = r * r;
    float p = 1.9875691500E-4f;
    p = fmaf(p, r, 1.3981999507E-3f);
    p = fmaf(p, r, 8.3334519073E-3f);
    p = fmaf(p, r, 4.1665795894E-2f);
    p = fmaf(p, r, 1.6666665459E-1f);
    p = fmaf(p, r, 5.0000001201E-1f);
    float e = fmaf(p, z, r) + 1.0f;
    int ki = (int)k;                                   /* exact, |k| <= 127 */
    float scale = orc_bits_to_float((uint32_t)(ki + 127) << 23);
    return e * scale;
}

/* sigmoid(x) = 1 / (1 + exp(-x)), true division (retinanet.py:73, create_pb.py:74) */
static inline float orc_sigmoidf(float x)
{
    return 1.0f / (1.0f + orc_expf(-x));
}

/* float -> bfloat16 -> float, round to nearest even (NaN kept quiet) */
static inline float orc_round_bf16(float x)
{
    uint32_t u = orc_float_to_bits(x);
    if ((u & 0x7fffffffu) > 0x7f800000u) return orc_bits_to_float(u | 0x00400000u);
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    u &= 0xffff0000u;
    return orc_bits_to_float(u);
}

#endif
