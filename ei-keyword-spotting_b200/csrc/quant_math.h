// eikws-b200: integer requantisation arithmetic shared by host planning code and device kernels.
// Semantics follow the TFLite/gemmlowp definitions the reference's int8 kernels are built on:
//   SaturatingRoundingDoublingHighMul   third_party/gemmlowp/fixedpoint/fixedpoint.h:329-340
//   RoundingDivideByPOT                 third_party/gemmlowp/fixedpoint/fixedpoint.h:357-368
//   MultiplyByQuantizedMultiplier*      TFL/kernels/internal/common.h:138-162
//   exp_on_negative_values, one_over_one_plus_x_for_x_in_0_1   fixedpoint.h:738-790, 843-862
// (paths relative to edge-impulse-sdk/).  Everything is int32/int64 arithmetic => bit-exact on any machine.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define EIKWS_HD __host__ __device__ __forceinline__
#else
#define EIKWS_HD inline
#endif

namespace eikws {
namespace qm {

// (a*b + nudge) / 2^31 with C++ truncating division, saturating the single overflow case.
EIKWS_HD int32_t srdhm(int32_t a, int32_t b) {
    if (a == b && a == INT32_MIN) return INT32_MAX;
    const int64_t ab = static_cast<int64_t>(a) * static_cast<int64_t>(b);
    const int64_t nudged = ab + (ab >= 0 ? (1ll << 30) : (1ll - (1ll << 30)));
    // truncation toward zero == arithmetic shift of the magnitude
    const int64_t q = nudged >= 0 ? (nudged >> 31) : -((-nudged) >> 31);
    return static_cast<int32_t>(q);
}

// rounding arithmetic right shift (ties away from zero)
EIKWS_HD int32_t rdiv_pot(int32_t x, int exponent) {
    const int32_t mask = static_cast<int32_t>((1ll << exponent) - 1);
    const int32_t remainder = x & mask;
    const int32_t threshold = (mask >> 1) + (x < 0 ? 1 : 0);
    return (x >> exponent) + (remainder > threshold ? 1 : 0);
}

EIKWS_HD int32_t mul_by_quantized_multiplier(int32_t x, int32_t mult, int shift) {
    const int left = shift > 0 ? shift : 0;
    const int right = shift > 0 ? 0 : -shift;
    return rdiv_pot(srdhm(static_cast<int32_t>(static_cast<uint32_t>(x) << left), mult), right);
}
EIKWS_HD int32_t mul_smaller_than_one(int32_t x, int32_t mult, int left_shift) {
    return rdiv_pot(srdhm(x, mult), -left_shift);
}
EIKWS_HD int32_t mul_greater_than_one(int32_t x, int32_t mult, int left_shift) {
    return srdhm(static_cast<int32_t>(static_cast<uint32_t>(x) << left_shift), mult);
}

// SaturatingRoundingMultiplyByPOT<e> (fixedpoint.h:375-416)
EIKWS_HD int32_t sat_mul_pot(int32_t x, int e) {
    if (e == 0) return x;
    if (e < 0) return rdiv_pot(x, -e);
    const int32_t thr = static_cast<int32_t>((1ll << (31 - e)) - 1);
    if (x > thr) return INT32_MAX;
    if (x < -thr) return INT32_MIN;
    return static_cast<int32_t>(static_cast<uint32_t>(x) << e);
}

EIKWS_HD int32_t rounding_half_sum(int32_t a, int32_t b) {
    const int64_t s = static_cast<int64_t>(a) + static_cast<int64_t>(b);
    const int64_t t = s + (s >= 0 ? 1 : -1);
    return static_cast<int32_t>(t >= 0 ? (t >> 1) : -((-t) >> 1));
}

// exp(x) for x in [-1/4, 0), Q0.31 -> Q0.31
EIKWS_HD int32_t exp_quarter_interval(int32_t a) {
    const int32_t c_term = 1895147668;  // exp(-1/8)
    const int32_t c_third = 715827883;  // 1/3
    const int32_t x = a + (1 << 28);
    const int32_t x2 = srdhm(x, x);
    const int32_t x3 = srdhm(x2, x);
    const int32_t x4 = srdhm(x2, x2);
    const int32_t x4_over_4 = sat_mul_pot(x4, -2);
    const int32_t poly = sat_mul_pot(srdhm(x4_over_4 + x3, c_third) + x2, -1);
    return c_term + srdhm(c_term, x + poly);
}

// exp(x) for x <= 0 given as Q5.26, result Q0.31
EIKWS_HD int32_t exp_on_negative_q5_26(int32_t a) {
    const int32_t quarter = 1 << 24;
    const int32_t a_mod = (a & (quarter - 1)) - quarter;
    int32_t result = exp_quarter_interval(sat_mul_pot(a_mod, 5));
    const int32_t remainder = a_mod - a;
    const int32_t mult[7] = {1672461947, 1302514674, 790015084, 290630308, 39332535, 720401, 242};
    for (int e = -2; e <= 4; e++)
        if (remainder & (1 << (26 + e))) result = srdhm(result, mult[e + 2]);
    return a == 0 ? INT32_MAX : result;
}

// 1/(1+x) for x in (0,1), Q0.31 -> Q0.31, three Newton-Raphson steps in Q2.29
EIKWS_HD int32_t one_over_one_plus_x(int32_t a) {
    const int32_t half_den = rounding_half_sum(a, INT32_MAX);
    int32_t x = 1515870810 + srdhm(half_den, -1010580540);
    for (int i = 0; i < 3; i++) {
        const int32_t hdx = srdhm(half_den, x);
        const int32_t one_minus = (1 << 29) - hdx;
        x = x + sat_mul_pot(srdhm(x, one_minus), 2);
    }
    return sat_mul_pot(x, 1);
}

}  // namespace qm
}  // namespace eikws
