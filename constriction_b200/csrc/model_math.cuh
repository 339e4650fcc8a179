// model_math.cuh -- float -> 24-bit fixed-point quantisation of entropy models, evaluated once per
// *model entry* (table build), never per coded symbol.  `__host__ __device__`: the tabulation
// kernels in model_tables.cu and the host unit-test harness compile the same code.
//
// Must be compiled WITHOUT floating-point contraction (nvcc -fmad=false, gcc -ffp-contract=off):
// the reference is Rust, which never fuses a*b+c, and the truncating casts below are sensitive to
// the last ulp.
//
// Reference behaviour being matched:
//   LeakyQuantizer<f64,i32,u32,24>      src/stream/model/quantize.rs:284-308, 475-486, 525-568
//   fast_quantized_cdf                  src/stream/model/categorical.rs:16-54
//   Gaussian CDF                        crate `probability` 0.20.3 (Cargo.lock:510-517):
//                                       (1 + erf((x - mu) / (sigma * sqrt 2))) / 2
//   erf / exp                           crate `libm` 0.2.16 (Cargo.lock:358-361), itself the
//                                       FreeBSD msun s_erf.c / e_exp.c algorithm (Sun, 1993):
//                                       piecewise rational approximations, restated here as
//                                       coefficient arrays + Horner loops.
#pragma once
#include <stdint.h>
#include <string.h>

#include "coder_math.cuh"

namespace ctr {
namespace mm {

CTR_HD uint32_t high_word(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (uint32_t)(b >> 32);
#endif
}

CTR_HD double with_zero_low_word(double x) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(x), 0);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    b &= 0xffffffff00000000ull;
    memcpy(&x, &b, 8);
    return x;
#endif
}

CTR_HD double from_bits(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

CTR_HD double absd(double x) { return x < 0.0 ? -x : (x == 0.0 ? 0.0 : x); }

// Horner with the highest coefficient innermost: c[0] + x*(c[1] + x*(... + x*c[n-1])).
template <int N>
CTR_HD double horner(const double (&c)[N], double x) {
    double acc = c[N - 1];
#pragma unroll
    for (int i = N - 2; i >= 0; --i) acc = c[i] + x * acc;
    return acc;
}

// 2^n * x for the exponent range exp() needs (|n| <= 1100); msun scalbn.
CTR_HD double scale_by_pow2(double x, int n) {
    const double two_p1023 = from_bits(0x7fe0000000000000ull);
    const double two_m969 = from_bits(0x0360000000000000ull);  // 2^-1022 * 2^53
    double y = x;
    if (n > 1023) {
        y *= two_p1023;
        n -= 1023;
        if (n > 1023) {
            y *= two_p1023;
            n -= 1023;
            if (n > 1023) n = 1023;
        }
    } else if (n < -1022) {
        y *= two_m969;
        n += 1022 - 53;
        if (n < -1022) {
            y *= two_m969;
            n += 1022 - 53;
            if (n < -1022) n = -1022;
        }
    }
    return y * from_bits((uint64_t)(0x3ff + n) << 52);
}

// msun e_exp.c
CTR_HD double exp_msun(double x) {
    const double ln2_hi = 6.93147180369123816490e-01;
    const double ln2_lo = 1.90821492927058770002e-10;
    const double inv_ln2 = 1.44269504088896338700e+00;
    const double P[5] = {1.66666666666666019037e-01, -2.77777777770155933842e-03, 6.61375632143793436117e-05,
                         -1.65339022054652515390e-06, 4.13813679705723846039e-08};
    uint32_t hx = high_word(x);
    const int sign = (int)(hx >> 31);
    hx &= 0x7fffffffu;
    if (hx >= 0x4086232bu) {
        if (x != x) return x;
        if (x > 709.782712893383973096) return x * from_bits(0x7fe0000000000000ull);
        if (x < -745.13321910194110842) return 0.0;
    }
    double hi, lo;
    int k;
    if (hx > 0x3fd62e42u) {
        if (hx >= 0x3ff0a2b2u)
            k = (int)(inv_ln2 * x + (sign ? -0.5 : 0.5));
        else
            k = 1 - sign - sign;
        hi = x - (double)k * ln2_hi;
        lo = (double)k * ln2_lo;
        x = hi - lo;
    } else if (hx > 0x3e300000u) {
        k = 0;
        hi = x;
        lo = 0.0;
    } else {
        return 1.0 + x;
    }
    const double xx = x * x;
    const double c = x - xx * horner(P, xx);
    const double y = 1.0 + (x * c / (2.0 - c) - lo + hi);
    return k == 0 ? y : scale_by_pow2(y, k);
}

// msun s_erf.c, |x| in [0.84375, 6): 1 - erf.
CTR_HD double erfc_mid(uint32_t ix, double x) {
    const double erx = 8.45062911510467529297e-01;
    const double ax = absd(x);
    if (ix < 0x3ff40000u) {  // |x| < 1.25
        const double PA[7] = {-2.36211856075265944077e-03, 4.14856118683748331666e-01, -3.72207876035701323847e-01,
                              3.18346619901161753674e-01,  -1.10894694282396677476e-01, 3.54783043256182359371e-02,
                              -2.16637559486879084300e-03};
        const double QA[7] = {1.0,
                              1.06420880400844228286e-01,
                              5.40397917702171048937e-01,
                              7.18286544141962662868e-02,
                              1.26171219808761642112e-01,
                              1.36370839120290507362e-02,
                              1.19844998467991074170e-02};
        const double s = ax - 1.0;
        return 1.0 - erx - horner(PA, s) / horner(QA, s);
    }
    const double s = 1.0 / (ax * ax);
    double R, S;
    if (ix < 0x4006db6du) {  // |x| < 1/0.35
        const double RA[8] = {-9.86494403484714822705e-03, -6.93858572707181764372e-01, -1.05586262253232909814e+01,
                              -6.23753324503260060396e+01, -1.62396669462573470355e+02, -1.84605092906711035994e+02,
                              -8.12874355063065934246e+01, -9.81432934416914548592e+00};
        const double SA[9] = {1.0,
                              1.96512716674392571292e+01,
                              1.37657754143519042600e+02,
                              4.34565877475229228821e+02,
                              6.45387271733267880336e+02,
                              4.29008140027567833386e+02,
                              1.08635005541779435134e+02,
                              6.57024977031928170135e+00,
                              -6.04244152148580987438e-02};
        R = horner(RA, s);
        S = horner(SA, s);
    } else {
        const double RB[7] = {-9.86494292470009928597e-03, -7.99283237680523006574e-01, -1.77579549177547519889e+01,
                              -1.60636384855821916062e+02, -6.37566443368389627722e+02, -1.02509513161107724954e+03,
                              -4.83519191608651397019e+02};
        const double SB[8] = {1.0,
                              3.03380607434824582924e+01,
                              3.25792512996573918826e+02,
                              1.53672958608443695994e+03,
                              3.19985821950859553908e+03,
                              2.55305040643316442583e+03,
                              4.74528541206955367215e+02,
                              -2.24409524465858183362e+01};
        R = horner(RB, s);
        S = horner(SB, s);
    }
    const double z = with_zero_low_word(ax);
    return exp_msun(-z * z - 0.5625) * exp_msun((z - ax) * (z + ax) + R / S) / ax;
}

CTR_HD double erf_msun(double x) {
    uint32_t ix = high_word(x);
    const bool negative = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    if (ix >= 0x7ff00000u) return 1.0 - 2.0 * (negative ? 1.0 : 0.0) + 1.0 / x;
    if (ix < 0x3feb0000u) {  // |x| < 0.84375
        if (ix < 0x3e300000u) return 0.125 * (8.0 * x + 1.02703333676410069053e+00 * x);
        const double PP[5] = {1.28379167095512558561e-01, -3.25042107247001499370e-01, -2.84817495755985104766e-02,
                              -5.77027029648944159157e-03, -2.37630166566501626084e-05};
        const double QQ[6] = {1.0,
                              3.97917223959155352819e-01,
                              6.50222499887672944485e-02,
                              5.08130628187576562776e-03,
                              1.32494738004321644526e-04,
                              -3.96022827877536812320e-06};
        const double z = x * x;
        const double y = horner(PP, z) / horner(QQ, z);
        return x + x * y;
    }
    double y;
    if (ix < 0x40180000u)
        y = 1.0 - erfc_mid(ix, x);
    else
        y = 1.0 - from_bits(0x0010000000000000ull);  // 1 - 2^-1022
    return negative ? -y : y;
}

CTR_HD double gaussian_cdf(double x, double mean, double std) {
    const double sqrt2 = 1.41421356237309504880168872420969808;
    return (1.0 + erf_msun((x - mean) / (std * sqrt2))) / 2.0;
}

CTR_HD uint64_t to_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return b;
#endif
}

// msun s_log1p.c (what libm 0.2.16's log1p implements; reference call sites: categorical.rs:11,98-126,166-172)
CTR_HD double log1p_msun(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
                          1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01};
    uint64_t ui = to_bits(x);
    const uint32_t hx = (uint32_t)(ui >> 32);
    int k = 1;
    double f = 0.0, c = 0.0;
    if (hx < 0x3fda827au || (hx >> 31)) {  // 1 + x < sqrt(2)+
        if (hx >= 0xbff00000u) {           // x <= -1
            const double zero = 0.0;
            return x == -1.0 ? x / zero : (x - x) / zero;
        }
        if ((hx << 1) < (0x3ca00000u << 1)) return x;  // |x| < 2^-53
        if (hx <= 0xbfd2bec4u) {                       // sqrt(2)/2- <= 1 + x < sqrt(2)+
            k = 0;
            f = x;
        }
    } else if (hx >= 0x7ff00000u) {
        return x;
    }
    if (k) {
        const double u = 1.0 + x;
        ui = to_bits(u);
        uint32_t hu = (uint32_t)(ui >> 32);
        hu += 0x3ff00000u - 0x3fe6a09eu;
        k = (int)(hu >> 20) - 0x3ff;
        if (k < 54) {  // correction term ~ log(1 + x) - log(u)
            c = k >= 2 ? 1.0 - (u - x) : x - (u - 1.0);
            c /= u;
        }
        hu = (hu & 0x000fffffu) + 0x3fe6a09eu;  // reduce u into [sqrt(2)/2, sqrt(2)]
        f = from_bits(((uint64_t)hu << 32) | (ui & 0xffffffffull)) - 1.0;
    }
    const double hfsq = 0.5 * f * f;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (Lg[1] + w * (Lg[3] + w * Lg[5]));
    const double t2 = z * (Lg[0] + w * (Lg[2] + w * (Lg[4] + w * Lg[6])));
    const double r = t2 + t1;
    const double dk = (double)k;
    return s * (hfsq + r) + (dk * ln2_lo + c) - hfsq + f + dk * ln2_hi;
}

// msun s_atan.c
CTR_HD double atan_msun(double x) {
    const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                          1.57079632679489655800e+00};
    const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                          6.12323399573676603587e-17};
    const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                           -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                           6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                           -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    uint32_t ix = high_word(x);
    const bool negative = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    int id;
    if (ix >= 0x44100000u) {  // |x| >= 2^66
        if (x != x) return x;
        const double z = hi[3] + 7.52316384526264005e-37;  // + 2^-120
        return negative ? -z : z;
    }
    if (ix < 0x3fdc0000u) {  // |x| < 0.4375
        if (ix < 0x3e400000u) return x;
        id = -1;
    } else {
        x = absd(x);
        if (ix < 0x3ff30000u) {
            if (ix < 0x3fe60000u) {
                id = 0;
                x = (2.0 * x - 1.0) / (2.0 + x);
            } else {
                id = 1;
                x = (x - 1.0) / (x + 1.0);
            }
        } else if (ix < 0x40038000u) {
            id = 2;
            x = (x - 1.5) / (1.0 + 1.5 * x);
        } else {
            id = 3;
            x = -1.0 / x;
        }
    }
    const double z = x * x;
    const double w = z * z;
    const double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const double r = hi[id] - (x * (s1 + s2) - lo[id] - x);
    return negative ? -r : r;
}

// CDFs of the other closed-form models of the Python API (pybindings/stream/model.rs:740-900).  They live in the
// crate `probability` 0.20.3, whose source is not part of the reference tree and for which the reference holds no
// golden vectors: textbook definitions, parity with a Rust build UNPINNED.
CTR_HD double laplace_cdf(double x, double mu, double b) {
    if (x <= mu) return 0.5 * exp_msun((x - mu) / b);
    return 1.0 - 0.5 * exp_msun(-(x - mu) / b);
}
CTR_HD double cauchy_cdf(double x, double x0, double gamma) {
    const double frac_1_pi = 0.318309886183790671537767526745028724;
    return frac_1_pi * atan_msun((x - x0) / gamma) + 0.5;
}
// kind: 0 Gaussian(mean, std), 1 Laplace(mean, scale), 2 Cauchy(location, scale)
CTR_HD double two_parameter_cdf(int kind, double x, double p0, double p1) {
    return kind == 1 ? laplace_cdf(x, p0, p1) : (kind == 2 ? cauchy_cdf(x, p0, p1) : gaussian_cdf(x, p0, p1));
}

// Rust `as u32` from a float: truncate toward zero, saturate, NaN -> 0.
CTR_HD uint32_t f64_to_u32_sat(double v) {
    if (!(v > 0.0)) return 0u;
    if (v >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)v;
}
CTR_HD uint32_t f32_to_u32_sat(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)v;
}

// quantize.rs:284-308: weight that is distributed according to the float CDF; the remaining
// (max-min+1) units guarantee a nonzero probability for every symbol in the support.
CTR_HD bool leaky_free_weight(int32_t min_symbol, int32_t max_symbol, double &free_weight) {
    if (!(max_symbol > min_symbol)) return false;
    const uint32_t support_minus_one = (uint32_t)max_symbol - (uint32_t)min_symbol;
    if (support_minus_one > kQuantileMask) return false;
    free_weight = (double)(kQuantileMask - support_minus_one);
    return true;
}

// quantize.rs:539-547: left-sided cumulative of `symbol` (index i = symbol - min_symbol).
CTR_HD uint32_t leaky_gaussian_left(double free_weight, int32_t min_symbol, double mean, double std, uint32_t i) {
    if (i == 0) return 0u;
    const int32_t symbol = (int32_t)((uint32_t)min_symbol + i);
    return f64_to_u32_sat(free_weight * gaussian_cdf((double)symbol - 0.5, mean, std)) + i;
}

}  // namespace mm
}  // namespace ctr
