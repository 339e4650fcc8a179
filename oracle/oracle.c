/*
 * oracle.c -- CPU restatement of constriction's ANS / range-coder hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Plain C11, compiled with
 *   gcc -O2 -ffp-contract=off   (Rust never contracts a*b+c into an fma)
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Parity is pinned by the reference's own golden
 * vectors (tests/test_oracle_golden.py).
 *
 * Third-party arithmetic that is NOT under /root/reference:
 *   probability 0.20.3 (Cargo.lock:510-517)  Gaussian::distribution
 *     -> special 0.10.3 (Cargo.lock:779-785) Error::error
 *     -> libm 0.2.16    (Cargo.lock:358-361) erf / exp
 *   libm's erf.rs / exp.rs are ports of FreeBSD msun s_erf.c / e_exp.c (via
 *   musl).  orc_erf / orc_exp restate that published algorithm (Sun
 *   Microsystems 1993 rational approximations), operation for operation.  The
 *   coefficients were cross-checked (decimal vs. IEEE hex form) and the result
 *   is within 1 ulp of glibc's erf/exp on random and special arguments
 *   (tests/test_host_math.py, 1.4e5 points).  Parity of erf at the last-ulp level against
 *   the real Rust build is pinned only through the goldens (G1-G8).
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* math                                                                 */
/* ------------------------------------------------------------------ */

static inline uint32_t hi_word(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    return (uint32_t)(b >> 32);
}
static inline double clear_low_word(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    b &= 0xffffffff00000000ull;
    memcpy(&x, &b, 8);
    return x;
}

/* libm scalbn.rs (musl scalbn.c) */
static double orc_scalbn(double x, int n) {
    const double x1p1023 = 0x1p1023, x1p53 = 0x1p53, x1p_1022 = 0x1p-1022;
    double y = x;
    if (n > 1023) {
        y *= x1p1023;
        n -= 1023;
        if (n > 1023) {
            y *= x1p1023;
            n -= 1023;
            if (n > 1023) n = 1023;
        }
    } else if (n < -1022) {
        y *= x1p_1022 * x1p53;
        n += 1022 - 53;
        if (n < -1022) {
            y *= x1p_1022 * x1p53;
            n += 1022 - 53;
            if (n < -1022) n = -1022;
        }
    }
    uint64_t u = (uint64_t)(0x3ff + n) << 52;
    double s;
    memcpy(&s, &u, 8);
    return y * s;
}

/* libm exp.rs  <-  FreeBSD e_exp.c */
double orc_exp(double x) {
    static const double ln2hi = 6.93147180369123816490e-01, /* 0x3fe62e42, 0xfee00000 */
        ln2lo = 1.90821492927058770002e-10,                 /* 0x3dea39ef, 0x35793c76 */
        invln2 = 1.44269504088896338700e+00,                /* 0x3ff71547, 0x652b82fe */
        P1 = 1.66666666666666019037e-01,                    /* 0x3FC55555, 0x5555553E */
        P2 = -2.77777777770155933842e-03,                   /* 0xBF66C16C, 0x16BEBD93 */
        P3 = 6.61375632143793436117e-05,                    /* 0x3F11566A, 0xAF25DE2C */
        P4 = -1.65339022054652515390e-06,                   /* 0xBEBBBD41, 0xC5D26BF1 */
        P5 = 4.13813679705723846039e-08;                    /* 0x3E663769, 0x72BEA4D0 */
    double hi, lo, c, xx, y;
    int k, sign;
    uint32_t hx = hi_word(x);
    sign = (int)(hx >> 31);
    hx &= 0x7fffffff;

    if (hx >= 0x4086232b) { /* |x| >= 708.39 or NaN */
        if (isnan(x)) return x;
        if (x > 709.782712893383973096) return x * 0x1p1023;
        if (x < -708.39641853226410622) {
            if (x < -745.13321910194110842) return 0.0;
        }
    }
    if (hx > 0x3fd62e42) {     /* |x| > 0.5 ln2 */
        if (hx >= 0x3ff0a2b2) { /* |x| >= 1.5 ln2 */
            k = (int)(invln2 * x + (sign ? -0.5 : 0.5));
        } else {
            k = 1 - sign - sign;
        }
        hi = x - (double)k * ln2hi;
        lo = (double)k * ln2lo;
        x = hi - lo;
    } else if (hx > 0x3e300000) { /* |x| > 2**-28 */
        k = 0;
        hi = x;
        lo = 0.0;
    } else {
        return 1.0 + x;
    }
    xx = x * x;
    c = x - xx * (P1 + xx * (P2 + xx * (P3 + xx * (P4 + xx * P5))));
    y = 1.0 + (x * c / (2.0 - c) - lo + hi);
    if (k == 0) return y;
    return orc_scalbn(y, k);
}

/* libm erf.rs  <-  FreeBSD s_erf.c (musl layout: erfc1 / erfc2 helpers) */
static const double erx = 8.45062911510467529297e-01, /* 0x3FEB0AC1, 0x60000000 */
    efx8 = 1.02703333676410069053e+00,                /* 0x3FF06EBA, 0x8214DB69 */
    pp0 = 1.28379167095512558561e-01, pp1 = -3.25042107247001499370e-01,
    pp2 = -2.84817495755985104766e-02, pp3 = -5.77027029648944159157e-03,
    pp4 = -2.37630166566501626084e-05, qq1 = 3.97917223959155352819e-01,
    qq2 = 6.50222499887672944485e-02, qq3 = 5.08130628187576562776e-03,
    qq4 = 1.32494738004321644526e-04, qq5 = -3.96022827877536812320e-06,
    /* erf in [0.84375,1.25] */
    pa0 = -2.36211856075265944077e-03, pa1 = 4.14856118683748331666e-01,
    pa2 = -3.72207876035701323847e-01, pa3 = 3.18346619901161753674e-01,
    pa4 = -1.10894694282396677476e-01, pa5 = 3.54783043256182359371e-02,
    pa6 = -2.16637559486879084300e-03, qa1 = 1.06420880400844228286e-01,
    qa2 = 5.40397917702171048937e-01, qa3 = 7.18286544141962662868e-02,
    qa4 = 1.26171219808761642112e-01, qa5 = 1.36370839120290507362e-02,
    qa6 = 1.19844998467991074170e-02,
    /* erfc in [1.25,1/0.35] */
    ra0 = -9.86494403484714822705e-03, ra1 = -6.93858572707181764372e-01,
    ra2 = -1.05586262253232909814e+01, ra3 = -6.23753324503260060396e+01,
    ra4 = -1.62396669462573470355e+02, ra5 = -1.84605092906711035994e+02,
    ra6 = -8.12874355063065934246e+01, ra7 = -9.81432934416914548592e+00,
    sa1 = 1.96512716674392571292e+01, sa2 = 1.37657754143519042600e+02,
    sa3 = 4.34565877475229228821e+02, sa4 = 6.45387271733267880336e+02,
    sa5 = 4.29008140027567833386e+02, sa6 = 1.08635005541779435134e+02,
    sa7 = 6.57024977031928170135e+00, sa8 = -6.04244152148580987438e-02,
    /* erfc in [1/.35,28] */
    rb0 = -9.86494292470009928597e-03, rb1 = -7.99283237680523006574e-01,
    rb2 = -1.77579549177547519889e+01, rb3 = -1.60636384855821916062e+02,
    rb4 = -6.37566443368389627722e+02, rb5 = -1.02509513161107724954e+03,
    rb6 = -4.83519191608651397019e+02, sb1 = 3.03380607434824582924e+01,
    sb2 = 3.25792512996573918826e+02, sb3 = 1.53672958608443695994e+03,
    sb4 = 3.19985821950859553908e+03, sb5 = 2.55305040643316442583e+03,
    sb6 = 4.74528541206955367215e+02, sb7 = -2.24409524465858183362e+01;

static double erfc1(double x) {
    double s = fabs(x) - 1.0;
    double P = pa0 + s * (pa1 + s * (pa2 + s * (pa3 + s * (pa4 + s * (pa5 + s * pa6)))));
    double Q = 1.0 + s * (qa1 + s * (qa2 + s * (qa3 + s * (qa4 + s * (qa5 + s * qa6)))));
    return 1.0 - erx - P / Q;
}

static double erfc2(uint32_t ix, double x) {
    double s, R, S, z;
    if (ix < 0x3ff40000) /* |x| < 1.25 */
        return erfc1(x);
    x = fabs(x);
    s = 1.0 / (x * x);
    if (ix < 0x4006db6d) { /* |x| < 1/.35 ~ 2.85714 */
        R = ra0 + s * (ra1 + s * (ra2 + s * (ra3 + s * (ra4 + s * (ra5 + s * (ra6 + s * ra7))))));
        S = 1.0 + s * (sa1 + s * (sa2 + s * (sa3 + s * (sa4 + s * (sa5 + s * (sa6 + s * (sa7 + s * sa8)))))));
    } else { /* |x| > 1/.35 */
        R = rb0 + s * (rb1 + s * (rb2 + s * (rb3 + s * (rb4 + s * (rb5 + s * rb6)))));
        S = 1.0 + s * (sb1 + s * (sb2 + s * (sb3 + s * (sb4 + s * (sb5 + s * (sb6 + s * sb7))))));
    }
    z = clear_low_word(x);
    return orc_exp(-z * z - 0.5625) * orc_exp((z - x) * (z + x) + R / S) / x;
}

double orc_erf(double x) {
    double r, s, z, y;
    uint32_t ix = hi_word(x);
    int sign = (int)(ix >> 31);
    ix &= 0x7fffffff;
    if (ix >= 0x7ff00000) /* erf(nan)=nan, erf(+-inf)=+-1 */
        return 1.0 - 2.0 * (double)sign + 1.0 / x;
    if (ix < 0x3feb0000) {   /* |x| < 0.84375 */
        if (ix < 0x3e300000) /* |x| < 2**-28 */
            return 0.125 * (8.0 * x + efx8 * x);
        z = x * x;
        r = pp0 + z * (pp1 + z * (pp2 + z * (pp3 + z * pp4)));
        s = 1.0 + z * (qq1 + z * (qq2 + z * (qq3 + z * (qq4 + z * qq5))));
        y = r / s;
        return x + x * y;
    }
    if (ix < 0x40180000) /* 0.84375 <= |x| < 6 */
        y = 1.0 - erfc2(ix, x);
    else
        y = 1.0 - 0x1p-1022;
    return sign ? -y : y;
}

/* probability-0.20.3 src/distribution/gaussian.rs Distribution::distribution
 * (call sites: quantize.rs:546,558) */
double orc_gaussian_cdf(double x, double mean, double std) {
    const double sqrt2 = 1.41421356237309504880168872420969808; /* core::f64::consts::SQRT_2 */
    return (1.0 + orc_erf((x - mean) / (std * sqrt2))) / 2.0;
}

/* Rust `f64 as u32` / `f32 as u32`: truncating, saturating, NaN -> 0 */
static inline uint32_t f64_as_u32(double v) {
    if (!(v > 0.0)) return 0; /* also NaN */
    if (v >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)v;
}
static inline uint32_t f32_as_u32(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)v;
}

/* ------------------------------------------------------------------ */
/* QuantizedGaussian (LeakyQuantizer<f64,i32,u32,24>)                   */
/* ------------------------------------------------------------------ */

/* quantize.rs:284-308: free_weight = (2^24 - 1) - (max - min), as f64 */
static int qgauss_free_weight(int32_t min_sym, int32_t max_sym, double *fw) {
    if (!(max_sym > min_sym)) return ORC_ERR_BAD_MODEL;
    uint32_t support_minus_one = (uint32_t)max_sym - (uint32_t)min_sym; /* wrapping_sub(..).as_() */
    uint32_t max_probability = 0xffffffffu >> (32 - ORC_PRECISION);
    if (support_minus_one > max_probability) return ORC_ERR_BAD_MODEL;
    *fw = (double)(max_probability - support_minus_one);
    return ORC_OK;
}

/* quantize.rs:475-486 slack() for Symbol=i32, Probability=u32: mask is all ones */
static inline uint32_t qslack(int32_t symbol, int32_t min_sym) { return (uint32_t)symbol - (uint32_t)min_sym; }

static inline uint32_t qgauss_left(double fw, int32_t min_sym, double mean, double std, int32_t s) {
    if (s == min_sym) return 0; /* quantize.rs:539-543 */
    return f64_as_u32(fw * orc_gaussian_cdf((double)s - 0.5, mean, std)) + qslack(s, min_sym); /* :545-547 */
}
static inline uint32_t qgauss_right(double fw, int32_t min_sym, int32_t max_sym, double mean, double std,
                                    int32_t s) {
    if (s == max_sym) return ORC_TOTAL; /* quantize.rs:550-555 */
    return f64_as_u32(fw * orc_gaussian_cdf((double)s + 0.5, mean, std)) + qslack(s, min_sym) + 1u; /* :557-559 */
}

int orc_qgauss_left_prob(int32_t min_sym, int32_t max_sym, double mean, double std, int32_t symbol,
                         uint32_t *left, uint32_t *prob) {
    double fw;
    int rc = qgauss_free_weight(min_sym, max_sym, &fw);
    if (rc) return rc;
    if (!(std > 0.0)) return ORC_ERR_BAD_MODEL; /* pybindings/stream/model.rs:654-657 */
    if (symbol < min_sym || symbol > max_sym) return ORC_ERR_IMPOSSIBLE_SYMBOL; /* quantize.rs:533-535 */
    uint32_t l = qgauss_left(fw, min_sym, mean, std, symbol);
    uint32_t r = qgauss_right(fw, min_sym, max_sym, mean, std, symbol);
    uint32_t p = r - l; /* wrapping_sub, quantize.rs:562-565 */
    if (p == 0) return ORC_ERR_BAD_MODEL;
    *left = l;
    *prob = p;
    return ORC_OK;
}

int orc_qgauss_cdf(int32_t min_sym, int32_t max_sym, double mean, double std, uint32_t *cdf) {
    double fw;
    int rc = qgauss_free_weight(min_sym, max_sym, &fw);
    if (rc) return rc;
    if (!(std > 0.0)) return ORC_ERR_BAD_MODEL;
    size_t n = (size_t)((int64_t)max_sym - (int64_t)min_sym) + 1;
    for (size_t i = 0; i < n; i++) cdf[i] = qgauss_left(fw, min_sym, mean, std, (int32_t)((int64_t)min_sym + (int64_t)i));
    cdf[n] = ORC_TOTAL;
    return ORC_OK;
}

/* quantize.rs:580-779.  The reference starts from an approximate inverse CDF and
 * steps / bisects with the formulas above until left <= q < right; that bin is
 * unique (left is strictly increasing in the symbol thanks to the slack term), so
 * the starting guess does not influence the result.  Here: plain bisection. */
int orc_qgauss_quantile(int32_t min_sym, int32_t max_sym, double mean, double std, uint32_t quantile,
                        int32_t *symbol, uint32_t *left, uint32_t *prob) {
    double fw;
    int rc = qgauss_free_weight(min_sym, max_sym, &fw);
    if (rc) return rc;
    if (!(std > 0.0)) return ORC_ERR_BAD_MODEL;
    if (quantile >= ORC_TOTAL) return ORC_ERR_INVALID_DATA;
    int64_t lo = min_sym, hi = max_sym; /* invariant: left(lo) <= q */
    while (lo < hi) {
        int64_t mid = lo + (hi - lo + 1) / 2;
        if (qgauss_left(fw, min_sym, mean, std, (int32_t)mid) <= quantile)
            lo = mid;
        else
            hi = mid - 1;
    }
    uint32_t l = qgauss_left(fw, min_sym, mean, std, (int32_t)lo);
    uint32_t r = qgauss_right(fw, min_sym, max_sym, mean, std, (int32_t)lo);
    *symbol = (int32_t)lo;
    *left = l;
    *prob = r - l;
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* log1p, atan (libm 0.2.16 <- musl <- FreeBSD msun s_log1p.c, s_atan.c)  */
/* ------------------------------------------------------------------ */
static inline uint64_t f64_bits(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}
static inline double f64_from_bits(uint64_t u) {
    double x;
    memcpy(&x, &u, 8);
    return x;
}

/* call sites: categorical.rs:11,98-126,166-172 */
double orc_log1p(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                        Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    uint64_t ui = f64_bits(x);
    uint32_t hx = (uint32_t)(ui >> 32);
    int k = 1;
    double f = 0.0, c = 0.0;
    if (hx < 0x3fda827au || (hx >> 31)) { /* 1+x < sqrt(2)+ */
        if (hx >= 0xbff00000u) {          /* x <= -1.0 */
            if (x == -1.0) return x / 0.0; /* log1p(-1) = -inf */
            return (x - x) / 0.0;          /* log1p(x<-1) = NaN */
        }
        if ((hx << 1) < (0x3ca00000u << 1)) return x; /* |x| < 2**-53 */
        if (hx <= 0xbfd2bec4u) {                      /* sqrt(2)/2- <= 1+x < sqrt(2)+ */
            k = 0;
            c = 0.0;
            f = x;
        }
    } else if (hx >= 0x7ff00000u) {
        return x;
    }
    if (k) {
        ui = f64_bits(1.0 + x);
        uint32_t hu = (uint32_t)(ui >> 32);
        hu += 0x3ff00000u - 0x3fe6a09eu;
        k = (int)(hu >> 20) - 0x3ff;
        /* correction term ~ log(1+x)-log(u), avoid underflow in c/u */
        if (k < 54) {
            c = k >= 2 ? 1.0 - (f64_from_bits(ui) - x) : x - (f64_from_bits(ui) - 1.0);
            c /= f64_from_bits(ui);
        } else {
            c = 0.0;
        }
        /* reduce u into [sqrt(2)/2, sqrt(2)] */
        hu = (hu & 0x000fffffu) + 0x3fe6a09eu;
        ui = ((uint64_t)hu << 32) | (ui & 0xffffffffull);
        f = f64_from_bits(ui) - 1.0;
    }
    double hfsq = 0.5 * f * f;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    double R = t2 + t1;
    double dk = (double)k;
    return s * (hfsq + R) + (dk * ln2_lo + c) - hfsq + f + dk * ln2_hi;
}

double orc_atan(double x) {
    static const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01,
                                     1.57079632679489655800e+00};
    static const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17,
                                     6.12323399573676603587e-17};
    static const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                                  -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                                  6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                                  -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    uint32_t ix = (uint32_t)(f64_bits(x) >> 32);
    const uint32_t sign = ix >> 31;
    ix &= 0x7fffffffu;
    int id;
    if (ix >= 0x44100000u) { /* |x| >= 2^66 */
        if (x != x) return x;
        double z = atanhi[3] + 0x1p-120f;
        return sign ? -z : z;
    }
    if (ix < 0x3fdc0000u) {     /* |x| < 0.4375 */
        if (ix < 0x3e400000u) return x; /* |x| < 2^-27 */
        id = -1;
    } else {
        x = fabs(x);
        if (ix < 0x3ff30000u) {     /* |x| < 1.1875 */
            if (ix < 0x3fe60000u) { /* 7/16 <= |x| < 11/16 */
                id = 0;
                x = (2.0 * x - 1.0) / (2.0 + x);
            } else { /* 11/16 <= |x| < 19/16 */
                id = 1;
                x = (x - 1.0) / (x + 1.0);
            }
        } else if (ix < 0x40038000u) { /* |x| < 2.4375 */
            id = 2;
            x = (x - 1.5) / (1.0 + 1.5 * x);
        } else { /* 2.4375 <= |x| < 2^66 */
            id = 3;
            x = -1.0 / x;
        }
    }
    double z = x * x;
    double w = z * z;
    /* break sum from i=0 to 10 aT[i]z**(i+1) into odd and even poly */
    double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - (x * (s1 + s2) - atanlo[id] - x);
    return sign ? -z : z;
}

/* ------------------------------------------------------------------ */
/* Other leakily quantised distributions (pybindings/stream/model.rs:740-960).  Their CDFs live in the crate
 * `probability 0.20.3`, whose source is not available here and for which the reference holds NO golden vectors:
 * the formulas below are the textbook definitions (PARITY UNPINNED for these three models). */
/* ------------------------------------------------------------------ */
double orc_laplace_cdf(double x, double mu, double b) {
    if (x <= mu) return 0.5 * orc_exp((x - mu) / b);
    return 1.0 - 0.5 * orc_exp(-(x - mu) / b);
}
double orc_cauchy_cdf(double x, double x0, double gamma) {
    const double frac_1_pi = 0.318309886183790671537767526745028724; /* core::f64::consts::FRAC_1_PI */
    return frac_1_pi * orc_atan((x - x0) / gamma) + 0.5;
}

/* LeakyQuantizer table for any CDF of two parameters (quantize.rs:525-568 with D = that distribution) */
int orc_qdist_cdf(int kind, int32_t min_sym, int32_t max_sym, double p0, double p1, uint32_t *cdf) {
    double fw;
    int rc = qgauss_free_weight(min_sym, max_sym, &fw);
    if (rc) return rc;
    if (!(p1 > 0.0)) return ORC_ERR_BAD_MODEL;
    size_t n = (size_t)((int64_t)max_sym - (int64_t)min_sym) + 1;
    cdf[0] = 0;
    for (size_t i = 1; i < n; i++) {
        double x = (double)(int32_t)((int64_t)min_sym + (int64_t)i) - 0.5;
        double c = kind == 1 ? orc_laplace_cdf(x, p0, p1) : (kind == 2 ? orc_cauchy_cdf(x, p0, p1) : orc_gaussian_cdf(x, p0, p1));
        cdf[i] = f64_as_u32(fw * c) + (uint32_t)i;
    }
    cdf[n] = ORC_TOTAL;
    return ORC_OK;
}

/* Binomial(n, p) over {0..n}: CDF(x + 0.5) = sum_{i <= x} pmf(i), pmf by the recurrence
 * pmf(i+1) = pmf(i) * (n-i)/(i+1) * p/(1-p) started from the mode in log space (PARITY UNPINNED: the
 * reference evaluates the regularised incomplete beta function of the `special` crate). */
int orc_binomial_cdf(int32_t n, double p, uint32_t *cdf) {
    if (n < 1 || !(p >= 0.0) || !(p <= 1.0)) return ORC_ERR_BAD_MODEL;
    double fw;
    int rc = qgauss_free_weight(0, n, &fw);
    if (rc) return rc;
    double *pmf = (double *)malloc(((size_t)n + 1) * sizeof(double));
    if (!pmf) return ORC_ERR_BAD_MODEL;
    const double q = 1.0 - p;
    int64_t mode = (int64_t)(((double)n + 1.0) * p);
    if (mode > n) mode = n;
    pmf[mode] = 1.0;
    for (int64_t i = mode; i < n; i++) pmf[i + 1] = q > 0.0 ? pmf[i] * ((double)(n - i) / (double)(i + 1)) * (p / q) : 0.0;
    for (int64_t i = mode; i > 0; i--) pmf[i - 1] = p > 0.0 ? pmf[i] * ((double)i / (double)(n - i + 1)) * (q / p) : 0.0;
    double norm = 0.0;
    for (int64_t i = 0; i <= n; i++) norm = norm + pmf[i];
    double cum = 0.0;
    cdf[0] = 0;
    for (int64_t i = 1; i <= n; i++) {
        cum = cum + pmf[i - 1];
        cdf[i] = f64_as_u32(fw * (cum / norm)) + (uint32_t)i;
    }
    cdf[(size_t)n + 1] = ORC_TOTAL;
    free(pmf);
    return ORC_OK;
}

/* categorical.rs:56-177 perfectly_quantized_probabilities (PRECISION = 24, Probability = u32), computed in f64
 * whatever the caller's float type (`F: Into<f64>`), followed by contiguous.rs:301-312 (weights -> CDF).
 * Order-sensitive details that decide ties: `sort_by` is stable (descending win), `max_by` returns the LAST maximum,
 * `min_by` the FIRST minimum, and the slots stay in sorted order until the final sort by original index. */
typedef struct {
    size_t original_index;
    double prob;
    uint32_t weight;
    double win, loss;
} orc_slot;

static void slots_stable_sort_desc_win(orc_slot *a, orc_slot *tmp, size_t n) {
    for (size_t width = 1; width < n; width *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * width) {
            size_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                /* take from the right run only if it is strictly greater: equal elements keep their order */
                if (a[j].win > a[i].win)
                    tmp[k++] = a[j++];
                else
                    tmp[k++] = a[i++];
            }
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, n * sizeof *a);
    }
}

static int cat_perfect_weights(const double *probs, size_t n, uint32_t *weights) {
    if (n < 2 || n > 0xffffffffull) return ORC_ERR_BAD_MODEL;
    uint32_t remaining = ORC_TOTAL - (uint32_t)n; /* wrapping_sub in the reference */
    double norm = 0.0;
    for (size_t i = 0; i < n; i++) norm = norm + probs[i];
    if (!isnormal(norm) || signbit(norm)) return ORC_ERR_BAD_MODEL;
    const double scale = (double)remaining / norm;
    orc_slot *slots = (orc_slot *)malloc(2 * n * sizeof *slots);
    if (!slots) return ORC_ERR_BAD_MODEL;
    orc_slot *tmp = slots + n;
    for (size_t i = 0; i < n; i++) {
        const double prob = probs[i];
        if (prob < 0.0) {
            free(slots);
            return ORC_ERR_BAD_MODEL;
        }
        const uint32_t current = f64_as_u32(prob * scale);
        remaining -= current;
        const uint32_t weight = current + 1u;
        slots[i].original_index = i;
        slots[i].prob = prob;
        slots[i].weight = weight;
        slots[i].win = prob * orc_log1p(1.0 / (double)weight);
        slots[i].loss = weight == 1u ? INFINITY : -prob * orc_log1p(-1.0 / (double)weight);
    }
    while (remaining != 0u) {
        slots_stable_sort_desc_win(slots, tmp, n);
        const size_t batch = remaining < n ? remaining : n;
        for (size_t i = 0; i < batch; i++) {
            slots[i].weight += 1u;
            slots[i].win = slots[i].prob * orc_log1p(1.0 / (double)slots[i].weight);
            slots[i].loss = -slots[i].prob * orc_log1p(-1.0 / (double)slots[i].weight);
        }
        remaining -= (uint32_t)batch;
    }
    for (;;) {
        size_t buyer = 0, seller = 0;
        for (size_t i = 1; i < n; i++) {
            if (slots[i].win >= slots[buyer].win) buyer = i;   /* max_by: last maximum */
            if (slots[i].loss < slots[seller].loss) seller = i; /* min_by: first minimum */
        }
        if (buyer == seller) break;
        if (slots[buyer].win <= slots[seller].loss) break;
        slots[seller].weight -= 1u;
        slots[seller].win = -INFINITY;
        slots[seller].loss = slots[seller].weight == 1u ? INFINITY : -slots[seller].prob * orc_log1p(-1.0 / (double)slots[seller].weight);
        slots[buyer].weight += 1u;
        slots[buyer].loss = INFINITY;
        slots[buyer].win = slots[buyer].prob * orc_log1p(1.0 / (double)slots[buyer].weight);
    }
    for (size_t i = 0; i < n; i++) weights[slots[i].original_index] = slots[i].weight;
    free(slots);
    return ORC_OK;
}

int orc_cat_perfect_cdf_f64(const double *pmf, size_t n, uint32_t *cdf) {
    uint32_t *w = (uint32_t *)malloc((n ? n : 1) * sizeof *w);
    if (!w) return ORC_ERR_BAD_MODEL;
    int rc = cat_perfect_weights(pmf, n, w);
    if (!rc) { /* contiguous.rs:471-497 from_nonzero_fixed_point_probabilities */
        uint32_t acc = 0;
        for (size_t i = 0; i < n; i++) {
            cdf[i] = acc;
            acc += w[i];
        }
        cdf[n] = ORC_TOTAL;
        if (acc != ORC_TOTAL) rc = ORC_ERR_BAD_MODEL;
    }
    free(w);
    return rc;
}
int orc_cat_perfect_cdf_f32(const float *pmf, size_t n, uint32_t *cdf) {
    double *d = (double *)malloc((n ? n : 1) * sizeof *d);
    if (!d) return ORC_ERR_BAD_MODEL;
    for (size_t i = 0; i < n; i++) d[i] = (double)pmf[i]; /* `prob.into()` */
    int rc = orc_cat_perfect_cdf_f64(d, n, cdf);
    free(d);
    return rc;
}

/* ------------------------------------------------------------------ */
/* Categorical                                                          */
/* ------------------------------------------------------------------ */

#define DEFINE_CAT(SUFFIX, F, AS_U32, ISNORMAL, EPS)                                                          \
    /* categorical.rs:30-45 / lazy_contiguous.rs:146-160: scale = F(2^24 - n) / sum(pmf) */                   \
    static int cat_scale_##SUFFIX(const F *pmf, size_t n, F *scale) {                                         \
        if (n < 2 || n >= (size_t)ORC_TOTAL - 1) return ORC_ERR_BAD_MODEL;                                    \
        uint32_t free_weight = ORC_TOTAL - (uint32_t)n;                                                       \
        F norm = (F)0;                                                                                        \
        for (size_t i = 0; i < n; i++) norm = norm + pmf[i]; /* Iterator::sum: sequential, in F */            \
        if (!ISNORMAL(norm) || signbit(norm)) return ORC_ERR_BAD_MODEL;                                       \
        *scale = (F)free_weight / norm;                                                                       \
        return ORC_OK;                                                                                        \
    }                                                                                                         \
    /* categorical.rs:47-53 + contiguous.rs:499-512 */                                                        \
    int orc_cat_cdf_##SUFFIX(const F *pmf, size_t n, uint32_t *cdf) {                                         \
        F scale;                                                                                              \
        int rc = cat_scale_##SUFFIX(pmf, n, &scale);                                                          \
        if (rc) return rc;                                                                                    \
        F cum = (F)0;                                                                                         \
        uint32_t slack = 0;                                                                                   \
        for (size_t i = 0; i < n; i++) {                                                                      \
            cdf[i] = AS_U32(cum * scale) + slack;                                                             \
            cum = cum + pmf[i];                                                                               \
            slack += 1u;                                                                                      \
        }                                                                                                     \
        cdf[n] = ORC_TOTAL;                                                                                   \
        return ORC_OK;                                                                                        \
    }                                                                                                         \
    /* lazy_contiguous.rs:228-257 */                                                                          \
    int orc_cat_lazy_left_prob_##SUFFIX(const F *pmf, size_t n, int32_t symbol, uint32_t *left,               \
                                        uint32_t *prob) {                                                     \
        F scale;                                                                                              \
        int rc = cat_scale_##SUFFIX(pmf, n, &scale);                                                          \
        if (rc) return rc;                                                                                    \
        /* internals.rs:528-531 `symbol as usize`: negatives become huge -> None */                           \
        if (symbol < 0 || (size_t)symbol >= n) return ORC_ERR_IMPOSSIBLE_SYMBOL;                              \
        size_t s = (size_t)symbol;                                                                            \
        F lc = (F)0;                                                                                          \
        for (size_t i = 0; i < s; i++) lc = lc + pmf[i];                                                      \
        uint32_t l = AS_U32(lc * scale) + (uint32_t)s;                                                        \
        F rcum = lc + pmf[s];                                                                                 \
        uint32_t r = (s == n - 1) ? ORC_TOTAL : AS_U32(rcum * scale) + (uint32_t)s + 1u;                      \
        uint32_t p = r - l;                                                                                   \
        if (p == 0) return ORC_ERR_BAD_MODEL;                                                                 \
        *left = l;                                                                                            \
        *prob = p;                                                                                            \
        return ORC_OK;                                                                                        \
    }                                                                                                         \
    /* lazy_contiguous.rs:268-330, statement for statement */                                                 \
    int orc_cat_lazy_quantile_##SUFFIX(const F *pmf, size_t n, uint32_t quantile, int32_t *symbol,            \
                                       uint32_t *left, uint32_t *prob) {                                      \
        F scale;                                                                                              \
        int rc = cat_scale_##SUFFIX(pmf, n, &scale);                                                          \
        if (rc) return rc;                                                                                    \
        F lcf = (F)0, rcf = (F)0;                                                                             \
        F enlarged = ((F)1 + EPS + EPS) * scale;                                                              \
        uint32_t sat = quantile > (uint32_t)n ? quantile - (uint32_t)n : 0;                                   \
        F lower_bound = (F)sat / enlarged;                                                                    \
        size_t it = 0, next_symbol = 0;                                                                       \
        while (it < n) {                                                                                      \
            F p = pmf[it++];                                                                                  \
            next_symbol += 1;                                                                                 \
            lcf = rcf;                                                                                        \
            rcf = rcf + p;                                                                                    \
            if (rcf >= lower_bound) break;                                                                    \
        }                                                                                                     \
        uint32_t lcum = AS_U32(lcf * scale) + (uint32_t)(next_symbol - 1);                                    \
        while (it < n) {                                                                                      \
            F p = pmf[it++];                                                                                  \
            uint32_t rcum = AS_U32(rcf * scale) + (uint32_t)next_symbol;                                      \
            if (rcum > quantile) {                                                                            \
                *symbol = (int32_t)(next_symbol - 1);                                                         \
                *left = lcum;                                                                                 \
                *prob = rcum - lcum;                                                                          \
                return ORC_OK;                                                                                \
            }                                                                                                 \
            lcum = rcum;                                                                                      \
            rcf = rcf + p;                                                                                    \
            next_symbol += 1;                                                                                 \
        }                                                                                                     \
        *symbol = (int32_t)(next_symbol - 1);                                                                 \
        *left = lcum;                                                                                         \
        *prob = ORC_TOTAL - lcum;                                                                             \
        return ORC_OK;                                                                                        \
    }

#define ISNORMAL_F(x) isnormal(x)
DEFINE_CAT(f32, float, f32_as_u32, ISNORMAL_F, 1.1920929e-07f)
DEFINE_CAT(f64, double, f64_as_u32, ISNORMAL_F, 2.2204460492503131e-16)

/* contiguous.rs:673-700 */
int orc_cdf_left_prob(const uint32_t *cdf, size_t n, int64_t index, uint32_t *left, uint32_t *prob) {
    if (index < 0 || (uint64_t)index >= n) return ORC_ERR_IMPOSSIBLE_SYMBOL;
    *left = cdf[index];
    *prob = cdf[index + 1] - cdf[index];
    return ORC_OK;
}

/* contiguous.rs:628-665: partition point of `cdf[..n] <= q`, minus one */
void orc_cdf_quantile(const uint32_t *cdf, size_t n, uint32_t q, size_t *index, uint32_t *left, uint32_t *prob) {
    size_t lo = 0, hi = n; /* first i in [0,n) with cdf[i] > q, or n */
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (cdf[mid] <= q)
            lo = mid + 1;
        else
            hi = mid;
    }
    size_t s = lo - 1; /* cdf[0] == 0 <= q, so lo >= 1 */
    *index = s;
    *left = cdf[s];
    *prob = cdf[s + 1] - cdf[s];
}

/* uniform.rs:44-77 (constructor) + :91-112 */
int orc_uniform_left_prob(uint32_t size, int32_t symbol, uint32_t *left, uint32_t *prob) {
    if (size < 2 || size > ORC_TOTAL) return ORC_ERR_BAD_MODEL;
    uint32_t per_bin = ORC_TOTAL / size, last = size - 1;
    if (symbol < 0) return ORC_ERR_IMPOSSIBLE_SYMBOL;
    uint32_t s = (uint32_t)symbol;
    uint32_t l = s * per_bin;
    if (s < last) {
        *left = l;
        *prob = per_bin;
        return ORC_OK;
    } else if (s == last) {
        *left = l;
        *prob = ORC_TOTAL - l;
        return ORC_OK;
    }
    return ORC_ERR_IMPOSSIBLE_SYMBOL;
}

/* uniform.rs:119-146 */
void orc_uniform_quantile(uint32_t size, uint32_t q, int32_t *symbol, uint32_t *left, uint32_t *prob) {
    uint32_t per_bin = ORC_TOTAL / size, last = size - 1;
    uint32_t guess = q / per_bin, rem = q % per_bin;
    if (guess < last) {
        *symbol = (int32_t)guess;
        *left = q - rem;
        *prob = per_bin;
    } else {
        *symbol = (int32_t)last;
        *left = last * per_bin;
        *prob = ORC_TOTAL - last * per_bin;
    }
}

/* ------------------------------------------------------------------ */
/* word vector (backends.rs:470-557 Vec<u32>)                           */
/* ------------------------------------------------------------------ */

static void vec_push(uint32_t **buf, size_t *len, size_t *cap, uint32_t w) {
    if (*len == *cap) {
        size_t nc = *cap ? *cap * 2 : 16;
        *buf = (uint32_t *)realloc(*buf, nc * sizeof(uint32_t));
        *cap = nc;
    }
    (*buf)[(*len)++] = w;
}

/* lib.rs:719-730 bit_array_to_chunks_truncated::<u64,u32>: number of chunks */
static inline int state_chunks(uint64_t state) {
    if (state == 0) return 0;
    return (state >> 32) ? 2 : 1;
}

/* ------------------------------------------------------------------ */
/* ANS coder                                                            */
/* ------------------------------------------------------------------ */

void orc_ans_init(orc_ans *c) { memset(c, 0, sizeof *c); }
void orc_ans_free(orc_ans *c) {
    free(c->bulk);
    memset(c, 0, sizeof *c);
}
void orc_ans_clear(orc_ans *c) { /* stack.rs:474-478 */
    c->len = 0;
    c->state = 0;
}

/* stack.rs:299-318 from_compressed + :440-462 read_initial_state (reads pop from the end) */
int orc_ans_from_compressed(orc_ans *c, const uint32_t *words, size_t n) {
    orc_ans_clear(c);
    for (size_t i = 0; i < n; i++) vec_push(&c->bulk, &c->len, &c->cap, words[i]);
    if (c->len == 0) {
        c->state = 0;
        return ORC_OK;
    }
    uint32_t first = c->bulk[--c->len];
    if (first == 0) return ORC_ERR_TRAILING_ZERO;
    uint64_t state = first;
    while (c->len > 0) {
        state = (state << 32) | c->bulk[--c->len];
        if (state >= (1ull << 32)) break;
    }
    c->state = state;
    return ORC_OK;
}

/* stack.rs:341-360 from_binary */
void orc_ans_from_binary(orc_ans *c, const uint32_t *words, size_t n) {
    orc_ans_clear(c);
    for (size_t i = 0; i < n; i++) vec_push(&c->bulk, &c->len, &c->cap, words[i]);
    uint64_t state = 1;
    while (state < (1ull << 32)) {
        if (c->len == 0) break;
        state = (state << 32) | c->bulk[--c->len];
    }
    c->state = state;
}

/* stack.rs:1014-1048 */
void orc_ans_encode(orc_ans *c, uint32_t left, uint32_t prob) {
    if ((c->state >> (64 - ORC_PRECISION)) >= (uint64_t)prob) {
        vec_push(&c->bulk, &c->len, &c->cap, (uint32_t)c->state);
        c->state >>= 32;
    }
    uint32_t remainder = (uint32_t)(c->state % prob);
    uint64_t prefix = c->state / prob;
    uint32_t quantile = left + remainder;
    c->state = (prefix << ORC_PRECISION) | (uint64_t)quantile;
}

/* stack.rs:1086 */
uint32_t orc_ans_peek_quantile(const orc_ans *c) { return (uint32_t)(c->state % (1ull << ORC_PRECISION)); }

/* stack.rs:1088-1097 */
void orc_ans_decode_advance(orc_ans *c, uint32_t left, uint32_t prob) {
    uint32_t quantile = orc_ans_peek_quantile(c);
    uint32_t remainder = quantile - left;
    c->state = (c->state >> ORC_PRECISION) * (uint64_t)prob + (uint64_t)remainder;
    if (c->state < (1ull << 32)) {
        if (c->len > 0) c->state = (c->state << 32) | c->bulk[--c->len];
    }
}

size_t orc_ans_num_words(const orc_ans *c) { return c->len + (size_t)state_chunks(c->state); }

size_t orc_ans_num_valid_bits(const orc_ans *c) { /* stack.rs:624-630 */
    size_t lz = c->state ? (size_t)__builtin_clzll(c->state) : 64;
    size_t v = 64 - lz;
    if (v < 1) v = 1;
    return 32 * c->len + v - 1;
}

int orc_ans_is_empty(const orc_ans *c) { return c->len == 0 && c->state == 0; } /* stack.rs:262-267 */

/* stack.rs:891-895 into_compressed / :1164-1177 CoderGuard: bulk ++ state chunks, low word first */
size_t orc_ans_get_compressed(const orc_ans *c, uint32_t *out) {
    memcpy(out, c->bulk, c->len * sizeof(uint32_t));
    size_t n = c->len;
    int k = state_chunks(c->state);
    if (k >= 1) out[n++] = (uint32_t)c->state;
    if (k == 2) out[n++] = (uint32_t)(c->state >> 32);
    return n;
}

/* stack.rs:944-955 into_binary (== CoderGuard<SEALED=true>, :1164-1171) */
int orc_ans_get_binary(const orc_ans *c, uint32_t *out, size_t *n_out) {
    if (c->state == 0) return ORC_ERR_NOT_SEALED;
    size_t valid_bits = 63 - (size_t)__builtin_clzll(c->state);
    if (valid_bits % 32 != 0) return ORC_ERR_NOT_SEALED;
    uint64_t truncated = c->state ^ (1ull << valid_bits);
    (void)truncated;
    memcpy(out, c->bulk, c->len * sizeof(uint32_t));
    size_t n = c->len;
    /* The Python API goes through get_binary = CoderGuard<SEALED=true> (:1164-1171): it checks
     * that the top chunk of the state is exactly 1 and appends the remaining chunks of the
     * untruncated iterator, i.e. exactly the low word when the state has two chunks (even
     * if that word is zero) and nothing when state == 1. */
    if (valid_bits == 32) out[n++] = (uint32_t)c->state;
    *n_out = n;
    return ORC_OK;
}

/* stack.rs:1117-1127 + backends.rs Vec seek: only forward (truncate) */
int orc_ans_seek(orc_ans *c, size_t pos, uint64_t state) {
    if (pos > c->len) return ORC_ERR_SEEK;
    c->len = pos;
    c->state = state;
    return ORC_OK;
}

int orc_ans_encode_iid_reverse(orc_ans *c, const int32_t *symbols, size_t n, const uint32_t *cdf,
                               int32_t min_sym, size_t alphabet) {
    for (size_t i = n; i-- > 0;) { /* stack.rs:835-849 `.rev()` + mod.rs:592-607 */
        uint32_t left, prob;
        int rc = orc_cdf_left_prob(cdf, alphabet, (int64_t)symbols[i] - (int64_t)min_sym, &left, &prob);
        if (rc) return rc;
        orc_ans_encode(c, left, prob);
    }
    return ORC_OK;
}

void orc_ans_decode_iid(orc_ans *c, int32_t *symbols, size_t n, const uint32_t *cdf, int32_t min_sym,
                        size_t alphabet) {
    for (size_t i = 0; i < n; i++) { /* mod.rs:1274-1297 */
        size_t idx;
        uint32_t left, prob;
        orc_cdf_quantile(cdf, alphabet, orc_ans_peek_quantile(c), &idx, &left, &prob);
        orc_ans_decode_advance(c, left, prob);
        symbols[i] = (int32_t)((int64_t)min_sym + (int64_t)idx);
    }
}

int orc_ans_encode_indexed_reverse(orc_ans *c, const int32_t *symbols, const uint32_t *model_idx, size_t n,
                                   const uint32_t *cdfs, size_t stride, int32_t min_sym, size_t alphabet) {
    for (size_t i = n; i-- > 0;) {
        uint32_t left, prob;
        const uint32_t *cdf = cdfs + (size_t)model_idx[i] * stride;
        int rc = orc_cdf_left_prob(cdf, alphabet, (int64_t)symbols[i] - (int64_t)min_sym, &left, &prob);
        if (rc) return rc;
        orc_ans_encode(c, left, prob);
    }
    return ORC_OK;
}

void orc_ans_decode_indexed(orc_ans *c, int32_t *symbols, const uint32_t *model_idx, size_t n,
                            const uint32_t *cdfs, size_t stride, int32_t min_sym, size_t alphabet) {
    for (size_t i = 0; i < n; i++) {
        size_t idx;
        uint32_t left, prob;
        const uint32_t *cdf = cdfs + (size_t)model_idx[i] * stride;
        orc_cdf_quantile(cdf, alphabet, orc_ans_peek_quantile(c), &idx, &left, &prob);
        orc_ans_decode_advance(c, left, prob);
        symbols[i] = (int32_t)((int64_t)min_sym + (int64_t)idx);
    }
}

int orc_ans_encode_qgauss_lazy_reverse(orc_ans *c, const int32_t *symbols, size_t n, int32_t min_sym,
                                       int32_t max_sym, const double *means, const double *stds,
                                       int per_symbol_params) {
    for (size_t i = n; i-- > 0;) {
        uint32_t left, prob;
        double m = per_symbol_params ? means[i] : means[0];
        double s = per_symbol_params ? stds[i] : stds[0];
        int rc = orc_qgauss_left_prob(min_sym, max_sym, m, s, symbols[i], &left, &prob);
        if (rc) return rc;
        orc_ans_encode(c, left, prob);
    }
    return ORC_OK;
}

/* Inverse of the standard normal CDF for the STARTING GUESS of the guided search below.  The reference calls
 * probability-0.20.3 `Gaussian::inverse` (algorithm AS241, source not available here); the guess never influences
 * the result (see orc_qgauss_quantile), only the number of search steps, so a rational approximation of comparable
 * cost and 1e-9 relative accuracy (P. J. Acklam's) stands in for it in the "reference-shaped" CPU baseline. */
static double inv_std_normal_cdf(double p) {
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00};
    if (p < 0.02425) {
        double q = sqrt(-2.0 * log(p));
        return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    if (p > 1.0 - 0.02425) {
        double q = sqrt(-2.0 * log(1.0 - p));
        return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    }
    double q = p - 0.5, r = q * q;
    return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
           (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
}

/* quantize.rs:580-779 in the reference's own shape: guess from the inverse CDF (:596-599), clamp to the support
 * (:601-615), then search downwards (:620-688) or upwards (:689-767) with a doubling step followed by bisection,
 * evaluating one Gaussian CDF per probe.  Same result as orc_qgauss_quantile (tests/test_oracle_golden.py); this
 * variant exists so that the single-thread "lazy model" CPU baseline does the reference's amount of work. */
int orc_qgauss_quantile_guided(int32_t min_sym, int32_t max_sym, double mean, double std, uint32_t quantile,
                               int32_t *symbol_out, uint32_t *left_out, uint32_t *prob_out) {
    double fw;
    int rc = qgauss_free_weight(min_sym, max_sym, &fw);
    if (rc) return rc;
    if (!(std > 0.0)) return ORC_ERR_BAD_MODEL;
    if (quantile >= ORC_TOTAL) return ORC_ERR_INVALID_DATA;
    const double guess = mean + std * inv_std_normal_cdf(((double)quantile + 0.5) * (1.0 / (double)ORC_TOTAL));
    int64_t symbol = guess >= 2147483647.0 ? 2147483647 : (guess <= -2147483648.0 ? -2147483648ll : (int64_t)guess); /* `as i32` */
    uint32_t left, right;
    if (symbol <= min_sym) {
        symbol = min_sym;
        left = 0;
    } else {
        if (symbol > max_sym) symbol = max_sym;
        left = qgauss_left(fw, min_sym, mean, std, (int32_t)symbol);
    }
    int64_t step = 1;
    if (left > quantile) { /* guess too high: :620-688 */
        symbol -= step;
        int found_lower = 0;
        for (;;) {
            const uint32_t old_left = left;
            left = qgauss_left(fw, min_sym, mean, std, (int32_t)symbol);
            if (symbol == min_sym && step <= 1) {
                right = old_left;
                break;
            }
            if (left <= quantile) {
                found_lower = 1;
                if (step <= 1) {
                    right = qgauss_right(fw, min_sym, max_sym, mean, std, (int32_t)symbol);
                    break;
                }
                step >>= 1;
                symbol += step;
            } else if (found_lower) {
                if (step > 1) step >>= 1;
                symbol -= step;
            } else {
                step <<= 1;
                while (symbol - step < min_sym) step >>= 1;
                symbol -= step;
            }
        }
    } else { /* guess right or too low: :689-767 */
        int found_upper = 0;
        for (;;) {
            right = qgauss_right(fw, min_sym, max_sym, mean, std, (int32_t)symbol);
            if (symbol == max_sym && step <= 1) {
                left = qgauss_left(fw, min_sym, mean, std, (int32_t)symbol);
                break;
            }
            if (right > quantile) {
                found_upper = 1;
                if (step <= 1) {
                    left = qgauss_left(fw, min_sym, mean, std, (int32_t)symbol); /* re-evaluated, :728-735 */
                    if (left <= quantile || symbol == min_sym) break;
                } else {
                    step >>= 1;
                }
                symbol -= step;
            } else if (found_upper) {
                if (step > 1) step >>= 1;
                symbol += step;
            } else {
                step <<= 1;
                while (symbol + step > max_sym) step >>= 1;
                symbol += step;
            }
        }
    }
    *symbol_out = (int32_t)symbol;
    *left_out = left;
    *prob_out = right - left;
    return ORC_OK;
}

/* decode_iid_symbols / decode_symbols with lazily evaluated Gaussians (stack.rs:1070-1100 + quantize.rs:580-779) */
int orc_ans_decode_qgauss_lazy(orc_ans *c, int32_t *symbols, size_t n, int32_t min_sym, int32_t max_sym,
                               const double *means, const double *stds, int per_symbol_params) {
    for (size_t i = 0; i < n; i++) {
        uint32_t left, prob;
        double m = per_symbol_params ? means[i] : means[0];
        double s = per_symbol_params ? stds[i] : stds[0];
        int rc = orc_qgauss_quantile_guided(min_sym, max_sym, m, s, orc_ans_peek_quantile(c), &symbols[i], &left, &prob);
        if (rc) return rc;
        orc_ans_decode_advance(c, left, prob);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* Range coder                                                          */
/* ------------------------------------------------------------------ */

void orc_renc_init(orc_renc *e) {
    memset(e, 0, sizeof *e);
    e->range = UINT64_MAX; /* queue.rs:98-106 */
}
void orc_renc_free(orc_renc *e) {
    free(e->bulk);
    memset(e, 0, sizeof *e);
}
void orc_renc_clear(orc_renc *e) {
    e->len = 0;
    e->lower = 0;
    e->range = UINT64_MAX;
    e->num_inverted = 0;
    e->first_inverted = 0;
}

/* queue.rs:612-705 */
int orc_renc_encode(orc_renc *e, uint32_t left, uint32_t prob) {
    uint64_t scale = e->range >> ORC_PRECISION;
    uint64_t new_range = scale * (uint64_t)prob;
    if (new_range == 0) return ORC_ERR_IMPOSSIBLE_SYMBOL; /* :640-642 */
    e->range = new_range;
    uint64_t new_lower = e->lower + scale * (uint64_t)left; /* wrapping */

    if (e->num_inverted) { /* :647-666 */
        if (new_lower + e->range > new_lower) {
            uint32_t first_word, consecutive;
            if (new_lower < e->lower) {
                first_word = e->first_inverted + 1u;
                consecutive = 0;
            } else {
                first_word = e->first_inverted;
                consecutive = 0xffffffffu;
            }
            vec_push(&e->bulk, &e->len, &e->cap, first_word);
            for (size_t i = 1; i < e->num_inverted; i++) vec_push(&e->bulk, &e->len, &e->cap, consecutive);
            e->num_inverted = 0;
        }
    }
    e->lower = new_lower;

    if (e->range < (1ull << 32)) { /* :670-702 */
        e->range <<= 32;
        uint32_t lower_word = (uint32_t)(e->lower >> 32);
        e->lower <<= 32;
        if (e->num_inverted) {
            e->num_inverted += 1;
        } else if (e->lower + e->range > e->lower) {
            vec_push(&e->bulk, &e->len, &e->cap, lower_word);
        } else {
            e->num_inverted = 1;
            e->first_inverted = lower_word;
        }
    }
    return ORC_OK;
}

/* queue.rs:357-376 */
size_t orc_renc_num_seal_words(const orc_renc *e) {
    if (e->range == UINT64_MAX) return 0;
    uint64_t point = e->lower + ((1ull << 32) - 1);
    uint32_t point_word = (uint32_t)(point >> 32);
    uint32_t upper_word = (uint32_t)((e->lower + e->range) >> 32);
    size_t count = (upper_word == point_word) ? 2 : 1;
    count += e->num_inverted;
    return count;
}

size_t orc_renc_num_words(const orc_renc *e) { return e->len + orc_renc_num_seal_words(e); } /* queue.rs:383-388 */

/* queue.rs:349-355 seal + :458-523 iter_seal / seal_words */
size_t orc_renc_get_compressed(const orc_renc *e, uint32_t *out) {
    memcpy(out, e->bulk, e->len * sizeof(uint32_t));
    size_t n = e->len;
    if (e->range == UINT64_MAX) return n;
    uint64_t point = e->lower + ((1ull << 32) - 1);
    if (e->num_inverted) {
        uint32_t first, consecutive;
        if (point >= e->lower) {
            first = e->first_inverted;
            consecutive = 0xffffffffu;
        } else {
            first = e->first_inverted + 1u;
            consecutive = 0;
        }
        out[n++] = first;
        for (size_t i = 1; i < e->num_inverted; i++) out[n++] = consecutive;
    }
    uint32_t point_word = (uint32_t)(point >> 32);
    out[n++] = point_word;
    uint32_t upper_word = (uint32_t)((e->lower + e->range) >> 32);
    if (upper_word == point_word) out[n++] = 0;
    return n;
}

int orc_renc_encode_iid(orc_renc *e, const int32_t *symbols, size_t n, const uint32_t *cdf, int32_t min_sym,
                        size_t alphabet) {
    for (size_t i = 0; i < n; i++) {
        uint32_t left, prob;
        int rc = orc_cdf_left_prob(cdf, alphabet, (int64_t)symbols[i] - (int64_t)min_sym, &left, &prob);
        if (rc) return rc;
        rc = orc_renc_encode(e, left, prob);
        if (rc) return rc;
    }
    return ORC_OK;
}

/* queue.rs:847-868 read_point */
static uint64_t rdec_read_point(orc_rdec *d) {
    int num_read = 0;
    uint64_t point = 0;
    while (d->pos < d->len) {
        point = (point << 32) | d->bulk[d->pos++];
        if (++num_read == 2) break;
    }
    if (num_read < 2 && num_read != 0) point <<= (64 - num_read * 32);
    return point;
}

void orc_rdec_init(orc_rdec *d, const uint32_t *words, size_t n) {
    d->bulk = words;
    d->len = n;
    d->pos = 0;
    d->lower = 0;
    d->range = UINT64_MAX;
    d->point = rdec_read_point(d);
}

int orc_rdec_peek_quantile(const orc_rdec *d, uint32_t *q) {
    uint64_t scale = d->range >> ORC_PRECISION;
    uint64_t quantile = (d->point - d->lower) / scale;
    if (quantile >= (1ull << ORC_PRECISION)) return ORC_ERR_INVALID_DATA;
    *q = (uint32_t)quantile;
    return ORC_OK;
}

void orc_rdec_advance(orc_rdec *d, uint32_t left, uint32_t prob) {
    uint64_t scale = d->range >> ORC_PRECISION;
    d->lower += scale * (uint64_t)left;
    d->range = scale * (uint64_t)prob;
    if (d->range < (1ull << 32)) {
        d->lower <<= 32;
        d->range <<= 32;
        d->point <<= 32;
        if (d->pos < d->len) d->point |= d->bulk[d->pos++];
    }
}

int orc_rdec_maybe_exhausted(const orc_rdec *d) {
    uint64_t max_difference = ((1ull << 32) << 1) - 1;
    return d->pos >= d->len && (d->range == UINT64_MAX || (d->point - d->lower) < max_difference);
}

/* queue.rs:911-928 Seek (+ backends.rs:1596-1608 Cursor::seek, queue.rs:66-77 state check) */
int orc_rdec_seek(orc_rdec *d, size_t pos, uint64_t lower, uint64_t range) {
    if ((range >> 32) == 0) return ORC_ERR_SEEK;
    if (pos > d->len) return ORC_ERR_SEEK;
    d->pos = pos;
    d->point = rdec_read_point(d);
    d->lower = lower;
    d->range = range;
    return ORC_OK;
}

int orc_rdec_decode_iid(orc_rdec *d, int32_t *symbols, size_t n, const uint32_t *cdf, int32_t min_sym,
                        size_t alphabet) {
    for (size_t i = 0; i < n; i++) {
        uint32_t q, left, prob;
        size_t idx;
        int rc = orc_rdec_peek_quantile(d, &q);
        if (rc) return rc;
        orc_cdf_quantile(cdf, alphabet, q, &idx, &left, &prob);
        orc_rdec_advance(d, left, prob);
        symbols[i] = (int32_t)((int64_t)min_sym + (int64_t)idx);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* multi-stream helpers                                                 */
/* ------------------------------------------------------------------ */

typedef struct {
    int kind; /* 0 ans enc, 1 ans dec, 2 range enc, 3 range dec */
    const int32_t *symbols;
    int32_t *symbols_out;
    uint64_t n_total, K;
    int interleaved;
    const uint64_t *sym_off;
    const uint32_t *cdf;
    int32_t min_sym;
    size_t alphabet;
    uint32_t **per_stream_words; /* enc: malloc'd per stream */
    uint64_t *per_stream_len;
    const uint32_t *words;
    const uint64_t *offsets;
    uint64_t k_begin, k_end;
    int rc;
} multi_job;

static uint64_t stream_len(const multi_job *j, uint64_t k) {
    if (j->interleaved) return (j->n_total > k) ? (j->n_total - k + j->K - 1) / j->K : 0;
    return j->sym_off[k + 1] - j->sym_off[k];
}

static void *multi_worker(void *arg) {
    multi_job *j = (multi_job *)arg;
    uint64_t maxlen = 0;
    for (uint64_t k = j->k_begin; k < j->k_end; k++) {
        uint64_t l = stream_len(j, k);
        if (l > maxlen) maxlen = l;
    }
    int32_t *tmp = (int32_t *)malloc((maxlen ? maxlen : 1) * sizeof(int32_t));
    for (uint64_t k = j->k_begin; k < j->k_end; k++) {
        uint64_t n = stream_len(j, k);
        if (j->kind == 0 || j->kind == 2) {
            const int32_t *src;
            if (j->interleaved) {
                for (uint64_t t = 0; t < n; t++) tmp[t] = j->symbols[t * j->K + k];
                src = tmp;
            } else {
                src = j->symbols + j->sym_off[k];
            }
            if (j->kind == 0) {
                orc_ans c;
                orc_ans_init(&c);
                int rc = orc_ans_encode_iid_reverse(&c, src, n, j->cdf, j->min_sym, j->alphabet);
                if (rc) j->rc = rc;
                uint32_t *w = (uint32_t *)malloc((c.len + 2) * sizeof(uint32_t));
                j->per_stream_len[k] = orc_ans_get_compressed(&c, w);
                j->per_stream_words[k] = w;
                orc_ans_free(&c);
            } else {
                orc_renc e;
                orc_renc_init(&e);
                int rc = orc_renc_encode_iid(&e, src, n, j->cdf, j->min_sym, j->alphabet);
                if (rc) j->rc = rc;
                uint32_t *w = (uint32_t *)malloc((orc_renc_num_words(&e) + 1) * sizeof(uint32_t));
                j->per_stream_len[k] = orc_renc_get_compressed(&e, w);
                j->per_stream_words[k] = w;
                orc_renc_free(&e);
            }
        } else {
            const uint32_t *w = j->words + j->offsets[k];
            size_t nw = (size_t)(j->offsets[k + 1] - j->offsets[k]);
            int32_t *dst = j->interleaved ? tmp : j->symbols_out + j->sym_off[k];
            if (j->kind == 1) {
                orc_ans c;
                orc_ans_init(&c);
                int rc = orc_ans_from_compressed(&c, w, nw);
                if (rc) j->rc = rc;
                orc_ans_decode_iid(&c, dst, n, j->cdf, j->min_sym, j->alphabet);
                orc_ans_free(&c);
            } else {
                orc_rdec d;
                orc_rdec_init(&d, w, nw);
                int rc = orc_rdec_decode_iid(&d, dst, n, j->cdf, j->min_sym, j->alphabet);
                if (rc) j->rc = rc;
            }
            if (j->interleaved)
                for (uint64_t t = 0; t < n; t++) j->symbols_out[t * j->K + k] = tmp[t];
        }
    }
    free(tmp);
    return NULL;
}

static int run_multi(multi_job *proto, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > proto->K) threads = (int)(proto->K ? proto->K : 1);
    multi_job *jobs = (multi_job *)malloc(sizeof(multi_job) * (size_t)threads);
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    uint64_t per = (proto->K + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 0; t < threads; t++) {
        jobs[t] = *proto;
        jobs[t].k_begin = (uint64_t)t * per < proto->K ? (uint64_t)t * per : proto->K;
        jobs[t].k_end = jobs[t].k_begin + per < proto->K ? jobs[t].k_begin + per : proto->K;
        jobs[t].rc = 0;
        if (threads == 1)
            multi_worker(&jobs[t]);
        else
            pthread_create(&tids[t], NULL, multi_worker, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < threads; t++) {
        if (threads > 1) pthread_join(tids[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(jobs);
    free(tids);
    return rc;
}

static int multi_encode(int kind, const int32_t *symbols, uint64_t n_total, uint64_t K, int interleaved,
                        const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym, size_t alphabet,
                        uint32_t **words_out, uint64_t *offsets_out, int threads) {
    multi_job j;
    memset(&j, 0, sizeof j);
    j.kind = kind;
    j.symbols = symbols;
    j.n_total = n_total;
    j.K = K;
    j.interleaved = interleaved;
    j.sym_off = sym_off;
    j.cdf = cdf;
    j.min_sym = min_sym;
    j.alphabet = alphabet;
    j.per_stream_words = (uint32_t **)calloc(K ? K : 1, sizeof(uint32_t *));
    j.per_stream_len = (uint64_t *)calloc(K ? K : 1, sizeof(uint64_t));
    int rc = run_multi(&j, threads);
    uint64_t total = 0;
    for (uint64_t k = 0; k < K; k++) {
        offsets_out[k] = total;
        total += j.per_stream_len[k];
    }
    offsets_out[K] = total;
    uint32_t *out = (uint32_t *)malloc((total ? total : 1) * sizeof(uint32_t));
    for (uint64_t k = 0; k < K; k++) {
        memcpy(out + offsets_out[k], j.per_stream_words[k], j.per_stream_len[k] * sizeof(uint32_t));
        free(j.per_stream_words[k]);
    }
    free(j.per_stream_words);
    free(j.per_stream_len);
    *words_out = out;
    return rc;
}

static int multi_decode(int kind, const uint32_t *words, const uint64_t *offsets, uint64_t n_total, uint64_t K,
                        int interleaved, const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym,
                        size_t alphabet, int32_t *symbols_out, int threads) {
    multi_job j;
    memset(&j, 0, sizeof j);
    j.kind = kind;
    j.symbols_out = symbols_out;
    j.n_total = n_total;
    j.K = K;
    j.interleaved = interleaved;
    j.sym_off = sym_off;
    j.cdf = cdf;
    j.min_sym = min_sym;
    j.alphabet = alphabet;
    j.words = words;
    j.offsets = offsets;
    return run_multi(&j, threads);
}

int orc_multi_ans_encode(const int32_t *symbols, uint64_t n_total, uint64_t K, int interleaved,
                         const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym, size_t alphabet,
                         uint32_t **words_out, uint64_t *offsets_out, int threads) {
    return multi_encode(0, symbols, n_total, K, interleaved, sym_off, cdf, min_sym, alphabet, words_out,
                        offsets_out, threads);
}
int orc_multi_ans_decode(const uint32_t *words, const uint64_t *offsets, uint64_t n_total, uint64_t K,
                         int interleaved, const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym,
                         size_t alphabet, int32_t *symbols_out, int threads) {
    return multi_decode(1, words, offsets, n_total, K, interleaved, sym_off, cdf, min_sym, alphabet, symbols_out,
                        threads);
}
int orc_multi_range_encode(const int32_t *symbols, uint64_t n_total, uint64_t K, int interleaved,
                           const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym, size_t alphabet,
                           uint32_t **words_out, uint64_t *offsets_out, int threads) {
    return multi_encode(2, symbols, n_total, K, interleaved, sym_off, cdf, min_sym, alphabet, words_out,
                        offsets_out, threads);
}
int orc_multi_range_decode(const uint32_t *words, const uint64_t *offsets, uint64_t n_total, uint64_t K,
                           int interleaved, const uint64_t *sym_off, const uint32_t *cdf, int32_t min_sym,
                           size_t alphabet, int32_t *symbols_out, int threads) {
    return multi_decode(3, words, offsets, n_total, K, interleaved, sym_off, cdf, min_sym, alphabet, symbols_out,
                        threads);
}

void orc_free(void *p) { free(p); }
