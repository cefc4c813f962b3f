// zz_math.h -- scalar primitives shared by the sm_100a kernels, the host library and
// (for zz_log / zz_u01 only) the CPU oracle, so that "same seed => same bits" holds by
// construction on both sides of the PCIe bus.
//
// Rules that make host and device agree bit for bit:
//   * only + - * / sqrt and integer ops (all IEEE-754 correctly rounded on x86-64 SSE2 and sm_100a);
//   * no fused multiply-add: device code is compiled with -fmad=false, host code with
//     -ffp-contract=off;
//   * the logarithm is OUR routine (zz_log), never libm's / libdevice's.
//
// Reference semantics restated here:
//   pos            src/common.jl:8
//   poisson_time   src/poissontime.jl:8-30 (two-parameter form), :39-65 (three-parameter form)
//   Boomerang flow src/sfact.jl:29-38, src/dynamics.jl:29-36 (zz_boom_at, with our own zz_sincos)
#ifndef ZZ_MATH_H
#define ZZ_MATH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define ZZ_HD __host__ __device__ __forceinline__
#define ZZ_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ZZ_HD static inline
#define ZZ_HD_NOINLINE static inline
#endif

#if defined(__CUDA_ARCH__)
#define ZZ_INF __longlong_as_double(0x7ff0000000000000LL)
ZZ_HD double zz_sqrt(double x) { return __dsqrt_rn(x); }
ZZ_HD uint64_t zz_d2u(double x) { return (uint64_t)__double_as_longlong(x); }
ZZ_HD double zz_u2d(uint64_t u) { return __longlong_as_double((long long)u); }
#else
#include <math.h>
#include <string.h>
#define ZZ_INF ((double)INFINITY)
ZZ_HD double zz_sqrt(double x) { return sqrt(x); }
ZZ_HD uint64_t zz_d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
ZZ_HD double zz_u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#endif

ZZ_HD double zz_pos(double x) { return x > 0.0 ? x : 0.0; }  // max(0,x); NaN -> 0 never reached on this path

// Natural logarithm for finite normal x > 0 (the uniforms below are in [2^-54, 1)).
// Classic argument reduction x = 2^k (1+f), s = f/(2+f), log(1+f) = 2s + s*R(s^2) with a
// degree-14 minimax polynomial (the well-known Sun/fdlibm scheme, error < 1 ulp); written
// out with explicit operation order so every compiler produces the same roundings.
ZZ_HD double zz_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01;
    const double ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01;
    const double Lg2 = 3.999999999940941908e-01;
    const double Lg3 = 2.857142874366239149e-01;
    const double Lg4 = 2.222219843214978396e-01;
    const double Lg5 = 1.818357216161805012e-01;
    const double Lg6 = 1.531383769920937332e-01;
    const double Lg7 = 1.479819860511658591e-01;

    uint64_t ux = zz_d2u(x);
    int32_t hx = (int32_t)(ux >> 32);
    uint32_t lx = (uint32_t)ux;
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    // normalise x into [sqrt(2)/2, sqrt(2))
    ux = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx;
    x = zz_u2d(ux);
    k += (i >> 20);
    double f = x - 1.0;
    double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {  // |f| < 2^-20
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return dk * ln2_hi + dk * ln2_lo;
        }
        double R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    double s = f / (2.0 + f);
    double z = s * s;
    i = hx - 0x6147a;
    double w = z * z;
    int32_t j = 0x6b851 - hx;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    double R = t2 + t1;
    if (i > 0) {
        double hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

// Exponential for finite x (the logistic target: sigmoid(x) = inv(1 + exp(-x)), scripts/logistic.jl:34).  Like zz_log
// this is OUR routine so that host and device agree bit for bit: the Sun/fdlibm scheme -- x = k ln2 + r with ln2 split in
// two pieces, exp(r) = 1 + 2r / (2 - c(r) ... ) with a degree-5 minimax polynomial in r^2 (error < 1 ulp), scaling by 2^k
// through the exponent field; only + - * /, one double -> int conversion and integer operations.
ZZ_HD double zz_exp(double x)
{
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10;
    const double invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05;
    const double P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    const double twom1000 = 9.33263618503218878990e-302;   // 2^-1000
    if (x > 7.09782712893383973096e+02) return ZZ_INF;      // overflow
    if (x < -7.45133219101941108420e+02) return 0.0;        // underflow
    const double ax = x < 0.0 ? -x : x;
    double hi = 0.0, lo = 0.0;
    int32_t k = 0;
    if (ax > 0.34657359027997264) {                         // |x| > ln2 / 2
        k = (int32_t)(invln2 * x + (x < 0.0 ? -0.5 : 0.5));
        const double t = (double)k;
        hi = x - t * ln2HI;                                 // t * ln2HI is exact (ln2HI has 21 trailing zero bits)
        lo = t * ln2LO;
        x = hi - lo;
    } else if (ax < 3.7252902984619141e-09) {               // |x| < 2^-28
        return 1.0 + x;
    }
    const double t = x * x;
    const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return zz_u2d(zz_d2u(y) + ((uint64_t)(int64_t)k << 52));
    return zz_u2d(zz_d2u(y) + ((uint64_t)(int64_t)(k + 1000) << 52)) * twom1000;
}

// scripts/logistic.jl:34,56-57 -- sigmoid(x) = inv(one(x) + exp(-x)), sigmoidn(x) = sigmoid(-x), nsigmoid(x) = -sigmoid(x)
ZZ_HD double zz_sigmoid(double x) { return 1.0 / (1.0 + zz_exp(-x)); }
ZZ_HD double zz_sigmoidn(double x) { return zz_sigmoid(-x); }
ZZ_HD double zz_nsigmoid(double x) { return -zz_sigmoid(x); }

// ---------------------------------------------------------------------------------------------
// Counter-based uniforms.  u(i,k) is the k-th draw of coordinate i's private stream:
// two rounds of the splitmix64 finaliser over (seed, coordinate, counter).  The value is
// strictly inside (0,1) -- in [2^-53, 1 - 2^-53] -- so log(u) is always finite.  (The reference draws from one
// sequential Xoroshiro128Plus stream in global event order, src/sfact.jl:121,134,139,186,
// which no parallel schedule can reproduce; see DESIGN.md "RNG contract".)
ZZ_HD uint64_t zz_mix64(uint64_t z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}

ZZ_HD double zz_u01(uint64_t seed0, uint64_t seed1, uint64_t coord, uint64_t ctr)
{
    uint64_t z = zz_mix64(seed0 + (coord + 1ULL) * 0x9E3779B97F4A7C15ULL);
    z = zz_mix64(z ^ (seed1 + ctr * 0xD1342543DE82EF95ULL));
    // 52 random bits as the mantissa of a double in [1,2), shifted to (0,1): u = m 2^-52 + 2^-53 (exact arithmetic;
    // two integer and two floating-point instructions instead of a 64-bit integer-to-double conversion)
    return (zz_u2d((z >> 12) | 0x3FF0000000000000ULL) - 1.0) + 1.1102230246251565404e-16;
}

// First arrival time of an inhomogeneous Poisson process with rate (a + b t)^+ given u ~ U(0,1).
// Expression shapes follow src/poissontime.jl:8-30 term by term ((a/b)^2 is (a/b)*(a/b)).
// zz_poisson_time_L takes L = log(u) so that callers can compute the logarithm early (it does not depend on a, b)
// and overlap its latency with the evaluation of the rates.
ZZ_HD double zz_poisson_time_L(double a, double b, double L)
{
    // One square root and the two quotients shared by all branches (SIMT lanes that take different branches of the
    // reference formula then diverge only in cheap selects).  Bit-identical to evaluating the reference expressions
    // branch by branch: (-L)*2/b == -(L*2/b) exactly, negation commutes with rounding.  L = log(u) < 0.
    if (b == 0.0) return a > 0.0 ? -L / a : ZZ_INF;
    const double q = a / b;
    const double t2 = L * 2.0 / b;
    int ok = b > 0.0;
    if (!ok && a > 0.0) ok = (-L <= -(a * a) / b + (a * a) / (2.0 * b));
    if (!ok) return ZZ_INF;
    const double arg = (b > 0.0 && a < 0.0) ? -t2 : q * q - t2;
    const double root = zz_sqrt(arg);
    return (b > 0.0 ? root : -root) - q;
}

ZZ_HD double zz_poisson_time(double a, double b, double u) { return zz_poisson_time_L(a, b, zz_log(u)); }

// Rate c + (a + b t)^+, c > 0 (src/poissontime.jl:39-65); used by the sticky variants.
ZZ_HD double zz_poisson_time3(double a, double b, double c, double u)
{
    double lu = zz_log(u);
    if (b > 0.0) {
        if (a < 0.0) {
            if (-c * a / b + lu < 0.0)
                return zz_sqrt(-2.0 * b * lu + c * c + 2.0 * a * c) / b - (a + c) / b;
            return -lu / c;
        }
        return zz_sqrt(-lu * 2.0 * b + (a + c) * (a + c)) / b - (a + c) / b;
    } else if (b == 0.0) {
        if (a > 0.0) return -lu / (a + c);
        return -lu / c;
    } else {
        if (a <= 0.0) return -lu / c;
        if (-c * a / b - (a * a) / (2.0 * b) + lu > 0.0)
            return zz_sqrt((a + c) * (a + c) - 2.0 * lu * b) / b - (a + c) / b;
        return (-lu + (a * a) / (2.0 * b)) / c;
    }
}

// ---------------------------------------------------------------------------------------------
// sin and cos together (the Boomerang flow rotates (x - mu, theta), src/sfact.jl:29-38 `sincos`).  Like zz_log this is
// OUR routine so that host and device agree bit for bit: Cody-Waite reduction by pi/2 in two 33-bit pieces (exact
// products for |n| < 2^20, i.e. |x| < 1.6e6), then the classic Sun/fdlibm kernels on [-pi/4, pi/4] (error < 1 ulp each);
// only + - * and one double -> integer conversion.  Beyond |x| ~ 1.6e6 the result stays deterministic but loses accuracy.
ZZ_HD void zz_sincos(double x, double* sn, double* cs)
{
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    const double pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04;
    const double S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05;
    const double C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    (void)pio2_1t;
    const double v = x * invpio2;
    const long long n = (long long)(v + (v < 0.0 ? -0.5 : 0.5));   // nearest integer, ties away from zero
    const double fn = (double)n;
    const double t = x - fn * pio2_1;
    double w = fn * pio2_2;
    const double r = t - w;
    w = fn * pio2_2t - ((t - r) - w);
    const double y0 = r - w;
    const double y1 = (r - y0) - w;
    const double z = y0 * y0;
    // kernel sin
    const double vv = z * y0;
    const double rs = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    const double ks = y0 - ((z * (0.5 * y1 - vv * rs) - y1) - vv * S1);
    // kernel cos
    const double ww = z * z;
    const double rc = z * (C1 + z * (C2 + z * C3)) + (ww * ww) * (C4 + z * (C5 + z * C6));
    const double hz = 0.5 * z;
    const double w1 = 1.0 - hz;
    const double kc = w1 + (((1.0 - w1) - hz) + (z * rc - y0 * y1));
    const int q = (int)(n & 3LL);
    *sn = (q == 0) ? ks : (q == 1) ? kc : (q == 2) ? -ks : -kc;
    *cs = (q == 0) ? kc : (q == 1) ? -ks : (q == 2) ? -kc : ks;
}

// Standard normal from two uniforms (Box-Muller): the velocity refreshment `randn(rng)` of src/sfact.jl:102.  (Julia
// draws it with a ziggurat from its own stream; equal in law, and the stream is ours anyway -- DESIGN.md "RNG contract".)
ZZ_HD double zz_randn(double u1, double u2)
{
    double sn, cs;
    zz_sincos(6.283185307179586232 * u2, &sn, &cs);
    return zz_sqrt(-2.0 * zz_log(u1)) * cs;
}

// Boomerang flow of one coordinate from its anchor (tf, xf, thf) to time s (src/sfact.jl:29-38, dynamics.jl:29-36):
//   x = (xf - mu) cos(tau) + thf sin(tau) + mu,   theta = -(xf - mu) sin(tau) + thf cos(tau),   tau = s - tf.
// tau == 0 returns the anchor itself (so that re-reading a coordinate at the time it was anchored is exact).
ZZ_HD void zz_boom_at(double tf, double xf, double thf, double mu, double s, double* x, double* th)
{
    const double tau = s - tf;
    double sn, cs;
    zz_sincos(tau, &sn, &cs);
    const double xm = xf - mu;
    const double xr = xm * cs + thf * sn + mu;
    const double tr = -xm * sn + thf * cs;
    *x = (tau == 0.0) ? xf : xr;
    *th = (tau == 0.0) ? thf : tr;
}

#endif  // ZZ_MATH_H
