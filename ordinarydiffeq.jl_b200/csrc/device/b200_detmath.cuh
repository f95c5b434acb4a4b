// b200_detmath.cuh — correctly-rounded log10 / exp10 in double-double arithmetic.
//
// Why: the automatic initial step (lib/OrdinaryDiffEqCore/src/initdt.jl:451-455)
// evaluates 10^(-(2 + log10(max(d1,d2)))/order).  libm, CUDA libdevice and Julia
// Base each round log10/pow slightly differently (all <1 ulp, none identical), and
// one ulp of dt0 reseeds the whole adaptive step sequence.  These
// routines return the correctly rounded double (error of the dd evaluation is
// ~1e-30 relative, so a misrounding needs an input within ~1e-14 ulp of a rounding
// boundary), which is what any faithful libm returns in all but a vanishing
// fraction of cases — and they are bit-identical on host and device.  The CPU
// oracle does NOT use this file: it uses libquadmath (log10q/powq) rounded to
// double, an independent route to the same correctly rounded value.
//
// Requires fmad=false / -ffp-contract=off (error-free transforms).
#pragma once
#include "b200_base.cuh"

struct b200_dd { double hi, lo; };

B200_HD b200_dd b200_dd_make(double hi, double lo) { b200_dd r; r.hi = hi; r.lo = lo; return r; }
B200_HD b200_dd b200_quick_two_sum(double a, double b) {
    double s = a + b; double e = b - (s - a); return b200_dd_make(s, e);
}
B200_HD b200_dd b200_two_sum(double a, double b) {
    double s = a + b; double bb = s - a; double e = (a - (s - bb)) + (b - bb);
    return b200_dd_make(s, e);
}
B200_HD b200_dd b200_two_prod(double a, double b) {
    double p = a * b; double e = fma(a, b, -p); return b200_dd_make(p, e);
}
B200_HD b200_dd b200_dd_add(b200_dd x, b200_dd y) {
    b200_dd s = b200_two_sum(x.hi, y.hi);
    b200_dd t = b200_two_sum(x.lo, y.lo);
    s.lo += t.hi;
    s = b200_quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return b200_quick_two_sum(s.hi, s.lo);
}
B200_HD b200_dd b200_dd_add_d(b200_dd x, double y) {
    b200_dd s = b200_two_sum(x.hi, y);
    s.lo += x.lo;
    return b200_quick_two_sum(s.hi, s.lo);
}
B200_HD b200_dd b200_dd_neg(b200_dd x) { return b200_dd_make(-x.hi, -x.lo); }
B200_HD b200_dd b200_dd_mul(b200_dd x, b200_dd y) {
    b200_dd p = b200_two_prod(x.hi, y.hi);
    p.lo += x.hi * y.lo + x.lo * y.hi;
    return b200_quick_two_sum(p.hi, p.lo);
}
B200_HD b200_dd b200_dd_mul_d(b200_dd x, double y) {
    b200_dd p = b200_two_prod(x.hi, y);
    p.lo += x.lo * y;
    return b200_quick_two_sum(p.hi, p.lo);
}
B200_HD b200_dd b200_dd_div(b200_dd x, b200_dd y) {
    double q1 = x.hi / y.hi;
    b200_dd r = b200_dd_add(x, b200_dd_neg(b200_dd_mul_d(y, q1)));
    double q2 = r.hi / y.hi;
    r = b200_dd_add(r, b200_dd_neg(b200_dd_mul_d(y, q2)));
    double q3 = r.hi / y.hi;
    b200_dd q = b200_quick_two_sum(q1, q2);
    return b200_dd_add_d(q, q3);
}

// log10(x), x finite > 0 (normal or subnormal).  Correctly rounded (see header).
B200_HD double b200_log10_cr(double x) {
    // x = 2^k * f with f in [sqrt(1/2), sqrt(2))
    uint64_t b = b200_d2u(x);
    int k = (int)((b >> 52) & 0x7FFull);
    if (k == 0) {                          // subnormal: scale by 2^64
        x = x * 18446744073709551616.0;
        b = b200_d2u(x);
        k = (int)((b >> 52) & 0x7FFull) - 64;
    }
    k -= 1023;
    double f = b200_u2d((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);   // [1,2)
    if (f > 1.4142135623730951) { f *= 0.5; k += 1; }
    // ln f = 2 atanh(z), z = (f-1)/(f+1); f-1 is exact (Sterbenz)
    b200_dd num = b200_dd_make(f - 1.0, 0.0);
    b200_dd den = b200_two_sum(f, 1.0);
    b200_dd z = b200_dd_div(num, den);
    b200_dd z2 = b200_dd_mul(z, z);
    // sum_{j=0}^{22} z2^j/(2j+1), Horner from the top; |z| <= 0.1716 -> z^47/47 < 3e-38.  The coefficients 1/(2j+1) are
    // double-double constants (rounded from 300-bit values): no division inside the series
    const double LH[23] = {0x1.0000000000000p+0, 0x1.5555555555555p-2, 0x1.999999999999ap-3, 0x1.2492492492492p-3, 0x1.c71c71c71c71cp-4,
                           0x1.745d1745d1746p-4, 0x1.3b13b13b13b14p-4, 0x1.1111111111111p-4, 0x1.e1e1e1e1e1e1ep-5, 0x1.af286bca1af28p-5,
                           0x1.8618618618618p-5, 0x1.642c8590b2164p-5, 0x1.47ae147ae147bp-5, 0x1.2f684bda12f68p-5, 0x1.1a7b9611a7b96p-5,
                           0x1.0842108421084p-5, 0x1.f07c1f07c1f08p-6, 0x1.d41d41d41d41dp-6, 0x1.bacf914c1bad0p-6, 0x1.a41a41a41a41ap-6,
                           0x1.8f9c18f9c18fap-6, 0x1.7d05f417d05f4p-6, 0x1.6c16c16c16c17p-6};
    const double LL[23] = {0x0.0p+0, 0x1.5555555555555p-56, -0x1.999999999999ap-57, 0x1.2492492492492p-57, 0x1.c71c71c71c71cp-58,
                           -0x1.745d1745d1746p-59, -0x1.3b13b13b13b14p-58, 0x1.1111111111111p-60, 0x1.e1e1e1e1e1e1ep-61, 0x1.af286bca1af28p-59,
                           0x1.8618618618618p-59, 0x1.642c8590b2164p-60, -0x1.eb851eb851eb8p-61, 0x1.2f684bda12f68p-59, 0x1.1a7b9611a7b96p-61,
                           0x1.0842108421084p-60, -0x1.f07c1f07c1f08p-61, 0x1.0750750750750p-60, -0x1.bacf914c1bad0p-60, 0x1.0690690690690p-60,
                           -0x1.f3831f3831f38p-61, 0x1.7d05f417d05f4p-62, -0x1.f49f49f49f49fp-61};
    b200_dd s = b200_dd_make(LH[22], LL[22]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 21; j >= 0; --j) s = b200_dd_add(b200_dd_mul(s, z2), b200_dd_make(LH[j], LL[j]));
    b200_dd lnf = b200_dd_mul(z, s);
    lnf = b200_dd_make(lnf.hi * 2.0, lnf.lo * 2.0);
    const b200_dd INV_LN10 = b200_dd_make(0x1.bcb7b1526e50ep-2, 0x1.95355baaafad3p-57);
    const b200_dd LOG10_2 = b200_dd_make(0x1.34413509f79ffp-2, -0x1.9dc1da994fd21p-59);
    b200_dd r = b200_dd_mul(lnf, INV_LN10);
    r = b200_dd_add(r, b200_dd_mul_d(LOG10_2, (double)k));
    return r.hi;
}

// 10^y for finite y with the result in the normal range.  Correctly rounded.
B200_HD double b200_exp10_cr(double y) {
    const b200_dd LOG2_10 = b200_dd_make(0x1.a934f0979a371p+1, 0x1.7f2495fb7fa6dp-53);
    const b200_dd LN2 = b200_dd_make(0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56);
    b200_dd t = b200_dd_mul_d(LOG2_10, y);          // y*log2(10)
    double n = rint(t.hi);
    b200_dd r = b200_dd_add_d(t, -n);               // |r| <= 0.5 (+tiny)
    b200_dd w = b200_dd_mul(r, LN2);                // 2^r = exp(w), |w| <= 0.347
    w = b200_dd_make(w.hi * 0.125, w.lo * 0.125);   // exact scaling
    // exp(w) Taylor to degree 17 (|w|<=0.0434: w^18/18! < 1e-40), Horner over the double-double constants 1/j!
    // (rounded from 300-bit values): no division — the r1 form `s = 1 + (w/j) s` spent 51 IEEE divisions here
    const double EH[18] = {0x1.0000000000000p+0, 0x1.0000000000000p+0, 0x1.0000000000000p-1, 0x1.5555555555555p-3, 0x1.5555555555555p-5,
                           0x1.1111111111111p-7, 0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-16, 0x1.71de3a556c734p-19,
                           0x1.27e4fb7789f5cp-22, 0x1.ae64567f544e4p-26, 0x1.1eed8eff8d898p-29, 0x1.6124613a86d09p-33, 0x1.93974a8c07c9dp-37,
                           0x1.ae7f3e733b81fp-41, 0x1.ae7f3e733b81fp-45, 0x1.952c77030ad4ap-49};
    const double EL[18] = {0x0.0p+0, 0x0.0p+0, 0x0.0p+0, 0x1.5555555555555p-57, 0x1.5555555555555p-59, 0x1.1111111111111p-63,
                           -0x1.f49f49f49f49fp-65, 0x1.a01a01a01a01ap-73, 0x1.a01a01a01a01ap-76, -0x1.c154f8ddc6c00p-73, 0x1.cbbc05b4fa99ap-76,
                           -0x1.c062e06d1f209p-80, -0x1.2aec959e14c06p-83, 0x1.f28e0cc748ebep-87, 0x1.05d6f8a2efd1fp-92, 0x1.1d8656b0ee8cbp-97,
                           0x1.1d8656b0ee8cbp-101, 0x1.ac981465ddc6cp-103};
    b200_dd s = b200_dd_make(EH[17], EL[17]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 16; j >= 0; --j) s = b200_dd_add(b200_dd_mul(s, w), b200_dd_make(EH[j], EL[j]));
    s = b200_dd_mul(s, s); s = b200_dd_mul(s, s); s = b200_dd_mul(s, s);
    // scale by 2^n, n integer in the normal range
    int ni = (int)n;
    if (ni < -1021) return 0.0;
    if (ni > 1023) return b200_u2d(0x7FF0000000000000ull);
    double sc = b200_u2d((uint64_t)(ni + 1023) << 52);
    return s.hi * sc;
}
