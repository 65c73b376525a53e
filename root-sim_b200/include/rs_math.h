/*
 * rs_math.h -- portable FP64 elementary functions with bit-identical results on host and device.
 *
 * Why: the reference draws exponential timestamps as -mean*log(1-u) with glibc's log
 * (reference src/lib/numerical.c:120-134).  glibc's libm is not available on the GPU and CUDA's
 * log() cannot be run on the CPU, so a device engine that called CUDA's log() could be compared with
 * a CPU run only up to a ULP tolerance.  These routines use nothing but IEEE-754 binary64
 * + - * / and integer bit manipulation, in a fixed evaluation order, so that the SAME bits come out of
 * gcc (-ffp-contract=off) and of nvcc (-fmad=false): the GPU engine can be checked bit-exactly
 * (timestamps included) against the CPU oracle, and the oracle separately checks them against libm
 * (tests/test_rs_math.py: <= 1 ULP of glibc log over the tested range).
 *
 * rs_log: classic argument reduction x = 2^k * (1+f), sqrt(1/2) <= 1+f < sqrt(2), then
 * log(1+f) = 2s + s*R(s^2), s = f/(2+f), with a degree-14 odd minimax polynomial (the reduction and
 * coefficient set are the well known ones of Sun's freely distributable fdlibm e_log.c; restated
 * here, not taken from the reference tree, which contains no libm).
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD static inline
#endif

RS_HD uint64_t rs_d2u(double x)
{
	union { double d; uint64_t u; } c;
	c.d = x;
	return c.u;
}

RS_HD double rs_u2d(uint64_t u)
{
	union { double d; uint64_t u; } c;
	c.u = u;
	return c.d;
}

RS_HD double rs_log(double x)
{
	const double ln2_hi = 6.93147180369123816490e-01;	/* 0x3fe62e42fee00000 */
	const double ln2_lo = 1.90821492927058770002e-10;	/* 0x3dea39ef35793c76 */
	const double L1 = 6.666666666666735130e-01;
	const double L2 = 3.999999999940941908e-01;
	const double L3 = 2.857142874366239149e-01;
	const double L4 = 2.222219843214978396e-01;
	const double L5 = 1.818357216161805012e-01;
	const double L6 = 1.531383769920937332e-01;
	const double L7 = 1.479819860511658591e-01;

	uint64_t ux = rs_d2u(x);
	int k = 0;
	if ((ux >> 63) || ux == 0 || (ux >> 52) >= 0x7ff) {
		if ((ux << 1) == 0)
			return rs_u2d(0xfff0000000000000ULL);	/* log(+-0) = -inf */
		if (ux == 0x7ff0000000000000ULL)
			return x;				/* log(+inf) = +inf */
		return rs_u2d(0x7ff8000000000000ULL);		/* negative or NaN */
	}
	if ((ux >> 52) == 0) {					/* subnormal: scale by 2^54 */
		x = x * 18014398509481984.0;
		ux = rs_d2u(x);
		k = -54;
	}
	uint32_t hx = (uint32_t)(ux >> 32);
	k += (int)(hx >> 20) - 1023;
	hx &= 0x000fffffu;
	uint32_t i = (hx + 0x95f64u) & 0x100000u;		/* 1 when mantissa >= sqrt(2) */
	ux = ((uint64_t)(hx | (i ^ 0x3ff00000u)) << 32) | (ux & 0xffffffffULL);
	k += (int)(i >> 20);
	const double f = rs_u2d(ux) - 1.0;
	const double dk = (double)k;
	const double s = f / (2.0 + f);
	const double z = s * s;
	const double w = z * z;
	const double t1 = w * (L2 + w * (L4 + w * L6));
	const double t2 = z * (L1 + w * (L3 + w * (L5 + w * L7)));
	const double R = t2 + t1;
	const double hfsq = 0.5 * f * f;
	return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

/*
 * rs_exp: x = k*ln2 + r with |r| <= ln2/2 (r kept as hi - lo), exp(r) = 1 + r + r*c/(2 - c) with
 * c = r - r^2 * P(r^2), P a degree-5 minimax polynomial (reduction and coefficients: fdlibm e_exp.c),
 * result scaled by 2^k with exact power-of-two multiplications.  Used where a model calls exp() and the
 * result steers the trace (models/traffic/normal_cdf.c:62,76: the accident probability that a coin is
 * compared with); <= 1 ULP of glibc exp over the tested range (tests/test_rs_math.py).
 */
RS_HD double rs_exp(double x)
{
	const double ln2_hi = 6.93147180369123816490e-01;	/* 0x3fe62e42fee00000 */
	const double ln2_lo = 1.90821492927058770002e-10;	/* 0x3dea39ef35793c76 */
	const double inv_ln2 = 1.44269504088896338700e+00;
	const double P1 = 1.66666666666666019037e-01;
	const double P2 = -2.77777777770155933842e-03;
	const double P3 = 6.61375632143793436117e-05;
	const double P4 = -1.65339022054652515390e-06;
	const double P5 = 4.13813679705723846039e-08;

	const uint64_t ux = rs_d2u(x);
	const uint64_t ax = ux & 0x7fffffffffffffffULL;
	const int neg = (int)(ux >> 63);
	if (ax > 0x7ff0000000000000ULL)
		return x + x;						/* NaN */
	if (x > 7.09782712893383973096e+02)
		return rs_u2d(0x7ff0000000000000ULL);			/* overflow */
	if (x < -7.45133219101941108420e+02)
		return 0.0;						/* underflow */
	double hi = 0.0, lo = 0.0;
	int k = 0;
	if (ax > 0x3fd62e42fefa39efULL) {				/* |x| > ln2/2 */
		if (ax < 0x3ff0a2b23f3bab73ULL) {			/* |x| < 3 ln2/2 */
			hi = neg ? x + ln2_hi : x - ln2_hi;
			lo = neg ? -ln2_lo : ln2_lo;
			k = neg ? -1 : 1;
		} else {
			k = (int)(inv_ln2 * x + (neg ? -0.5 : 0.5));
			const double t = (double)k;
			hi = x - t * ln2_hi;
			lo = t * ln2_lo;
		}
		x = hi - lo;
	} else if (ax < 0x3e30000000000000ULL) {			/* |x| < 2^-28 */
		return 1.0 + x;
	}
	const double t = x * x;
	const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
	if (k == 0)
		return 1.0 - ((x * c) / (c - 2.0) - x);
	double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
	if (k >= -1021) {
		if (k == 1024)
			return y * 2.0 * rs_u2d(0x7fe0000000000000ULL);
		return y * rs_u2d((uint64_t)(1023 + k) << 52);
	}
	return y * rs_u2d((uint64_t)(1023 + k + 1000) << 52) * rs_u2d((uint64_t)(1023 - 1000) << 52);
}
