/*
 * jdde_math.h — the handful of scalar routines that must give the same bits on
 * the CPU (gcc, C11 or C++) and on the GPU (nvcc, -fmad=false verification build).
 *
 * Only +, -, * (correctly rounded under IEEE-754) and exact exponent
 * manipulation are used, so the results do not depend on a libm.
 *
 *  jdde_ipow(x,e)    x**e for integer e>=1 as a fixed square-and-multiply chain.
 *                    The reference prints `pow(x, e)` (SymEngine C printer) and
 *                    lets libm decide; SURVEY.md §7 hard part 1.
 *  jdde_inv_root3(p) p**(-1/3)  — step-size decrease factor, reference
 *                    /root/reference/jitcdde/_jitcdde.py:778  (`p**(-1/self.q)`, q = 3)
 *  jdde_inv_root4(p) p**(-1/4)  — step-size increase factor, reference
 *                    /root/reference/jitcdde/_jitcdde.py:783  (`p**(-1/(self.q+1))`)
 *
 * The two roots agree with libm's pow to ~1 ulp (checked in tests/test_oracle.py);
 * the oracle can be switched between them (root_mode) to show that the choice
 * does not change accepted-step counts.
 */
#ifndef JDDE_MATH_H
#define JDDE_MATH_H

#include <math.h>

#if defined(__CUDACC__)
#  define JDDE_HD __host__ __device__ __forceinline__
#else
#  define JDDE_HD static inline
#endif

JDDE_HD double jdde_ipow(double x, unsigned e)
{
	/* e = 10: x2 = x*x, x4 = x2*x2, x8 = x4*x4, result = x2*x8 */
	double result = 1.0;
	double base = x;
	int have = 0;
	while (e)
	{
		if (e & 1u)
		{
			result = have ? result*base : base;
			have = 1;
		}
		e >>= 1;
		if (e)
			base = base*base;
	}
	return result;
}

/* p**(-1/r) for r = 3, 4 by Newton's iteration on y**(-r) = m,  y <- y + y·(1 − m·y^r)/r,  written with
 * multiplications only (1/3 as a rounded constant, 1/4 exact): relative error e -> ((r+1)/2)·e², so five
 * steps take the cubic starting polynomial (error < 8 %) below one ulp. p = m·2^e is split exactly with
 * frexp/ldexp such that e is a multiple of r. */
JDDE_HD double jdde_inv_root4(double p)
{
	int e;
	double m = frexp(p, &e);
	int r = e % 4;
	if (r < 0)
		r += 4;
	int const k = (e - r)/4;
	m = ldexp(m, r);                 /* m in [0.5,8) */
	double y = 1.217525705905916 + m*(-0.25141147967130534 + m*(0.03938573046995773 + m*-0.002237737611829044));
	for (int i = 0; i < 5; i++)
	{
		double const y2 = y*y;
		double const y4 = y2*y2;
		y = y + y*(1.0 - m*y4)*0.25;
	}
	return ldexp(y, -k);
}

JDDE_HD double jdde_inv_root3(double p)
{
	int e;
	double m = frexp(p, &e);
	int r = e % 3;
	if (r < 0)
		r += 3;
	int const k = (e - r)/3;
	m = ldexp(m, r);                 /* m in [0.5,4) */
	double y = 1.4841220496057164 + m*(-0.6341672442901666 + m*(0.1816221035823464 + m*-0.01927416318870279));
	for (int i = 0; i < 5; i++)
	{
		double const y3 = y*y*y;
		y = y + y*(1.0 - m*y3)*0x1.5555555555555p-2;
	}
	return ldexp(y, -k);
}

/* Both roots through ONE code path: p**(-1/4) if `four`, else p**(-1/3). Bit-identical with jdde_inv_root4 / jdde_inv_root3
 * (tests/test_oracle.py): the same polynomial, the same five Newton steps, the same exact exponent split — but the split
 * is made on the bits of p (normal numbers; anything else takes the functions above) instead of frexp/ldexp, and the
 * two variants differ only in selected constants. In the kernels every lane of a warp that decreases OR increases its
 * step runs these ≈ 45 instructions together, where two separate functions with libm's frexp/ldexp were ≈ 130
 * instructions each, executed one after the other by a few lanes (profiles/: 4.9 and 6.7 of 32 lanes active). */
JDDE_HD double jdde_inv_root(double p, int four)
{
#if defined(__CUDA_ARCH__)
	unsigned long long const bits = (unsigned long long)__double_as_longlong(p);
#else
	union { double d; unsigned long long u; } pun;
	pun.d = p;
	unsigned long long const bits = pun.u;
#endif
	unsigned const be = (unsigned)(bits >> 52);           /* sign and biased exponent */
	if (be-1u >= 0x7feu)                                  /* zero, subnormal, negative, infinite or NaN */
		return four ? jdde_inv_root4(p) : jdde_inv_root3(p);
	/* p = m·2^e with m in [0.5,1); e+1032 > 0 and 1032 is a multiple of 3 and of 4 */
	unsigned const eo = be-1022u+1032u;
	unsigned const q3 = (eo*43691u) >> 17;                /* eo/3 for eo < 2^15 */
	unsigned const quotient = four ? eo >> 2 : q3;
	unsigned const rem = four ? eo & 3u : eo-3u*q3;
	int const k = (int)quotient - (four ? 258 : 344);
	unsigned long long const m_bits = (bits & 0x000fffffffffffffull) | ((unsigned long long)(1022u+rem) << 52);   /* m·2^rem, exact */
	unsigned long long const scale_bits = (unsigned long long)(1023-k) << 52;                                       /* 2^-k */
#if defined(__CUDA_ARCH__)
	double const m = __longlong_as_double((long long)m_bits);
	double const scale = __longlong_as_double((long long)scale_bits);
#else
	pun.u = m_bits;
	double const m = pun.d;
	pun.u = scale_bits;
	double const scale = pun.d;
#endif
	double const c0 = four ? 1.217525705905916 : 1.4841220496057164;
	double const c1 = four ? -0.25141147967130534 : -0.6341672442901666;
	double const c2 = four ? 0.03938573046995773 : 0.1816221035823464;
	double const c3 = four ? -0.002237737611829044 : -0.01927416318870279;
	double const step = four ? 0.25 : 0x1.5555555555555p-2;
	double y = c0 + m*(c1 + m*(c2 + m*c3));
	for (int i = 0; i < 5; i++)
	{
		double const y2 = y*y;
		double const yr = y2*(four ? y2 : y);
		y = y + y*(1.0 - m*yr)*step;
	}
	return y*scale;
}

/* ------------------------------------------------------------------------------------------------
 * Elementary functions for right-hand sides of the VERIFICATION build: sin, cos, exp, tanh, log with
 * the same bits on the CPU (gcc -ffp-contract=off: oracle, reference core with the "chain" right-hand
 * side) and on the GPU (nvcc -fmad=false). The reference prints libm calls; CUDA's and glibc's libm
 * differ in the last bit now and then, which a DDE with sensitive dynamics amplifies to a different
 * accept/reject decision — so a bit-for-bit comparison needs one implementation for both sides.
 * Only correctly rounded operations (+ − × ÷, rint, exact exponent arithmetic) in a fixed order;
 * classical argument reductions (Cody–Waite with a three-part π/2, ln 2 in two parts) and the
 * well-known minimax polynomials for the reduced ranges. Accuracy: ≤ 1 ulp on the tested ranges
 * (tests/test_oracle.py), which is libm's own class. The production build calls CUDA's functions.
 * Outside the range where the reduction is exact (|x| ≥ 2^20·π/2 for sin/cos) libm takes over.
 * ------------------------------------------------------------------------------------------------ */

JDDE_HD double jdde_bits_to_double(unsigned long long u)
{
#if defined(__CUDA_ARCH__)
	return __longlong_as_double((long long)u);
#else
	union { double d; unsigned long long u; } pun;
	pun.u = u;
	return pun.d;
#endif
}

JDDE_HD unsigned long long jdde_double_to_bits(double d)
{
#if defined(__CUDA_ARCH__)
	return (unsigned long long)__double_as_longlong(d);
#else
	union { double d; unsigned long long u; } pun;
	pun.d = d;
	return pun.u;
#endif
}

/* x − n·π/2 as head + tail for |n| < 2^20: π/2 = P1 + P2 + P3 (+ …) with 33-bit P1, P2, so that n·P1 and n·P2 are exact */
JDDE_HD int jdde_reduce_pio2(double x, double * head, double * tail)
{
	double const fn = rint(x*6.36619772367581382433e-01);
	double const r1 = x - fn*1.57079632673412561417e+00;               /* P1 */
	double const w1 = fn*6.07710050630396597660e-11;                   /* P2 */
	double const r2 = r1 - w1;
	double const w2 = fn*2.02226624879595063154e-21 - ((r1 - r2) - w1);  /* P3 and what the last subtraction lost */
	double const y0 = r2 - w2;
	*head = y0;
	*tail = (r2 - y0) - w2;
	return (int)fn;
}

/* sin on [−π/4, π/4] for the argument x + y (y: tail) */
JDDE_HD double jdde_kernel_sin(double x, double y)
{
	double const z = x*x;
	double const v = z*x;
	double const r = 8.33333333332248946124e-03 + z*(-1.98412698298579493134e-04 + z*(2.75573137070700676789e-06 + z*(-2.50507602534068634195e-08 + z*1.58969099521155010221e-10)));
	return x - ((z*(0.5*y - v*r) - y) - v*-1.66666666666666324348e-01);
}

/* cos on [−π/4, π/4] for the argument x + y */
JDDE_HD double jdde_kernel_cos(double x, double y)
{
	double const z = x*x;
	double const r = z*(4.16666666666666019037e-02 + z*(-1.38888888888741095749e-03 + z*(2.48015872894767294178e-05 + z*(-2.75573143513906633035e-07 + z*(2.08757232129817482790e-09 + z*-1.13596475577881948265e-11)))));
	double const hz = 0.5*z;
	double const w = 1.0 - hz;
	return w + (((1.0 - w) - hz) + (z*r - x*y));
}

JDDE_HD double jdde_sin(double x)
{
	if (!(fabs(x) < 1647099.0))        /* 2^20·π/2; also NaN and infinity */
		return sin(x);
	if (x == 0.0)
		return x;                      /* keeps the sign of zero */
	double head, tail;
	int const n = jdde_reduce_pio2(x, &head, &tail);
	switch (n & 3)
	{
		case 0: return jdde_kernel_sin(head, tail);
		case 1: return jdde_kernel_cos(head, tail);
		case 2: return -jdde_kernel_sin(head, tail);
		default: return -jdde_kernel_cos(head, tail);
	}
}

JDDE_HD double jdde_cos(double x)
{
	if (!(fabs(x) < 1647099.0))
		return cos(x);
	double head, tail;
	int const n = jdde_reduce_pio2(x, &head, &tail);
	switch (n & 3)
	{
		case 0: return jdde_kernel_cos(head, tail);
		case 1: return -jdde_kernel_sin(head, tail);
		case 2: return -jdde_kernel_cos(head, tail);
		default: return jdde_kernel_sin(head, tail);
	}
}

/* exp(x) = 2^k·exp(r), r = x − k·ln 2 in [−ln2/2, ln2/2], exp(r) = 1 + r + r·c/(2 − c) with c = r − r²·P(r²) */
JDDE_HD double jdde_exp(double x)
{
	if (!(x > -708.0 && x < 709.0))    /* results below the normal range or above it, NaN: libm's business */
		return exp(x);
	double const fk = rint(x*1.44269504088896338700e+00);
	int const k = (int)fk;
	double const hi = x - fk*6.93147180369123816490e-01;      /* ln 2, upper 32 bits: k·hi-part exact */
	double const lo = fk*1.90821492927058770002e-10;
	double const r = hi - lo;
	double const t = r*r;
	double const c = r - t*(1.66666666666666019037e-01 + t*(-2.77777777770155933842e-03 + t*(6.61375632143793436117e-05 + t*(-1.65339022054652515390e-06 + t*4.13813679705723846039e-08))));
	double const y = 1.0 - ((lo - (r*c)/(2.0 - c)) - hi);
	/* y in (0.7, 1.5), k in [−1022, 1023]: the scaling is exact (two factors, so that each is a normal number) */
	int const k1 = k/2, k2 = k - k1;
	return (y*jdde_bits_to_double((unsigned long long)(1023+k1) << 52))*jdde_bits_to_double((unsigned long long)(1023+k2) << 52);
}

/* tanh: continued fraction x/(1 + x²/(3 + x²/(5 + …))) for |x| < 0.5, (1 − e)/(1 + e) with e = exp(−2|x|) beyond */
JDDE_HD double jdde_tanh(double x)
{
	double const a = fabs(x);
	if (a != a)
		return x;
	double result;
	if (a < 0.5)
	{
		double const z = a*a;
		double f = 21.0;
		f = 19.0 + z/f;
		f = 17.0 + z/f;
		f = 15.0 + z/f;
		f = 13.0 + z/f;
		f = 11.0 + z/f;
		f = 9.0 + z/f;
		f = 7.0 + z/f;
		f = 5.0 + z/f;
		f = 3.0 + z/f;
		f = 1.0 + z/f;
		result = a/f;
	}
	else if (a < 20.0)
	{
		double const e = jdde_exp(-2.0*a);
		result = (1.0 - e)/(1.0 + e);
	}
	else
		result = 1.0;
	return (jdde_double_to_bits(x) >> 63) ? -result : result;
}

/* log(x) = k·ln 2 + log(m), m in [√½, √2): log(m) = 2s + s·(z·P(z)) with s = (m−1)/(m+1), z = s² — in the compensated form
 * f − (hfsq − (s·(hfsq + R) + (k·ln2_lo + …))) that keeps the error below an ulp */
JDDE_HD double jdde_log(double x)
{
	unsigned long long const bits = jdde_double_to_bits(x);
	unsigned const be = (unsigned)(bits >> 52);
	if (be-1u >= 0x7feu)               /* zero, subnormal, negative, infinite, NaN */
		return log(x);
	int k = (int)be - 1023;
	double m = jdde_bits_to_double((bits & 0x000fffffffffffffull) | (1023ull << 52));    /* [1, 2) */
	if (m > 1.41421356237309514547)
	{
		m = m*0.5;
		k += 1;
	}
	double const f = m - 1.0;
	double const s = f/(2.0 + f);
	double const z = s*s;
	double const w = z*z;
	double const t1 = w*(3.999999999940941908e-01 + w*(2.222219843214978396e-01 + w*1.531383769920937332e-01));
	double const t2 = z*(6.666666666666735130e-01 + w*(2.857142874366239149e-01 + w*(1.818357216161805012e-01 + w*1.479819860511658591e-01)));
	double const R = t2 + t1;
	double const hfsq = 0.5*f*f;
	double const dk = (double)k;
	return dk*6.93147180369123816490e-01 - ((hfsq - (s*(hfsq + R) + dk*1.90821492927058770002e-10)) - f);
}

/* A uniform number in [0, 1) that is a pure function of (seed, stream, counter): what pws_fuzzy_increase draws
 * (/root/reference/jitcdde/_jitcdde.py:730-731, `self.rng.random() < p`). The reference consumes a NumPy generator
 * sequentially; an ensemble needs a draw that does not depend on which trajectory asks first, so every trajectory
 * (stream) indexes its own sequence by the number of right-hand-side evaluations it has made (counter). Two rounds of the
 * SplitMix64 finaliser over the combined words; the same bits on CPU and GPU. */
JDDE_HD double jdde_uniform(unsigned seed, unsigned long long stream, unsigned long long counter)
{
	unsigned long long z = (unsigned long long)seed*0x9E3779B97F4A7C15ull ^ stream*0xD1B54A32D192ED03ull ^ counter*0x8CB92BA72F3D8DD7ull;
	for (int round = 0; round < 2; round++)
	{
		z += 0x9E3779B97F4A7C15ull;
		z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27))*0x94D049BB133111EBull;
		z ^= z >> 31;
	}
	return (double)(z >> 11)*0x1.0p-53;
}

/* what the generated right-hand sides call (jitcdde_b200/_codegen.py) */
#if defined(JD_FAST) || defined(JDDE_LIBM)
#  define JDDE_SIN(x) sin(x)
#  define JDDE_COS(x) cos(x)
#  define JDDE_EXP(x) exp(x)
#  define JDDE_TANH(x) tanh(x)
#  define JDDE_LOG(x) log(x)
#else
#  define JDDE_SIN(x) jdde_sin(x)
#  define JDDE_COS(x) jdde_cos(x)
#  define JDDE_EXP(x) jdde_exp(x)
#  define JDDE_TANH(x) jdde_tanh(x)
#  define JDDE_LOG(x) jdde_log(x)
#endif

#endif
