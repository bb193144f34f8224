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

#endif
