/*
 * jdde_math.h — the handful of scalar routines that must give the same bits on
 * the CPU (gcc, C11 or C++) and on the GPU (nvcc, -fmad=false verification build).
 *
 * Only +, -, *, /, sqrt (all correctly rounded under IEEE-754) and exact
 * exponent manipulation are used, so the results do not depend on a libm.
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

JDDE_HD double jdde_inv_root4(double p)
{
	return 1.0/sqrt(sqrt(p));
}

JDDE_HD double jdde_inv_root3(double p)
{
	/* p = m·2^e with m in [0.5,1); fold e mod 3 into m so that m in [0.5,4) */
	int e;
	double m = frexp(p, &e);
	int r = e % 3;
	if (r < 0)
		r += 3;
	int const k = (e - r)/3;
	m = ldexp(m, r);
	/* quadratic start for m**(-1/3) on [0.5,4), relative error < 6 % */
	double y = 1.4774 + m*(-0.4930 + m*0.0706);
	/* Newton on y**(-3) = m:  y <- y + y·(1 − m·y³)/3, error e -> 2e² */
	for (int i = 0; i < 6; i++)
	{
		double const y3 = y*y*y;
		y = y + y*(1.0 - m*y3)/3.0;
	}
	return ldexp(y, -k);
}

#endif
