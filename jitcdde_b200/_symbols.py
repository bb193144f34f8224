"""
Symbols used to write down a delay differential equation.

Same names and meaning as the reference's symbol layer
(/root/reference/jitcdde/_jitcdde.py:24-58 and the SymPy spelling in
/root/reference/jitcdde/sympy_symbols.py:1-50). The reference builds them on
SymEngine; this front-end is SymPy-based because the code generator
(`_codegen.py`) walks SymPy trees.
"""

import sympy

#: the symbol for time
t = sympy.Symbol("t", real=True)

#: current state, `current_y(i)` = i-th component of y(t)
current_y = sympy.Function("current_y", real=True)

#: past state, `past_y(time, i, anchors(time))`
past_y = sympy.Function("past_y", real=True)

#: past derivative, `past_dy(time, i, anchors(time))`
past_dy = sympy.Function("past_dy", real=True)

#: the pair of anchors bracketing a time, `anchors(time)`
anchors = sympy.Function("anchors", real=True)


def y(index, time=t):
	"""i-th component of the state at `time` (default: now).

	Expands to `current_y` or `past_y(…, anchors(…))` exactly as the reference
	does (/root/reference/jitcdde/_jitcdde.py:26-33)."""
	if time == t:
		return current_y(index)
	return past_y(time, index, anchors(time))


def dy(index, time):
	"""i-th component of the derivative at a past `time` (neutral DDEs;
	/root/reference/jitcdde/_jitcdde.py:35-46)."""
	if time == t:
		raise ValueError("Do not use `dy` to compute the current derivative. Use your dynamical equations instead.")
	return past_dy(time, index, anchors(time))


def quadrature(integrand, variable, lower, upper, nsteps=20, method="gauss"):
	"""Estimator of an integral over past points from evaluations at discrete
	points (/root/reference/jitcdde/_jitcdde.py:88-126): midpoint rule or
	Gauß–Legendre with `nsteps` nodes."""
	sample = lambda pos: integrand.subs(variable, pos)

	if method == "midpoint":
		half_step = sympy.simplify(sympy.sympify(upper-lower)/sympy.sympify(nsteps)/2)
		return 2*half_step*sum(
				sample(lower+(1+2*i)*half_step)
				for i in range(nsteps)
			)
	elif method == "gauss":
		from sympy.integrals.quadrature import gauss_legendre
		factor = sympy.simplify(sympy.sympify(upper-lower)/2)
		return factor*sum(
				weight*sample(lower+(1+pos)*factor)
				for pos, weight in zip(*gauss_legendre(nsteps, 20), strict=True)
			)
	else:
		raise NotImplementedError(f"I know no integration method named {method}.")


#: placeholders resolved by `jitcdde_input` (reference: _jitcdde.py:1509-1510)
input_shift = sympy.Symbol("external_input", real=True, negative=False)
input_base_n = sympy.Symbol("input_base_n", integer=True, negative=False)


def input(index, time=t):  # noqa: A001
	"""External input for `jitcdde_input` (/root/reference/jitcdde/_jitcdde.py:1511-1517): component `index` of the
	input at `time` (default: now). Expands to a delayed lookup of a dummy dynamical variable."""
	if time != t:
		from warnings import warn
		warn("Do not use delayed inputs unless you also have undelayed inputs (otherwise you can just shift your time frame). If you use delayed and undelayed inputs, you have to rely on automatic delay detection or explicitly add `input[-1].time+delay` to the delays (and consider it as a max_delay) for each delayed input.", stacklevel=2)
	return y(index+input_base_n, time-input_shift)
