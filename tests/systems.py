"""
The DDE systems used by tests, golden-vector generation and bench.py. Each is the system of
one of the reference's tests or examples (file:line cited), written with this package's
SymPy symbols.
"""

import numpy as np
import sympy
from sympy import Symbol, exp, sin, sqrt, tanh

from jitcdde_b200 import anchors, current_y, input, past_dy, past_y, t, y  # noqa: A004
from jitcdde_b200 import _codegen


def mackey_glass(control=True):
	"""examples/mackey_glass.py:64-72; with control parameters (β, τ) as in BASELINE.json configs[1]."""
	if control:
		beta, tau = sympy.symbols("beta tau")
		return dict(f=[beta*y(0, t-tau)/(1+y(0, t-tau)**10) - 0.1*y(0)], control_pars=[beta, tau])
	return dict(f=[0.25*y(0, t-15)/(1+y(0, t-15)**10) - 0.1*y(0)], control_pars=[])


def mackey_glass_kat():
	"""tests/test_python_core.py:13-27"""
	return dict(f=[0.25*y(0, t-15)/(1.0+y(0, t-15)**10) - 0.1*y(0, t)], control_pars=[])


def two_roessler(k=0.25, control=False):
	"""tests/test_jitcdde.py:30-41 (also tests/test_lyaps.py:16-27, examples/two_Roesslers.py)"""
	omega = np.array([0.88167179, 0.87768425])
	delay = 4.5
	control_pars = []
	if control:
		k = Symbol("k")
		control_pars = [k]
	f = [
		omega[0] * (-y(1) - y(2)),
		omega[0] * (y(0) + 0.165 * y(1)),
		omega[0] * (0.2 + y(2) * (y(0) - 10.0)),
		omega[1] * (-y(4) - y(5)) + k * (y(0, t-delay) - y(3)),
		omega[1] * (y(3) + 0.165 * y(4)),
		omega[1] * (0.2 + y(5) * (y(3) - 10.0)),
	]
	return dict(f=f, control_pars=control_pars, delay=delay)


ROESSLER_TEST_PARAMETERS = dict(rtol=1e-4, atol=1e-7, pws_rtol=1e-3, pws_atol=1e-3, first_step=30, max_step=100, min_step=1e-16)   # tests/test_jitcdde.py:51-59
LYAP_TEST_PARAMETERS = dict(rtol=1e-7, atol=1e-7, pws_rtol=1e-3, pws_atol=1e-3, first_step=30, max_step=100, min_step=1e-30)     # tests/test_lyaps.py:28-36


def fitzhugh_nagumo_ode():
	"""tests/test_ode.py:19-46: an ODE (no delay)"""
	a, b1, b2, c, k = -0.025794, 0.0065, 0.0135, 0.02, 0.128
	f = [
		y(0) * (a-y(0)) * (y(0)-1.0) - y(1) + k * (y(2) - y(0)),
		b1*y(0) - c*y(1),
		y(2) * (a-y(2)) * (y(2)-1.0) - y(3) + k * (y(0) - y(2)),
		b2*y(2) - c*y(3),
	]
	y0 = np.array([-0.00338158, -0.00223185, 0.01524253, -0.00613449])
	y1 = np.array([0.0011789485114731, -0.0021947158873226, 0.0195744683782066, -0.0057801623466600])
	f_of_y0 = np.array([0.0045396904008868, 0.00002265673, 0.0043665702488807, 0.000328463955])
	return dict(f=f, control_pars=[], y0=y0, y1=y1, f_of_y0=f_of_y0)


def neutral():
	"""tests/test_neutral.py:16-56 (examples/neutral.py): past_dy and an anchor helper"""
	sech = lambda x: 2/(exp(x)+exp(-x))
	eps = 1e-5
	abs_ = lambda x: sqrt(x**2+eps**2)
	ε = [0.03966, 0.03184, 0.02847]
	ν = [1, 2.033, 3.066]
	μ = [0.16115668456085775, 0.14093420256851111, 0.11465065353644151]
	ybar_0 = 0
	τ = 1.7735
	ζ = [0.017940997406325931, 0.015689701773967984, 0.012763648066925721]
	anchors_past = Symbol("anchors_past")
	difference = Symbol("difference")
	factor_μ = Symbol("factor_μ")
	factor_ζ = Symbol("factor_ζ")
	ydot = [y(i) for i in range(3, 6)]
	y_tot = sum(current_y(i) for i in range(3))
	ydot_tot = sum(current_y(i) for i in range(3, 6))
	y_past = sum(past_y(t-τ, i, anchors_past) for i in range(3))
	ydot_past = sum(past_y(t-τ, i, anchors_past) for i in range(3, 6))
	yddot_past = sum(past_dy(t-τ, i, anchors_past) for i in range(3, 6))
	helpers = [
		(anchors_past, anchors(t-τ)),
		(difference, ybar_0-y_past),
		(factor_μ, sech(difference)**2 * (yddot_past + 2*ydot_past**2*tanh(difference))),
		(factor_ζ, 2*abs_(y_tot)*ydot_tot),
	]
	f = {y(i): ydot[i] for i in range(3)}
	f.update({ydot[i]: μ[i]*factor_μ - ζ[i]*factor_ζ - ε[i]*ν[i]*ydot[i] - ν[i]**2*y(i) for i in range(3)})
	initial = [0.66698806, 0.02581308, -0.77761941, 0.94863382, 0.70167179, -1.05108156]
	control = [-0.72689205, 0.09982525, 0.41167269, 0.86527044, -0.50028942, 0.7227548]
	return dict(f=f, helpers=helpers, control_pars=[], delay=τ, initial=initial, control=control, T=5)


def state_dependent():
	"""examples/state_dependent.py:14-22"""
	f = [-y(0, t-2-y(0)**2) + 5]
	eps = 0.1
	past = [(-2.0, [4.5], [0.0]), (-1.0-eps, [4.5], [0.0]), (-1.0+eps, [-0.5], [0.0]), (0.0, [-0.5], [0.0])]
	return dict(f=f, control_pars=[], past=past, max_delay=1e10)


def kuramoto(n=6, seed=42):
	"""examples/kuramoto_network.py:42-58, small: many lookup sites with different delays"""
	rng = np.random.default_rng(seed)
	ω, c, q = 1, 42, 0.5
	A = rng.choice([1, 0], size=(n, n), p=[q, 1-q])
	τ = rng.uniform(np.pi/5, 2*np.pi, size=(n, n))
	f = [ω + c/(n-1)*sum(sin(y(j, t-τ[i, j])-y(i)) for j in range(n) if A[j, i]) for i in range(n)]
	return dict(f=f, control_pars=[], delays=τ.flatten(), max_delay=float(np.max(τ)), initial=rng.uniform(0, 2*np.pi, n))


def data_input(seed=42):
	"""examples/data_input.py:19-50 of the reference: two delay-coupled relaxators driven by irregularly sampled data"""
	rng = np.random.default_rng(seed)
	times = np.sort(rng.uniform(0, 70, 40))
	times[0], times[-1] = 0.0, 70.0
	data = np.stack([np.sin(0.3*times)+0.1*rng.normal(size=40), np.cos(0.2*times)+0.1*rng.normal(size=40)], axis=1)
	slopes = np.gradient(data, times, axis=0)
	tau = 10.0
	f = [-y(0) + y(1, t-tau) + input(0), -y(1) + y(0, t-tau) + input(1)]
	return dict(f=f, input=[(times[k], data[k], slopes[k]) for k in range(40)], tau=tau, initial=np.zeros(2))


def data_input_program(system=None):
	"""(program, joined past, max_delay) of the extended system jitcdde_input hands to the core"""
	from jitcdde_b200 import jitcdde_input
	system = system or data_input()
	DDE = jitcdde_input(system["f"], system["input"], verbose=False)
	DDE.constant_past(system["initial"])
	program = _codegen.generate(list(DDE.f_sym()), helpers=DDE.helpers, control_pars=DDE.control_pars)
	return program, [(a.time, a.state, a.diff) for a in DDE.past], DDE.max_delay


def synchronised_fhn():
	"""tests/test_transversal_lyap.py:22-52, scenario "grouped by variables": three FitzHugh–Nagumo-like oscillators
	on a ring with a delayed inhibitor, synchronised; coupling k is a control parameter"""
	a, b, c, tau = -0.025794, 0.01, 0.02, 1.0
	k = sympy.Symbol("k")
	f = [
		y(0)*(a-y(0))*(y(0)-1.0) - y(3) + k*(y(1)-y(0)),
		y(1)*(a-y(1))*(y(1)-1.0) - y(4) + k*(y(2)-y(1)),
		y(2)*(a-y(2))*(y(2)-1.0) - y(5) + k*(y(0)-y(2)),
		b*y(0) - c*y(3, t-tau),
		b*y(1) - c*y(4, t-tau),
		b*y(2) - c*y(5, t-tau),
	]
	zero = [0.]*6
	vectors = [([1., 1., 1., 0., 0., 0.], zero), (zero, [1., 1., 1., 0., 0., 0.]), ([0., 0., 0., 1., 1., 1.], zero), (zero, [0., 0., 0., 1., 1., 1.])]
	return dict(f=f, control_pars=[k], vectors=vectors, groups=([0, 1, 2], [3, 4, 5]), tau=tau)


def projected_program(kind, system=None, **kw):
	"""(DDE object, program, projector description) of jitcdde_restricted_lyap / jitcdde_transversal_lyap for the
	synchronised oscillators, generated exactly as the class's compile_C does"""
	from jitcdde_b200 import jitcdde_restricted_lyap, jitcdde_transversal_lyap
	system = system or synchronised_fhn()
	if kind == "restricted":
		DDE = jitcdde_restricted_lyap(system["f"], vectors=system["vectors"], control_pars=system["control_pars"], verbose=False, **kw)
		simplify = False      # as the reference's test: simplification breaks the symmetry that keeps the dynamics on the manifold
		projector = dict(mode="restricted", vectors=DDE.vectors, state_components=DDE.state_components, diff_components=DDE.diff_components)
	else:
		DDE = jitcdde_transversal_lyap(system["f"], groups=system["groups"], control_pars=system["control_pars"], verbose=False, **kw)
		simplify = False
		projector = dict(mode="transversal", indices=DDE.tangent_indices)
	blocks = (DDE.n_basic, DDE.n//DDE.n_basic-1) if DDE.n_basic != DDE.n else None
	program = _codegen.generate(list(DDE.f_sym()), helpers=DDE.helpers, control_pars=DDE.control_pars, simplify=simplify, blocks=blocks)
	return DDE, program, projector


def as_list(f):
	"""f as a list of expressions (dict keyed by y(i) → ordered list)"""
	if isinstance(f, dict):
		out = [None]*len(f)
		for key, value in f.items():
			out[int(sympy.sympify(key).args[0])] = value
		return out
	return list(f)


def program(system, **kwargs):
	return _codegen.generate(as_list(system["f"]), helpers=system.get("helpers", ()), control_pars=system.get("control_pars", ()), **kwargs)


def lyap_program(system, n_lyap, simplify=True):
	"""(program, n_basic) of the extended system of jitcdde_lyap"""
	basic = as_list(system["f"])
	helpers = system.get("helpers", ())
	delays = _codegen.get_delays(basic, helpers)
	extended = _codegen.tangent_vector_f(basic, helpers, len(basic), n_lyap, delays, simplify=simplify)
	return _codegen.generate(extended, control_pars=system.get("control_pars", ()), blocks=(len(basic), n_lyap)), len(basic)


def mg_grid(N, seed_stride=True):
	"""
	Deterministic (β, τ) members of the BASELINE.json configs[1] grid:
	β_i = 0.15 + 0.20·i/2499 (i < 2500), τ_j = 10 + 15·j/3999 (j < 4000), member = i·4000 + j.
	For N < 10^7 the members are taken evenly strided through the grid.
	"""
	total = 2500*4000
	members = (np.arange(N, dtype=np.int64)*total)//N if N < total else np.arange(N, dtype=np.int64) % total
	i, j = members//4000, members % 4000
	return np.stack([0.15 + 0.20*i/2499.0, 10 + 15*j/3999.0], axis=1)
