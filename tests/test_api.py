"""
The reference's own API tests, restated for this package (tests/test_jitcdde.py, tests/test_past_from_function.py of
the reference): every way of defining the same two-Rössler system — list, generator function, dictionary, helpers,
control parameters (tuple and list form), common-subexpression off, a tiny extra delay that forces the
past-within-step iteration — must reproduce the reference's golden vector y(T=10) within the reference's tolerance
(rtol 1e-3, atol 1e-6), through all the ways of getting there (samples, one big step, adjust_diff first, zero jump,
annihilating jumps, many tiny samples). Runs on the host twin here and on the GPU with -m gpu.
"""

import os

import numpy as np
import pytest
import sympy

from jitcdde_b200 import UnsuccessfulIntegration, _find_max_delay, _get_delays, jitcdde, quadrature, t, y
from tests import scenarios, systems

omega = np.array([0.88167179, 0.87768425])
delay = 4.5
T = 10
test_parameters = dict(rtol=1e-4, atol=1e-7, pws_rtol=1e-3, pws_atol=1e-3, first_step=30, max_step=100, min_step=1e-16)
tolerance = dict(rtol=1e-3, atol=1e-6)

f = [
	omega[0] * (-y(1) - y(2)),
	omega[0] * (y(0) + 0.165 * y(1)),
	omega[0] * (0.2 + y(2) * (y(0) - 10.0)),
	omega[1] * (-y(4) - y(5)) + 0.25 * (y(0, t-delay) - y(3)),
	omega[1] * (y(3) + 0.165 * y(4)),
	omega[1] * (0.2 + y(5) * (y(3) - 10.0)),
]


def f_generator():
	yield from f


tiny_delay = 1e-30
f_with_tiny_delay = [
	omega[0] * (-y(1) - y(2)),
	omega[0] * (y(0) + 0.165 * y(1, t-tiny_delay)),
	omega[0] * (0.2 + y(2) * (y(0) - 10.0)),
	omega[1] * (-y(4) - y(5)) + 0.25 * (y(0, t-delay) - y(3, t-tiny_delay)),
	omega[1] * (y(3) + 0.165 * y(4)),
	omega[1] * (0.2 + y(5) * (y(3) - 10.0)),
]

delayed_y, y3m10, coupling_term = sympy.symbols("delayed_y y3m10 coupling_term")
f_alt_helpers = [(delayed_y, y(0, t-delay)), (coupling_term, 0.25 * (delayed_y - y(3))), (y3m10, y(3)-10)]
f_alt = [
	omega[0] * (-y(1) - y(2)),
	omega[0] * (y(0) + 0.165 * y(1)),
	omega[0] * (0.2 + y(2) * (y(0) - 10.0)),
	omega[1] * (-y(4) - y(5)) + coupling_term,
	omega[1] * (y3m10 + 10 + 0.165 * y(4)),
	omega[1] * (0.2 + y(5) * y3m10),
]

a, b, c, k, tau = sympy.symbols("a b c k tau")
parameters = [0.165, 0.2, 10.0, 0.25, 4.5]
f_params = [
	omega[0] * (-y(1) - y(2)),
	omega[0] * (y(0) + a * y(1)),
	omega[0] * (b + y(2) * (y(0) - c)),
	omega[1] * (-y(4) - y(5)) + k * (y(0, t-tau) - y(3)),
	omega[1] * (y(3) + a * y(4)),
	omega[1] * (b + y(5) * (y(3) - c)),
]

VARIANTS = {
	"list": lambda: jitcdde(f, verbose=False),
	"generator": lambda: jitcdde(f_generator, verbose=False),
	"dictionary": lambda: jitcdde({y(i): entry for i, entry in enumerate(f)}, verbose=False),
	"helpers": lambda: jitcdde(f_alt, helpers=f_alt_helpers, verbose=False),
	"automatic_anchor_helpers": lambda: jitcdde(f, automatic_anchor_helpers=True, verbose=False),
	"no_cse": lambda: jitcdde(f, verbose=False),
	"parameters": lambda: jitcdde(f_params, control_pars=[a, b, c, k, tau], max_delay=parameters[-1], verbose=False),
	"parameters_list": lambda: jitcdde(f_params, control_pars=[a, b, c, k, tau], max_delay=parameters[-1], verbose=False),
	"past_within_step": lambda: jitcdde(f_with_tiny_delay, verbose=False),
}
NO_JUMPS = {"past_within_step"}      # "Because the minimum delay must be larger than the jump width" (tests/test_jitcdde.py:171-174)


GPU_VARIANTS = ["list", "helpers", "parameters", "past_within_step"]


def prebuild_images():
	"""called by __graft_entry__.build(): compile the kernel images of the GPU variants on the CPU box"""
	for name in GPU_VARIANTS:
		VARIANTS[name]().compile_C()


def prepared(name):
	g = scenarios.golden("roessler")
	DDE = VARIANTS[name]()
	DDE.set_integration_parameters(**test_parameters)
	DDE.add_past_points(scenarios.past_of(g))
	DDE.check()
	if name == "no_cse":
		DDE.compile_C(do_cse=False)
	if name == "parameters":
		DDE.set_parameters(*parameters)
	if name == "parameters_list":
		DDE.set_parameters(parameters)
	DDE.initial_discontinuities_handled = True
	return DDE, g["y10_reference"]


def run_variant(name):
	results = {}
	DDE, y10 = prepared(name)
	for time in np.linspace(0, T, 10, endpoint=True):
		value = DDE.integrate(time)
	results["samples"] = value
	DDE, _ = prepared(name)
	results["one_big_step"] = DDE.integrate(T)
	DDE, _ = prepared(name)
	for time in np.linspace(0.0, T, 1000, endpoint=True):
		value = DDE.integrate(time)
	results["tiny_steps"] = value
	if name not in NO_JUMPS:
		DDE, _ = prepared(name)
		DDE.adjust_diff()
		results["adjust_diff"] = DDE.integrate(T)
		DDE, _ = prepared(name)
		DDE.integrate(T/2)
		DDE.jump(np.zeros(6), T/2)
		results["zero_jump"] = DDE.integrate(T)
		DDE, _ = prepared(name)
		rng = np.random.default_rng(seed=42)
		DDE.integrate(T/2)
		change = rng.normal(0, 10, 6)
		DDE.jump(change, T/2)
		DDE.jump(-change, T/2)
		results["annihilating_jumps"] = DDE.integrate(T)
	return results, y10


def check_variant(name):
	results, y10 = run_variant(name)
	for how, value in results.items():
		np.testing.assert_allclose(value, y10, err_msg=f"{name}/{how}", **tolerance)
	# sampling grid does not influence the integration
	assert np.array_equal(results["samples"], results["one_big_step"])
	assert np.array_equal(results["samples"], results["tiny_steps"])


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_input_variants_reproduce_golden_vector(twin, name):
	check_variant(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GPU_VARIANTS)
def test_input_variants_reproduce_golden_vector_gpu(name):
	check_variant(name)


def test_jump_places_two_anchors(twin):
	"""tests/test_jitcdde.py:353-371"""
	g = scenarios.golden("roessler")
	DDE = jitcdde(f, verbose=False)
	DDE.set_integration_parameters(**test_parameters)
	DDE.add_past_points(scenarios.past_of(g))
	width, change = 1e-5, np.random.default_rng(1).normal(0, 5, 6)
	before = DDE.integrate(T)
	past = DDE.get_state()
	DDE.jump(change, T, width)
	past = DDE.get_state()
	assert past[-1].time == T+width
	assert past[-2].time == T
	np.testing.assert_allclose(past[-1].state, before+change, rtol=1e-3)
	np.testing.assert_allclose(past[-2].state, before, rtol=1e-3)


def test_unsuccessful_integration(twin):
	"""tests/test_jitcdde.py:332-351"""
	g = scenarios.golden("roessler")
	for pars in (dict(min_step=1.0), dict(min_step=1e-3, rtol=1e-10, atol=0), dict(min_step=1e-3, rtol=0, atol=1e-10)):
		DDE = jitcdde(f, verbose=False)
		DDE.add_past_points(scenarios.past_of(g))
		DDE.set_integration_parameters(**pars)
		with pytest.raises(UnsuccessfulIntegration):
			DDE.integrate(1000)


def test_find_max_delay_and_delays():
	"""tests/test_jitcdde.py:373-385"""
	assert _find_max_delay([0, 1, 4.5, 2]) == 4.5
	assert sorted(float(d) for d in _get_delays(f)) == [0.0, 4.5]
	with pytest.raises(ValueError):
		_find_max_delay([0, sympy.Symbol("tau")])


def test_check_rejects_bad_input():
	"""tests/test_jitcdde.py:387-410"""
	for bad in ([y(1)], [y(0, t-1)+y(2)], [sympy.Symbol("z")*y(0)], []):
		with pytest.raises(ValueError):
			jitcdde(bad, verbose=False).check()


def test_quadrature():
	"""tests/test_jitcdde.py:412-425: ∫_0^1 x² dx with both rules"""
	x = sympy.Symbol("x")
	for method, places in (("midpoint", 3), ("gauss", 10)):
		value = float(quadrature(x**2, x, 0, 1, nsteps=20, method=method))
		assert abs(value-1/3) < 10**-places


def test_past_from_function(twin):
	"""tests/test_past_from_function.py: a past given as callable, as symbolic expressions, or as explicit anchors
	leads to the same trajectory (double Mackey–Glass; atol 0.01 as in the reference)"""
	tau1, tau2 = 15, 10
	f_mg = [0.25*y(0, t-tau1)/(1.0+y(0, t-tau1)**10) - 0.1*y(0) + 0.05*y(0, t-tau2)]
	expressions = [sympy.cos(t)*0.1+0.9]
	function = lambda time: [np.cos(time)*0.1+0.9]
	results = []
	for how in ("callable", "symbolic", "anchors"):
		DDE = jitcdde(f_mg, verbose=False, ring_capacity=256)
		if how == "callable":
			DDE.past_from_function(function)
		elif how == "symbolic":
			DDE.past_from_function(expressions)
		else:
			for time in np.linspace(-15, 0, 200):
				DDE.add_past_point(time, function(time), [-np.sin(time)*0.1])
		DDE.step_on_discontinuities()
		results.append(np.array([DDE.integrate(DDE.t+s) for s in (10, 50, 100)]))
	np.testing.assert_allclose(results[0], results[1], atol=0.01)
	np.testing.assert_allclose(results[0], results[2], atol=0.01)


def test_input(twin):
	"""tests/test_jitcdde.py:427-452 (as intended there: a subset of the components is replaced by the recorded
	trajectory given as external input; the remaining ones must still reach the golden vector) and
	examples/data_input.py (constant past + adjust_diff with input)."""
	from itertools import combinations
	from jitcdde_b200 import input, jitcdde_input  # noqa: A004
	g = scenarios.golden("roessler")
	DDE = jitcdde(f, verbose=False, max_delay=2*T)  # nothing is forgotten: the recording has to begin at t = 0
	DDE.set_integration_parameters(**test_parameters)
	DDE.add_past_points(scenarios.past_of(g))
	DDE.integrate(T)
	recorded = [(a.time, a.state.copy(), a.diff.copy()) for a in DDE.get_state()]
	assert recorded[0][0] <= 0 and recorded[-1][0] >= T

	rng = np.random.default_rng(seed=42)
	combos = [combo for k in range(1, 6) for combo in combinations(range(6), k)]
	for index in rng.choice(len(combos), 3, replace=False):
		combo = combos[index]
		substitutions = {y(i): input(i) for i in combo}
		f_input = [expression.subs(substitutions) for expression in f]
		DDE = jitcdde_input(f_input, recorded, verbose=False)
		DDE.set_integration_parameters(**test_parameters)
		DDE.add_past_points(scenarios.past_of(g))
		DDE.initial_discontinuities_handled = True
		value = DDE.integrate(T)
		assert value.shape == (6,)
		free = [i for i in range(6) if i not in combo]
		np.testing.assert_allclose(value[free], g["y10_reference"][free], **tolerance)

	# examples/data_input.py:19-50
	times = np.linspace(0, 70, 20)
	data = np.random.default_rng(42).normal(size=(20, 2))
	slopes = np.gradient(data, times, axis=0)
	spline = [(times[k], data[k], slopes[k]) for k in range(20)]
	tau_in = 10
	f_in = [-y(0) + y(1, t-tau_in) + input(0), -y(1) + y(0, t-tau_in) + input(1)]
	DDE = jitcdde_input(f_in, spline, verbose=False)
	DDE.constant_past(np.zeros(2))
	DDE.adjust_diff()
	result = np.vstack([DDE.integrate(time) for time in np.linspace(DDE.t, 70, 30)])
	assert result.shape == (30, 2) and np.all(np.isfinite(result))
	assert DDE.max_delay == 70


def test_get_state_round_trip(twin):
	"""get_state() → add_past_points() on a fresh integrator continues the same trajectory (_jitcdde.py:357-371, 318-331):
	the anchors carry everything the integration depends on except the controller's step size."""
	g = scenarios.golden("roessler")
	DDE = jitcdde(f, verbose=False)
	DDE.set_integration_parameters(**test_parameters)
	DDE.add_past_points(scenarios.past_of(g))
	DDE.initial_discontinuities_handled = True
	DDE.integrate(T/2)
	state = [(a.time, a.state.copy(), a.diff.copy()) for a in DDE.get_state()]
	assert state[0][0] <= DDE.t-delay <= state[1][0] and state[-1][0] == DDE.t
	dt = DDE.dt
	expected = DDE.integrate(T)

	DDE2 = jitcdde(f, verbose=False)
	DDE2.set_integration_parameters(**{**test_parameters, "first_step": float(dt)})
	DDE2.add_past_points(state)
	DDE2.initial_discontinuities_handled = True
	assert DDE2.t == state[-1][0]
	np.testing.assert_allclose(DDE2.integrate(T), expected, rtol=1e-6)
	np.testing.assert_allclose(expected, g["y10_reference"], **tolerance)


def _pipelined_against_synchronous(N=6):
	"""integrate(..., wait=False) / wait_result / synchronize deliver the same samples as synchronous calls"""
	from jitcdde_b200 import pinned_empty
	from tests import systems
	pars = systems.mg_grid(N)
	runs = []
	for pipelined in (False, True):
		DDE = scenarios.make(systems.mackey_glass(), max_delay=25.0, n_members=N, ring_capacity=256)   # pipelined calls cannot grow the ring
		DDE.constant_past([1.0])
		DDE.set_parameters(pars)
		DDE.delays = pars[:, 1:2]
		DDE.step_on_discontinuities()
		times = 30+10.0*np.arange(1, 7)
		if not pipelined:
			runs.append(np.array([DDE.integrate(T) for T in times]))
			continue
		outs = [pinned_empty((N, 1)), pinned_empty((N, 1))]
		samples = []
		for i, T in enumerate(times):
			assert DDE.integrate(T, out=outs[i & 1], wait=False) is outs[i & 1]
			if i:
				DDE.wait_result(1)
				samples.append(outs[(i-1) & 1].copy())
		DDE.synchronize()
		samples.append(outs[(len(times)-1) & 1].copy())
		runs.append(np.array(samples))
		assert DDE.t.shape == (N,) and np.all(DDE.t >= times[-1])
		# a synchronous call after queued ones drains them first
		DDE.integrate(times[-1]+10, out=outs[0], wait=False)
		after = DDE.integrate(times[-1]+20)
		assert np.all(np.isfinite(after)) and np.all(np.isfinite(outs[0]))
	assert np.array_equal(runs[0], runs[1])


def test_pipelined_sampling(twin):
	_pipelined_against_synchronous()


@pytest.mark.gpu
def test_pipelined_sampling_gpu():
	_pipelined_against_synchronous(N=4096)


def test_symbols_from_anywhere_give_the_same_program():
	"""tests/test_sympy_input.py: symbols imported from the package, from its sympy_symbols module, or defined by hand with
	SymPy (the reference's users do all three) lead to the same generated statements"""
	import sympy
	import jitcdde_b200
	import jitcdde_b200.sympy_symbols
	from jitcdde_b200 import _codegen
	manual_t = sympy.Symbol("t", real=True)
	manual_current_y, manual_past_y, manual_anchors = (sympy.Function(name, real=True) for name in ("current_y", "past_y", "anchors"))
	def manual_y(index, time=manual_t):
		return manual_current_y(index) if time == manual_t else manual_past_y(time, index, manual_anchors(time))
	bodies = set()
	for t_, y_ in ((jitcdde_b200.t, jitcdde_b200.y), (jitcdde_b200.sympy_symbols.t, jitcdde_b200.sympy_symbols.y), (manual_t, manual_y),
			(jitcdde_b200.sympy_symbols.t, manual_y)):
		f = [sympy.cos(t_)*y_(1, t_-3.5) - y_(0), -y_(0, t_-1.25)*y_(1)]
		bodies.add(_codegen.generate(f).body)
	assert len(bodies) == 1


def test_save_compiled_and_module_location():
	"""tests/test_jitcdde.py:138-147 of the reference: save_compiled() writes a module file, jitcdde(n=…, module_location=…,
	delays=…) loads it without any symbolic input. Host side only here (file format, metadata, refusals); the GPU test
	integrates with the loaded image."""
	import tempfile
	from jitcdde_b200 import _build, jitcdde
	system = systems.two_roessler()
	DDE = jitcdde(system["f"], verbose=False)
	with tempfile.TemporaryDirectory() as directory:
		path = DDE.save_compiled(directory)
		assert os.path.dirname(path) == directory and os.path.getsize(path) > 1000
		with pytest.raises(OSError):
			DDE.save_compiled(path)
		assert DDE.save_compiled(path, overwrite=True) == path
		meta, image = _build.load_image_file(path)
		assert (meta["n"], meta["n_basic"], meta["n_sites"], meta["n_params"], meta["mode"]) == (6, 6, 1, 0, "fast")
		with open(image, "rb") as f, open(DDE._image, "rb") as g:
			assert f.read() == g.read()
		loaded = jitcdde(n=6, module_location=path, delays=[system["delay"]], verbose=False)
		assert loaded.n == 6 and loaded.max_delay == system["delay"] and loaded._program.n_sites == 1
		with pytest.raises(RuntimeError):
			loaded.compile_C()
		with pytest.raises(ValueError):
			jitcdde(n=5, module_location=path, delays=[system["delay"]])
		with pytest.raises(ValueError):
			jitcdde(module_location=path)
		with open(os.path.join(directory, "garbage"), "wb") as f:
			f.write(b"not a module")
		with pytest.raises(ValueError):
			jitcdde(module_location=os.path.join(directory, "garbage"), delays=[1.0])
