"""
Pins the CPU oracle (oracle/dde_oracle.c, the plain-C restatement used as checker on the GPU
box) against
  (1) the reference's own golden vectors and known-answer tests, and
  (2) outputs of the reference's unmodified C core (tests/golden/*.npz, made by
      tests/golden/make_golden.py from /root/reference/jitcdde/jitced_template.c) — bit for bit.
"""

import ctypes

import numpy as np
import pytest

from oracle import oracle
from tests import scenarios, systems


def _traj(system, g, n_basic=None, prog=None, max_delay=None, **pars):
	prog = prog or systems.program(system)
	lib = oracle.build(prog, n_basic=n_basic)
	return oracle.Trajectory(lib, scenarios.past_of(g), max_delay=max_delay if max_delay is not None else system.get("delay", system.get("max_delay", 0.0)), **pars)


def test_known_answer_single_step():
	"""tests/test_python_core.py:13-48: one Bogacki–Shampine step of Mackey–Glass"""
	prog = systems.program(systems.mackey_glass_kat())
	lib = oracle.build(prog)
	y0, dy0 = 0.8, -0.0794952762375263
	T = oracle.Trajectory(lib, [(-1.0, [y0-dy0], [dy0]), (0.0, [y0], [dy0])], max_delay=15)
	T.get_next_step(1.0)
	assert T.t == 0.0
	assert abs(T.error(0) - (-1.34096023725590e-5)) < 5e-8        # assertAlmostEqual: 7 places
	T.accept_step()
	assert abs(T.get_current_state()[0] - 0.724447497727209) < 5e-8
	assert T.t == 1.0


def test_two_roessler_golden_vector():
	"""tests/test_jitcdde.py:33-97: y(10) within rtol 1e-3 / atol 1e-6; and bit-identical with the reference core"""
	g = scenarios.golden("roessler")
	T = _traj(systems.two_roessler(), g, **systems.ROESSLER_TEST_PARAMETERS)
	samples = np.array([T.integrate(t) for t in g["times"]])
	np.testing.assert_allclose(samples[-1], g["y10_reference"], rtol=1e-3, atol=1e-6)
	assert np.array_equal(samples, g["samples"])
	assert (T.counters["accepted"], T.counters["rejected"]) == (int(g["accepted"]), int(g["rejected"]))
	assert T.t == float(g["final_t"])


def test_sampling_grid_does_not_change_steps():
	"""tests/test_jitcdde.py:88-130: one big step, 10 samples and 10000 tiny samples give the same trajectory"""
	g = scenarios.golden("roessler")
	results = []
	for times in ([10.0], np.linspace(0, 10, 10), np.linspace(0, 10, 10000)):
		T = _traj(systems.two_roessler(), g, **systems.ROESSLER_TEST_PARAMETERS)
		for t in times:
			value = T.integrate(t)
		results.append((value, T.counters["accepted"], T.counters["rejected"]))
	for value, acc, rej in results[1:]:
		assert np.array_equal(value, results[0][0]) and (acc, rej) == results[0][1:]


def test_jumps_and_adjust_diff():
	g = scenarios.golden("roessler")
	T = _traj(systems.two_roessler(), g, **systems.ROESSLER_TEST_PARAMETERS)
	mi, ma = T.adjust_diff()
	assert np.array_equal(mi, g["adjust_diff_min"]) and np.array_equal(ma, g["adjust_diff_max"])
	assert np.array_equal(T.integrate(5.0), g["mid"])
	mi, ma = T.jump(g["change"], 5.0)
	assert np.array_equal(mi, g["jump_min"]) and np.array_equal(ma, g["jump_max"])
	assert np.array_equal(T.integrate(7.0), g["after_jump"])
	T.forget(4.5)
	assert len(T.get_state()) == int(g["n_anchors_after"])


def test_annihilating_jumps():
	"""tests/test_jitcdde.py:110-118"""
	g = scenarios.golden("roessler")
	T = _traj(systems.two_roessler(), g, **systems.ROESSLER_TEST_PARAMETERS)
	T.integrate(5.0)
	change = np.random.default_rng(42).normal(0, 10, 6)
	T.jump(change, 5.0)
	T.jump(-change, 5.0)
	np.testing.assert_allclose(T.integrate(10.0), g["y10_reference"], rtol=1e-3, atol=1e-6)


def test_mackey_glass_ensemble_and_flavours():
	g = scenarios.golden("mg_ensemble")
	prog = systems.program(systems.mackey_glass())
	lib = oracle.build(prog)
	pars = g["parameters"]
	past = (g["past_times"], g["past_states"], g["past_diffs"])
	res = oracle.run_ensemble(lib, past, pars, 25.0, 1, pars[:, 1:2].copy(), g["offsets"])
	assert np.array_equal(res["out"], g["samples_chain"])
	assert np.array_equal(res["counters"][:, :2], g["counts_chain"])
	assert np.array_equal(res["t"], g["final_t_chain"])
	# what the reference would literally generate (nested anchors() calls, libm pow, libm roots)
	# takes the same steps and differs only by rounding
	assert np.array_equal(g["counts_literal"], g["counts_chain"])
	np.testing.assert_allclose(g["samples_literal"], g["samples_chain"], rtol=1e-10)
	# libm roots in the oracle's controller: same counts
	res0 = oracle.run_ensemble(lib, past, pars, 25.0, 1, pars[:, 1:2].copy(), g["offsets"], root_mode=0)
	assert np.array_equal(res0["counters"][:, :2], g["counts_chain"])
	np.testing.assert_allclose(res0["out"], g["samples_chain"], rtol=1e-10)


def test_shared_roots_match_libm():
	lib = oracle.build(systems.program(systems.mackey_glass()))
	lib.jo_inv_root.restype = ctypes.c_double
	lib.jo_inv_root.argtypes = [ctypes.c_double, ctypes.c_int]
	for p in 10.0**np.linspace(-12, 12, 4001):
		assert abs(lib.jo_inv_root(p, 3)/p**(-1/3.)-1) < 1e-15
		assert abs(lib.jo_inv_root(p, 4)/p**(-1/4.)-1) < 5e-16


def test_ode_limit():
	"""tests/test_ode.py:19-69"""
	system = systems.fitzhugh_nagumo_ode()
	g = scenarios.golden("ode")
	T = _traj(system, g, max_delay=0.0)
	value = T.integrate(1.0)
	np.testing.assert_allclose(value, system["y1"], rtol=1e-5)
	assert np.array_equal(value, g["adaptive"])
	T = _traj(system, g, max_delay=0.0)
	assert np.array_equal(T.integrate_blindly(0.98, 0.01), g["blind"])
	value = T.integrate(1.0)
	np.testing.assert_allclose(value, system["y1"], rtol=1e-5)
	assert np.array_equal(value, g["after_blind"])


def test_neutral():
	"""tests/test_neutral.py:16-71"""
	system = systems.neutral()
	g = scenarios.golden("neutral")
	T = _traj(system, g, rtol=1e-5)
	mi, ma = T.adjust_diff()
	assert np.array_equal(mi, g["adjust_diff_min"]) and np.array_equal(ma, g["adjust_diff_max"])
	samples = np.array([T.integrate(t) for t in g["times"]])
	assert np.array_equal(samples, g["samples"])
	np.testing.assert_allclose(samples[-1], system["control"], rtol=1e-4)


def test_neutral_is_ulp_sensitive():
	"""Why GPU results for this system are compared at the reference's own tolerance (1e-4): changing one
	initial value by one ulp changes the reference algorithm's accept/reject sequence and its result by ~1e-5."""
	system = systems.neutral()
	g = scenarios.golden("neutral")
	lib = oracle.build(systems.program(system))
	outcomes = []
	for factor in (1.0, 1.0+2.3e-16):
		past = [(a[0], a[1].copy(), a[2].copy()) for a in scenarios.past_of(g)]
		for anchor in past:
			anchor[1][0] *= factor
		T = oracle.Trajectory(lib, past, max_delay=system["delay"], rtol=1e-5)
		T.adjust_diff()
		outcomes.append((np.array([T.integrate(t) for t in g["times"]]), T.counters))
	(a, ca), (b, cb) = outcomes
	deviation = np.max(np.abs(a-b)/np.abs(a))
	assert 1e-9 < deviation < 1e-4
	np.testing.assert_allclose(b[-1], system["control"], rtol=1e-4)


def test_state_dependent_delay():
	system = systems.state_dependent()
	g = scenarios.golden("state_dependent")
	T = _traj(system, g)
	assert np.array_equal(T.integrate_blindly(0.01), g["blind"])
	samples = np.array([T.integrate(t) for t in g["times"]])
	assert np.array_equal(samples, g["samples"])
	assert (T.counters["accepted"], T.counters["rejected"]) == (int(g["accepted"]), int(g["rejected"]))


def test_many_lookup_sites():
	system = systems.kuramoto()
	g = scenarios.golden("kuramoto")
	T = _traj(system, g, rtol=0, atol=1e-5)
	assert np.array_equal(T.integrate_blindly(system["max_delay"], 0.1), g["blind"])
	t0 = T.t
	samples = np.array([T.integrate(t0+o) for o in g["offsets"]])
	assert np.array_equal(samples, g["samples"])


def test_external_input():
	"""jitcdde_input (examples/data_input.py): extended system + joined past, against the reference core"""
	prog, past, max_delay = systems.data_input_program()
	g = scenarios.golden("data_input")
	assert np.array_equal(np.array([a[1] for a in past]), g["past_states"]) and np.array_equal(np.array([a[0] for a in past]), g["past_times"])
	T = oracle.Trajectory(oracle.build(prog), past, max_delay=max_delay)
	mi, ma = T.adjust_diff()
	assert np.array_equal(mi[:2], g["adjust_diff_min"]) and np.array_equal(ma[:2], g["adjust_diff_max"])
	samples = np.array([T.integrate(t)[:2] for t in g["times"]])
	assert np.array_equal(samples, g["samples"])
	assert (T.counters["accepted"], T.counters["rejected"]) == (int(g["accepted"]), int(g["rejected"]))


@pytest.mark.parametrize("name", ["mg_lyap", "roessler_lyap"])
def test_lyapunov_bitwise(name):
	g = scenarios.golden(name)
	if name == "mg_lyap":
		system, n_lyap, pars, steps = systems.mackey_glass(control=False), int(g["n_lyap"]), {}, [15.0]
		system["delay"] = 15.0
	else:
		system, n_lyap, pars, steps = systems.two_roessler(), 4, systems.LYAP_TEST_PARAMETERS, [4.5]
	prog, n_basic = systems.lyap_program(system, n_lyap)
	T = _traj(system, g, n_basic=n_basic, prog=prog, **pars)
	T.step_on_discontinuities(steps)
	t0 = T.t
	for k, o in enumerate(g["offsets"]):
		state, lyaps, dt = T.integrate_lyap(t0+o)
		assert np.array_equal(state, g["states"][k]) and dt == g["weights"][k]
		# the reference takes log(norms) with NumPy, whose vectorised log differs from libm's by an ulp now and then
		np.testing.assert_allclose(lyaps, g["lyaps"][k], rtol=4e-16, atol=1e-17)


def test_lyapunov_step_count_is_perturbation_sensitive():
	"""Why Lyapunov runs of the production build (FMA contraction) are compared statistically: perturbing one
	tangent component of the initial past by 1e-15 changes the reference algorithm's own accepted-step count
	by several percent (the sub-leading separation functions are nearly dependent at first, Gram–Schmidt
	amplifies rounding, and the error norm over all components feeds this back into the step size)."""
	g = scenarios.golden("mg_lyap")
	system = systems.mackey_glass(control=False)
	system["delay"] = 15.0
	prog, n_basic = systems.lyap_program(system, 3)
	counts = []
	for eps in (0.0, 1e-15):
		past = [(a[0], a[1].copy(), a[2].copy()) for a in scenarios.past_of(g)]
		past[1][1][2] *= 1+eps
		lib = oracle.build(prog, n_basic=n_basic)
		T = oracle.Trajectory(lib, past, max_delay=15.0)
		T.step_on_discontinuities([15.0])
		t0 = T.t
		for o in g["offsets"]:
			T.integrate_lyap(t0+o)
		counts.append(T.counters["accepted"])
	assert counts[0] == int(g["accepted"])
	assert 0.01 < abs(counts[1]-counts[0])/counts[0] < 0.15


def test_orthonormality_invariants():
	"""tests/test_past.py:9-26: after orthonormalise the separation functions have norm 1 and vanishing scalar products"""
	g = scenarios.golden("roessler_lyap")
	system = systems.two_roessler()
	prog, n_basic = systems.lyap_program(system, 4)
	T = _traj(system, g, n_basic=n_basic, prog=prog, **systems.LYAP_TEST_PARAMETERS)
	T.step_on_discontinuities([4.5])
	T.integrate(T.t+20)
	T.orthonormalise(4, 4.5)
	norms = T.orthonormalise(4, 4.5)      # second pass: already orthonormal
	np.testing.assert_allclose(norms, 1.0, atol=1e-7)


def test_lyapunov_spectrum_statistical():
	"""tests/test_lyaps.py:38-75 with the reference's acceptance rule (3.5 standard errors); a shorter run than the
	reference's 1000 time units keeps the CPU suite fast, so the error bars are wider, not the rule looser."""
	from scipy.stats import sem
	g = scenarios.golden("roessler_lyap")
	system = systems.two_roessler()
	prog, n_basic = systems.lyap_program(system, 4)
	T = _traj(system, g, n_basic=n_basic, prog=prog, **systems.LYAP_TEST_PARAMETERS)
	T.step_on_discontinuities([4.5])
	lyaps, weights = [], []
	for target in np.arange(T.t, T.t+1000, 10):
		_, lyap, weight = T.integrate_lyap(target)
		lyaps.append(lyap); weights.append(weight)
	lyaps = np.vstack(lyaps)
	start = 40
	for i, control in enumerate(g["lyap_controls"]):
		lyap = np.average(lyaps[start:, i], weights=weights[start:])
		assert abs(lyap-control) < 3.5*sem(lyaps[start:, i])


def test_unsuccessful_integration():
	"""tests/test_jitcdde.py:332-351"""
	g = scenarios.golden("roessler")
	for pars in (dict(min_step=1.0), dict(min_step=1e-3, rtol=1e-10, atol=0), dict(min_step=1e-3, rtol=0, atol=1e-10)):
		T = _traj(systems.two_roessler(), g, **pars)
		with pytest.raises(RuntimeError):
			T.integrate(1000)
		assert T.status == 1
