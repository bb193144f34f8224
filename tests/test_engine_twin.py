"""
CPU tests of the product's Python classes and engine logic (no GPU): the engine header compiled
for the host (tests/host_twin) behind the real C ABI and the real `jitcdde` / `jitcdde_lyap`
classes must reproduce, bit for bit, what the reference's own C core produced
(tests/golden/*.npz). The CUDA build of the same code is checked in test_gpu_parity.py.
"""

import numpy as np
import pytest

from tests import scenarios


@pytest.mark.parametrize("name", sorted(scenarios.ALL))
def test_bitwise_against_reference_core(twin, name):
	g = scenarios.golden(scenarios.FIXTURE.get(name, name))
	result = scenarios.ALL[name](g)
	scenarios.compare(result, g, exact=True, suffix="_chain" if name == "mg_ensemble" else "")


@pytest.mark.parametrize("window", ["0", "1"])
@pytest.mark.parametrize("name", ["mg_ensemble", "state_dependent", "kuramoto", "roessler", "roessler_jump", "neutral", "mg_lyap"])
def test_register_window_does_not_change_results(twin, monkeypatch, name, window):
	"""The register-resident anchor window (JD_WINDOW=1) is an access-path optimisation: forced on or off,
	also for systems where the build policy would not choose it, results stay bit-identical. (The
	shared-memory window, JD_WINDOW=2, is device-only and is covered by the GPU tests.)"""
	monkeypatch.setenv("JITCDDE_B200_WINDOW", window)
	g = scenarios.golden(scenarios.FIXTURE.get(name, name))
	result = scenarios.ALL[name](g)
	scenarios.compare(result, g, exact=True, suffix="_chain" if name == "mg_ensemble" else "")


@pytest.mark.parametrize("capacity", ["4", "7", "40", "4096"])
@pytest.mark.parametrize("name", ["mg_lyap", "roessler_lyap"])
def test_team_orthonormalisation_is_bitwise_identical(twin, monkeypatch, name, capacity):
	"""The block-per-member Gram–Schmidt of csrc/jdde_team.cuh (staged in shared memory when the live anchors fit the
	capacity, streamed through it in tiles otherwise; here with a one-thread team on the host) against the reference core's vectors."""
	monkeypatch.setenv("JITCDDE_TWIN_TEAM", capacity)
	g = scenarios.golden(name)
	scenarios.compare(scenarios.ALL[name](g), g, exact=True)


def test_reference_golden_vector(twin):
	"""tests/test_jitcdde.py:88-97: state at T=10 within the reference's own tolerance"""
	g = scenarios.golden("roessler")
	result = scenarios.roessler(g)
	np.testing.assert_allclose(result["samples"][-1], g["y10_reference"], rtol=1e-3, atol=1e-6)


def test_ensembles_of_other_configurations(twin):
	"""C3 (Lyapunov ensemble over the coupling), C5a (state-dependent delay, per-member pasts), C5b (neutral):
	product API on the host twin against the oracle's ensemble driver, bit for bit (same libm on the host)."""
	for product, reference, N in ((scenarios.product_c3, scenarios.oracle_c3, 6), (scenarios.product_c5a, scenarios.oracle_c5a, 9), (scenarios.product_c5b, scenarios.oracle_c5b, 5)):
		ours, theirs = product(N), reference(N)
		for key in ours:
			if key == "lyaps":
				np.testing.assert_allclose(ours[key], theirs[key], rtol=4e-16, atol=1e-17)
			else:
				assert np.array_equal(ours[key], theirs[key]), (product.__name__, key, np.max(np.abs(ours[key]-theirs[key])))


def test_lyapunov_resume_after_partial_ring_overflow(twin):
	"""Mackey–Glass with per-member delays and 2 exponents: with a ring of 32 anchors the members with long delays
	fill it during a call, the ring grows and the call is resumed. Members that had already reached the target
	(and were orthonormalised) must keep their exponents: same results as with a ring that never overflows."""
	import sympy
	from jitcdde_b200 import jitcdde_lyap, t, y
	beta, tau = sympy.symbols("beta tau")
	f = [beta*y(0, t-tau)/(1+y(0, t-tau)**10) - 0.1*y(0)]
	pars = np.stack([np.full(12, 0.25), np.linspace(4.0, 26.0, 12)], axis=1)
	results = []
	for capacity in (32, 512):
		DDE = jitcdde_lyap(f, n_lyap=2, control_pars=[beta, tau], max_delay=26.0, delays=[0, tau], n_members=12, verbose=False,
			verification=True, ring_capacity=capacity)
		DDE.constant_past([1.0])
		DDE.set_parameters(pars)
		DDE.initial_discontinuities_handled = True
		out = [DDE.integrate(T) for T in (30.0, 60.0, 90.0)]
		results.append((out, DDE._engine.ring_capacity))
	assert results[0][1] > 32, "the small ring was meant to overflow"
	for small, big in zip(results[0][0], results[1][0]):
		for a, b in zip(small, big):
			assert np.array_equal(a, b)


@pytest.mark.parametrize("kind", sorted(scenarios.PROJECTED))
def test_restricted_and_transversal_lyapunov_bitwise_against_reference_core(twin, kind):
	"""jitcdde_restricted_lyap (remove_state/diff_component, remove_projections) and jitcdde_transversal_lyap
	(normalise_indices) against the same calls of the reference's C core (tests/golden/restricted.npz, transversal.npz)"""
	result, expected = scenarios.PROJECTED[kind](scenarios.golden(kind), verification=True)
	scenarios.compare(result, expected, exact=True)


def test_restricted_and_transversal_lyapunov_agree(twin):
	"""The reference's own integration test (tests/test_transversal_lyap.py), shortened: both classes give the same largest
	transversal exponent (within their standard errors) with the sign the coupling dictates, and the restricted run stays on
	the synchronisation manifold."""
	from scipy.stats import sem
	from tests import systems
	system = systems.synchronised_fhn()
	results = {}
	for kind in ("restricted", "transversal"):
		DDE, _, _ = systems.projected_program(kind, verification=True)
		rng = np.random.default_rng(42)
		for k, sign in ((-0.1, 1), (0.1, -1), (0.0, 0)):
			DDE.purge_past()
			single = rng.random(2)
			initial = np.array([single[j] for j, group in enumerate(system["groups"]) for _ in group])
			DDE.constant_past(initial if kind == "restricted" else single, 0.0)
			DDE.set_parameters(k)
			for time in range(1, 100):
				DDE.integrate_blindly(time)
			lyaps, weights = [], []
			for time in DDE.t + np.arange(100, 3100, 100):
				state, lyap, weight = DDE.integrate(time)
				lyaps.append(lyap)
				weights.append(weight)
			if kind == "restricted":
				for anchor in DDE.get_state():
					for group in system["groups"]:
						assert len({anchor[1][i] for i in group}) == 1 and len({anchor[2][i] for i in group}) == 1
			results[kind, k] = (np.average(lyaps, weights=weights), sem(lyaps), sign)
	for k in (-0.1, 0.1, 0.0):
		(l1, m1, sign), (l2, m2, _) = results["restricted", k], results["transversal", k]
		assert (np.sign(l1) if abs(l1) > m1 else 0) == sign
		assert (np.sign(l2) if abs(l2) > m2 else 0) == sign
		assert abs(l1-l2) < max(m1, m2), (k, l1, m1, l2, m2)


def test_installation_check(twin):
	"""jitcdde.test() of the reference (_jitcdde.py:1661-1677): finishes silently"""
	import jitcdde_b200
	assert jitcdde_b200.test() is None


def test_orthonormalisation_invariants(twin):
	"""tests/test_past.py:10-27 on the engine's Gram–Schmidt: after orthonormalising, the separation functions have norm 1 and
	are mutually orthogonal over the delay window — a second pass finds nothing to remove and returns norms of 1."""
	from jitcdde_b200 import jitcdde_lyap, t, y
	rng = np.random.default_rng(seed=42)
	m = 4
	DDE = jitcdde_lyap([-y(0, t-2)], n_lyap=m-1, verbose=False, verification=True)
	for time in sorted(rng.uniform(-3, 0, 5)):
		DDE.add_past_point(time, rng.normal(0, 1, m), rng.normal(0, 1, m))
	delay = rng.uniform(0.0, 2.0)
	DDE._initiate()
	first = DDE._engine.orthonormalise(m-1, delay)
	assert np.all(first > 0) and not np.allclose(first, 1.0)
	np.testing.assert_allclose(DDE._engine.orthonormalise(m-1, delay), 1.0, rtol=1e-12)


def test_remove_projections_invariants(twin):
	"""tests/test_past.py:29-86: removing the projections twice changes nothing the second time and leaves a normalised
	separation function; unit vectors are removed by zeroing the component."""
	from jitcdde_b200 import jitcdde_restricted_lyap, t, y
	rng = np.random.default_rng(seed=42)
	n_basic = 3
	f = [-y(1, t-1), -y(2, t-1), -y(0, t-1)]
	random_vector = lambda: rng.random(n_basic)
	DDE = jitcdde_restricted_lyap(f, vectors=[(random_vector(), random_vector()), (random_vector(), random_vector())], verbose=False, verification=True)
	assert DDE.n == 6*n_basic
	time = rng.uniform(-10, 10)
	for _ in range(rng.integers(3, 10)):
		DDE.add_past_point(time, rng.random(DDE.n), rng.random(DDE.n))
		time += 0.1 + rng.random()
	DDE.remove_projections()
	before = DDE.get_state()
	assert abs(DDE.remove_projections()-1.0) < 1e-12
	for a, b in zip(before, DDE.get_state()):
		np.testing.assert_allclose(a[1], b[1], rtol=1e-12, atol=1e-14)
		np.testing.assert_allclose(a[2], b[2], rtol=1e-12, atol=1e-14)
		assert not a[1][2*n_basic:].any() and not a[2][2*n_basic:].any()      # the dummies are cleared

	unit = np.zeros(n_basic)
	unit[1] = 1
	DDE = jitcdde_restricted_lyap(f, vectors=[(unit, np.zeros(n_basic)), (np.zeros(n_basic), unit)], verbose=False, verification=True)
	assert DDE.n == 2*n_basic and DDE.state_components == [1] and DDE.diff_components == [1]
	DDE.add_past_point(-1.0, rng.random(DDE.n), rng.random(DDE.n))
	DDE.add_past_point(0.0, rng.random(DDE.n), rng.random(DDE.n))
	DDE.remove_projections()
	for anchor in DDE.get_state():
		assert anchor[1][n_basic+1] == 0.0 and anchor[2][n_basic+1] == 0.0


@pytest.mark.parametrize("capacity", [16, 256])
def test_sampling_in_one_launch_equals_separate_calls(twin, capacity):
	"""integrate_samples(times) is the loop `[integrate(T) for T in times]` inside one launch: the same steps and the same
	states, bit for bit — also when the ring (16 anchors here) fills up on the way, grows and the call is resumed, and with
	per-member sample times."""
	import sympy
	from jitcdde_b200 import jitcdde, t, y
	beta, tau = sympy.symbols("beta tau")
	f = [beta*y(0, t-tau)/(1+y(0, t-tau)**10) - 0.1*y(0)]
	N = 7
	pars = np.stack([np.linspace(0.2, 0.3, N), np.linspace(8.0, 20.0, N)], axis=1)
	times = 25.0 + 7.5*np.arange(1, 9)
	per_member = times[None, :] + 0.1*np.arange(N)[:, None]

	def fresh():
		DDE = jitcdde(f, control_pars=[beta, tau], max_delay=20.0, n_members=N, verbose=False, verification=True, ring_capacity=capacity)
		DDE.constant_past([1.0])
		DDE.set_parameters(pars)
		DDE.delays = pars[:, 1:2]
		DDE.max_delay = 20.0
		DDE.step_on_discontinuities()
		return DDE

	for sample_times, separate in ((times, lambda k: times[k]), (per_member, lambda k: per_member[:, k])):
		one = fresh()
		series = one.integrate_samples(sample_times)
		many = fresh()
		loop = np.stack([many.integrate(separate(k)) for k in range(len(times))])
		assert series.shape == (len(times), N, 1)
		assert np.array_equal(series, loop)
		for a, b in zip(one.counters(), many.counters()):
			assert np.array_equal(a, b)
		assert np.array_equal(one.t, many.t) and np.array_equal(one.dt, many.dt)
		if capacity == 16:
			assert one._engine.ring_capacity > 16, "the small ring was meant to overflow"

	single = jitcdde([0.25*y(0, t-15)/(1+y(0, t-15)**10) - 0.1*y(0)], verbose=False, verification=True)
	single.constant_past([1.0])
	single.step_on_discontinuities()
	series = single.integrate_samples(single.t + np.arange(1, 6)*10.0)
	assert series.shape == (5, 1)


def test_pws_fuzzy_increase(twin):
	"""pws_fuzzy_increase=True (/root/reference/jitcdde/_jitcdde.py:730-731; reference test: tests/test_jitcdde.py:182-186): after
	a past-within-step event (first_step 30 > delay 4.5) the step size is increased with the reference's probability, drawn per
	member from a counter-based generator. The engine and the oracle draw the same numbers: bit-identical; the draws matter
	(other seeds, other steps); two runs agree; and the result meets the reference's golden vector within its tolerance."""
	from oracle import oracle
	from tests import systems
	g = scenarios.golden("roessler")
	past = scenarios.past_of(g)
	N = 8
	k = np.linspace(0.2, 0.3, N)[:, None]
	k[0, 0] = 0.25      # the coupling of tests/test_jitcdde.py:30-41
	times = np.linspace(0, 10, 10, endpoint=True)
	system = systems.two_roessler(control=True)

	def run(seed):
		DDE = scenarios.make(system, n_members=N, verification=True, rng=np.random.default_rng(seed))
		DDE.add_past_points(past)
		DDE.set_integration_parameters(pws_fuzzy_increase=True, **systems.ROESSLER_TEST_PARAMETERS)
		DDE.set_parameters(k)
		DDE.initial_discontinuities_handled = True
		out = np.array([DDE.integrate(T) for T in times])
		return out, DDE.counters(), DDE._int_params["pws_fuzzy_seed"]

	out, counters, seed = run(1)
	again, counters_again, _ = run(1)
	assert seed != 0 and np.array_equal(out, again) and all(np.array_equal(a, b) for a, b in zip(counters, counters_again))
	other, counters_other, other_seed = run(2)
	assert other_seed != seed and not np.array_equal(counters[0], counters_other[0])
	deterministic = scenarios.make(system, n_members=N, verification=True)
	deterministic.add_past_points(past)
	deterministic.set_integration_parameters(**systems.ROESSLER_TEST_PARAMETERS)
	deterministic.set_parameters(k)
	deterministic.initial_discontinuities_handled = True
	deterministic.integrate(10.0)
	assert not np.array_equal(counters[0], deterministic.counters()[0])

	lib = oracle.build(systems.program(system))
	arrays = (np.array([a[0] for a in past]), np.array([a[1] for a in past]), np.array([a[2] for a in past]))
	ref = oracle.run_ensemble(lib, arrays, k, 4.5, 0, None, times, absolute_samples=True, fuzzy_seed=seed, n_members=N, **systems.ROESSLER_TEST_PARAMETERS)
	assert np.array_equal(out, ref["out"])
	assert np.array_equal(counters[0], ref["counters"][:, 0]) and np.array_equal(counters[1], ref["counters"][:, 1])
	# member 0 is the reference's own test case: tests/test_jitcdde.py:61-64, 88-97
	np.testing.assert_allclose(out[-1, 0], g["y10_reference"], rtol=1e-3, atol=1e-6)


def test_blind_integration_grows_the_ring_beforehand(twin):
	"""The reference-style call integrate_blindly(max_delay, small step) needs more anchors than the default ring of 64 holds; a
	blind call cannot be resumed after an overflow, so the ring is sized before it. Same result as with a ring that
	is large enough from the start; jump and adjust_diff on a full ring likewise."""
	from jitcdde_b200 import jitcdde, t, y
	f = [0.25*y(0, t-15)/(1+y(0, t-15)**10) - 0.1*y(0)]
	results = []
	for capacity in (None, 1024):
		DDE = jitcdde(f, verbose=False, verification=True, ring_capacity=capacity)
		DDE.constant_past([1.0])
		blind = DDE.integrate_blindly(30.0, 0.1)
		mi, ma = DDE.jump(0.1, DDE.t)
		after = DDE.integrate_blindly(DDE.t+10.0, 0.05)
		mi2, ma2 = DDE.adjust_diff()
		results.append((blind, mi, ma, after, mi2, ma2, DDE.t, len(DDE.get_state())))
	assert all(np.array_equal(a, b) for a, b in zip(*results))
