"""
GPU parity tests (run with -m gpu on a B200): the CUDA engine, called through the C ABI by the
product's Python classes, against
  - golden vectors produced by the reference's own C core (tests/golden/*.npz),
  - the CPU oracle (oracle/dde_oracle.c) on seeded ensembles the oracle finishes in seconds,
  - size-independent properties at the full per-GPU ensemble size of BASELINE.json configs[1].

Bars: verification build (nvcc -fmad=false): bit for bit — identical accepted/rejected step counts
and identical states. Production build (FMA contraction): identical step counts and states within
1e-9 relative for the non-chaotic systems (north_star asks 1e-10 of the verification build, which is
exact), 1e-6 for the chaotic Rössler systems over their short horizons.
"""

import numpy as np
import pytest

from jitcdde_b200 import UnsuccessfulIntegration, _build, _capi, jitcdde
from oracle import oracle
from tests import scenarios, systems

pytestmark = pytest.mark.gpu

CHAOTIC = {"roessler", "roessler_jump", "roessler_lyap"}
#: right-hand sides with elementary functions (sin, exp): the verification build, the oracle and the reference core's
#: "chain" right-hand side all call the bit-reproducible functions of csrc/jdde_math.h, so these systems are held to the
#: same bar as the others — bit for bit. The production build calls CUDA's libm: north_star's tolerance there.
LIBM = {"kuramoto", "neutral"}


def _compare_neutral(result, g):
	"""Production build only. The neutral DDE (tests/test_neutral.py) amplifies a 1-ulp perturbation to ~1e-5 and flips
	accept/reject decisions — in the reference itself (tests/test_oracle.py::test_neutral_is_ulp_sensitive).
	With FMA contraction and CUDA's exp the bar is therefore the reference's own: rtol 1e-4."""
	# (components oscillate through zero: the tolerance is relative to the amplitude, ~1, as well)
	np.testing.assert_allclose(result["samples"], g["samples"], rtol=1e-4, atol=1e-4)
	np.testing.assert_allclose(result["samples"][-1], g["control"], rtol=1e-4, atol=1e-4)
	np.testing.assert_allclose(result["adjust_diff_min"], g["adjust_diff_min"], rtol=1e-12)
	np.testing.assert_allclose(result["adjust_diff_max"], g["adjust_diff_max"], rtol=1e-12)
	assert abs(int(result["accepted"])-int(g["accepted"])) <= 0.05*int(g["accepted"])


def _compare_lyap_statistically(result, g, rtol):
	"""Local exponents of the sub-leading separation functions react to 1e-15 perturbations (Gram–Schmidt
	of nearly dependent vectors; the oracle shows the same), and through the error norm over
	all n components they feed back into the step sizes. With FMA contraction the run is therefore compared
	as the reference's own test does (tests/test_lyaps.py:63-75): weighted averages within 3.5 standard
	errors; the state at the integration tolerance; step counts within 15 % (a 1e-15 perturbation of a tangent
	vector changes the oracle's own count by 10 %: tests/test_oracle.py::test_lyapunov_step_count_is_perturbation_sensitive)."""
	from scipy.stats import sem
	np.testing.assert_allclose(result["states"], g["states"], rtol=1e-4)
	np.testing.assert_allclose(np.sum(result["weights"]), np.sum(g["weights"]), rtol=1e-2)
	assert abs(int(result["accepted"])-int(g["accepted"])) <= 0.15*int(g["accepted"])
	for i in range(g["lyaps"].shape[1]):
		ours = np.average(result["lyaps"][:, i], weights=result["weights"])
		theirs = np.average(g["lyaps"][:, i], weights=g["weights"])
		assert abs(ours-theirs) <= 3.5*sem(g["lyaps"][:, i]) + 1e-9, (i, ours, theirs)


@pytest.mark.parametrize("name", sorted(scenarios.ALL))
def test_verification_build_is_bitwise_identical_to_reference_core(name):
	g = scenarios.golden(scenarios.FIXTURE.get(name, name))
	result = scenarios.ALL[name](g, verification=True)
	scenarios.compare(result, g, exact=True, suffix="_chain" if name == "mg_ensemble" else "")


@pytest.mark.parametrize("name", sorted(scenarios.ALL))
def test_production_build_matches_reference_core(name):
	g = scenarios.golden(scenarios.FIXTURE.get(name, name))
	result = scenarios.ALL[name](g, verification=False)
	rtol = 1e-6 if name in CHAOTIC else 1e-9
	if name == "neutral":
		return _compare_neutral(result, g)
	if name.endswith("lyap"):
		return _compare_lyap_statistically(result, g, rtol)
	scenarios.compare(result, g, exact=False, rtol=rtol, suffix="_chain" if name == "mg_ensemble" else "")


@pytest.mark.parametrize("kind", sorted(scenarios.PROJECTED))
def test_restricted_and_transversal_lyapunov(kind):
	"""jitcdde_restricted_lyap / jitcdde_transversal_lyap: the projector kernels against the same calls of
	the reference's C core — verification build bit for bit, production build to rounding (the system is not chaotic)"""
	g = scenarios.golden(kind)
	result, expected = scenarios.PROJECTED[kind](g, verification=True)
	scenarios.compare(result, expected, exact=True)
	result, expected = scenarios.PROJECTED[kind](g, verification=False)
	for key in ("accepted", "rejected"):
		# FMA contraction may flip an accept/reject decision that sits on the threshold (seen: 304 instead of 305 rejections)
		assert abs(int(result[key])-int(expected[key])) <= 2+0.01*int(expected[key]), key
	for key in ("blind_states", "states", "blind_totals"):
		np.testing.assert_allclose(result[key], expected[key], rtol=1e-6, atol=1e-10, err_msg=key)
	np.testing.assert_allclose(result["blind_lyaps"], expected["blind_lyaps"], rtol=1e-6, atol=1e-9)
	# after a flipped decision the steps, hence the overshoot of every sample (the weights), differ: compare the averages
	np.testing.assert_allclose(np.sum(result["weights"]), np.sum(expected["weights"]), rtol=1e-2)
	ours, theirs = np.average(result["lyaps"], weights=result["weights"]), np.average(expected["lyaps"], weights=expected["weights"])
	assert abs(ours-theirs) < 2e-3, (ours, theirs)


def test_reference_golden_vector():
	"""tests/test_jitcdde.py:88-97 with the reference's own tolerance, production build"""
	g = scenarios.golden("roessler")
	result = scenarios.roessler(g)
	np.testing.assert_allclose(result["samples"][-1], g["y10_reference"], rtol=1e-3, atol=1e-6)


def _mg_engine(N, mode, cap=128):
	prog = systems.program(systems.mackey_glass())
	pars = systems.mg_grid(N)
	E = _capi.Engine(1, 1, prog.n_sites, 2, N, ring_capacity=cap)
	E.load_image(_build.build_image(prog, mode=mode))
	E.set_past(np.array([-1.0, 0.0]), np.ones((2, 1)), np.zeros((2, 1)))
	E.set_parameters(pars)
	E.set_integration_parameters(25.0, **oracle.DEFAULTS)
	return prog, pars, E


@pytest.mark.parametrize("mode", [_build.VERIFY, _build.FAST])
def test_ensemble_against_oracle(mode):
	"""4096 members of the (β, τ) grid over 300 time units: every member's step counts and samples"""
	N = 4096
	prog, pars, E = _mg_engine(N, mode)
	offsets = np.arange(0, 300, 10.0)
	disc, status = E.step_on_discontinuities(pars[:, 1:2].copy())
	t0 = E.get_t()
	samples = np.array([E.integrate(t0+o)[0] for o in offsets])
	accepted, rejected, rhs = E.get_counters()
	assert not E.get_status().any()

	lib = oracle.build(prog, flags=oracle.FAST_FLAGS if False else oracle.VERIFY_FLAGS)
	ref = oracle.run_ensemble(lib, (np.array([-1.0, 0.0]), np.ones((2, 1)), np.zeros((2, 1))), pars, 25.0, 1, pars[:, 1:2].copy(), offsets)
	assert np.array_equal(accepted, ref["counters"][:, 0])
	assert np.array_equal(rejected, ref["counters"][:, 1])
	assert np.array_equal(rhs, ref["counters"][:, 2])
	if mode == _build.VERIFY:
		assert np.array_equal(samples, ref["out"])
		assert np.array_equal(E.get_t(), ref["t"])
	else:
		np.testing.assert_allclose(samples, ref["out"], rtol=1e-9)
	E.close()


@pytest.mark.parametrize("mode", [_build.VERIFY, _build.FAST])
def test_sampling_in_one_launch(mode):
	"""jdde_integrate_samples: 30 samples of 4096 members in ONE launch — the same steps and states as 30 calls of
	integrate, bit for bit in either build (same kernel code, same inputs), hence equal to the oracle like them; results
	written straight into page-locked host memory by the kernel and through a device buffer + copy."""
	from jitcdde_b200 import pinned_empty
	N = 4096
	offsets = np.arange(0, 300, 10.0)
	runs = []
	for how in ("loop", "samples", "samples_pinned"):
		prog, pars, E = _mg_engine(N, mode)
		E.step_on_discontinuities(pars[:, 1:2].copy())
		t0 = E.get_t()
		if how == "loop":
			samples = np.array([E.integrate(t0+o)[0] for o in offsets])
		else:
			times = t0[:, None]+offsets[None, :]
			out = pinned_empty((len(offsets), N, 1)) if how == "samples_pinned" else None
			samples, status = E.integrate_samples(times, out=out)
			samples = np.array(samples)
		assert not E.get_status().any()
		runs.append((samples, E.get_t(), E.get_dt(), *E.get_counters()))
		E.close()
	for other in runs[1:]:
		for a, b in zip(runs[0], other):
			assert np.array_equal(a, b)
	lib = oracle.build(prog)
	ref = oracle.run_ensemble(lib, (np.array([-1.0, 0.0]), np.ones((2, 1)), np.zeros((2, 1))), pars, 25.0, 1, pars[:, 1:2].copy(), offsets)
	assert np.array_equal(runs[1][3], ref["counters"][:, 0]) and np.array_equal(runs[1][4], ref["counters"][:, 1])
	if mode == _build.VERIFY:
		assert np.array_equal(runs[1][0], ref["out"])
	else:
		np.testing.assert_allclose(runs[1][0], ref["out"], rtol=1e-9)


def test_full_size_properties():
	"""Per-GPU share of the 10^7-member ensemble (2^20 members): determinism, sampling-grid
	independence, and a random subsample checked bit for bit against the oracle."""
	N = 1 << 20
	horizon = 100.0
	results = []
	for n_calls in (1, 10):
		prog, pars, E = _mg_engine(N, _build.VERIFY)
		E.step_on_discontinuities(pars[:, 1:2].copy())
		t0 = E.get_t()
		for k in range(1, n_calls+1):
			out, status = E.integrate(t0+horizon*k/n_calls)
		assert not status.any()
		results.append((out.copy(), E.get_t(), *E.get_counters()))
		E.close()
	# sampling grid does not change the steps (the integrator overshoots and interpolates, _jitcdde.py:830-834)
	for a, b in zip(results[0][1:], results[1][1:]):
		assert np.array_equal(a, b)
	np.testing.assert_array_equal(results[0][0], results[1][0])

	out, t_end, accepted, rejected, rhs = results[0]
	assert accepted.min() > 5 and (accepted+rejected).max() < 5000
	assert np.array_equal(rhs, 3*(accepted+rejected))       # FSAL: 3 evaluations per attempt, no pws iterations here

	rng = np.random.default_rng(1)
	pick = np.sort(rng.choice(N, 512, replace=False))
	lib = oracle.build(prog)
	ref = oracle.run_ensemble(lib, (np.array([-1.0, 0.0]), np.ones((2, 1)), np.zeros((2, 1))), pars[pick], 25.0, 1, pars[pick, 1:2].copy(), [horizon])
	assert np.array_equal(out[pick], ref["out"][0])
	assert np.array_equal(accepted[pick], ref["counters"][:, 0])
	assert np.array_equal(t_end[pick], ref["t"])


def test_ring_growth_and_resume():
	"""A ring that is too small overflows, grows and the call resumes: same result as a large ring."""
	g = scenarios.golden("mg_ensemble")
	big = scenarios.mg_ensemble(g, verification=True, ring_capacity=256)
	small = scenarios.mg_ensemble(g, verification=True, ring_capacity=8)
	for key in big:
		assert np.array_equal(big[key], small[key]), key


def test_unsuccessful_integration():
	"""tests/test_jitcdde.py:332-351"""
	g = scenarios.golden("roessler")
	for pars in (dict(min_step=1.0), dict(min_step=1e-3, rtol=1e-10, atol=0), dict(min_step=1e-3, rtol=0, atol=1e-10)):
		DDE = scenarios.make(systems.two_roessler())
		DDE.add_past_points(scenarios.past_of(g))
		DDE.set_integration_parameters(**pars)
		with pytest.raises(UnsuccessfulIntegration):
			DDE.integrate(1000)


def test_call_order_errors():
	prog = systems.program(systems.mackey_glass())
	E = _capi.Engine(1, 1, prog.n_sites, 2, 8)
	with pytest.raises(_capi.EngineError, match="no kernel image"):
		E.set_integration_parameters(25.0, **oracle.DEFAULTS)
	E.load_image(_build.build_image(prog))
	with pytest.raises(_capi.EngineError, match="no past set"):
		E.integrate(1.0)
	with pytest.raises(_capi.EngineError, match="two anchors"):
		E.set_past(np.array([0.0]), np.ones((1, 1)), np.zeros((1, 1)))
	E.set_past(np.array([-1.0, 0.0]), np.ones((2, 1)), np.zeros((2, 1)))
	with pytest.raises(_capi.EngineError, match="integration parameters not set"):
		E.integrate(1.0)
	E.set_integration_parameters(25.0, **oracle.DEFAULTS)
	with pytest.raises(_capi.EngineError, match="control parameters not set"):
		E.integrate(1.0)
	E.close()


def test_watchdog_budget():
	prog, pars, E = _mg_engine(64, _build.FAST)
	E.set_attempt_budget(5)
	out, status = E.integrate(100.0)
	assert (status == _capi.STATUS_BUDGET).all()
	E.set_attempt_budget(1 << 24)
	E.clear_status(_capi.STATUS_BUDGET)
	out, status = E.resume()
	assert not status.any()
	E.close()


def test_lyapunov_ensemble_c3():
	"""BASELINE.json configs[2] shape at test size: two Rösslers + 4 exponents over a grid of couplings;
	verification build against the oracle (bit for bit; exponents to an ulp of CUDA's log)."""
	N = 256
	ours, theirs = scenarios.product_c3(N, verification=True), scenarios.oracle_c3(N)
	for key in ("out", "weights", "accepted", "rejected"):
		assert np.array_equal(ours[key], theirs[key]), key
	np.testing.assert_allclose(ours["lyaps"], theirs["lyaps"], rtol=1e-13, atol=1e-15)


def test_lane_group_stepper_equals_one_thread_per_member(monkeypatch):
	"""Lyapunov systems are integrated by 1 + n_lyap lanes per member (csrc/jdde_group.cuh); with
	JITCDDE_B200_GROUP=0 by one thread per member. Same statements per component → bit-identical in the verification build
	(N is no multiple of the 6 members a warp holds; step_on_discontinuities and past-within-step included)."""
	N = 250
	for verification in (True, False):
		group = scenarios.product_c3(N, verification=verification)
		monkeypatch.setenv("JITCDDE_B200_GROUP", "0")
		solo = scenarios.product_c3(N, verification=verification)
		monkeypatch.delenv("JITCDDE_B200_GROUP")
		for key in group:
			if verification:
				assert np.array_equal(group[key], solo[key]), key
			elif key in ("out", "weights"):
				# FMA contraction is the compiler's choice per kernel; the system is chaotic (λ ≈ 0.08 over 50 time units)
				np.testing.assert_allclose(group[key], solo[key], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("ring_capacity", [1024, 2048, 4096])
def test_cluster_orthonormalisation_is_bitwise_identical(ring_capacity):
	"""Rings too long for one block's shared memory are orthonormalised by a thread-block cluster per member (2, 4, 8
	blocks, sums through distributed shared memory; csrc/jdde_team.cuh). The verification build hands the running sums
	from block to block in the reference's order: bit-identical with the oracle; the production build agrees to rounding."""
	N = 40
	theirs = scenarios.oracle_c3(N)
	ours = scenarios.product_c3(N, verification=True, ring_capacity=ring_capacity)
	for key in ("out", "weights", "accepted", "rejected"):
		assert np.array_equal(ours[key], theirs[key]), key
	np.testing.assert_allclose(ours["lyaps"], theirs["lyaps"], rtol=1e-13, atol=1e-15)
	fast = scenarios.product_c3(N, verification=False, ring_capacity=ring_capacity)
	np.testing.assert_allclose(fast["lyaps"][:2], theirs["lyaps"][:2], rtol=1e-6, atol=1e-8)


def _lyap_sweep(n_lyap, N):
	from jitcdde_b200 import jitcdde_lyap
	rng = np.random.default_rng(3)
	system = systems.two_roessler(control=True)
	DDE = jitcdde_lyap(system["f"], n_lyap=n_lyap, control_pars=system["control_pars"], n_members=N, verbose=False, verification=True, rng=np.random.default_rng(4))
	DDE.add_past_point(-4.5, rng.random(6), rng.random(6))
	DDE.add_past_point(0.0, rng.random(6), rng.random(6))
	DDE.set_integration_parameters(**systems.LYAP_TEST_PARAMETERS)
	DDE.set_parameters((0.5*np.arange(N)/(N-1)).reshape(N, 1))
	DDE.step_on_discontinuities()
	out = [DDE.integrate(DDE.t+5.0) for _ in range(3)]
	return [np.asarray(x) for sample in out for x in sample] + list(DDE.counters())


@pytest.mark.parametrize("n_lyap", [1, 8])
def test_lane_groups_of_other_sizes(monkeypatch, n_lyap):
	"""2 lanes per member (16 members per warp) and 9 lanes per member (3 members, 27 of 32 lanes): bit-identical with one
	thread per member in the verification build"""
	group = _lyap_sweep(n_lyap, 50)
	monkeypatch.setenv("JITCDDE_B200_GROUP", "0")
	solo = _lyap_sweep(n_lyap, 50)
	for a, b in zip(group, solo):
		assert np.array_equal(a, b)


@pytest.mark.parametrize("preparation", ["adjust_diff", "integrate_blindly", "step_on_discontinuities"])
def test_lyapunov_spectrum_of_two_roesslers(preparation):
	"""The reference's own test of jitcdde_lyap (tests/test_lyaps.py:38-76) on the GPU, production build (lane-group stepper, team
	Gram–Schmidt): after each way of handling the initial discontinuity, 100 samples of 10 time units; the weighted averages of
	the four local exponents from sample 40 on lie within 3.5 standard errors of [0.0806, 0, −0.0368, −0.1184]."""
	from scipy.stats import sem
	from jitcdde_b200 import jitcdde_lyap
	controls = scenarios.golden("roessler_lyap")["lyap_controls"]
	system = systems.two_roessler()
	rng = np.random.default_rng(seed=42)
	DDE = jitcdde_lyap(system["f"], n_lyap=len(controls), verbose=False, ring_capacity=2048)
	DDE.add_past_point(-system["delay"], rng.random(6), rng.random(6))
	DDE.add_past_point(0.0, rng.random(6), rng.random(6))
	DDE.set_integration_parameters(**systems.LYAP_TEST_PARAMETERS)
	if preparation == "adjust_diff":
		DDE.adjust_diff()
	elif preparation == "integrate_blindly":
		DDE.integrate_blindly(100.0, 0.1)
	else:
		DDE.step_on_discontinuities()
	lyaps, weights = [], []
	for T in np.arange(DDE.t, DDE.t+1000, 10):
		_, lyap, weight = DDE.integrate(T)
		lyaps.append(lyap)
		weights.append(weight)
	lyaps = np.vstack(lyaps)
	start = 40
	for i, control in enumerate(controls):
		lyap = np.average(lyaps[start:, i], weights=weights[start:])
		assert abs(lyap-control) < 3.5*sem(lyaps[start:, i]), (i, lyap, control, sem(lyaps[start:, i]))


def test_full_size_lyapunov_ensemble_c3():
	"""BASELINE.json configs[2] at its full size — two Rösslers + 4 exponents, 10^5 coupling strengths — on the lane-group stepper
	and the team Gram–Schmidt (verification build, three samples of 10 time units; ring of 512 anchors: 26 GB): all members
	finish, the exponents are finite, and a random subsample agrees with the oracle bit for bit."""
	from jitcdde_b200 import jitcdde_lyap
	N = 100_000
	past, k, _ = scenarios.c3_inputs(N)
	offsets = np.array([10.0, 20.0, 30.0])
	system = systems.two_roessler(control=True)
	DDE = jitcdde_lyap(system["f"], n_lyap=4, control_pars=system["control_pars"], n_members=N, verbose=False, verification=True, ring_capacity=512)
	DDE.add_past_points(past)
	DDE.set_integration_parameters(**systems.LYAP_TEST_PARAMETERS)
	DDE.set_parameters(k)
	DDE.step_on_discontinuities()
	t0 = DDE.t
	states, lyaps, weights = zip(*[DDE.integrate(t0+o) for o in offsets])
	accepted, rejected, _ = DDE.counters()
	assert DDE._engine.ring_capacity == 512 and np.all(np.isfinite(lyaps)) and accepted.min() > 1000

	pick = np.sort(np.random.default_rng(2).choice(N, 48, replace=False))
	prog, n_basic = systems.lyap_program(system, 4)
	lib = oracle.build(prog, n_basic=n_basic)
	arrays = (np.array([a[0] for a in past]), np.array([a[1] for a in past]), np.array([a[2] for a in past]))
	ref = oracle.run_ensemble(lib, arrays, k[pick], 4.5, 1, np.array([4.5]), offsets, lyap=True, **systems.LYAP_TEST_PARAMETERS)
	assert np.array_equal(np.array(states)[:, pick], ref["out"])
	assert np.array_equal(np.array(weights)[:, pick], ref["weights"])
	assert np.array_equal(accepted[pick], ref["counters"][:, 0]) and np.array_equal(rejected[pick], ref["counters"][:, 1])
	np.testing.assert_allclose(np.array(lyaps)[:, pick], ref["lyaps"], rtol=1e-13, atol=1e-15)


def test_full_size_state_dependent_and_neutral_ensembles_c5():
	"""BASELINE.json configs[4] at full size (2^20 members each): state-dependent delay — verification build, random subsample bit for
	bit against the oracle — and the neutral DDE on the pointer window — production build, the reference's own tolerance."""
	N = 1 << 20
	rng = np.random.default_rng(3)
	ours = scenarios.product_c5a(N, verification=True)
	pick = np.sort(rng.choice(N, 256, replace=False))
	times, states, diffs, sample_times = scenarios.c5a_inputs(N)
	lib = oracle.build(systems.program(systems.state_dependent()))
	ref = oracle.run_ensemble(lib, (times[pick], states[pick], diffs[pick]), None, 1e10, 3, np.array([0.01, 0.0]), sample_times-0.01, per_member_past=True, n_members=len(pick))
	assert np.array_equal(ours["out"][:, pick], ref["out"])
	assert np.array_equal(ours["accepted"][pick], ref["counters"][:, 0]) and np.array_equal(ours["rejected"][pick], ref["counters"][:, 1])
	assert np.all(np.isfinite(ours["out"]))
	del ours

	ours = scenarios.product_c5b(N)
	pick = np.sort(rng.choice(N, 128, replace=False))
	initial, sample_times = scenarios.c5b_inputs(N)
	system = systems.neutral()
	lib = oracle.build(systems.program(system))
	times = np.broadcast_to(np.array([-1.0, 0.0]), (len(pick), 2)).copy()
	states = np.stack([initial[pick], initial[pick]], axis=1)
	ref = oracle.run_ensemble(lib, (times, states, np.zeros_like(states)), None, system["delay"], 2, None, sample_times-0.0, per_member_past=True, n_members=len(pick), rtol=1e-5)
	np.testing.assert_allclose(ours["out"][:, pick], ref["out"], rtol=1e-4, atol=1e-4)
	assert np.all(np.isfinite(ours["out"]))
	assert np.all(np.abs(ours["accepted"][pick]-ref["counters"][:, 0]) <= 0.05*ref["counters"][:, 0]+2)


def test_production_build_is_deterministic():
	"""Lane groups, shared-memory staging by cp.async, cluster sums through distributed shared memory: no result may depend on
	the timing of warps or blocks — two runs of the production build agree bit for bit (C3 with cluster Gram–Schmidt, C5b with
	the pointer window)."""
	a, b = (scenarios.product_c3(100, ring_capacity=1024) for _ in range(2))
	for key in a:
		assert np.array_equal(a[key], b[key]), key
	a, b = (scenarios.product_c5b(512) for _ in range(2))
	for key in a:
		assert np.array_equal(a[key], b[key]), key


def test_state_dependent_ensemble_c5a():
	"""BASELINE.json configs[4], state-dependent delay with per-member pasts (divergent lookups, backward
	moves of the cursor, max_delay = 1e10 so nothing is ever forgotten and rings grow)."""
	N = 4096
	ours, theirs = scenarios.product_c5a(N, verification=True), scenarios.oracle_c5a(N)
	for key in ours:
		assert np.array_equal(ours[key], theirs[key]), key
	fast = scenarios.product_c5a(N, verification=False)
	assert np.array_equal(fast["accepted"], theirs["accepted"])
	np.testing.assert_allclose(fast["out"], theirs["out"], rtol=1e-9)


def test_neutral_ensemble_c5b():
	"""BASELINE.json configs[4], neutral DDE (past_dy, anchor helper, exp in the right-hand side). Verification build: bit for
	bit against the oracle — the elementary functions are the shared ones of csrc/jdde_math.h —, identical step counts
	included. Production build (FMA contraction, CUDA's exp): the system amplifies an ulp to 1e-5 in the reference itself,
	so the reference's own tolerance (tests/test_neutral.py: rtol 1e-4)."""
	N = 1024
	theirs = scenarios.oracle_c5b(N)
	exact = scenarios.product_c5b(N, verification=True)
	for key in exact:
		assert np.array_equal(exact[key], theirs[key]), key
	ours = scenarios.product_c5b(N)
	# components oscillate through zero: tolerance relative to the amplitude (~1) as well
	np.testing.assert_allclose(ours["out"], theirs["out"], rtol=1e-4, atol=1e-4)
	assert np.all(np.abs(ours["accepted"]-theirs["accepted"]) <= 0.05*theirs["accepted"]+2)


def test_pws_fuzzy_increase_on_gpu():
	"""pws_fuzzy_increase=True (_jitcdde.py:730-731; reference test tests/test_jitcdde.py:182-186): the draws come from the
	counter-based generator shared with the oracle — verification build bit for bit, reference's golden vector within its tolerance."""
	g = scenarios.golden("roessler")
	past = scenarios.past_of(g)
	N = 64
	k = np.linspace(0.2, 0.3, N)[:, None]
	k[0, 0] = 0.25
	times = np.linspace(0, 10, 10, endpoint=True)
	system = systems.two_roessler(control=True)
	DDE = scenarios.make(system, n_members=N, verification=True, rng=np.random.default_rng(5))
	DDE.add_past_points(past)
	DDE.set_integration_parameters(pws_fuzzy_increase=True, **systems.ROESSLER_TEST_PARAMETERS)
	DDE.set_parameters(k)
	DDE.initial_discontinuities_handled = True
	out = np.array([DDE.integrate(T) for T in times])
	accepted, rejected, _ = DDE.counters()
	lib = oracle.build(systems.program(system))
	arrays = (np.array([a[0] for a in past]), np.array([a[1] for a in past]), np.array([a[2] for a in past]))
	ref = oracle.run_ensemble(lib, arrays, k, 4.5, 0, None, times, absolute_samples=True, fuzzy_seed=DDE._int_params["pws_fuzzy_seed"], n_members=N, **systems.ROESSLER_TEST_PARAMETERS)
	assert np.array_equal(out, ref["out"])
	assert np.array_equal(accepted, ref["counters"][:, 0]) and np.array_equal(rejected, ref["counters"][:, 1])
	np.testing.assert_allclose(out[-1, 0], g["y10_reference"], rtol=1e-3, atol=1e-6)


def test_integration_with_saved_module(tmp_path):
	"""tests/test_jitcdde.py:138-147 (TestIntegrationSaveAndLoad): save_compiled(), then an integrator built from the module file
	alone — no symbolic input — reproduces the reference core's vectors bit for bit (the file carries the verification image)."""
	g = scenarios.golden("roessler")
	original = scenarios.make(systems.two_roessler(), verification=True)
	path = original.save_compiled(str(tmp_path), overwrite=True)
	DDE = jitcdde(n=6, module_location=path, delays=[systems.two_roessler()["delay"]], verbose=False)
	DDE.set_integration_parameters(**systems.ROESSLER_TEST_PARAMETERS)
	DDE.add_past_points(scenarios.past_of(g))
	DDE.initial_discontinuities_handled = True
	samples = np.array([DDE.integrate(T) for T in g["times"]])
	assert np.array_equal(samples, g["samples"])
	accepted, rejected, _ = DDE.counters()
	assert (accepted[0], rejected[0]) == (int(g["accepted"]), int(g["rejected"]))


def test_blind_integration_grows_the_ring_beforehand_on_gpu():
	"""integrate_blindly(max_delay, small step) with the default ring of 64 anchors, jump and adjust_diff on a full ring"""
	from jitcdde_b200 import t, y
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


def test_past_from_function_and_state_round_trip_on_gpu():
	"""past_from_function (anchor placement is this package's own heuristic: the reference's lives in chspy and
	has no golden vector — parity-unpinned) and the round trip get_state() → add_past_points() of
	examples/mackey_glass_parameter_jump.py:55-57: the second integrator continues bit for bit where the first one stands."""
	from jitcdde_b200 import t, y
	f = [0.25*y(0, t-15)/(1+y(0, t-15)**10) - 0.1*y(0)]
	DDE = jitcdde(f, verbose=False, verification=True)
	DDE.past_from_function(lambda time: [1.0+0.1*np.sin(0.3*time)])
	DDE.step_on_discontinuities()
	a = DDE.integrate(DDE.t+50.0)
	state = DDE.get_state()
	assert len(state) > 10 and all(np.isfinite(anchor[1]).all() for anchor in state)
	direct = DDE.integrate(DDE.t+25.0)
	second = jitcdde(f, verbose=False, verification=True)
	second.add_past_points(state)
	second.initial_discontinuities_handled = True
	second.set_integration_parameters(first_step=float(DDE.dt))
	continued = second.integrate(second.t+25.0)
	assert second.t > 0 and np.all(np.isfinite(continued))
	# the anchors round-trip exactly; the second integrator starts with its own first step, so the states agree to the tolerance
	np.testing.assert_allclose(continued, direct, rtol=1e-4)
	again = second.get_state()
	assert np.array_equal(np.array([anchor[0] for anchor in again])[:3] >= state[0][0], [True]*3)
