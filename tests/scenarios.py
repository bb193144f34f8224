"""
Scenarios shared by the CPU tests (engine host twin) and the GPU tests (CUDA engine): each one
drives the product's public Python API exactly as a user of the reference would, on the inputs
stored in a golden fixture (tests/golden/*.npz, outputs of the reference's own C core), and
returns what the fixture recorded.
"""

import os

import numpy as np

from jitcdde_b200 import jitcdde, jitcdde_lyap
from tests import systems

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
	return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def past_of(g):
	return [(g["past_times"][k], g["past_states"][k], g["past_diffs"][k]) for k in range(len(g["past_times"]))]


def make(system, cls=jitcdde, **kwargs):
	options = dict(verbose=False)
	for key in ("helpers", "control_pars", "max_delay", "delays"):
		if key in system and system[key] is not None and len(np.shape(system[key])) >= 0:
			if key == "control_pars" and not system[key]:
				continue
			options[key] = system[key]
	options.update(kwargs)
	return cls(system["f"], **options)


def roessler(g, **kw):
	DDE = make(systems.two_roessler(), **kw)
	DDE.set_integration_parameters(**systems.ROESSLER_TEST_PARAMETERS)
	DDE.add_past_points(past_of(g))
	DDE.initial_discontinuities_handled = True
	samples = np.array([DDE.integrate(T) for T in g["times"]])
	accepted, rejected, _ = DDE.counters()
	return dict(samples=samples, accepted=accepted[0], rejected=rejected[0], final_t=DDE.t)


def roessler_jump(g, **kw):
	DDE = make(systems.two_roessler(), **kw)
	DDE.set_integration_parameters(**systems.ROESSLER_TEST_PARAMETERS)
	DDE.add_past_points(past_of(g))
	ad_min, ad_max = DDE.adjust_diff()
	mid = DDE.integrate(5.0)
	j_min, j_max = DDE.jump(g["change"], 5.0)
	after = DDE.integrate(7.0)
	n_anchors = len(DDE.get_state())
	accepted, rejected, _ = DDE.counters()
	return dict(adjust_diff_min=ad_min, adjust_diff_max=ad_max, mid=mid, jump_min=j_min, jump_max=j_max, after_jump=after,
		n_anchors_after=n_anchors, jump_accepted=accepted[0], jump_rejected=rejected[0])


def mg_ensemble(g, **kw):
	pars = g["parameters"]
	N = len(pars)
	DDE = make(systems.mackey_glass(), max_delay=25.0, n_members=N, **kw)
	DDE.constant_past([1.0])
	DDE.set_parameters(pars)
	DDE.delays = pars[:, 1:2]            # one numeric delay per member (τ is a control parameter)
	DDE.max_delay = 25.0
	disc = DDE.step_on_discontinuities()
	t0 = DDE.t
	samples = np.array([DDE.integrate(t0+off) for off in g["offsets"]])
	accepted, rejected, _ = DDE.counters()
	n_anchors = np.array([len(DDE.get_state(m)) for m in range(N)])
	return dict(samples=samples, counts=np.stack([accepted, rejected], axis=1), final_t=DDE.t, disc=disc, n_anchors=n_anchors)


def neutral(g, **kw):
	system = systems.neutral()
	DDE = make(system, **kw)
	DDE.set_integration_parameters(rtol=1e-5)
	DDE.constant_past(system["initial"])
	mi, ma = DDE.adjust_diff()
	samples = np.array([DDE.integrate(T) for T in g["times"]])
	accepted, rejected, _ = DDE.counters()
	return dict(adjust_diff_min=mi, adjust_diff_max=ma, samples=samples, accepted=accepted[0], rejected=rejected[0])


def state_dependent(g, **kw):
	system = systems.state_dependent()
	DDE = make(system, **kw)
	for anchor in system["past"]:
		DDE.add_past_point(*anchor)
	blind = DDE.integrate_blindly(0.01)
	samples = np.array([DDE.integrate(T) for T in g["times"]])
	accepted, rejected, _ = DDE.counters()
	return dict(blind=blind, samples=samples, accepted=accepted[0], rejected=rejected[0], n_anchors=len(DDE.get_state()))


def ode(g, **kw):
	system = systems.fitzhugh_nagumo_ode()
	DDE = make(system, **kw)
	DDE.add_past_point(-1.0, system["y0"], system["f_of_y0"])
	DDE.add_past_point(0.0, system["y0"], system["f_of_y0"])
	adaptive = DDE.integrate(1.0)
	accepted, rejected, _ = DDE.counters()
	DDE2 = make(system, **kw)
	DDE2.add_past_point(-1.0, system["y0"], system["f_of_y0"])
	DDE2.add_past_point(0.0, system["y0"], system["f_of_y0"])
	blind = DDE2.integrate_blindly(0.98, 0.01)
	after_blind = DDE2.integrate(1.0)
	return dict(adaptive=adaptive, blind=blind, after_blind=after_blind, accepted=accepted[0], rejected=rejected[0])


def kuramoto(g, **kw):
	system = systems.kuramoto()
	DDE = make(system, **kw)
	DDE.set_integration_parameters(rtol=0, atol=1e-5)
	DDE.constant_past(system["initial"], time=0.0)
	blind = DDE.integrate_blindly(system["max_delay"], 0.1)
	t0 = DDE.t
	samples = np.array([DDE.integrate(t0+o) for o in g["offsets"]])
	accepted, rejected, _ = DDE.counters()
	return dict(blind=blind, samples=samples, accepted=accepted[0], rejected=rejected[0])


def _lyap_run(DDE, g):
	t0 = DDE.t
	states, lyaps, weights = [], [], []
	for o in g["offsets"]:
		s, l, w = DDE.integrate(t0+o)
		states.append(s); lyaps.append(l); weights.append(w)
	accepted, rejected, _ = DDE.counters()
	return dict(states=np.array(states), lyaps=np.array(lyaps), weights=np.array(weights), accepted=accepted[0], rejected=rejected[0])


def mg_lyap(g, **kw):
	DDE = make(systems.mackey_glass(control=False), cls=jitcdde_lyap, n_lyap=int(g["n_lyap"]), **kw)
	DDE.add_past_points(past_of(g))      # full anchors incl. the tangent vectors the fixture was made with
	DDE.step_on_discontinuities()
	return _lyap_run(DDE, g)


def roessler_lyap(g, **kw):
	DDE = make(systems.two_roessler(), cls=jitcdde_lyap, n_lyap=4, **kw)
	DDE.add_past_points(past_of(g))
	DDE.set_integration_parameters(**systems.LYAP_TEST_PARAMETERS)
	DDE.step_on_discontinuities()
	return _lyap_run(DDE, g)


def data_input(g, **kw):
	"""jitcdde_input (examples/data_input.py): the input spline lives in dummy components of the anchors"""
	from jitcdde_b200 import jitcdde_input
	system = systems.data_input()
	DDE = jitcdde_input(system["f"], system["input"], verbose=False, **kw)
	DDE.constant_past(system["initial"])
	mi, ma = DDE.adjust_diff()
	samples = np.array([DDE.integrate(T) for T in g["times"]])
	accepted, rejected, _ = DDE.counters()
	return dict(adjust_diff_min=mi, adjust_diff_max=ma, samples=samples, accepted=accepted[0], rejected=rejected[0], max_delay=DDE.max_delay)


def _projected(kind, g, **kw):
	"""jitcdde_restricted_lyap / jitcdde_transversal_lyap on three synchronised oscillators (tests/test_transversal_lyap.py):
	20 blind calls of one step, then adaptive samples; the past (incl. the random separation function) comes from the fixture"""
	DDE, _, _ = systems.projected_program(kind, **kw)
	DDE.add_past_points(past_of(g))
	DDE.set_parameters(float(g["k"]))
	select = (lambda s: s[:DDE.n_basic]) if kind == "restricted" else (lambda s: s[DDE.main_indices])
	blind = [DDE.integrate_blindly(float(T)) for T in range(1, 21)]
	samples = [DDE.integrate(T) for T in g["times"]]
	accepted, rejected, _ = DDE.counters()
	result = dict(blind_states=np.array([b[0] for b in blind]), blind_lyaps=np.array([b[1] for b in blind]), blind_totals=np.array([b[2] for b in blind]),
		states=np.array([s[0] for s in samples]), lyaps=np.array([s[1] for s in samples]), weights=np.array([s[2] for s in samples]),
		accepted=accepted[0], rejected=rejected[0])
	# the fixture holds the whole state of the last anchor; the classes return the basic / main components
	expected = dict(g)
	expected["blind_states"] = np.array([select(s) for s in g["blind_states"]])
	expected["states"] = np.array([select(s) for s in g["states"]])
	return result, expected


def restricted(g, **kw):
	return _projected("restricted", g, **kw)


def transversal(g, **kw):
	return _projected("transversal", g, **kw)


PROJECTED = dict(restricted=restricted, transversal=transversal)

ALL = dict(roessler=roessler, roessler_jump=roessler_jump, mg_ensemble=mg_ensemble, neutral=neutral,
	state_dependent=state_dependent, ode=ode, kuramoto=kuramoto, mg_lyap=mg_lyap, roessler_lyap=roessler_lyap, data_input=data_input)
FIXTURE = dict(roessler_jump="roessler")


def compare(result, g, exact, rtol=1e-10, suffix=""):
	"""exact: bit for bit; otherwise counts identical and values within rtol (relative)."""
	for key, value in result.items():
		expected = g[key+suffix] if key+suffix in g else g[key]
		value = np.asarray(value)
		if key in ("lyaps", "blind_lyaps") and exact:
			# log(norms)/Δt: the reference takes the logarithm with NumPy (_jitcdde.py:1168), whose vectorised log
			# differs from libm's by an ulp now and then; the norms themselves are compared through the states
			np.testing.assert_allclose(value, expected, rtol=4e-16, atol=1e-17, err_msg=key)
		elif exact or value.dtype.kind in "iu":
			assert np.array_equal(value, np.asarray(expected).reshape(value.shape)), f"{key}: not identical (max abs diff {np.max(np.abs(value-np.asarray(expected).reshape(value.shape)))})"
		else:
			np.testing.assert_allclose(value, np.asarray(expected).reshape(value.shape), rtol=rtol, atol=1e-300, err_msg=key)


# ----------------------------------------------------------------------------------------------------
# Ensemble versions of the other BASELINE.json configurations (C3, C5a, C5b), run through the
# product API (`product_*`) and through the oracle's ensemble driver (`oracle_*`) on the same inputs.

def c3_inputs(N, seed=5):
	"""two Rösslers + 4 exponents, coupling k_m = 0.5·m/(N−1) (tests/test_lyaps.py:16-45)"""
	from jitcdde_b200._past import Past
	rng = np.random.default_rng(seed)
	P = Past(n=30, n_basic=6, rng=np.random.default_rng(seed+1))
	P.add((-4.5, rng.random(6), rng.random(6)))
	P.add((0.0, rng.random(6), rng.random(6)))
	k = (0.5*np.arange(N)/max(N-1, 1)).reshape(N, 1)
	return [(a.time, a.state, a.diff) for a in P], k, np.arange(10, 60, 10.0)


def product_c3(N, **kw):
	past, k, offsets = c3_inputs(N)
	system = systems.two_roessler(control=True)
	DDE = jitcdde_lyap(system["f"], n_lyap=4, control_pars=system["control_pars"], n_members=N, verbose=False, **kw)
	DDE.add_past_points(past)
	DDE.set_integration_parameters(**systems.LYAP_TEST_PARAMETERS)
	DDE.set_parameters(k)
	DDE.step_on_discontinuities()
	t0 = DDE.t
	states, lyaps, weights = [], [], []
	for o in offsets:
		s, l, w = DDE.integrate(t0+o)
		states.append(np.atleast_2d(s)); lyaps.append(np.atleast_2d(l)); weights.append(np.atleast_1d(w))
	accepted, rejected, _ = DDE.counters()
	return dict(out=np.array(states), lyaps=np.array(lyaps), weights=np.array(weights), accepted=accepted, rejected=rejected)


def oracle_c3(N):
	from oracle import oracle
	past, k, offsets = c3_inputs(N)
	system = systems.two_roessler(control=True)
	prog, n_basic = systems.lyap_program(system, 4)
	lib = oracle.build(prog, n_basic=n_basic)
	arrays = (np.array([a[0] for a in past]), np.array([a[1] for a in past]), np.array([a[2] for a in past]))
	r = oracle.run_ensemble(lib, arrays, k, 4.5, 1, np.array([4.5]), offsets, lyap=True, **systems.LYAP_TEST_PARAMETERS)
	return dict(out=r["out"], lyaps=r["lyaps"], weights=r["weights"], accepted=r["counters"][:, 0], rejected=r["counters"][:, 1])


def c5a_inputs(N):
	"""examples/state_dependent.py:14-29 with initial states shifted by 1e-3·m/N"""
	shift = 1e-3*np.arange(N)/N
	times = np.broadcast_to(np.array([-2.0, -1.1, -0.9, 0.0]), (N, 4)).copy()
	states = (np.array([4.5, 4.5, -0.5, -0.5])[None, :]+shift[:, None]).reshape(N, 4, 1)
	return times, states, np.zeros((N, 4, 1)), np.linspace(0.9, 2, 12)


def product_c5a(N, **kw):
	times, states, diffs, sample_times = c5a_inputs(N)
	system = systems.state_dependent()
	DDE = jitcdde(system["f"], max_delay=1e10, n_members=N, verbose=False, **kw)
	for j in range(4):
		DDE.add_past_point(times[:, j] if N > 1 else times[0, j], states[:, j] if N > 1 else states[0, j], diffs[:, j] if N > 1 else diffs[0, j])
	DDE.integrate_blindly(0.01)
	out = np.array([np.atleast_2d(DDE.integrate(T)) for T in sample_times])
	accepted, rejected, _ = DDE.counters()
	return dict(out=out, accepted=accepted, rejected=rejected)


def oracle_c5a(N):
	from oracle import oracle
	times, states, diffs, sample_times = c5a_inputs(N)
	lib = oracle.build(systems.program(systems.state_dependent()))
	# integrate_blindly(0.01) from t = 0 with the default step (max_step = 10 → one step of 0.01), then absolute sample times
	r = oracle.run_ensemble(lib, (times, states, diffs), None, 1e10, 3, np.array([0.01, 0.0]), sample_times-0.01, per_member_past=True, n_members=N)
	return dict(out=r["out"], accepted=r["counters"][:, 0], rejected=r["counters"][:, 1])


def c5b_inputs(N):
	system = systems.neutral()
	shift = 1e-3*np.arange(N)/N
	return np.array(system["initial"])[None, :]+shift[:, None], np.linspace(1, 5, 5)


def product_c5b(N, **kw):
	initial, sample_times = c5b_inputs(N)
	system = systems.neutral()
	DDE = jitcdde(system["f"], helpers=system["helpers"], n_members=N, verbose=False, **kw)
	DDE.set_integration_parameters(rtol=1e-5)
	DDE.constant_past(initial if N > 1 else initial[0])
	DDE.adjust_diff()
	out = np.array([np.atleast_2d(DDE.integrate(T)) for T in sample_times])
	accepted, rejected, _ = DDE.counters()
	return dict(out=out, accepted=accepted, rejected=rejected)


def oracle_c5b(N):
	from oracle import oracle
	initial, sample_times = c5b_inputs(N)
	system = systems.neutral()
	lib = oracle.build(systems.program(system))
	times = np.broadcast_to(np.array([-1.0, 0.0]), (N, 2)).copy()
	states = np.stack([initial, initial], axis=1)
	r = oracle.run_ensemble(lib, (times, states, np.zeros_like(states)), None, system["delay"], 2, None, sample_times-0.0, per_member_past=True, n_members=N, rtol=1e-5)
	return dict(out=r["out"], accepted=r["counters"][:, 0], rejected=r["counters"][:, 1], t=r["t"])
