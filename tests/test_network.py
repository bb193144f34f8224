"""
The network engine (`jitcdde_network`: one large delay-coupled system, node-parallel, structured graph input)
against the symbolic path (`jitcdde`, one expression per node as in the reference's
examples/kuramoto_network.py:42-63) on a small Kuramoto network. Both run on the host twin here (CPU); the GPU
version of the same comparison is in test_gpu_network.py. The symbolic path itself is pinned to the reference's C
core (golden "kuramoto"). The two paths add the coupling terms in a different order (SymPy sorts sums), so they
agree to rounding, not bit for bit: states within 1e-9, identical step counts.
"""

import os

import numpy as np
import pytest
import sympy

from tests import systems


def kuramoto_graph(n, seed=3, q=0.3, c=1.5):
	rng = np.random.default_rng(seed)
	A = rng.choice([1, 0], size=(n, n), p=[q, 1-q])
	np.fill_diagonal(A, 0)
	tau = rng.uniform(np.pi/5, 2*np.pi, size=(n, n))
	omega = rng.uniform(0.8, 1.2, n)
	initial = rng.uniform(0, 2*np.pi, n)
	src, dst = np.nonzero(A)           # A[j, i]: edge j → i
	return dict(n=n, A=A, tau=tau, omega=omega, initial=initial, src=src, dst=dst, edge_tau=tau[dst, src], weight=c/(n-1))


def run_network(g, sample_times, blind_to, **kw):
	from jitcdde_b200 import jitcdde_network
	net = jitcdde_network(g["n"], local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=g["weight"], node_parameters=[g["omega"]], verbose=False, **kw)
	net.set_integration_parameters(rtol=0, atol=1e-6, max_step=float(np.min(g["edge_tau"])))      # (as run_symbolic)
	net.constant_past(g["initial"], time=0.0)
	blind = net.integrate_blindly(blind_to, 0.1)
	samples = np.array([net.integrate(T) for T in sample_times])
	return blind, samples, net.counters(), net.t


def symbolic_f(g):
	from jitcdde_b200 import t, y
	n, A, tau = g["n"], g["A"], g["tau"]
	return [g["omega"][i] + g["weight"]*sum(sympy.sin(y(j, t-tau[i, j])-y(i)) for j in range(n) if A[j, i]) for i in range(n)]


def prebuild_images():
	"""called by __graft_entry__.build(): compile, on the CPU box, the kernel images the GPU tests of the network
	engine need (the symbolic path has one lookup site per edge; its image takes minutes to compile)"""
	from jitcdde_b200 import _build, _codegen
	_build.build_image(_codegen.generate(symbolic_f(kuramoto_graph(12))), mode=_build.VERIFY)
	for mode in (_build.FAST, _build.VERIFY):
		_build.build_net_image(_codegen.generate_network(lambda y, t, w: w, lambda yd, y: sympy.sin(yd-y), 1), mode=mode)


def run_symbolic(g, sample_times, blind_to, **kw):
	from jitcdde_b200 import jitcdde
	f = symbolic_f(g)
	DDE = jitcdde(f, verbose=False, max_delay=float(np.max(g["edge_tau"])), ring_capacity=256, **kw)
	DDE.set_integration_parameters(rtol=0, atol=1e-6, max_step=float(np.min(g["edge_tau"])))
	DDE.constant_past(g["initial"], time=0.0)
	blind = DDE.integrate_blindly(blind_to, 0.1)
	samples = np.array([DDE.integrate(T) for T in sample_times])
	accepted, rejected, _ = DDE.counters()
	return blind, samples, (int(accepted[0]), int(rejected[0])), DDE.t


def test_network_engine_matches_symbolic_path(twin):
	g = kuramoto_graph(12)
	blind_to = float(np.max(g["edge_tau"]))
	times = blind_to + np.arange(0.5, 5.5, 0.5)
	nb, ns, nc, nt = run_network(g, times, blind_to)
	sb, ss, sc, st = run_symbolic(g, times, blind_to)
	np.testing.assert_allclose(nb, sb, rtol=1e-10)
	np.testing.assert_allclose(ns, ss, rtol=1e-9)
	assert (nc[0], nc[1]) == sc
	assert abs(nt-st) < 1e-9


def run_network_hub_golden(**kw):
	"""tests/golden/network_hub.npz: a hub with 1100 in-edges, nodes without in-edges (reference core)"""
	from jitcdde_b200 import jitcdde_network
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "network_hub.npz"))
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False, verification=True, **kw)
	net.set_integration_parameters(rtol=0, atol=1e-5, max_step=float(np.min(g["edge_tau"])))
	net.constant_past(g["initial"], time=0.0)
	blind = net.integrate_blindly(float(g["blind_to"]), 0.1)
	samples = np.array([net.integrate(T) for T in g["times"]])
	return g, blind, samples, net.counters(), net.t


def run_network_golden(**kw):
	"""the 128-node network of tests/golden/network.npz (outputs of the REFERENCE'S C core, tests/golden/make_golden.py::golden_network)
	on the network engine, verification build"""
	from jitcdde_b200 import jitcdde_network
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "network.npz"))
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False, verification=True, **kw)
	net.set_integration_parameters(rtol=0, atol=1e-5, max_step=float(np.min(g["edge_tau"])))      # (as make_golden.py::golden_network)
	net.constant_past(g["initial"], time=0.0)
	blind = net.integrate_blindly(float(g["blind_to"]), 0.1)
	samples = np.array([net.integrate(T) for T in g["times"]])
	return g, blind, samples, net.counters(), net.t


def test_network_engine_against_reference_core(twin):
	"""The reference's C core has integrated this network itself (one expression per node, one anchors() call per edge, as
	examples/kuramoto_network.py:50-62 builds it): identical accepted and rejected step counts, identical bits of every state."""
	g, blind, samples, counters, t_end = run_network_golden()
	assert np.array_equal(blind, g["blind"])
	assert np.array_equal(samples, g["samples"])
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert t_end == float(g["final_t"])


def run_network_pws_golden(**kw):
	"""the 96-node network of tests/golden/network_pws.npz — every tenth delay shorter than the steps — on the network engine,
	verification build, default integration parameters but the tolerance (as the reference core was driven: make_golden.py::golden_network_pws)"""
	from jitcdde_b200 import jitcdde_network
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "network_pws.npz"))
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False, verification=True, **kw)
	net.set_integration_parameters(rtol=0, atol=1e-4)
	net.constant_past(g["initial"], time=0.0)
	samples, dts = [], []
	for T in g["times"]:
		samples.append(net.integrate(T))
		dts.append(net.dt)
	return g, np.array(samples), np.array(dts), net.counters(), net.t


def test_network_past_within_step_against_reference_core(twin):
	"""Delays shorter than the step (_jitcdde.py:838-861): lookups reach into the step under way and interpolate towards the
	result of the attempt before; the step is repeated until two successive results agree or shrunk, the step size then
	grows by the credit rule. Same accepted and rejected counts, same step sizes, same bits as the reference's C core under
	the reference's controller."""
	g, samples, dts, counters, t_end = run_network_pws_golden()
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert counters[2] > counters[0]+counters[1]      # repetitions: more calls of get_next_step than steps
	assert np.array_equal(dts, g["dts"])
	assert np.array_equal(samples, g["samples"])
	assert t_end == float(g["final_t"])


def run_network_adjust_diff_golden(**kw):
	"""tests/golden/network_adjust_diff.npz (reference core: adjust_diff after a constant past, then integrate) on the network engine"""
	from jitcdde_b200 import jitcdde_network
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "network_adjust_diff.npz"))
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False, verification=True, **kw)
	net.set_integration_parameters(rtol=0, atol=1e-5)
	net.constant_past(g["initial"], time=0.0)
	minima, maxima = net.adjust_diff()
	t_after = net.t
	samples = np.array([net.integrate(T) for T in g["times"]])
	return g, minima, maxima, t_after, samples, net.counters(), net.t


def test_network_adjust_diff_against_reference_core(twin):
	"""adjust_diff (_jitcdde.py:863-878, apply_jump + truncate_past of the C core) for a network: same extrema over the jump
	interval, same anchors (the integration that follows takes the same steps to the same bits)"""
	g, minima, maxima, t_after, samples, counters, t_end = run_network_adjust_diff_golden()
	assert np.array_equal(minima, g["minima"]) and np.array_equal(maxima, g["maxima"])
	assert t_after == float(g["t_after"])
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert np.array_equal(samples, g["samples"])
	assert t_end == float(g["final_t"])


def run_network_saturated_golden(**kw):
	"""tests/golden/network_saturated.npz: a ring whose step size saturates at max_step = the delay, then blind steps of that length"""
	from jitcdde_b200 import jitcdde_network
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "network_saturated.npz"))
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False, verification=True, **kw)
	net.set_integration_parameters(rtol=0, atol=1e-3, max_step=0.1, first_step=0.1)
	net.constant_past(g["initial"], time=0.0)
	samples, dts = [], []
	for T in g["times"]:
		samples.append(net.integrate(T))
		dts.append(net.dt)
	blind = net.integrate_blindly(net.t+5.0, 0.1)
	return g, np.array(samples), np.array(dts), blind, net.counters(), net.t


def test_step_size_saturated_at_the_smallest_delay(twin):
	"""ADVICE.md (round 1): once dt has saturated at max_step = smallest delay, (t+h)−tau rounds one ulp above t in a few percent of
	the attempts. That is a past-within-step event of one ulp — in the reference, too, which shrinks the step by pws_factor and
	rejects the attempt (12 of 474 here) — not a reason to stop; blind steps of exactly the delay extrapolate the last segment by
	that ulp. Same counts, step sizes and bits as the reference's C core."""
	g, samples, dts, blind, counters, t_end = run_network_saturated_golden()
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert np.array_equal(dts, g["dts"]) and dts.max() == 0.1 and dts.min() < 0.1
	assert np.array_equal(samples, g["samples"])
	assert np.array_equal(blind, g["blind"])
	assert t_end == float(g["final_t"])


def test_network_with_a_hub_against_reference_core(twin):
	"""in-degrees from 0 to 1100: the sums of a node run over its in-edges in input order, however many"""
	g, blind, samples, counters, t_end = run_network_hub_golden()
	assert np.array_equal(blind, g["blind"])
	assert np.array_equal(samples, g["samples"])
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert t_end == float(g["final_t"])


def run_network_jump_golden(**kw):
	"""tests/golden/network_jump.npz (reference core): a forward jump with one amplitude per node, a backward jump with a common one"""
	from jitcdde_b200 import jitcdde_network
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "network_jump.npz"))
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False, verification=True, **kw)
	net.set_integration_parameters(rtol=0, atol=1e-5, max_step=float(np.min(g["edge_tau"])))
	net.constant_past(g["initial"], time=0.0)
	b = float(g["blind_to"])
	net.integrate_blindly(b, 0.1)
	out = {"first": net.integrate(b+2.0)}
	out["minima_1"], out["maxima_1"] = net.jump(g["amplitude"], net.t, 1e-5, forward=True)
	out["t_1"] = net.t
	out["second"] = np.array([net.integrate(T) for T in b+np.array([2.5, 3.0, 4.0])])
	out["minima_2"], out["maxima_2"] = net.jump(0.3, net.t, 1e-4, forward=False)
	out["t_2"] = net.t
	out["third"] = np.array([net.integrate(T) for T in b+np.array([4.5, 5.0, 6.0])])
	return g, out, net.counters(), net.t


def check_network_jump(g, out, counters, t_end):
	for key in ("first", "minima_1", "maxima_1", "second", "minima_2", "maxima_2", "third"):
		assert np.array_equal(out[key], g[key]), key
	assert out["t_1"] == float(g["t_1"]) and out["t_2"] == float(g["t_2"])
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert t_end == float(g["final_t"])


def test_network_jump_against_reference_core(twin):
	"""jump (_jitcdde.py:935-974, apply_jump of the C core) for a network, forward and backward: same extrema, same anchors, the
	integration goes on with the same steps to the same bits"""
	check_network_jump(*run_network_jump_golden())


def run_network_discontinuities_golden(**kw):
	"""tests/golden/network_discontinuities.npz (reference core): step_on_discontinuities with two propagations, then integrate"""
	from jitcdde_b200 import jitcdde_network
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "network_discontinuities.npz"))
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False, verification=True, **kw)
	net.set_integration_parameters(rtol=0, atol=1e-6)
	net.constant_past(g["initial"], time=0.0)
	after = net.step_on_discontinuities(propagations=2)
	t_after = net.t
	samples = np.array([net.integrate(T) for T in g["times"]])
	return g, after, t_after, samples, net.counters(), net.t


def check_network_discontinuities(g, after, t_after, samples, counters, t_end):
	assert np.array_equal(after, g["after"]) and t_after == float(g["t_after"])
	assert np.array_equal(samples, g["samples"])
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert t_end == float(g["final_t"])


def test_network_step_on_discontinuities_against_reference_core(twin):
	"""step_on_discontinuities (_jitcdde.py:935-999) for a network with two distinct delays: adaptive steps capped at the times at
	which the initial discontinuity arrives (two propagations: 0.7, 1.3, 1.4, 2.0, 2.6 after each other, as the reference adds them up)"""
	check_network_discontinuities(*run_network_discontinuities_golden())


def checkpoint_round_trip(**kw):
	"""integrate, take the state (get_state), start a NEW network from it with the step size the first one had reached, go on with
	both: the same trajectory, bit for bit (reference: get_state / add_past_points, _jitcdde.py:298-309, 357-371;
	examples/mackey_glass_parameter_jump.py:55-57 switches integrators this way)"""
	from jitcdde_b200 import jitcdde_network
	g = kuramoto_graph(12)
	make = lambda: jitcdde_network(g["n"], local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=g["weight"], node_parameters=[g["omega"]], verbose=False, **kw)
	blind_to = float(np.max(g["edge_tau"]))
	net = make()
	net.set_integration_parameters(rtol=0, atol=1e-6)
	net.constant_past(g["initial"], time=0.0)
	net.integrate_blindly(blind_to, 0.1)
	net.integrate(blind_to+2.0)
	anchors, dt = net.get_state(), net.dt
	assert anchors[-1][0] == net.t and anchors[-1][0]-anchors[0][0] >= blind_to      # the whole delay window is there
	assert all(a[0] < b[0] for a, b in zip(anchors, anchors[1:]))
	resumed = make()
	resumed.set_integration_parameters(rtol=0, atol=1e-6, first_step=dt)
	resumed.add_past_points(anchors)
	targets = blind_to+np.array([2.5, 3.0, 4.0])
	assert np.array_equal(np.array([resumed.integrate(T) for T in targets]), np.array([net.integrate(T) for T in targets]))
	assert resumed.t == net.t


def test_network_checkpoint_and_resume(twin):
	checkpoint_round_trip()


@pytest.mark.parametrize("seed", [1, 2])
def test_network_matches_symbolic_path_with_short_delays_and_jumps(twin, seed):
	"""Two independent implementations of the same reference semantics — the network engine and the symbolic ensemble engine (one
	expression per node, pinned to the reference core by its own goldens) — on random small networks in which a third of the
	delays are shorter than the steps: adjust_diff, integration with past-within-step repetitions, a jump, more integration.
	The two add a node's coupling terms in different orders, so they agree to rounding: same extrema and states to 1e-8."""
	from jitcdde_b200 import jitcdde, jitcdde_network, t, y
	rng = np.random.default_rng(seed)
	n = 7
	A = rng.choice([1, 0], size=(n, n), p=[0.4, 0.6])
	np.fill_diagonal(A, 0)
	tau = np.where(rng.random((n, n)) < 0.33, rng.uniform(0.004, 0.03, (n, n)), rng.uniform(0.4, 1.5, (n, n)))
	omega, initial, weight = rng.uniform(0.8, 1.2, n), rng.uniform(0, 2*np.pi, n), 0.3
	src, dst = np.nonzero(A)      # A[j, i]: edge j → i
	parameters = dict(rtol=0, atol=1e-5)
	amplitude = rng.uniform(-0.3, 0.3, n)

	def drive(system):
		system.set_integration_parameters(**parameters)
		system.constant_past(initial, time=0.0)
		out = [np.concatenate(system.adjust_diff())]
		out.append(system.integrate(1.0))
		out.append(np.concatenate(system.jump(amplitude, system.t)))
		out.extend(system.integrate(T) for T in (1.5, 2.5))
		return [np.asarray(x, dtype=float).ravel() for x in out]

	net = jitcdde_network(n, local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y), edges=(src, dst, tau[dst, src]),
		weight=weight, node_parameters=[omega], verbose=False)
	f = [omega[i] + weight*sum(sympy.sin(y(j, t-tau[i, j])-y(i)) for j in range(n) if A[j, i]) for i in range(n)]
	symbolic = jitcdde(f, verbose=False, max_delay=float(tau[dst, src].max()), ring_capacity=256)
	for a, b in zip(drive(net), drive(symbolic)):
		np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-12)
	assert net.counters()[2] > sum(net.counters()[:2])      # steps were repeated (past within step)


def test_node_partition_exchange_protocol(twin):
	"""Two engine instances, each owning half of the nodes (what two ranks hold), driven through the per-attempt
	protocol attempt → exchange staged slices + max of the error norm → control: same trajectory as one instance."""
	from jitcdde_b200 import _build, _capi, _codegen
	g = kuramoto_graph(10)
	n = g["n"]
	blind_to = float(np.max(g["edge_tau"]))
	times = blind_to + np.arange(0.5, 3.5, 0.5)
	_, reference, counters, _ = run_network(g, times, blind_to)

	program = _codegen.generate_network(lambda y, t, w: w, lambda yd, y: sympy.sin(yd-y), 1)
	_build.build_net_image(program)
	order = np.argsort(g["dst"], kind="stable")
	src, dst, tau = g["src"][order], g["dst"][order], g["edge_tau"][order]
	from oracle import oracle
	pars = dict(oracle.DEFAULTS, rtol=0, atol=1e-6, max_step=float(tau.min()), first_step=min(1.0, float(tau.min())))
	engines = []
	for lo, hi in ((0, n//2), (n//2, n)):
		mask = (dst >= lo) & (dst < hi)
		row_ptr = np.concatenate([[0], np.cumsum(np.bincount(dst[mask]-lo, minlength=hi-lo))])
		E = _capi.NetEngine(n, lo, hi, int(mask.sum()), 1, ring_capacity=128)
		E.set_graph(row_ptr, src[mask], tau[mask], np.full(int(mask.sum()), g["weight"]))
		E.set_node_parameters(g["omega"][None, :])
		E.set_past([-1.0, 0.0], np.stack([g["initial"]]*2), np.zeros((2, n)))
		E.set_integration_parameters(float(tau.max()), **pars)
		engines.append((E, lo, hi))

	import ctypes
	def exchange():
		views = []
		for E, lo, hi in engines:
			s, p = E.exchange_pointers()
			views.append((np.ctypeslib.as_array(ctypes.cast(s, ctypes.POINTER(ctypes.c_double)), (2*n,)),
				np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), (1,)), lo, hi))
		for s, p, lo, hi in views:           # all-gather of the staged (state, diff) pairs
			for s2, p2, lo2, hi2 in views:
				s2[2*lo:2*hi] = s[2*lo:2*hi]
		pmax = max(v[1][0] for v in views)      # all-reduce(max)
		for v in views:
			v[1][0] = pmax

	def run(begin):
		for E, _, _ in engines:
			begin(E)
		while not engines[0][0].poll()[0]:
			for E, _, _ in engines:
				E.attempt()
			exchange()
			for E, _, _ in engines:
				E.control()

	number = round(blind_to/0.1)
	run(lambda E: E.begin(1, blind_to, blind_to/number, number))
	samples = []
	for T in times:
		run(lambda E: E.begin(0, T))
		samples.append(engines[0][0].recent_state(T))
		assert np.array_equal(samples[-1], engines[1][0].recent_state(T))
	assert np.array_equal(np.array(samples), reference)
	assert engines[0][0].counters() == counters


def test_ring_growth_gives_the_same_trajectory(twin):
	"""The reference keeps every anchor in an unbounded list; a ring that is too small for the delay window fills up, is doubled
	and the call goes on where it stopped (blind and adaptive): same states and step counts as with a ring that is large enough."""
	g = kuramoto_graph(12)
	blind_to = float(np.max(g["edge_tau"]))
	times = blind_to + np.arange(0.5, 3.5, 0.5)
	small = run_network(g, times, blind_to, ring_capacity=8)
	large = run_network(g, times, blind_to, ring_capacity=512)
	assert np.array_equal(small[0], large[0]) and np.array_equal(small[1], large[1])
	assert small[2] == large[2] and small[3] == large[3]
