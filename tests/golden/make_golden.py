"""
Generates the golden vectors in this directory from the REFERENCE'S OWN C core.

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py

For every system below, /root/reference/jitcdde/jitced_template.c is rendered and compiled
unmodified (oracle/build_ref.py) and driven by the restated Python controller
(oracle/controller.py), i.e. the outputs recorded here are outputs of the reference
implementation, not of this repository's engine or oracle. The fixtures also carry the inputs
(past, parameters, sample times), so that the tests need nothing from /root/reference.

Flavour: "chain" right-hand side + libm-free controller roots (root_mode=1) — the configuration
the CUDA verification build reproduces bit for bit. For the Mackey–Glass ensemble the
"literal" flavour (what the reference itself would print: nested anchors() calls, libm pow,
libm roots) is recorded as well, to pin how little the flavours differ.

Also copied here (data, not code): the reference's own golden vector for the two coupled
Rösslers, tests/two_Roessler_past.dat + tests/two_Roessler_y10.dat.
"""

import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from jitcdde_b200._past import Past  # noqa: E402
from oracle import build_ref, controller, oracle  # noqa: E402
from tests import systems  # noqa: E402

REF_TESTS = "/root/reference/tests"


def roots_from(lib):
	lib.jo_inv_root.restype = ctypes.c_double
	lib.jo_inv_root.argtypes = [ctypes.c_double, ctypes.c_int]
	lib.jo_tanh.restype = ctypes.c_double
	lib.jo_tanh.argtypes = [ctypes.c_double]
	roots = lambda p, order: lib.jo_inv_root(p, int(order))
	roots.tanh = lambda x: lib.jo_tanh(float(x))      # the controller's sigmoids (oracle/controller.py) with the same bits as on the GPU
	return roots


def make_core(prog, name, past, n_basic=None, flavour="chain", parameters=None):
	build_ref.build_ref(prog, name, n_basic=n_basic, flavour=flavour, force=True)
	module = build_ref.load_ref(name)
	core = module.dde_integrator([(float(a[0]), np.array(a[1], dtype=float), np.array(a[2], dtype=float)) for a in past])
	if parameters is not None:
		core.set_parameters(*[float(x) for x in parameters])
	return core


def past_arrays(past):
	return dict(
		past_times=np.array([a[0] for a in past], dtype=float),
		past_states=np.array([a[1] for a in past], dtype=float),
		past_diffs=np.array([a[2] for a in past], dtype=float),
	)


def save(name, **arrays):
	path = os.path.join(HERE, name + ".npz")
	np.savez_compressed(path, **arrays)
	print("wrote", path, {k: np.shape(v) for k, v in arrays.items()})


def golden_roessler():
	system = systems.two_roessler()
	prog = systems.program(system)
	roots = roots_from(oracle.build(prog))
	data = np.loadtxt(os.path.join(REF_TESTS, "two_Roessler_past.dat"))
	y10 = np.loadtxt(os.path.join(REF_TESTS, "two_Roessler_y10.dat"))
	past = [(r[0], r[1:7], r[7:13]) for r in data]
	times = np.linspace(0, 10, 10, endpoint=True)

	core = make_core(prog, "ref_roessler", past)
	C = controller.Controller(core, system["delay"], root_mode=1, roots=roots, **systems.ROESSLER_TEST_PARAMETERS)
	samples = np.array([C.integrate(T) for T in times])
	result = dict(samples=samples, accepted=C.accepted, rejected=C.rejected, final_t=C.t)

	# jump in the middle (tests/test_jitcdde.py:110-118) + adjust_diff at the start (:98-102)
	core = make_core(prog, "ref_roessler", past)
	C = controller.Controller(core, system["delay"], root_mode=1, roots=roots, **systems.ROESSLER_TEST_PARAMETERS)
	ad_min, ad_max = C.adjust_diff()
	mid = C.integrate(5.0)
	change = np.random.default_rng(42).normal(0, 0.1, 6)
	j_min, j_max = C.jump(change, 5.0)
	after_jump = C.integrate(7.0)
	state_anchors = C.get_state()
	save("roessler", **past_arrays(past), y10_reference=y10, times=times, **result,
		adjust_diff_min=ad_min, adjust_diff_max=ad_max, mid=mid, change=change, jump_min=j_min, jump_max=j_max,
		after_jump=after_jump, n_anchors_after=len(state_anchors), jump_accepted=C.accepted, jump_rejected=C.rejected)


def golden_mg_ensemble():
	system = systems.mackey_glass()
	prog = systems.program(system)
	roots = roots_from(oracle.build(prog))
	N = 24
	pars = systems.mg_grid(N)
	past = [(-1.0, [1.0], [0.0]), (0.0, [1.0], [0.0])]
	offsets = np.arange(0, 300, 10.0)
	out = {}
	for flavour, root_mode in (("chain", 1), ("literal", 0)):
		samples = np.empty((len(offsets), N, 1))
		counts = np.empty((N, 2), dtype=np.int64)
		final_t = np.empty(N)
		disc = np.empty((N, 1))
		n_anchors = np.empty(N, dtype=np.int64)
		for m in range(N):
			core = make_core(prog, f"ref_mg_{flavour}", past, flavour=flavour, parameters=pars[m])
			C = controller.Controller(core, 25.0, root_mode=root_mode, roots=roots)
			disc[m] = C.step_on_discontinuities(controller.discontinuity_steps([pars[m, 1]]))
			t0 = C.t
			for s, off in enumerate(offsets):
				samples[s, m] = C.integrate(t0+off)
			counts[m] = C.accepted, C.rejected
			final_t[m] = C.t
			n_anchors[m] = len(C.get_state())
		out[flavour] = dict(samples=samples, counts=counts, final_t=final_t, disc=disc, n_anchors=n_anchors)
	save("mg_ensemble", parameters=pars, offsets=offsets, **past_arrays(past),
		**{f"{k}_{fl}": v for fl, d in out.items() for k, v in d.items()})


def golden_neutral():
	system = systems.neutral()
	prog = systems.program(system)
	roots = roots_from(oracle.build(prog))
	past = [(-1.0, system["initial"], np.zeros(6)), (0.0, system["initial"], np.zeros(6))]
	core = make_core(prog, "ref_neutral", past)
	C = controller.Controller(core, system["delay"], root_mode=1, roots=roots, rtol=1e-5)
	mi, ma = C.adjust_diff()
	times = np.linspace(0.5, 5, 10)
	samples = np.array([C.integrate(T) for T in times])
	save("neutral", **past_arrays(past), adjust_diff_min=mi, adjust_diff_max=ma, times=times, samples=samples,
		control=np.array(system["control"]), accepted=C.accepted, rejected=C.rejected)


def golden_state_dependent():
	system = systems.state_dependent()
	prog = systems.program(system)
	roots = roots_from(oracle.build(prog))
	past = system["past"]
	core = make_core(prog, "ref_state_dependent", past)
	C = controller.Controller(core, system["max_delay"], root_mode=1, roots=roots)
	blind = C.integrate_blindly(0.01)
	times = np.linspace(0.9, 2, 30)
	samples = np.array([C.integrate(T) for T in times])
	save("state_dependent", **past_arrays(past), blind=blind, times=times, samples=samples,
		accepted=C.accepted, rejected=C.rejected, n_anchors=len(C.get_state()))


def golden_ode():
	system = systems.fitzhugh_nagumo_ode()
	prog = systems.program(system)
	roots = roots_from(oracle.build(prog))
	past = [(-1.0, system["y0"], system["f_of_y0"]), (0.0, system["y0"], system["f_of_y0"])]
	core = make_core(prog, "ref_ode", past)
	C = controller.Controller(core, 0.0, root_mode=1, roots=roots)
	adaptive = C.integrate(1.0)
	core = make_core(prog, "ref_ode", past)
	C2 = controller.Controller(core, 0.0, root_mode=1, roots=roots)
	blind = C2.integrate_blindly(0.98, 0.01)
	after_blind = C2.integrate(1.0)
	save("ode", **past_arrays(past), adaptive=adaptive, blind=blind, after_blind=after_blind, y1=system["y1"],
		accepted=C.accepted, rejected=C.rejected)


def golden_kuramoto():
	system = systems.kuramoto()
	prog = systems.program(system)
	roots = roots_from(oracle.build(prog))
	n = prog.n
	past = [(-1.0, system["initial"], np.zeros(n)), (0.0, system["initial"], np.zeros(n))]
	core = make_core(prog, "ref_kuramoto", past)
	C = controller.Controller(core, system["max_delay"], root_mode=1, roots=roots, rtol=0, atol=1e-5)
	blind = C.integrate_blindly(system["max_delay"], 0.1)
	t0 = C.t
	offsets = np.arange(0, 4, 0.2)
	samples = np.array([C.integrate(t0+o) for o in offsets])
	save("kuramoto", **past_arrays(past), blind=blind, offsets=offsets, samples=samples, accepted=C.accepted, rejected=C.rejected)


def network_graph(n=128, k=10, seed=11):
	"""fixed in-degree random graph of BASELINE.json configs[3] at a size the reference core can take (one search cursor per
	edge: 1280 call sites): tau ~ U(pi/5, 2pi), natural frequencies ~ U(0.8, 1.2), coupling c·q/k per edge"""
	rng = np.random.default_rng(seed)
	dst = np.repeat(np.arange(n, dtype=np.int64), k)
	src = rng.integers(0, n, size=n*k)
	tau = rng.uniform(np.pi/5, 2*np.pi, size=n*k)
	return dict(n=n, src=src, dst=dst, edge_tau=tau, weight=42*0.05/k, omega=rng.uniform(0.8, 1.2, n), initial=rng.uniform(0, 2*np.pi, n))


def _reference_network_core(g, name):
	"""the reference's C core for the network g: one `set_dy(i, …)` per node, one `anchors()` call per in-edge
	(/root/reference/examples/kuramoto_network.py:50-62) — with the terms of a node in the order and association of the
	engine (local term first, then the in-edges in input order, each `weight*coupling` added to the running sum), the
	coupling function printed by the same code generator, sin from csrc/jdde_math.h"""
	import types
	import sympy
	from jitcdde_b200 import _codegen
	n, E = g["n"], len(g["src"])
	net_program = _codegen.generate_network(lambda y, t, w: w, lambda yd, y: sympy.sin(yd-y), 1)
	lines = []
	for i in range(n):
		edges = np.nonzero(g["dst"] == i)[0]      # input order within a node: what the CSR build of jitcdde_network keeps
		terms = f"JDN_LOCAL(current_y({i}), t, par)"
		for e in edges:
			tau = repr(float(g["edge_tau"][e]))
			terms += f" + {float(g['weight'])!r}*JDN_COUPLING(past_y(t-{tau}, {int(g['src'][e])}, anchors(t-{tau})), current_y({i}))"
		lines.append(f"{{ double const par[1] = {{ {float(g['omega'][i])!r} }}; set_dy({i}, {terms}); }}")
	program = types.SimpleNamespace(n=n, n_sites=E, param_names=[], literal_f=lines, literal_sites=E, body="")
	definitions = f'#include "{os.path.join(ROOT, "jitcdde_b200", "csrc", "jdde_math.h")}"\n' + net_program.body
	build_ref.build_ref(program, name, flavour="literal", definitions=definitions, force=True)
	module = build_ref.load_ref(name)
	past = [(-1.0, g["initial"], np.zeros(n)), (0.0, g["initial"], np.zeros(n))]
	return module.dde_integrator([(a[0], np.array(a[1], dtype=float), np.array(a[2], dtype=float)) for a in past])


def golden_network():
	"""
	The structured network engine (jitcdde_network) against the REFERENCE'S C core on a 128-node delay-coupled Kuramoto
	network, driven by the reference's controller restated in oracle/controller.py with max_step = smallest delay (the
	step limit the network engine had before it could iterate on a past within the step; it is never reached here).
	"""
	g = network_graph()
	n = g["n"]
	core = _reference_network_core(g, "ref_network")
	roots = roots_from(oracle.build(systems.program(systems.mackey_glass())))
	max_delay, min_delay = float(np.max(g["edge_tau"])), float(np.min(g["edge_tau"]))
	C = controller.Controller(core, max_delay, root_mode=1, roots=roots, rtol=0, atol=1e-5, max_step=min_delay, first_step=min(1.0, min_delay))
	blind = C.integrate_blindly(max_delay, 0.1)
	times = max_delay + np.arange(0.5, 8.5, 0.5)
	samples = np.array([C.integrate(T) for T in times])
	save("network", src=g["src"], dst=g["dst"], edge_tau=g["edge_tau"], weight=g["weight"], omega=g["omega"], initial=g["initial"],
		blind_to=max_delay, blind=blind, times=times, samples=samples, accepted=C.accepted, rejected=C.rejected, final_t=C.t)


def golden_network_pws():
	"""
	The same with delays SHORTER than the steps (past within step, _jitcdde.py:838-861): every tenth edge has a delay of
	0.002–0.03 while the tolerance allows steps of ≈ 0.05–0.3, so lookups reach into the step under way, the step is
	repeated until two successive results agree, shrunk when it is far from explicit, and the step size grows by the
	credit rule afterwards. Default max_step (10), default pws parameters, no blind start: the integration starts at the
	end of the constant past with first_step = 1 and has to find its step size through rejections with a tentative
	anchor in place. Written by the reference's C core under the controller of oracle/controller.py.
	"""
	g = network_graph(n=96, k=6, seed=23)
	rng = np.random.default_rng(5)
	short = rng.random(len(g["edge_tau"])) < 0.1
	g["edge_tau"] = np.where(short, rng.uniform(0.002, 0.03, len(short)), g["edge_tau"])
	core = _reference_network_core(g, "ref_network_pws")
	roots = roots_from(oracle.build(systems.program(systems.mackey_glass())))
	max_delay = float(np.max(g["edge_tau"]))
	C = controller.Controller(core, max_delay, root_mode=1, roots=roots, rtol=0, atol=1e-4)
	times = np.arange(0.5, 12.5, 0.5)
	samples, dts = [], []
	for T in times:
		samples.append(C.integrate(T))
		dts.append(C.dt)
	print(f"network_pws: accepted {C.accepted}, rejected {C.rejected}, short delays {int(short.sum())}, dt {min(dts):.4f}..{max(dts):.4f}")
	save("network_pws", src=g["src"], dst=g["dst"], edge_tau=g["edge_tau"], weight=g["weight"], omega=g["omega"], initial=g["initial"],
		times=times, samples=np.array(samples), dts=np.array(dts), accepted=C.accepted, rejected=C.rejected, final_t=C.t)


def saturated_ring(n=8, tau=0.1):
	"""a ring of n nearly uncoupled phase oscillators, every delay equal to tau: with a loose tolerance the step size saturates at
	max_step = tau, and (t+h)−tau then lands one ulp above t in a few percent of the attempts"""
	rng = np.random.default_rng(2)
	dst = np.arange(n, dtype=np.int64)
	return dict(n=n, src=(dst-1) % n, dst=dst, edge_tau=np.full(n, tau), weight=0.05, omega=rng.uniform(0.9, 1.1, n), initial=rng.uniform(0, 2*np.pi, n))


def golden_network_saturated():
	"""step size saturated at max_step = the (only) delay, then blind steps of exactly that length — by the reference's C core:
	the one-ulp overlaps are past-within-step events there, too (the step is shrunk by pws_factor and the attempt rejected)"""
	g = saturated_ring()
	core = _reference_network_core(g, "ref_network_saturated")
	roots = roots_from(oracle.build(systems.program(systems.mackey_glass())))
	C = controller.Controller(core, 0.1, root_mode=1, roots=roots, rtol=0, atol=1e-3, max_step=0.1, first_step=0.1)
	times = np.arange(2.0, 42.0, 2.0)
	samples, dts = [], []
	for T in times:
		samples.append(C.integrate(T))
		dts.append(C.dt)
	blind = C.integrate_blindly(C.t+5.0, 0.1)
	print(f"network_saturated: accepted {C.accepted}, rejected {C.rejected}, dt {min(dts)}..{max(dts)}")
	save("network_saturated", src=g["src"], dst=g["dst"], edge_tau=g["edge_tau"], weight=g["weight"], omega=g["omega"], initial=g["initial"],
		times=times, samples=np.array(samples), dts=np.array(dts), blind=blind, accepted=C.accepted, rejected=C.rejected, final_t=C.t)


def hub_graph(n=200, hub=57, hub_degree=1100, seed=9):
	"""a hub with more in-edges than a tile of the persistent stepper keeps in shared memory (1024: that tile goes through the
	scratch array), nodes without in-edges, in-degree 0–4 elsewhere"""
	rng = np.random.default_rng(seed)
	degree = rng.integers(0, 5, n)
	degree[hub] = hub_degree
	degree[[3, 120]] = 0
	dst = np.repeat(np.arange(n, dtype=np.int64), degree)
	src = rng.integers(0, n, len(dst))
	tau = rng.uniform(0.5, 3.0, len(dst))
	return dict(n=n, src=src, dst=dst, edge_tau=tau, weight=0.002, omega=rng.uniform(0.8, 1.2, n), initial=rng.uniform(0, 2*np.pi, n))


def golden_network_hub():
	"""a network with a hub of 1100 in-edges (the tile with the hub does not fit the shared-memory staging of jdn_k_run) and
	nodes without in-edges — by the reference's C core"""
	g = hub_graph()
	core = _reference_network_core(g, "ref_network_hub")
	roots = roots_from(oracle.build(systems.program(systems.mackey_glass())))
	max_delay, min_delay = float(np.max(g["edge_tau"])), float(np.min(g["edge_tau"]))
	C = controller.Controller(core, max_delay, root_mode=1, roots=roots, rtol=0, atol=1e-5, max_step=min_delay, first_step=min(1.0, min_delay))
	blind = C.integrate_blindly(max_delay, 0.1)
	times = max_delay + np.arange(0.5, 4.5, 0.5)
	samples = np.array([C.integrate(T) for T in times])
	print(f"network_hub: {len(g['src'])} edges, accepted {C.accepted}, rejected {C.rejected}")
	save("network_hub", src=g["src"], dst=g["dst"], edge_tau=g["edge_tau"], weight=g["weight"], omega=g["omega"], initial=g["initial"],
		blind_to=max_delay, blind=blind, times=times, samples=samples, accepted=C.accepted, rejected=C.rejected, final_t=C.t)


def golden_network_adjust_diff():
	"""adjust_diff at the end of a constant past (the alternative to the blind start of examples/kuramoto_network.py), then an
	adaptive integration — by the reference's C core under the controller of oracle/controller.py"""
	g = network_graph(n=64, k=5, seed=3)
	core = _reference_network_core(g, "ref_network_adjust")
	roots = roots_from(oracle.build(systems.program(systems.mackey_glass())))
	max_delay = float(np.max(g["edge_tau"]))
	C = controller.Controller(core, max_delay, root_mode=1, roots=roots, rtol=0, atol=1e-5)
	minima, maxima = C.adjust_diff()
	t_after = C.t
	times = np.arange(0.5, 6.5, 0.5)
	samples = np.array([C.integrate(T) for T in times])
	print(f"network_adjust_diff: accepted {C.accepted}, rejected {C.rejected}")
	save("network_adjust_diff", src=g["src"], dst=g["dst"], edge_tau=g["edge_tau"], weight=g["weight"], omega=g["omega"], initial=g["initial"],
		minima=minima, maxima=maxima, t_after=t_after, times=times, samples=samples, accepted=C.accepted, rejected=C.rejected, final_t=C.t)


def golden_network_jump():
	"""jumps during the integration of a network (forward with one amplitude per node, backward with a common one) — by the
	reference's C core under the controller of oracle/controller.py"""
	g = network_graph(n=64, k=5, seed=3)
	core = _reference_network_core(g, "ref_network_jump")
	roots = roots_from(oracle.build(systems.program(systems.mackey_glass())))
	max_delay, min_delay = float(np.max(g["edge_tau"])), float(np.min(g["edge_tau"]))
	C = controller.Controller(core, max_delay, root_mode=1, roots=roots, rtol=0, atol=1e-5, max_step=min_delay, first_step=min(1.0, min_delay))
	C.integrate_blindly(max_delay, 0.1)
	first = C.integrate(max_delay+2.0)
	amplitude = np.random.default_rng(4).uniform(-0.5, 0.5, g["n"])
	extrema_1 = C.jump(amplitude, C.t, 1e-5, forward=True)
	t_1 = C.t
	second = np.array([C.integrate(T) for T in max_delay+np.array([2.5, 3.0, 4.0])])
	extrema_2 = C.jump(0.3, C.t, 1e-4, forward=False)
	t_2 = C.t
	third = np.array([C.integrate(T) for T in max_delay+np.array([4.5, 5.0, 6.0])])
	print(f"network_jump: accepted {C.accepted}, rejected {C.rejected}")
	save("network_jump", src=g["src"], dst=g["dst"], edge_tau=g["edge_tau"], weight=g["weight"], omega=g["omega"], initial=g["initial"],
		blind_to=max_delay, first=first, amplitude=amplitude, minima_1=extrema_1[0], maxima_1=extrema_1[1], t_1=t_1, second=second,
		minima_2=extrema_2[0], maxima_2=extrema_2[1], t_2=t_2, third=third, accepted=C.accepted, rejected=C.rejected, final_t=C.t)


def two_delay_graph(n=24, seed=6):
	"""in-degree 2, delays drawn from {0.7, 1.3}: few distinct delays, the case step_on_discontinuities is meant for"""
	rng = np.random.default_rng(seed)
	dst = np.repeat(np.arange(n, dtype=np.int64), 2)
	src = rng.integers(0, n, 2*n)
	tau = rng.choice([0.7, 1.3], 2*n)
	return dict(n=n, src=src, dst=dst, edge_tau=tau, weight=0.4, omega=rng.uniform(0.8, 1.2, n), initial=rng.uniform(0, 2*np.pi, n))


def golden_network_discontinuities():
	"""step_on_discontinuities (two propagations) for a network with two distinct delays, then an adaptive integration — by the
	reference's C core under the controller of oracle/controller.py"""
	g = two_delay_graph()
	core = _reference_network_core(g, "ref_network_disc")
	roots = roots_from(oracle.build(systems.program(systems.mackey_glass())))
	C = controller.Controller(core, 1.3, root_mode=1, roots=roots, rtol=0, atol=1e-6)
	steps = controller.discontinuity_steps([0.7, 1.3], propagations=2)
	after = C.step_on_discontinuities(steps)
	t_after = C.t
	times = t_after + np.arange(0.5, 4.5, 0.5)
	samples = np.array([C.integrate(T) for T in times])
	print(f"network_discontinuities: steps {steps}, accepted {C.accepted}, rejected {C.rejected}")
	save("network_discontinuities", src=g["src"], dst=g["dst"], edge_tau=g["edge_tau"], weight=g["weight"], omega=g["omega"], initial=g["initial"],
		steps=np.array(steps), after=after, t_after=t_after, times=times, samples=samples, accepted=C.accepted, rejected=C.rejected, final_t=C.t)


def golden_data_input():
	"""the extended system and joined past that jitcdde_input hands to the core (reference: _jitcdde.py:1519-1659)"""
	prog, past, max_delay = systems.data_input_program()
	roots = roots_from(oracle.build(prog))
	core = make_core(prog, "ref_data_input", past)
	C = controller.Controller(core, max_delay, root_mode=1, roots=roots)
	mi, ma = C.adjust_diff()
	times = np.linspace(0, 70, 36)[1:]
	samples = np.array([C.integrate(T)[:2] for T in times])
	save("data_input", **past_arrays(past), adjust_diff_min=mi[:2], adjust_diff_max=ma[:2], times=times, samples=samples,
		accepted=C.accepted, rejected=C.rejected, max_delay=max_delay)


def golden_lyap():
	# Mackey–Glass with 3 exponents (limit cycle: deterministic, non-chaotic), cf. examples/mackey_glass_lyap.py
	system = systems.mackey_glass(control=False)
	n_lyap = 3
	prog, n_basic = systems.lyap_program(system, n_lyap)
	roots = roots_from(oracle.build(prog, n_basic=n_basic))
	P = Past(n=prog.n, n_basic=n_basic, rng=np.random.default_rng(7))
	P.constant([1.0])
	past = [(a.time, a.state, a.diff) for a in P]
	core = make_core(prog, "ref_mg_lyap", past, n_basic=n_basic)
	C = controller.Controller(core, 15.0, root_mode=1, roots=roots)
	C.step_on_discontinuities(controller.discontinuity_steps([15.0]))
	t0 = C.t
	offsets = np.arange(10, 310, 10.0)
	states, lyaps, weights = [], [], []
	for o in offsets:
		s, l, w = C.integrate_lyap(t0+o, n_basic, n_lyap)
		states.append(s); lyaps.append(l); weights.append(w)
	save("mg_lyap", **past_arrays(past), n_lyap=n_lyap, offsets=offsets, states=np.array(states), lyaps=np.array(lyaps), weights=np.array(weights),
		accepted=C.accepted, rejected=C.rejected)

	# two Rösslers with 4 exponents (tests/test_lyaps.py:16-56), short horizon for the bitwise check
	system = systems.two_roessler()
	prog, n_basic = systems.lyap_program(system, 4)
	roots = roots_from(oracle.build(prog, n_basic=n_basic))
	rng = np.random.default_rng(seed=42)
	P = Past(n=prog.n, n_basic=n_basic, rng=np.random.default_rng(11))
	P.add((-system["delay"], rng.random(6), rng.random(6)))
	P.add((0.0, rng.random(6), rng.random(6)))
	past = [(a.time, a.state, a.diff) for a in P]
	core = make_core(prog, "ref_roessler_lyap", past, n_basic=n_basic)
	C = controller.Controller(core, system["delay"], root_mode=1, roots=roots, **systems.LYAP_TEST_PARAMETERS)
	C.step_on_discontinuities(controller.discontinuity_steps([system["delay"]]))
	t0 = C.t
	offsets = np.arange(10, 60, 10.0)
	states, lyaps, weights = [], [], []
	for o in offsets:
		s, l, w = C.integrate_lyap(t0+o, n_basic, 4)
		states.append(s); lyaps.append(l); weights.append(w)
	save("roessler_lyap", **past_arrays(past), offsets=offsets, states=np.array(states), lyaps=np.array(lyaps), weights=np.array(weights),
		accepted=C.accepted, rejected=C.rejected, lyap_controls=np.array([0.0806, 0, -0.0368, -0.1184]))


def golden_projected():
	# restricted and transversal Lyapunov exponents of three synchronised oscillators (tests/test_transversal_lyap.py), short
	# protocol: 20 blind calls of one step each, then 12 adaptive samples — all from the reference's C core
	# (remove_state_component / remove_diff_component / remove_projections / normalise_indices)
	system = systems.synchronised_fhn()
	for kind in ("restricted", "transversal"):
		DDE, prog, projector = systems.projected_program(kind)
		single = np.array([0.3, 0.1])
		initial = np.array([single[j] for j, group in enumerate(system["groups"]) for _ in group]) if kind == "restricted" else single
		DDE.constant_past(initial, 0.0)
		past = [(a.time, a.state, a.diff) for a in DDE.past]
		roots = roots_from(oracle.build(prog, n_basic=DDE.n_basic))
		name = "ref_"+kind
		build_ref.build_ref(prog, name, n_basic=DDE.n_basic, force=True, tangent_indices=projector.get("indices", ()))
		core = build_ref.load_ref(name).dde_integrator([(float(a[0]), np.array(a[1], dtype=float), np.array(a[2], dtype=float)) for a in past])
		core.set_parameters(-0.1)
		C = controller.Controller(core, system["tau"], root_mode=1, roots=roots)
		blind = [C.integrate_blindly_projected(float(T), projector) for T in range(1, 21)]
		times = C.t + np.arange(20, 260, 20.0)
		samples = [C.integrate_projected(T, projector) for T in times]
		save(kind, **past_arrays(past), k=-0.1, blind_states=np.array([b[0] for b in blind]), blind_lyaps=np.array([b[1] for b in blind]),
			blind_totals=np.array([b[2] for b in blind]), times=times, states=np.array([s[0] for s in samples]),
			lyaps=np.array([s[1] for s in samples]), weights=np.array([s[2] for s in samples]), accepted=C.accepted, rejected=C.rejected)


if __name__ == "__main__":
	assert build_ref.reference_available(), "run this where /root/reference exists"
	makers = dict(roessler=golden_roessler, mg_ensemble=golden_mg_ensemble, neutral=golden_neutral, state_dependent=golden_state_dependent,
		ode=golden_ode, kuramoto=golden_kuramoto, lyap=golden_lyap, data_input=golden_data_input, projected=golden_projected, network=golden_network, network_pws=golden_network_pws,
		network_adjust_diff=golden_network_adjust_diff, network_saturated=golden_network_saturated,
		network_hub=golden_network_hub, network_jump=golden_network_jump,
		network_discontinuities=golden_network_discontinuities)
	for name in sys.argv[1:] or makers:      # optional arguments: only these fixtures
		makers[name]()
