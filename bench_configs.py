"""
bench_configs.py — the legs of bench.py for the other BASELINE.json configurations:

  C1  examples/mackey_glass.py, ONE trajectory                       (configs[0])
  C3  two Rösslers + 4 Lyapunov exponents, ensemble over the coupling (configs[2])
  C4  kuramoto_network, one delay-coupled system, node-partitioned    (configs[3])  → key "network" of the bench line
  C5a state-dependent delay, C5b neutral DDE, ensembles               (configs[4])

Each leg runs the configuration through the public Python API (production build), times the sampling loop on the
device (CUDA events on the engine's stream; setup — code generation, image, upload, discontinuity handling — is
outside) and reports its metric with a `roofline` object of its own: ALGORITHMIC bytes per attempted step (SURVEY.md
§8d: read current anchor 8(1+2n) + 3 stages × L lookups × 2 anchors × 8(1+2m) + write tentative anchor 8(1+2n) + 32 B
of controller state) × attempts/s against the measured copy bandwidth. Parity of the same configurations:
tests/test_gpu_parity.py, tests/test_gpu_network.py. These are secondary lines; the headline is bench.py's own.
"""

import time

import numpy as np


def _events():
	import torch
	return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _roofline(bytes_per_attempt, attempts, seconds, peak, kernel, note=None):
	achieved = bytes_per_attempt*attempts/seconds/1e9
	r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak, "traffic": None,
		"kernel": kernel, "algorithmic_bytes_per_attempted_step": bytes_per_attempt}
	if note:
		r["note"] = note
	if achieved > peak:
		r["note"] = (r.get("note", "")+"; frac > 1: the algorithmic bytes count every anchor access as DRAM traffic, but the staged window and the L2 "
			"serve most of them here (few live anchors per member) — the kernel is not DRAM-bound on this workload").lstrip("; ")
	return r


def _timed_sampling(DDE, sample, n_warm, n_timed):
	"""sample(k) queues one sampling call; returns (seconds on the device, accepted, attempts) of the timed calls"""
	import torch
	engine = DDE._engine
	stream = torch.cuda.current_stream()
	engine.set_stream(stream.cuda_stream)
	for k in range(n_warm):
		sample(k)
	a0, r0, f0 = engine.sum_counters()
	ev0, ev1 = _events()
	torch.cuda.synchronize()
	ev0.record(stream)
	for k in range(n_warm, n_warm+n_timed):
		sample(k)
	ev1.record(stream)
	torch.cuda.synchronize()
	a1, r1, f1 = engine.sum_counters()
	if engine.count_failed():
		raise RuntimeError(f"members failed: status {np.unique(engine.get_status())}")
	return ev0.elapsed_time(ev1)*1e-3, a1-a0, (f1-f0)//3


def leg_c1(peak):
	"""examples/mackey_glass.py:64-79 — ONE trajectory (β = 0.25, τ = 15), sampled every 10 time units over 10 000."""
	from jitcdde_b200 import jitcdde, t, y
	DDE = jitcdde([0.25*y(0, t-15)/(1+y(0, t-15)**10) - 0.1*y(0)], verbose=False)
	DDE.constant_past([1.0])
	DDE.step_on_discontinuities()
	start = DDE.t
	# (a) the reference's loop: one call (one launch, one result) per sample
	w0 = time.perf_counter()
	for k in range(1, 201):
		DDE.integrate(start+10.0*k)
	per_call = (time.perf_counter()-w0)/200
	a0 = DDE.counters()[0][0]
	# (b) the whole series in one launch
	times = start+2000.0+10.0*np.arange(1, 801)
	w0 = time.perf_counter()
	DDE.integrate_samples(times)
	series = time.perf_counter()-w0
	a1 = DDE.counters()[0][0]
	return {
		"workload": "single Mackey-Glass trajectory, beta 0.25, tau 15, sampled every 10 time units (BASELINE.json configs[0])",
		"metric": "accepted steps/s of one trajectory", "unit": "steps/s",
		"value": float((a1-a0)/series), "value_per_call_loop": float(a0/(200*per_call)) if per_call else None,
		"us_per_integrate_call": 1e6*per_call, "us_per_sample_in_series": 1e6*series/len(times),
		"roofline": None,
		"note": "one thread of one SM: bound by the dependent FP64 chain of a single trajectory (latency), no throughput roofline applies; "
			"integrate_samples removes the launch + result round trip per sample",
	}


def leg_c3(peak, N):
	"""two Rösslers + 4 Lyapunov exponents (tests/test_lyaps.py:16-38) over a grid of couplings: lane-group stepper +
	Gram–Schmidt over the delay window per sample"""
	from jitcdde_b200 import jitcdde_lyap
	from tests import scenarios, systems
	past, k, offsets = scenarios.c3_inputs(N)
	system = systems.two_roessler(control=True)
	DDE = jitcdde_lyap(system["f"], n_lyap=4, control_pars=system["control_pars"], n_members=N, verbose=False, ring_capacity=2048)      # ≈ 1000 live anchors per member after the transient (delay 4.5, rtol 1e-7); 1 MB of ring per member
	DDE.add_past_points(past)
	DDE.set_integration_parameters(**systems.LYAP_TEST_PARAMETERS)
	DDE.set_parameters(k)
	DDE.step_on_discontinuities()
	t0 = DDE.t
	DDE.integrate(t0+20.0)      # transient: the ring grows to its working size outside the timed part
	engine = DDE._engine
	base = np.asarray(engine.get_t())
	seconds, accepted, attempts = _timed_sampling(DDE, lambda i: engine.integrate_lyap(base+10.0*(i+1)), 1, 4)
	return {
		"workload": f"two Roesslers + 4 Lyapunov exponents, {N} couplings (BASELINE.json configs[2]), 4 samples of 10 time units, orthonormalisation per sample",
		"metric": "accepted member-steps/s (extended system, n = 30)", "unit": "member-steps/s", "value": accepted/seconds,
		"members": N, "attempts_per_accepted": attempts/max(accepted, 1), "ms_per_sample": 1e3*seconds/4,
		"roofline": _roofline(1536, attempts, seconds, peak, "jdde_k_integrate_group (+ jdde_k_orthonormalise_team, not counted)",
			"n = 30, one distinct delay touching 5 components: 488 + 3·2·88 + 488 + 32 B per attempt"),
	}


def leg_c5a(peak, N):
	"""examples/state_dependent.py:14-29: delay depends on the state (lookups move back and forth), max_delay = 1e10: nothing is forgotten"""
	from jitcdde_b200 import jitcdde
	from tests import scenarios, systems
	times, states, diffs, sample_times = scenarios.c5a_inputs(N)
	system = systems.state_dependent()
	DDE = jitcdde(system["f"], max_delay=1e10, n_members=N, verbose=False, ring_capacity=512)      # nothing is forgotten: ≈ 100–200 anchors per member
	for j in range(4):
		DDE.add_past_point(times[:, j], states[:, j], diffs[:, j])
	DDE.integrate_blindly(0.01)
	engine = DDE._engine
	grid = np.linspace(0.9, 2, 12)
	seconds, accepted, attempts = _timed_sampling(DDE, lambda i: engine.integrate_samples(grid[4*i:4*i+4], fetch=False), 1, 2)
	return {
		"workload": f"state-dependent delay (examples/state_dependent.py), {N} members with shifted pasts (BASELINE.json configs[4]), 8 samples up to t = 2",
		"metric": "accepted member-steps/s", "unit": "member-steps/s", "value": accepted/seconds,
		"members": N, "attempts_per_accepted": attempts/max(accepted, 1), "ms_per_sample": 1e3*seconds/8,
		"roofline": _roofline(224, attempts, seconds, peak, "jdde_k_integrate", "n = 1, one lookup: as the Mackey-Glass ensemble, 224 B per attempt"),
	}


def leg_c5b(peak, N):
	"""examples/neutral.py:14-56: neutral DDE, n = 6, one anchor helper → 6 past_y + 3 past_dy"""
	from jitcdde_b200 import jitcdde
	from tests import scenarios, systems
	initial, sample_times = scenarios.c5b_inputs(N)
	system = systems.neutral()
	DDE = jitcdde(system["f"], helpers=system["helpers"], n_members=N, verbose=False, ring_capacity=256)
	DDE.set_integration_parameters(rtol=1e-5)
	DDE.constant_past(initial)
	DDE.adjust_diff()
	engine = DDE._engine
	grid = np.linspace(1, 5, 6)
	seconds, accepted, attempts = _timed_sampling(DDE, lambda i: engine.integrate_samples(grid[2*i:2*i+2], fetch=False), 1, 2)
	return {
		"workload": f"neutral DDE (examples/neutral.py), {N} members with shifted initial states (BASELINE.json configs[4]), 4 samples up to t = 5",
		"metric": "accepted member-steps/s", "unit": "member-steps/s", "value": accepted/seconds,
		"members": N, "attempts_per_accepted": attempts/max(accepted, 1), "ms_per_sample": 1e3*seconds/4,
		"roofline": _roofline(864, attempts, seconds, peak, "jdde_k_integrate", "n = 6, one lookup of whole anchors: 104 + 3·2·104 + 104 + 32 B per attempt"),
	}


def ensemble_legs(peak, quick=False):
	scale = 10 if quick else 1
	legs = {}
	for name, fn, args in (("c1", leg_c1, ()), ("c3", leg_c3, (100_000//scale,)), ("c5a", leg_c5a, ((1 << 20)//scale,)), ("c5b", leg_c5b, ((1 << 20)//scale,))):
		w0 = time.perf_counter()
		try:
			legs[name] = fn(peak, *args)
			legs[name]["leg_seconds"] = time.perf_counter()-w0
		except Exception as error:      # a secondary line must not take the headline down
			legs[name] = {"error": f"{type(error).__name__}: {error}"}
	return legs


# --------------------------------------------------------------------------------------- network (C4)
def network_graph(n, k=10, seed=42):
	"""fixed in-degree k random graph, tau ~ U(pi/5, 2pi), omega = 1, coupling of examples/kuramoto_network.py:43-63 (c·q = 2.1)
	spread over the k inputs"""
	rng = np.random.default_rng(seed)
	dst = np.repeat(np.arange(n, dtype=np.int64), k)
	src = rng.integers(0, n, size=n*k)
	tau = rng.uniform(np.pi/5, 2*np.pi, size=n*k)
	return dict(n=n, src=src, dst=dst, edge_tau=tau, weight=42*0.05/k, omega=np.ones(n), initial=rng.uniform(0, 2*np.pi, n))


def network_leg(peak, device, rank, world, nodes=1_000_000, in_degree=10, span=1.0):
	"""One delay-coupled Kuramoto system of `nodes` nodes, partitioned by node over the `world` GPUs (every rank keeps the
	whole history; each attempted step exchanges the new anchor through peer memory from inside the step kernel,
	DESIGN.md §5). Timed: an adaptive integrate over `span` time units after the blind start (wall clock around the
	call on every rank, barrier on both sides, max over ranks: the integration loop runs in the runtime library)."""
	import sympy
	import torch
	import torch.distributed as dist
	from jitcdde_b200 import jitcdde_network
	try:
		g = network_graph(nodes, in_degree)
		net = jitcdde_network(nodes, local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y), edges=(g["src"], g["dst"], g["edge_tau"]),
			weight=g["weight"], node_parameters=[g["omega"]], verbose=False, device=device, ring_capacity=512)
		net.set_integration_parameters(rtol=0, atol=1e-5, max_step=float(np.min(g["edge_tau"])))      # steps up to the smallest delay (the workload of round 1; beyond it the engine iterates on the past within the step)
		net.constant_past(g["initial"], time=0.0)
		net.integrate_blindly(float(np.max(g["edge_tau"])), 0.1)
		net.integrate(net.t+0.2)      # warm-up: step size settles, graphs are captured
		a0 = net.counters()
		launches0 = net._engine.launch_count
		torch.cuda.synchronize()
		if world > 1:
			dist.barrier()
		w0 = time.perf_counter()
		out = net.integrate(net.t+span)
		torch.cuda.synchronize()
		seconds = time.perf_counter()-w0
		if world > 1:
			tmax = torch.tensor([seconds], dtype=torch.float64, device="cuda")
			dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
			seconds = float(tmax.item())
		a1 = net.counters()
		accepted, attempts = a1[0]-a0[0], a1[2]-a0[2]
		edges = nodes*in_degree
		# per attempt and edge: 3 stage lookups on 2–3 consecutive anchors of the source node (16-byte pairs: 48 B), edge data
		# (source 4, delay 8, cursor 4+4) and the three interpolated values written and read once (48 B): ≈ 116 B; per node:
		# current pair 16, new pair 16 (× ranks when exchanged), parameters 8, row pointers 16
		bytes_per_attempt = edges*116 + nodes*56
		return {
			"workload": f"Kuramoto network, {nodes} nodes, in-degree {in_degree}, tau ~ U(pi/5, 2pi), rtol 0, atol 1e-5 (BASELINE.json configs[3]); adaptive integrate over {span:g} time units after integrate_blindly(max tau, 0.1)",
			"metric": "accepted steps/s of the whole network", "unit": "steps/s", "value": accepted/seconds,
			"node_steps_per_s": accepted*nodes/seconds, "delayed_lookups_per_s": attempts*edges*3/seconds,
			"accepted": int(accepted), "attempts": int(attempts), "ms_per_attempt": 1e3*seconds/max(attempts, 1),
			"n_gpus": world, "scaling": "strong", "partition": "contiguous node blocks, full history replica per rank",
			"exchange": net._exchange if isinstance(net._exchange, str) else ("nccl" if world > 1 else None),
			"exchange_bytes_per_attempt": int(16*nodes*(world-1)/world*world) if world > 1 else 0,
			"gpu_launches": int(net._engine.launch_count-launches0),
			"checksum": float(np.sum(out)),
			"roofline": {"bound": "hbm", "achieved": bytes_per_attempt*attempts/seconds/1e9/world, "peak": peak, "unit": "GB/s",
				"frac": bytes_per_attempt*attempts/seconds/1e9/world/peak, "traffic": None, "kernel": "jdn_k_run (persistent cooperative stepper: lookups and stages of a tile in one pass)",
				"algorithmic_bytes_per_attempt": bytes_per_attempt, "note": "random 16-byte gathers: DRAM moves 32-byte sectors / 64-byte units, see profiles/"},
		}
	except Exception as error:
		return {"error": f"{type(error).__name__}: {error}"}
