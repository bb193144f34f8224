"""GPU tests of the network engine: against the symbolic path of this package (itself pinned to the reference core)
on a small Kuramoto network, and determinism / partition-independence properties at a size where the symbolic path is
impossible (10^5 nodes)."""

import numpy as np
import pytest
import sympy

from tests.test_network import check_network_discontinuities, check_network_jump, checkpoint_round_trip, kuramoto_graph, run_network, run_network_adjust_diff_golden, run_network_discontinuities_golden, run_network_golden, run_network_hub_golden, run_network_jump_golden, run_network_pws_golden, run_network_saturated_golden, run_symbolic

pytestmark = pytest.mark.gpu


def large_graph(n, k=10, seed=42, short_delays=False):
	"""BASELINE.json configs[3] shape: fixed in-degree k random graph, tau ~ U(pi/5, 2pi), omega = 1,
	coupling strength of examples/kuramoto_network.py (c·q = 2.1) spread over the k inputs. short_delays: every tenth delay
	is shorter than the steps the tolerance allows (past within step)."""
	rng = np.random.default_rng(seed)
	dst = np.repeat(np.arange(n, dtype=np.int64), k)
	src = rng.integers(0, n, size=n*k)
	tau = rng.uniform(np.pi/5, 2*np.pi, size=n*k)
	if short_delays:
		tau = np.where(rng.random(n*k) < 0.1, rng.uniform(0.002, 0.03, n*k), tau)
	return dict(n=n, src=src, dst=dst, edge_tau=tau, weight=42*0.05/k, omega=np.ones(n), initial=rng.uniform(0, 2*np.pi, n))


def test_network_engine_matches_symbolic_path_on_gpu():
	g = kuramoto_graph(12)      # same graph as the CPU test; its symbolic image is prebuilt by __graft_entry__.build()
	blind_to = float(np.max(g["edge_tau"]))
	times = blind_to + np.arange(0.5, 5.5, 0.5)
	nb, ns, nc, nt = run_network(g, times, blind_to)
	sb, ss, sc, st = run_symbolic(g, times, blind_to, verification=True)
	np.testing.assert_allclose(nb, sb, rtol=1e-9)
	np.testing.assert_allclose(ns, ss, rtol=1e-8)
	assert (nc[0], nc[1]) == sc


def test_network_engine_against_reference_core_on_gpu():
	"""BASELINE.json configs[3] at the size the reference core can take (128 nodes, in-degree 10): the CUDA verification build of
	the network engine reproduces the reference's C core bit for bit — identical step counts, identical states (golden:
	tests/golden/network.npz, written by the reference core itself)."""
	g, blind, samples, counters, t_end = run_network_golden()
	assert np.array_equal(blind, g["blind"])
	assert np.array_equal(samples, g["samples"])
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert t_end == float(g["final_t"])
	# the production build (FMA contraction, CUDA's sin): identical counts, states to rounding (north_star: 1e-10)
	from jitcdde_b200 import jitcdde_network
	import sympy
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False)
	net.set_integration_parameters(rtol=0, atol=1e-5, max_step=float(np.min(g["edge_tau"])))
	net.constant_past(g["initial"], time=0.0)
	net.integrate_blindly(float(g["blind_to"]), 0.1)
	fast = np.array([net.integrate(T) for T in g["times"]])
	np.testing.assert_allclose(fast, g["samples"], rtol=1e-10)
	assert net.counters()[:2] == (int(g["accepted"]), int(g["rejected"]))


def test_network_past_within_step_on_gpu():
	"""delays shorter than the step (tests/golden/network_pws.npz, written by the reference's C core under the reference's
	controller): the persistent stepper repeats a step until two successive results agree, as _jitcdde.py:838-861 —
	identical step counts, step sizes and states, bit for bit"""
	g, samples, dts, counters, t_end = run_network_pws_golden()
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert counters[2] > counters[0]+counters[1]
	assert np.array_equal(dts, g["dts"])
	assert np.array_equal(samples, g["samples"])
	assert t_end == float(g["final_t"])
	# production build: the same decisions, states to rounding
	from jitcdde_b200 import jitcdde_network
	net = jitcdde_network(len(g["omega"]), local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y),
		edges=(g["src"], g["dst"], g["edge_tau"]), weight=float(g["weight"]), node_parameters=[g["omega"]], verbose=False)
	net.set_integration_parameters(rtol=0, atol=1e-4)
	net.constant_past(g["initial"], time=0.0)
	fast = np.array([net.integrate(T) for T in g["times"]])
	np.testing.assert_allclose(fast, g["samples"], rtol=1e-9)
	assert net.counters()[:2] == (int(g["accepted"]), int(g["rejected"]))


def test_network_step_on_discontinuities_on_gpu():
	"""step_on_discontinuities for a network with two distinct delays (tests/golden/network_discontinuities.npz, reference core)"""
	check_network_discontinuities(*run_network_discontinuities_golden())


def test_network_checkpoint_and_resume_on_gpu():
	"""get_state (the live anchors, unpacked from the ring on the host) → add_past_points of a new network → the same trajectory"""
	checkpoint_round_trip()


def test_network_jump_on_gpu():
	"""jumps during the integration of a network (tests/golden/network_jump.npz, reference core): bit-identical"""
	check_network_jump(*run_network_jump_golden())


def test_network_with_a_hub_on_gpu():
	"""a hub with 1100 in-edges — its tile exceeds the 1024 edges jdn_k_run stages in shared memory and goes through the scratch
	array — and nodes without in-edges: bit-identical with the reference core (tests/golden/network_hub.npz)"""
	g, blind, samples, counters, t_end = run_network_hub_golden()
	assert np.array_equal(blind, g["blind"])
	assert np.array_equal(samples, g["samples"])
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert t_end == float(g["final_t"])


def test_step_size_saturated_at_the_smallest_delay_on_gpu():
	"""the step size saturates at max_step = the delay: one-ulp overlaps are handled as by the reference (tests/golden/network_saturated.npz,
	reference core), blind steps of exactly the delay included"""
	g, samples, dts, blind, counters, t_end = run_network_saturated_golden()
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert np.array_equal(dts, g["dts"])
	assert np.array_equal(samples, g["samples"])
	assert np.array_equal(blind, g["blind"])
	assert t_end == float(g["final_t"])


def test_network_adjust_diff_on_gpu():
	"""adjust_diff for a network (tests/golden/network_adjust_diff.npz, reference core): same extrema, same anchors — the
	integration that follows takes the same steps to the same bits"""
	g, minima, maxima, t_after, samples, counters, t_end = run_network_adjust_diff_golden()
	assert np.array_equal(minima, g["minima"]) and np.array_equal(maxima, g["maxima"])
	assert t_after == float(g["t_after"])
	assert (counters[0], counters[1]) == (int(g["accepted"]), int(g["rejected"]))
	assert np.array_equal(samples, g["samples"])
	assert t_end == float(g["final_t"])


@pytest.mark.parametrize("n", [100_000, 1_000_000])      # 10^6 nodes, 10^7 edges: BASELINE.json configs[3]
def test_large_network_properties(n):
	g = large_graph(n)
	blind_to = float(np.max(g["edge_tau"]))
	times = blind_to + np.array([1.0, 2.0])
	runs = [run_network(g, times, blind_to) for _ in range(2)]
	assert np.array_equal(runs[0][1], runs[1][1]) and runs[0][2] == runs[1][2]      # deterministic (max-reduction is order-free)
	blind, samples, (accepted, rejected, attempts), t_end = runs[0]
	assert np.all(np.isfinite(samples)) and accepted > 60 and attempts == accepted+rejected
	# phases advance roughly with omega = 1 plus bounded coupling
	drift = (samples[-1]-g["initial"])/times[-1]
	assert np.all(np.abs(drift-1.0) < 2.2)


def _start(net, g, short_delays):
	"""the blind start of examples/kuramoto_network.py — or, in the run with short delays, adjust_diff (both halves of it meet
	between the ranks) — and what the network looks like afterwards"""
	if short_delays:
		minima, maxima = net.adjust_diff()
		return np.concatenate([minima, maxima])
	return net.integrate_blindly(float(np.max(g["edge_tau"])), 0.1)


def _peer_worker(rank, world, port, queue, n, ring_capacity, short_delays=False):
	"""one rank of a two-process run: rank r on device r when the box has two GPUs (stores and arrival counters then cross
	NVLink); with a single visible GPU both ranks share it (CUDA IPC works within a device, too)"""
	import os
	import warnings
	warnings.simplefilter("ignore")
	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
	import torch
	import torch.distributed as dist
	from jitcdde_b200 import jitcdde_network
	device = rank % torch.cuda.device_count()
	torch.cuda.set_device(device)
	dist.init_process_group("gloo", rank=rank, world_size=world)      # rendezvous only: carries the IPC handles
	g = large_graph(n, short_delays=short_delays)
	net = jitcdde_network(n, local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y), edges=(g["src"], g["dst"], g["edge_tau"]),
		weight=g["weight"], node_parameters=[g["omega"]], verbose=False, device=device, ring_capacity=ring_capacity, exchange="peer")
	net.set_integration_parameters(rtol=0, atol=1e-5)
	net.constant_past(g["initial"], time=0.0)
	blind = _start(net, g, short_delays)
	out = net.integrate(net.t+2.0)
	queue.put((rank, blind, out, net.counters(), device, str(getattr(torch.cuda.get_device_properties(device), "uuid", device))))
	dist.barrier()
	dist.destroy_process_group()


@pytest.mark.parametrize("ring_capacity,short_delays", [(512, False), (16, False), (512, True)])      # 16: the ring fills up and is doubled on both ranks while they integrate; short delays: past within step, the overlap and "results differ" are merged between the ranks
def test_fused_peer_exchange_equals_one_gpu(ring_capacity, short_delays):
	"""Two ranks that own half of the nodes each and exchange the tentative anchor through peer memory from inside the step
	kernel (device-side arrival barrier, csrc/jdn_kernels.cu: jdn_k_exchange) take the same steps and reach bit-identical
	states as one rank that owns all nodes."""
	import socket
	import torch.multiprocessing as mp
	from jitcdde_b200 import jitcdde_network
	try:
		import pynvml
		pynvml.nvmlInit()
		if pynvml.nvmlDeviceGetCount() < 2 and pynvml.nvmlDeviceGetComputeMode(pynvml.nvmlDeviceGetHandleByIndex(0)) != pynvml.NVML_COMPUTEMODE_DEFAULT:
			pytest.skip("one GPU in exclusive-process mode: two ranks cannot share it")
	except ImportError:
		pass
	n = 20_000
	g = large_graph(n, short_delays=short_delays)
	net = jitcdde_network(n, local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y), edges=(g["src"], g["dst"], g["edge_tau"]),
		weight=g["weight"], node_parameters=[g["omega"]], verbose=False, ring_capacity=512)
	net.set_integration_parameters(rtol=0, atol=1e-5)
	net.constant_past(g["initial"], time=0.0)
	blind = _start(net, g, short_delays)
	out = net.integrate(net.t+2.0)
	counters = net.counters()
	if short_delays:
		assert counters[2] > counters[0]+counters[1]      # steps were repeated
	del net

	sock = socket.socket()
	sock.bind(("127.0.0.1", 0))
	port = sock.getsockname()[1]
	sock.close()
	ctx = mp.get_context("spawn")
	queue = ctx.Queue()
	procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, queue, n, ring_capacity, short_delays)) for r in range(2)]
	for p in procs:
		p.start()
	results = [queue.get(timeout=600) for _ in procs]
	for p in procs:
		p.join(timeout=120)
		assert p.exitcode == 0
	for rank, peer_blind, peer_out, peer_counters, device, uuid in results:
		assert np.array_equal(peer_blind, blind), rank
		assert np.array_equal(peer_out, out), rank
		assert tuple(peer_counters) == tuple(counters), rank
	import torch
	devices = sorted({(r[4], r[5]) for r in results})
	print(f"fused peer exchange ran on device(s) {[d[0] for d in devices]} of {torch.cuda.device_count()} visible")
	if torch.cuda.device_count() >= 2:
		# two GPUs on the box: the two ranks MUST have been on different devices (peer stores over NVLink, system-scope atomics)
		assert len(devices) == 2, devices


def test_ring_growth_on_gpu():
	"""a ring of 16 anchors for a delay window of 63 steps: filled, doubled (twice over 48 KB of shared memory for the anchor times
	is not reached here), continued — identical to a ring that is large enough; the CUDA graph of the attempt batch is rebuilt"""
	g = large_graph(20_000)
	blind_to = float(np.max(g["edge_tau"]))
	times = blind_to + np.array([1.0, 2.0])
	small = run_network(g, times, blind_to, ring_capacity=16)
	large = run_network(g, times, blind_to, ring_capacity=512)
	assert np.array_equal(small[0], large[0]) and np.array_equal(small[1], large[1]) and small[2] == large[2]
