"""GPU tests of the network engine: against the symbolic path of this package (itself pinned to the reference core)
on a small Kuramoto network, and determinism / partition-independence properties at a size where the symbolic path is
impossible (10^5 nodes)."""

import numpy as np
import pytest
import sympy

from tests.test_network import kuramoto_graph, run_network, run_symbolic

pytestmark = pytest.mark.gpu


def large_graph(n, k=10, seed=42):
	"""BASELINE.json configs[3] shape: fixed in-degree k random graph, tau ~ U(pi/5, 2pi), omega = 1,
	coupling strength of examples/kuramoto_network.py (c·q = 2.1) spread over the k inputs."""
	rng = np.random.default_rng(seed)
	dst = np.repeat(np.arange(n, dtype=np.int64), k)
	src = rng.integers(0, n, size=n*k)
	tau = rng.uniform(np.pi/5, 2*np.pi, size=n*k)
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


def test_large_network_properties():
	g = large_graph(100_000)
	blind_to = float(np.max(g["edge_tau"]))
	times = blind_to + np.array([1.0, 2.0])
	runs = [run_network(g, times, blind_to) for _ in range(2)]
	assert np.array_equal(runs[0][1], runs[1][1]) and runs[0][2] == runs[1][2]      # deterministic (max-reduction is order-free)
	blind, samples, (accepted, rejected, attempts), t_end = runs[0]
	assert np.all(np.isfinite(samples)) and accepted > 60 and attempts == accepted+rejected
	# phases advance roughly with omega = 1 plus bounded coupling
	drift = (samples[-1]-g["initial"])/times[-1]
	assert np.all(np.abs(drift-1.0) < 2.2)
