"""
The N > 1 path without GPUs: two processes (gloo, world_size 2) each integrate their contiguous shard of a
Mackey–Glass ensemble on the host twin of the engine and gather the results; the concatenation must equal
the single-process run of the whole ensemble bit for bit (members are independent: no data-path collective).
"""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from tests import systems

N = 48
OFFSETS = [0.0, 10.0, 20.0]


def _run(pars):
	from jitcdde_b200 import jitcdde
	system = systems.mackey_glass()
	DDE = jitcdde(system["f"], control_pars=system["control_pars"], max_delay=25.0, n_members=len(pars), verbose=False)
	DDE.constant_past([1.0])
	DDE.set_parameters(pars)
	DDE.delays = pars[:, 1:2]
	DDE.max_delay = 25.0
	DDE.step_on_discontinuities()
	t0 = DDE.t
	samples = np.stack([DDE.integrate(t0+o) for o in OFFSETS], axis=1)     # (members, samples, n)
	accepted, rejected, _ = DDE.counters()
	return samples, accepted


def _worker(rank, world, port, queue):
	import warnings
	warnings.simplefilter("ignore")
	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
	import torch.distributed as dist
	from jitcdde_b200 import gather_members, shard
	from tests.twin import use_twin
	dist.init_process_group("gloo", rank=rank, world_size=world)
	start, stop = shard(N, rank, world)
	with use_twin():
		samples, accepted = _run(systems.mg_grid(N)[start:stop])
	all_samples = gather_members(samples, N)
	all_accepted = gather_members(accepted, N)
	total = np.array([accepted.sum()], dtype=np.int64)
	import torch
	t = torch.from_numpy(total)
	dist.all_reduce(t)
	if rank == 0:
		queue.put((all_samples, all_accepted, int(t.item())))
	dist.barrier()
	dist.destroy_process_group()


def test_shard_covers_all_members():
	from jitcdde_b200 import shard
	for n, world in ((10, 3), (1250000, 8), (7, 8)):
		pieces = [shard(n, r, world) for r in range(world)]
		assert pieces[0][0] == 0 and pieces[-1][1] == n
		assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
		assert max(b-a for a, b in pieces)-min(b-a for a, b in pieces) <= 1


def test_two_ranks_equal_one(twin):
	sock = socket.socket()
	sock.bind(("127.0.0.1", 0))
	port = sock.getsockname()[1]
	sock.close()
	ctx = mp.get_context("spawn")
	queue = ctx.Queue()
	procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
	for p in procs:
		p.start()
	all_samples, all_accepted, total = queue.get(timeout=300)
	for p in procs:
		p.join(timeout=60)
		assert p.exitcode == 0
	samples, accepted = _run(systems.mg_grid(N))
	assert np.array_equal(all_samples, samples)
	assert np.array_equal(all_accepted, accepted)
	assert total == accepted.sum()
