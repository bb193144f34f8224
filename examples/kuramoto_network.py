"""
Delay-coupled Kuramoto oscillators on a random network — the system of the reference's examples/kuramoto_network.py — in the
structured form of jitcdde_network: a local term, a coupling function (both symbolic, compiled into the kernel) and edge arrays
with one delay per edge. The reference writes one symbolic expression per node, which stops being practical beyond a few
hundred nodes; this runs 10^6 nodes and 10^7 edges on one B200 and is partitioned over the GPUs of a node when launched with
torchrun (one process per GPU; the new anchor is exchanged from inside the step kernel).

    python examples/kuramoto_network.py [nodes]
    python -m torch.distributed.run --nproc-per-node 8 examples/kuramoto_network.py 1000000
"""
import os
import sys

import numpy as np
import sympy

from jitcdde_b200 import jitcdde_network

nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
in_degree, coupling_strength = 10, 42*0.05

if "RANK" in os.environ:                 # under torchrun: one rank per GPU, nodes in contiguous blocks
	import torch
	import torch.distributed as dist
	torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
	dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
device = int(os.environ.get("LOCAL_RANK", 0))

rng = np.random.default_rng(42)
destination = np.repeat(np.arange(nodes), in_degree)
source = rng.integers(0, nodes, nodes*in_degree)
delay = rng.uniform(np.pi/5, 2*np.pi, nodes*in_degree)
omega = rng.normal(1.0, 0.05, nodes)

net = jitcdde_network(nodes,
	local=lambda y, t, w: w,                              # dy_i/dt = ω_i + …
	coupling=lambda y_delayed, y: sympy.sin(y_delayed-y),  # … Σ_j c/k · sin(y_j(t − τ_ij) − y_i)
	edges=(source, destination, delay), weight=coupling_strength/in_degree, node_parameters=[omega], device=device, verbose=False)
net.set_integration_parameters(rtol=0, atol=1e-5)
net.constant_past(rng.uniform(0, 2*np.pi, nodes), time=0.0)
net.integrate_blindly(float(delay.max()), 0.1)           # as the reference's example: past the initial discontinuities with fixed steps

for time in net.t+np.arange(10, 60, 10):
	phases = net.integrate(time)
	order = abs(np.exp(1j*phases).mean())
	if int(os.environ.get("RANK", 0)) == 0:
		print(f"t = {time:7.2f}   order parameter {order:.4f}   accepted steps {net.counters()[0]}")
