"""Throughput of the network engine (BASELINE configs[3] shape) on one GPU, or on several under torchrun."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np, sympy
from tests.test_gpu_network import large_graph

world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from jitcdde_b200 import jitcdde_network

for n in [int(x) for x in (sys.argv[1:] or ["100000", "1000000"])]:
    g = large_graph(n)
    net = jitcdde_network(n, local=lambda y, t, w: w, coupling=lambda yd, y: sympy.sin(yd-y), edges=(g["src"], g["dst"], g["edge_tau"]),
        weight=g["weight"], node_parameters=[g["omega"]], verbose=False, device=local, ring_capacity=512)
    net.set_integration_parameters(rtol=0, atol=1e-5)
    net.constant_past(g["initial"], time=0.0)
    t0 = time.perf_counter(); net.integrate_blindly(float(np.max(g["edge_tau"])), 0.1); t1 = time.perf_counter()
    a0 = net.counters()
    out = net.integrate(net.t+10.0); t2 = time.perf_counter()
    a1 = net.counters()
    if rank == 0:
        steps = a1[0]-a0[0]; attempts = a1[2]-a0[2]
        print(f"n={n} ranks={world}: blind {a0[0]} steps in {t1-t0:.3f}s ({a0[0]/(t1-t0):.1f} steps/s); adaptive {steps} accepted / {attempts} attempts in {t2-t1:.3f}s "
              f"→ {steps/(t2-t1):.1f} accepted steps/s, {steps*n/(t2-t1):.3e} node-steps/s, {attempts*n*10*3/(t2-t1):.3e} delayed lookups/s; checksum {float(np.sum(out)):.12e}", flush=True)
if world > 1:
    dist.destroy_process_group()
