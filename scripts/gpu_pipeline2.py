"""Does a concurrent D2H copy slow the integrate kernel down? Device-resident loop with and without unrelated 10 MB D2H copies on another stream."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np, torch
from jitcdde_b200 import jitcdde
from tests import systems
N = 1250000
pars = systems.mg_grid(N)
s = systems.mackey_glass()
DDE = jitcdde(s["f"], control_pars=s["control_pars"], max_delay=25.0, n_members=N, verbose=False, ring_capacity=128)
DDE.constant_past([1.0]); DDE.set_parameters(pars); DDE.delays = pars[:, 1:2]; DDE.step_on_discontinuities()
E = DDE._engine
k = 0
def T():
	global k
	k += 1
	return 30.0+10*k
for _ in range(3): E.integrate(T(), fetch=False)
E.synchronize()
src = torch.zeros(N, dtype=torch.float64, device="cuda"); dst = torch.empty(N, dtype=torch.float64).pin_memory()
side = torch.cuda.Stream()
K = 20
for copies in (0, 1, 4):
	E.synchronize(); torch.cuda.synchronize(); w = time.perf_counter()
	for _ in range(K):
		E.integrate(T(), fetch=False)
		with torch.cuda.stream(side):
			for _ in range(copies): dst.copy_(src, non_blocking=True)
	E.synchronize(); el = time.perf_counter()-w; torch.cuda.synchronize()
	print(f"{copies} unrelated 10 MB D2H copies per step: {el/K*1e3:.3f} ms/step", flush=True)
