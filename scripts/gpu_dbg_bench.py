import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jitcdde_b200 import jitcdde
from tests import systems
members = 1250000
pars = systems.mg_grid(members); system = systems.mackey_glass()
DDE = jitcdde(system["f"], control_pars=system["control_pars"], max_delay=25.0, n_members=members, ring_capacity=128, verbose=False)
DDE.constant_past([1.0]); DDE.set_parameters(pars); DDE.delays = pars[:, 1:2]; DDE.max_delay = 25.0
DDE.step_on_discontinuities()
E = DDE._engine
if "stream" in sys.argv: E.set_stream(torch.cuda.current_stream().cuda_stream)
k = 0
for n in (3, 10, 10, 10, 10, 5, 10):
    torch.cuda.synchronize(); E.synchronize(); w0 = time.perf_counter()
    E.integrate_samples(30.0+10.0*np.arange(k+1, k+n+1), fetch=False); k += n
    w1 = time.perf_counter(); E.synchronize(); torch.cuda.synchronize(); w2 = time.perf_counter()
    print(f"{n} samples: call {1e3*(w1-w0):.2f} ms, done {1e3*(w2-w0):.2f} ms -> {1e3*(w2-w0)/n:.3f} ms/sample", flush=True)
