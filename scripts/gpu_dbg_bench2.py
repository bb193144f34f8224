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
stream = torch.cuda.current_stream()
if "stream" in sys.argv: E.set_stream(stream.cuda_stream)
k = 0
def go(n):
    global k
    E.integrate_samples(30.0+10.0*np.arange(k+1, k+n+1), fetch=False); k += n
go(3)
if "sum" in sys.argv: print(E.sum_counters())
torch.cuda.synchronize(); E.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter()
if "stream" in sys.argv: ev0.record(stream)
go(10); w1 = time.perf_counter(); go(10); w2 = time.perf_counter()
if "stream" in sys.argv: ev1.record(stream)
E.synchronize(); torch.cuda.synchronize(); w3 = time.perf_counter()
print(sys.argv[1:], f"calls {1e3*(w1-w0):.2f} {1e3*(w2-w1):.2f} ms, total wall {1e3*(w3-w0):.2f} ms", (ev0.elapsed_time(ev1) if "stream" in sys.argv else None), flush=True)
