"""Where does the time of pipelined sampling go? MG sweep, 1.25e6 members: device-resident loop, queue-only loop, queue + per-sample wait."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
from jitcdde_b200 import jitcdde, pinned_empty
from tests import systems
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1250000
pars = systems.mg_grid(N)
s = systems.mackey_glass()
DDE = jitcdde(s["f"], control_pars=s["control_pars"], max_delay=25.0, n_members=N, verbose=False, ring_capacity=128)
DDE.constant_past([1.0]); DDE.set_parameters(pars); DDE.delays = pars[:, 1:2]; DDE.step_on_discontinuities()
E = DDE._engine
outs = [pinned_empty((N, 1)), pinned_empty((N, 1))]
k = 0; T0 = 30.0
def T():
	global k
	k += 1
	return T0+10*k
for _ in range(3): DDE.integrate(T(), out=outs[0])
K = 20
E.synchronize(); w = time.perf_counter()
for _ in range(K): E.integrate(T(), fetch=False)
E.synchronize(); print(f"device-resident: {(time.perf_counter()-w)/K*1e3:.3f} ms/step", flush=True)
w = time.perf_counter()
for i in range(K): DDE.integrate(T(), out=outs[i & 1], wait=False)
DDE.synchronize(); print(f"queue only:      {(time.perf_counter()-w)/K*1e3:.3f} ms/step", flush=True)
w = time.perf_counter()
for i in range(K):
	DDE.integrate(T(), out=outs[i & 1], wait=False)
	if i: DDE.wait_result(1)
DDE.synchronize(); print(f"queue + wait:    {(time.perf_counter()-w)/K*1e3:.3f} ms/step", flush=True)
w = time.perf_counter()
for i in range(K): DDE.integrate(T(), out=outs[0])
print(f"synchronous:     {(time.perf_counter()-w)/K*1e3:.3f} ms/step", flush=True)
w = time.perf_counter()
for i in range(K): E.synchronize(); 
print(f"empty synchronize: {(time.perf_counter()-w)/K*1e6:.1f} us", flush=True)
