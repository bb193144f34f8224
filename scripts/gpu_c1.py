"""C1 (BASELINE.json configs[0]): examples/mackey_glass.py — ONE trajectory, constant past, step_on_discontinuities, 1000 samples of 10 time
units — through the public API on the GPU, and the same protocol on the reference's own C core with its Python control loop on one host core."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
from jitcdde_b200 import jitcdde, y, t

f = [0.25*y(0, t-15)/(1+y(0, t-15)**10) - 0.1*y(0)]
DDE = jitcdde(f, verbose=False)
DDE.constant_past([1.0])
DDE.step_on_discontinuities()
times = DDE.t + np.arange(10, 10010, 10.0)
DDE.integrate(times[0])
t0 = time.perf_counter()
data = [DDE.integrate(T) for T in times[1:]]
el = time.perf_counter()-t0
acc = int(DDE.counters()[0][0])
print(f"GPU, N = 1: {len(times)-1} samples, {acc} accepted steps in {el:.3f} s → {acc/el:.3e} steps/s ({el/(len(times)-1)*1e6:.0f} µs per integrate() call); y(T) = {data[-1][0]:.12f}", flush=True)

try:
	from oracle import build_ref, controller, oracle
	from tests import systems
	prog = systems.program(systems.mackey_glass(control=False))
	name = "ref_mg_c1"
	build_ref.build_ref(prog, name, flavour="literal", flags=build_ref.FAST_FLAGS)
	core = build_ref.load_ref(name).dde_integrator([(-1.0, np.ones(1), np.zeros(1)), (0.0, np.ones(1), np.zeros(1))])
	C = controller.Controller(core, 15.0)
	C.step_on_discontinuities(controller.discontinuity_steps([15.0]))
	rt = C.t + np.arange(10, 10010, 10.0)
	C.integrate(rt[0])
	t0 = time.perf_counter()
	ref = [C.integrate(T) for T in rt[1:]]
	el = time.perf_counter()-t0
	print(f"reference C core + Python control loop, one host core: {C.accepted} accepted steps in {el:.3f} s → {C.accepted/el:.3e} steps/s; y(T) = {ref[-1][0]:.12f}", flush=True)
except Exception as e:
	print("reference arm unavailable:", e)
