"""A/B timing of kernel build variants on the bench workload (device-resident loop).
Usage: python scripts/gpu_variants.py "W=2,B=32,MB=32" "W=2,B=32,MB=32,S=4" ...
  W window mode, B block size, MB resident blocks per SM, R maxrregcount, D extra defines (a+b), S samples per launch, CAP ring capacity"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jitcdde_b200 import _build, _capi

from tests import systems
DEFAULTS = dict(atol=1e-10, rtol=1e-5, first_step=1.0, min_step=1e-10, max_step=10.0, decrease_threshold=1.1, increase_threshold=0.5, safety_factor=0.9, max_factor=5.0, min_factor=0.2, pws_factor=3, pws_atol=0.0, pws_rtol=1e-5, pws_max_iterations=10, pws_base_increase_chance=0.1)

N = int(os.environ.get("MEMBERS", 1250000)); STEPS = int(os.environ.get("STEPS", 24)); WARM = int(os.environ.get("WARM", 12))
prog = systems.program(systems.mackey_glass()); pars = systems.mg_grid(N)
ref = None
for spec in sys.argv[1:]:
    opts = dict(kv.split("=") for kv in spec.split(","))
    os.environ["JITCDDE_B200_WINDOW"] = opts.get("W", "2")
    flags = [f"-DJD_MIN_BLOCKS={opts.get('MB','32')}"] + ([f"-maxrregcount={opts['R']}"] if "R" in opts else []) + [f"-D{d}" for d in opts.get("D","").split("+") if d]
    block = int(opts.get("B", 32)); cap = int(opts.get("CAP", 128)); S = int(opts.get("S", 1))
    try:
        image = _build.build_image(prog, mode=opts.get("mode", "fast"), block=block, extra_flags=tuple(flags))
    except Exception as e:
        print(f"{spec:34s} build failed: {str(e)[:300]}", flush=True); continue
    E = _capi.Engine(1, 1, prog.n_sites, 2, N, ring_capacity=cap); E.load_image(image)
    E.set_past(np.array([-1.0, 0.0]), np.ones((2, 1)), np.zeros((2, 1))); E.set_parameters(pars)
    E.set_integration_parameters(25.0, **DEFAULTS)
    E.step_on_discontinuities(pars[:, 1:2].copy())
    T0 = 30.0
    def run(k0, count):
        k = k0
        while k < k0+count:
            s = min(S, k0+count-k)
            if S == 1: E.integrate(T0+10.0*(k+1), fetch=False)
            else: E.integrate_samples(T0+10.0*np.arange(k+1, k+s+1), fetch=False)
            k += s
        return k
    k = run(0, WARM)
    a0 = E.sum_counters()[0]; E.synchronize(); w0 = time.perf_counter()
    k = run(k, STEPS)
    E.synchronize(); el = time.perf_counter()-w0
    a1 = E.sum_counters()[0]
    out, st = E.integrate(T0+10.0*k)
    chk = float(out.sum())
    if ref is None: ref = chk
    print(f"{spec:34s} {1e3*el/STEPS:8.3f} ms/step  {(a1-a0)/el:.3e} steps/s  status={np.unique(st)} checksum_rel_dev={abs(chk-ref)/abs(ref):.2e}", flush=True)
    E.close()
