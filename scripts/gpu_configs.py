"""Throughput of the other BASELINE.json configurations (production build), for profiles/: C3, C5a, C5b."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
from tests import scenarios

def timed(name, fn, N):
    t0 = time.perf_counter(); r = fn(N); el = time.perf_counter()-t0
    acc = int(np.sum(r["accepted"])); rej = int(np.sum(r["rejected"]))
    print(f"{name:34s} N={N:8d} accepted={acc:12d} rejected={rej:10d} wall={el:8.3f}s  {acc/el:.3e} accepted steps/s (incl. JIT cache lookup, upload, all D2H)", flush=True)

which = sys.argv[1:] or ["c5a", "c5b", "c3"]
for w in which:
    if w == "c5a":
        timed("C5a state-dependent", scenarios.product_c5a, 1 << 12); timed("C5a state-dependent", scenarios.product_c5a, 1 << 20)
    if w == "c5b":
        timed("C5b neutral", scenarios.product_c5b, 1 << 12); timed("C5b neutral", scenarios.product_c5b, 1 << 20)
    if w == "c3":
        timed("C3 two Roesslers + 4 exponents", scenarios.product_c3, 1 << 10); timed("C3 two Roesslers + 4 exponents", scenarios.product_c3, 1 << 14)
