"""Small runs of the kernels with shared-memory staging, lane groups and clusters, to be run under compute-sanitizer
(memcheck / racecheck / synccheck): python scripts/gpu_sanitize.py [c3|c5b|mg|net]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
from tests import scenarios

which = sys.argv[1:] or ["c3", "c5b", "mg"]
if "c3" in which:
	r = scenarios.product_c3(14, ring_capacity=1024)      # lane-group stepper with staging, cluster Gram–Schmidt (2 blocks per member)
	print("c3", int(np.sum(r["accepted"])), np.asarray(r["lyaps"]).mean(axis=(0, 1)))
if "c5b" in which:
	r = scenarios.product_c5b(40)                         # pointer window in shared memory
	print("c5b", int(np.sum(r["accepted"])))
if "mg" in which:
	g = scenarios.golden("mg_ensemble")
	r = scenarios.mg_ensemble(g)                          # per-thread window, by value
	print("mg", {k: np.asarray(v).shape for k, v in r.items()})
