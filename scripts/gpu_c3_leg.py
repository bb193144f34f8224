"""C3 leg of bench.py alone: python scripts/gpu_c3_leg.py [members]"""
import os, sys, json, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import bench_configs
r = bench_configs.leg_c3(6538.9, int(sys.argv[1]) if len(sys.argv) > 1 else 100000)
print({k: r.get(k) for k in ("value", "ms_per_sample", "attempts_per_accepted")}, r["roofline"]["frac"])
