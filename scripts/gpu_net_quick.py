"""network leg only, one GPU: python scripts/gpu_net_quick.py [nodes]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.simplefilter("ignore")
import bench_configs
r = bench_configs.network_leg(6538.9, 0, 0, 1, nodes=int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)
print(os.environ.get("JITCDDE_B200_NET_FLAGS", ""), os.environ.get("JITCDDE_B200_NET_PERSISTENT", ""), {k: r.get(k) for k in ("value", "ms_per_attempt", "attempts", "gpu_launches", "checksum", "error")}, r.get("roofline", {}).get("frac"))
