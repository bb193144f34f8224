"""network leg of bench.py alone, under torchrun: python -m torch.distributed.run --nproc-per-node N scripts/gpu_net_scale.py [nodes] [exchange]"""
import os, sys, json, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import torch
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
	import torch.distributed as dist
	dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if len(sys.argv) > 2:
	os.environ["JITCDDE_B200_NET_EXCHANGE"] = sys.argv[2]
import bench_configs
r = bench_configs.network_leg(6538.9, local, rank, world, nodes=int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, span=2.0)
if rank == 0:
	print(json.dumps({k: r.get(k) for k in ("n_gpus", "value", "ms_per_attempt", "accepted", "attempts", "gpu_launches", "checksum", "exchange", "error")}), flush=True)
if world > 1:
	dist.destroy_process_group()
