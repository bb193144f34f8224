"""
Multi-GPU plumbing for ensembles: one process per GPU, members sharded contiguously, no
communication while integrating; results are gathered at the end. The reference has no
counterpart (it integrates one trajectory per object and has no multi-process code).

torch.distributed is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""

import numpy as np


def shard(n_members, rank, world_size):
	"""Contiguous block of members [start, stop) that process `rank` of `world_size` integrates."""
	base, extra = divmod(n_members, world_size)
	start = rank*base + min(rank, extra)
	return start, start + base + (1 if rank < extra else 0)


def gather_members(local, n_members=None):
	"""
	Concatenate the per-rank arrays (leading axis = members of this rank's shard) in rank order on every
	rank. With an uninitialised process group (single process) the array is returned unchanged.
	"""
	import torch
	import torch.distributed as dist
	local = np.ascontiguousarray(local)
	if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
		return local
	world, rank = dist.get_world_size(), dist.get_rank()
	device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
	counts = torch.zeros(world, dtype=torch.int64, device=device)
	counts[rank] = local.shape[0]
	dist.all_reduce(counts)
	counts = counts.cpu().numpy()
	if n_members is not None and counts.sum() != n_members:
		raise ValueError(f"shards hold {counts.sum()} members, expected {n_members}")
	longest = int(counts.max())
	padded = np.zeros((longest,)+local.shape[1:], dtype=local.dtype)
	padded[:local.shape[0]] = local
	mine = torch.from_numpy(padded).to(device)
	pieces = [torch.empty_like(mine) for _ in range(world)]
	dist.all_gather(pieces, mine)
	return np.concatenate([p.cpu().numpy()[:counts[r]] for r, p in enumerate(pieces)], axis=0)
