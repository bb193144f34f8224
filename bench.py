#!/usr/bin/env python
"""
bench.py — accepted trajectory-steps/s of the Mackey–Glass (β, τ) parameter-sweep ensemble
(BASELINE.json configs[1]) on N B200s, beside the reference's C core on the host cores.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA engine
    python bench.py --impl reference --steps K --warmup W     # the reference's own C core on the CPU
    (N > 1: python -m torch.distributed.run --nproc-per-node N … bench.py --gpus N …)

Workload: dy/dt = β·y(t−τ)/(1+y(t−τ)^10) − 0.1·y, control parameters (β, τ) on the
deterministic grid β_i = 0.15+0.20·i/2499, τ_j = 10+15·j/3999 (10^7 members on 8 GPUs ⇒ 1.25·10^6
members per GPU, weak scaling: contiguous member ranges per rank, no collective in the data path),
past = constant 1, default integration parameters of the reference, per-member
step_on_discontinuities (target τ_j), then the sampling loop `integrate(30 + 10·k)`.
One bench STEP = one `integrate` call advancing every member by 10 time units (≈ 23 accepted
steps per member). `value` = accepted steps of all members and ranks ÷ device time (CUDA events on the
launching stream, max over ranks) with all inputs resident in HBM; `e2e` = the same loop through the
public Python API with host buffers: per step the target time goes host→device and the states and
per-member status come back device→host into pinned memory.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

MEMBERS_PER_GPU = 1_250_000
ADVANCE = 10.0
MAX_DELAY = 25.0
BYTES_PER_STEP = 266      # algorithmic bytes per accepted step, SURVEY.md §8d / DESIGN.md
FLOPS_PER_STEP = 165      # algorithmic FP64 flops per accepted step, same source
METRIC = "accepted trajectory-steps/s, Mackey-Glass ensemble"
UNIT = "trajectory-steps/s"


def config(members_per_gpu, world):
	return {
		"workload": "mackey_glass_beta_tau_sweep (BASELINE.json configs[1])",
		"members_per_gpu": members_per_gpu,
		"members_total": members_per_gpu*world,
		"sharding": "members dealt round-robin to the ranks, no data-path collective",
		"step": f"integrate(): advance every member by {ADVANCE:g} time units",
		"integration_parameters": "reference defaults (atol 1e-10, rtol 1e-5)",
		"discontinuities": "step_on_discontinuities, target tau_j per member",
		"ring_capacity": 128,
		"l2": "inputs larger than L2 (anchor rings 5.1 GB per GPU)",
		"build": "production (FMA contraction); parity of this build: tests/test_gpu_parity.py",
	}


def measured_peaks():
	path = os.path.join(ROOT, "MEASURED_PEAKS.json")
	if os.path.exists(path):
		with open(path) as f:
			return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
	return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
	"""Samples SM clock and throttle reasons through NVML (≈0.2 ms per sample) in a background thread;
	the summary covers the samples taken between mark_start() and mark_end(), i.e. DURING the timed region."""

	def __init__(self, index):
		self.samples = []          # (time, sm_mhz, reasons mask)
		self.max_mhz = None
		self.window = [None, None]
		self._stop = threading.Event()
		try:
			import pynvml
			pynvml.nvmlInit()
			self.nv = pynvml
			self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
			self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
			self.thread = threading.Thread(target=self._run, daemon=True)
			self.thread.start()
		except Exception:
			self.nv = None

	def _run(self):
		nv = self.nv
		get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
		while not self._stop.is_set():
			try:
				self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM), get_reasons(self.handle)))
			except Exception:
				pass
			time.sleep(0.0005)

	def sample_now(self):
		"""one sample taken by the calling thread (the background thread may be starved for a short timed region)"""
		if not self.nv:
			return
		nv = self.nv
		get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
		try:
			self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM), get_reasons(self.handle)))
		except Exception:
			pass

	def mark_start(self):
		self.window[0] = time.perf_counter()

	def mark_end(self):
		self.window[1] = time.perf_counter()

	def stop(self):
		self._stop.set()
		if self.nv:
			self.thread.join()

	def summary(self):
		nv = self.nv
		if not nv:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
		bits = {
			"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
			"hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
			"sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
			"sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
		}
		inside = [s for s in self.samples if self.window[0] <= s[0] <= self.window[1]]
		mask = 0
		for s in inside:
			mask |= s[2]
		return {
			"sm_mhz": float(np.median([s[1] for s in inside])) if inside else None,
			"sm_max_mhz": self.max_mhz,
			"reasons": sorted(name for name, bit in bits.items() if mask & bit),
			"samples": len(inside),
		}


# --------------------------------------------------------------------------------------- CPU reference
_worker = {}


def _ref_worker_init(kind, pars_chunks):
	import warnings
	warnings.simplefilter("ignore")
	from oracle import build_ref, controller, oracle as oracle_mod
	from tests import systems
	_worker["kind"] = kind
	if kind == "reference":
		module = build_ref.load_ref("ref_mg_bench")
		_worker["make"] = lambda p: controller.Controller(_with_pars(module.dde_integrator(_PAST()), p), MAX_DELAY)
	_worker["pars_chunks"] = pars_chunks


def _PAST():
	return [(-1.0, np.array([1.0]), np.array([0.0])), (0.0, np.array([1.0]), np.array([0.0]))]


def _with_pars(core, p):
	core.set_parameters(float(p[0]), float(p[1]))
	return core


def _ref_worker_setup(chunk_index):
	from oracle import controller
	members = []
	for p in _worker["pars_chunks"][chunk_index]:
		C = _worker["make"](p)
		C.step_on_discontinuities(controller.discontinuity_steps([p[1]]))
		members.append([C, C.t])
	_worker["members"] = members
	return len(members)


def _ref_worker_step(k):
	t0 = time.perf_counter()
	before = sum(m[0].accepted for m in _worker["members"])
	for C, start in _worker["members"]:
		C.integrate(start + ADVANCE*k)
	return sum(m[0].accepted for m in _worker["members"])-before, time.perf_counter()-t0


def cpu_reference(steps, warmup, members_per_core, cores=None):
	"""
	The reference's own C core (oracle/_ref/ref_mg_bench, the unmodified jitced_template.c rendered for
	this system, reference-like flags) driven by the reference's Python control loop
	(oracle/controller.py), one process per core, members dealt evenly. Returns
	(steps/s, ms per step, cores, description). Falls back to the plain-C port of the oracle if the
	reference module was not prebuilt.
	"""
	import multiprocessing as mp
	from oracle import build_ref
	from tests import systems
	cores = cores or os.cpu_count()
	if not build_ref.have_ref("ref_mg_bench"):
		return cpu_port(steps, warmup, members_per_core*cores, cores) + ("port",)
	pars = systems.mg_grid(members_per_core*cores)
	chunks = [pars[c::cores] for c in range(cores)]
	ctx = mp.get_context("fork")
	# one single-worker pool per core keeps each member set pinned to its process
	pools = [ctx.Pool(1, initializer=_ref_worker_init, initargs=("reference", chunks)) for _ in range(cores)]
	try:
		for r in [pool.apply_async(_ref_worker_setup, (c,)) for c, pool in enumerate(pools)]:
			r.get()
		k = 0
		for _ in range(warmup):
			k += 1
			for r in [pool.apply_async(_ref_worker_step, (k,)) for pool in pools]:
				r.get()
		total, t0 = 0, time.perf_counter()
		for _ in range(steps):
			k += 1
			for r in [pool.apply_async(_ref_worker_step, (k,)) for pool in pools]:
				total += r.get()[0]
		elapsed = time.perf_counter()-t0
	finally:
		for pool in pools:
			pool.terminate()
	sample = f"{members_per_core*cores} members of the same grid ({members_per_core} per core), {steps} steps of {ADVANCE:g} time units after {warmup} warm-up steps"
	return total/elapsed, 1e3*elapsed/steps, cores, sample, "reference"


def cpu_port(steps, warmup, members, cores):
	"""The plain-C oracle port (controller in C as well, OpenMP over members) — the strongest CPU implementation here."""
	from oracle import oracle
	from tests import systems
	os.environ.setdefault("OMP_NUM_THREADS", str(cores))
	lib = oracle.build(systems.program(systems.mackey_glass()), flags=oracle.FAST_FLAGS, openmp=True, tag="_fast")
	pars = systems.mg_grid(members)
	past = (np.array([-1.0, 0.0]), np.ones((2, 1)), np.zeros((2, 1)))
	offsets_w = ADVANCE*np.arange(1, warmup+1)
	offsets = ADVANCE*np.arange(1, warmup+steps+1)
	t0 = time.perf_counter()
	rw = oracle.run_ensemble(lib, past, pars, MAX_DELAY, 1, pars[:, 1:2].copy(), offsets_w, keep_output=False)
	t1 = time.perf_counter()
	rf = oracle.run_ensemble(lib, past, pars, MAX_DELAY, 1, pars[:, 1:2].copy(), offsets, keep_output=False)
	t2 = time.perf_counter()
	accepted = int(rf["counters"][:, 0].sum()-rw["counters"][:, 0].sum())
	elapsed = (t2-t1)-(t1-t0)
	sample = f"{members} members of the same grid, {steps} steps of {ADVANCE:g} time units (run with minus run without them)"
	return accepted/elapsed, 1e3*elapsed/steps, cores, sample


# --------------------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
	import torch
	import torch.distributed as dist

	from jitcdde_b200 import jitcdde, pinned_empty
	from tests import systems

	rank = int(os.environ.get("RANK", 0))
	world = int(os.environ.get("WORLD_SIZE", 1))
	local_rank = int(os.environ.get("LOCAL_RANK", 0))
	if world != args.gpus:
		raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
	if not torch.cuda.is_available():
		raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
	torch.cuda.set_device(local_rank)
	if world > 1:
		dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

	members = args.members
	# Weak scaling: every rank integrates `members` members. Members are dealt to the ranks round-robin
	# (rank r takes r, r+world, … of the global grid), so every GPU sees the same mix of (β, τ) and hence of
	# step sizes; contiguous blocks would give the ranks with larger β systematically more steps.
	pars = systems.mg_grid(members*world)[rank::world]
	system = systems.mackey_glass()

	def fresh():
		DDE = jitcdde(system["f"], control_pars=system["control_pars"], max_delay=MAX_DELAY, n_members=members,
			device=local_rank, ring_capacity=128, verbose=False)
		DDE.constant_past([1.0])
		DDE.set_parameters(pars)
		DDE.delays = pars[:, 1:2]
		DDE.max_delay = MAX_DELAY
		DDE.step_on_discontinuities()
		return DDE

	def barrier():
		torch.cuda.synchronize()
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	def reduce_max(x):
		if world == 1:
			return x
		t = torch.tensor([x], dtype=torch.float64, device="cuda")
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		return float(t.item())

	def reduce_sum(x):
		if world == 1:
			return x
		t = torch.tensor([x], dtype=torch.float64, device="cuda")
		dist.all_reduce(t, op=dist.ReduceOp.SUM)
		return float(t.item())

	# ---------------- device-resident loop: `value`
	# After the discontinuity stepping member j stands at t = tau_j (10…25). All members are then sampled
	# on the common time grid T_k = 30 + 10 k, as a user of the reference samples a trajectory
	# (examples/mackey_glass.py:76-78). The only per-step input is the scalar T_k (8 bytes); the ensemble
	# state (anchor rings, controller state) is resident in HBM.
	setup_start = time.perf_counter()
	DDE = fresh()
	torch.cuda.synchronize()
	setup_seconds = time.perf_counter()-setup_start      # reported separately, outside every timed region
	engine = DDE._engine
	stream = torch.cuda.current_stream()
	engine.set_stream(stream.cuda_stream)
	T0 = 30.0
	k = 0
	clocks = ClockSampler(local_rank)      # started before the warm-up: NVML's first queries are slow
	for _ in range(args.warmup):
		k += 1
		engine.integrate(T0+ADVANCE*k, fetch=False)
	accepted_before = engine.sum_counters()[0]
	launches_before = engine.launch_count
	ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	barrier()
	clocks.mark_start()
	ev0.record(stream)
	for _ in range(args.steps):
		k += 1
		engine.integrate(T0+ADVANCE*k, fetch=False)
	ev1.record(stream)
	clocks.sample_now()      # the host is ahead of the GPU here: the launches above are still running
	barrier()
	clocks.mark_end()
	clocks.stop()
	ms = reduce_max(ev0.elapsed_time(ev1))
	launches = engine.launch_count-launches_before
	accepted = engine.sum_counters()[0]-accepted_before
	total_accepted = reduce_sum(accepted)
	status = engine.get_status()
	if status.any():
		raise SystemExit(f"members failed: {np.unique(status)}")
	value = total_accepted/(ms*1e-3)

	# ---------------- end to end through the public API with host buffers: `e2e`
	# per step: the target time goes host→device, the kernel runs, the states (N×8 B) come back device→host
	# into pinned memory; the number of failed members (8 B) is read once at the end (the status array follows only
	# if it is non-zero); all inside the timed region
	# Sampling is pipelined as the public API offers it (integrate(..., wait=False) / wait_result / synchronize):
	# while the kernel of sample k runs, the states of sample k-1 travel to the host and are read there.
	outs = [pinned_empty((members, 1)), pinned_empty((members, 1))]
	for i in range(2):
		k += 1
		DDE.integrate(T0+ADVANCE*k, out=outs[i])
	accepted_before = engine.sum_counters()[0]
	barrier()
	w0 = time.perf_counter()
	ev0.record(stream)
	checksum = 0.0
	for i in range(args.steps):
		k += 1
		DDE.integrate(T0+ADVANCE*k, out=outs[i & 1], wait=False)   # queue: H2D target, kernel, D2H states
		if i:
			DDE.wait_result(1)                                       # the previous sample has arrived …
			checksum += float(outs[(i-1) & 1][0, 0])                 # … and is read on the host
	DDE.synchronize()                                                # last sample + failed-member count (8 B) + status check
	checksum += float(outs[(args.steps-1) & 1][0, 0])
	ev1.record(stream)
	barrier()
	e2e_ms = reduce_max(max(ev0.elapsed_time(ev1), 1e3*(time.perf_counter()-w0)))
	e2e_accepted = reduce_sum(engine.sum_counters()[0]-accepted_before)
	e2e_value = e2e_accepted/(e2e_ms*1e-3)

	peak, peak_source = measured_peaks()
	from jitcdde_b200._capi import measure_fp64_peak
	fp64_peak = measure_fp64_peak(local_rank)      # FLOP/s, DFMA microbenchmark on this GPU (not in MEASURED_PEAKS.json)
	# dominant kernel = jdde_k_integrate, one launch per step: algorithmic bytes of this rank's launches ÷ their duration
	achieved = (total_accepted/world*BYTES_PER_STEP)/(ms*1e-3)/1e9
	line = {
		"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
		"ms_per_step": ms/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
		"dtype": "f64", "data": "synthetic", "config": config(members, world),
		"e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": int(members*8),
			"ms_per_step": e2e_ms/args.steps, "api": "DDE.integrate(T, out=pinned, wait=False) + wait_result(1) per step, synchronize() at the end"},
		"gpu_launches": int(launches),
		"clocks": clocks.summary(),
		"roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak, "traffic": TRAFFIC_PER_LAUNCH,
			"kernel": "jdde_k_integrate", "algorithmic_bytes_per_accepted_step": BYTES_PER_STEP, "peak_source": peak_source,
			"fp64_tflops_algorithmic": value/world*FLOPS_PER_STEP/1e12, "fp64_peak_tflops_measured": fp64_peak/1e12,
			"fp64_frac": value/world*FLOPS_PER_STEP/fp64_peak},
		"accepted_steps_per_member_per_step": total_accepted/(members*world*args.steps),
		"setup": {"seconds": setup_seconds, "what": "code generation, kernel image (cache hit or nvcc), engine creation, upload of past and parameters, step_on_discontinuities; not part of any timed region"},
	}
	if rank == 0 and world == 1 and not args.no_cpu:
		v, ms_cpu, cores, sample, kind = cpu_reference(args.cpu_steps, 1, args.cpu_members_per_core)
		line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
		pv, pms, pcores, psample = cpu_port(args.cpu_steps*4, 1, args.cpu_members_per_core*cores*8, cores)
		line["cpu_port"] = {"value": pv, "unit": UNIT, "cores": pcores, "kind": "port", "sample": psample,
			"note": "plain-C oracle incl. controller, OpenMP: no Python in the loop (the reference has Python in the loop)"}
	if rank == 0:
		print(json.dumps(line), flush=True)
	if world > 1:
		dist.destroy_process_group()


#: dram__bytes_read.sum + dram__bytes_write.sum of one jdde_k_integrate launch from the ncu capture under profiles/ (null until measured)
TRAFFIC_PER_LAUNCH = 3651599704   # 2.664 GB read + 0.988 GB written, profiles/ncu_full_r01_final_jdde_k_integrate.csv (launch #14 of the bench)


def reference_arm(args):
	rank = int(os.environ.get("RANK", 0))
	if rank != 0:
		return
	cores = os.cpu_count()
	value, ms, cores, sample, kind = cpu_reference(args.steps, args.warmup, args.cpu_members_per_core)
	line = {
		"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
		"ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": config(args.members, args.gpus),
		"cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
		"e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
	}
	print(json.dumps(line), flush=True)


def main():
	parser = argparse.ArgumentParser()
	parser.add_argument("--gpus", type=int, default=1)
	parser.add_argument("--steps", type=int, default=20)
	parser.add_argument("--warmup", type=int, default=3)
	parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
	parser.add_argument("--members", type=int, default=MEMBERS_PER_GPU, help="ensemble members per GPU")
	parser.add_argument("--cpu-members-per-core", type=int, default=512)
	parser.add_argument("--cpu-steps", type=int, default=10)
	parser.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
	args = parser.parse_args()
	if args.warmup < 3:
		args.warmup = 3
	if args.impl == "reference":
		reference_arm(args)
	else:
		gpu_arm(args)


if __name__ == "__main__":
	main()
