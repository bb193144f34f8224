"""
`jitcdde_network` — one large delay-coupled network (10^4…10^6 nodes) as a structured problem:

    dy_i/dt = local(y_i, t, p_i…) + Σ_{edges e into i} weight_e · coupling(y_{src_e}(t − tau_e), y_i)

The reference expresses such systems through its symbolic API, one expression per node with literal delays
(/root/reference/examples/kuramoto_network.py:42-63); that path exists here too (`jitcdde`) and is what this
class is tested against on small networks, but it cannot be code-generated for 10^6 nodes (hard
part 7). Method names and semantics follow `jitcdde`: constant_past / add_past_point(s) / get_state,
set_integration_parameters, step_on_discontinuities / adjust_diff / integrate_blindly, integrate, jump, t, dt. Delays shorter
than the step are handled as by the reference (past-within-step repetitions, _jitcdde.py:838-861).

On one GPU an integrate call is a handful of launches of a persistent cooperative kernel (csrc/jdn_kernels.cu: jdn_k_run).
Multi-GPU: launched with one process per GPU (torch.distributed initialised), the nodes are partitioned into contiguous
blocks; every rank keeps the whole history, integrates its block and, in every attempted step, pushes its part of the new
anchor into the peers' memory over NVLink from inside the step kernel; the ranks merge the error norm and meet in a
device-side barrier, so that all take the same accept/reject decision (exchange="peer", the default; SURVEY.md §8e,
DESIGN.md §5). exchange="nccl" instead all-gathers the staged anchor slices and max-reduces the error norm between an
attempt kernel and a control kernel. Without torch.distributed the whole network runs on one GPU.
"""

from warnings import warn

import numpy as np

from . import _build, _capi, _codegen
from ._jitcdde import UnsuccessfulIntegration, _default_min_step

MODE_ADAPTIVE, MODE_BLIND, MODE_CONTINUE = 0, 1, 2


class _DeviceArray:
	"""minimal __cuda_array_interface__ carrier so that torch can wrap engine buffers without copying"""
	def __init__(self, pointer, shape, typestr):
		self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (pointer, False), "version": 3}


class jitcdde_network:
	def __init__(self, n, local, coupling, edges, node_parameters=(), weight=None, max_delay=None, device=0, ring_capacity=512, verbose=True, poll_every=8, exchange=None, verification=False):
		"""
		n: number of nodes. local(y, t, *p) and coupling(y_delayed, y): functions returning SymPy expressions.
		edges: (src, dst, tau[, weight]) arrays, one entry per directed edge src → dst with delay tau > 0.
		node_parameters: sequence of arrays of length n, passed to `local` in this order.
		ring_capacity: anchors kept (power of two); must exceed max_delay / smallest step (2·8·n bytes each).
		exchange (several GPUs, one process each): "peer" (default) — the step kernel stores the new anchor into the
		peers' memory over NVLink and the GPUs meet in a device-side barrier (csrc/jdn_kernels.cu) — or "nccl": all-gather
		and all-reduce between the attempt and the control kernel. Environment: JITCDDE_B200_NET_EXCHANGE.
		"""
		import os
		self.exchange = exchange or os.environ.get("JITCDDE_B200_NET_EXCHANGE", "peer")
		if self.exchange not in ("peer", "nccl"):
			raise ValueError("exchange must be 'peer' or 'nccl'")
		self.n = int(n)
		self.verbose = verbose
		self.device = device
		self.poll_every = poll_every
		src, dst, tau = (np.asarray(a) for a in edges[:3])
		w = np.asarray(edges[3], dtype=float) if len(edges) > 3 else (np.full(len(src), 1.0 if weight is None else float(weight)))
		if not (len(src) == len(dst) == len(tau) == len(w)):
			raise ValueError("edge arrays have different lengths")
		if len(tau) and not np.all(tau > 0):
			raise ValueError("all delays must be positive")
		order = np.argsort(dst, kind="stable")       # CSR by destination, input order kept within a node
		self._src, self._dst, self._tau, self._w = src[order].astype(np.int32), dst[order].astype(np.int64), np.asarray(tau, dtype=float)[order], w[order]
		self.min_delay = float(self._tau.min()) if len(tau) else np.inf
		self.max_delay = float(max_delay if max_delay is not None else (self._tau.max() if len(tau) else 0.0))
		self._node_parameters = [np.asarray(p, dtype=float) for p in node_parameters]
		self._program = _codegen.generate_network(local, coupling, len(self._node_parameters))
		self.verification = verification      # kernels without FMA contraction, bit-reproducible elementary functions (as for jitcdde)
		self._image = _build.build_net_image(self._program, mode=_build.VERIFY if verification else _build.FAST)
		self.ring_capacity = ring_capacity
		self.max_ring_capacity = 1 << 14      # the anchor times of the ring are staged in shared memory (128 KB)

		self.rank, self.world = 0, 1
		try:
			import torch.distributed as dist
			if dist.is_available() and dist.is_initialized():
				self.rank, self.world = dist.get_rank(), dist.get_world_size()
		except ImportError:
			pass
		if self.world > 1 and self.n % self.world:
			raise ValueError("the number of nodes must be a multiple of the number of ranks")
		block = self.n//self.world
		self.lo, self.hi = self.rank*block, (self.rank+1)*block

		self._past = []
		self._engine = None
		self._int_params = None
		self.integration_parameters_set = False

	# ------------------------------------------------------------------ past
	def add_past_point(self, time, state, derivative):
		self._past.append((float(time), np.array(state, dtype=float), np.array(derivative, dtype=float)))
		self._past.sort(key=lambda a: a[0])
		self._drop()

	def constant_past(self, state, time=0.0):
		state = np.array(state, dtype=float)
		self._past = [(time-1.0, state, np.zeros_like(state)), (float(time), state, np.zeros_like(state))]
		self._drop()

	def add_past_points(self, anchors):
		"""anchors: (time, state, derivative) triples, e.g. what get_state returned (reference: _jitcdde.py:298-309)"""
		for time, state, derivative in anchors:
			self._past.append((float(time), np.array(state, dtype=float), np.array(derivative, dtype=float)))
		self._past.sort(key=lambda a: a[0])
		self._drop()

	def get_state(self):
		"""as jitcdde.get_state (_jitcdde.py:357-371): all anchors the integration still needs, as (time, state, derivative)
		triples — a checkpoint: hand them to add_past_points of a new jitcdde_network (and its step size, `dt`, to first_step)."""
		self._initiate()
		times, states, diffs = self._engine.get_history()
		return [(float(times[k]), states[k].copy(), diffs[k].copy()) for k in range(len(times))]

	def _drop(self):
		if self._engine is not None:
			self._engine.close()
			self._engine = None

	# ------------------------------------------------------------------ parameters
	def set_integration_parameters(self, atol=1e-10, rtol=1e-5, first_step=1.0, min_step=_default_min_step, max_step=10.0,
			decrease_threshold=1.1, increase_threshold=0.5, safety_factor=0.9, max_factor=5.0, min_factor=0.2,
			pws_factor=3, pws_atol=0.0, pws_rtol=1e-5, pws_max_iterations=10, pws_base_increase_chance=0.1, pws_fuzzy_increase=False):
		"""as jitcdde.set_integration_parameters (_jitcdde.py:621-743). A delay shorter than the step (past within step) is
		handled as by the reference — the step is repeated until two successive results agree, or shrunk — on one GPU and with
		the fused peer exchange; the all-gather exchange (exchange="nccl") does not iterate and limits max_step to the
		smallest delay instead. pws_fuzzy_increase (a random draw per step-size increase) is not available for networks."""
		if pws_fuzzy_increase:
			raise NotImplementedError("pws_fuzzy_increase is not available for jitcdde_network; the deterministic credit rule (the reference's default) is used.")
		if self.world > 1 and self.exchange == "nccl" and max_step > self.min_delay:
			max_step = self.min_delay
			if self.verbose:
				print(f"max_step limited to the smallest delay ({self.min_delay:g}): the all-gather exchange needs delay >= step.")
		if first_step > max_step:
			first_step = max_step
		if min_step > first_step:
			min_step = first_step
		self._int_params = dict(atol=atol, rtol=rtol, first_step=first_step, min_step=min_step, max_step=max_step,
			decrease_threshold=decrease_threshold, increase_threshold=increase_threshold, safety_factor=safety_factor,
			max_factor=max_factor, min_factor=min_factor, pws_factor=pws_factor, pws_atol=pws_atol, pws_rtol=pws_rtol,
			pws_max_iterations=pws_max_iterations, pws_base_increase_chance=pws_base_increase_chance)
		self.max_step = max_step
		self.integration_parameters_set = True
		if self._engine is not None:
			self._engine.set_integration_parameters(self.max_delay, **self._int_params)

	def _initiate(self):
		if not self.integration_parameters_set:
			self.set_integration_parameters()
		if self._engine is None:
			assert len(self._past) > 1, "You need to add at least two past points first."
			mask = (self._dst >= self.lo) & (self._dst < self.hi)
			src, tau, w, dst = self._src[mask], self._tau[mask], self._w[mask], self._dst[mask]
			row_ptr = np.zeros(self.hi-self.lo+1, dtype=np.int64)
			np.add.at(row_ptr, dst-self.lo+1, 1)
			row_ptr = np.cumsum(row_ptr)
			E = _capi.NetEngine(self.n, self.lo, self.hi, len(src), len(self._node_parameters), ring_capacity=self.ring_capacity, device=self.device)
			E.load_image(self._image)
			E.set_graph(row_ptr, src, tau, w)
			if self._node_parameters:
				E.set_node_parameters(np.stack(self._node_parameters))
			E.set_past([a[0] for a in self._past], np.stack([a[1] for a in self._past]), np.stack([a[2] for a in self._past]))
			E.set_integration_parameters(self.max_delay, **self._int_params)
			self._engine = E
			self._exchange = None
			if self.world > 1:
				self._setup_exchange()

	# ------------------------------------------------------------------ multi-GPU anchor exchange
	def _setup_exchange(self):
		import torch
		import torch.distributed as dist
		if self.exchange == "peer":
			# fused exchange: the step kernel stores into the peers' staging buffers (CUDA IPC over NVLink) and a device-side
			# arrival barrier completes it; torch.distributed only carries the 64-byte handles, once
			handles = [None]*self.world
			dist.all_gather_object(handles, self._engine.exchange_handle())
			self._engine.connect_peers(self.rank, self.world, handles)
			dist.barrier()
			self._exchange = "peer"
			return
		stage_ptr, p_ptr = self._engine.exchange_pointers()
		as_tensor = lambda ptr, count: torch.as_tensor(_DeviceArray(ptr, (count,), "<f8"), device=torch.device("cuda", self.device))
		self._exchange = (as_tensor(stage_ptr, 2*self.n), as_tensor(p_ptr, 1))
		self._engine.set_stream(torch.cuda.current_stream().cuda_stream)

	def _exchange_anchors(self):
		"""all-gather of the staged [state, diff] slices over NVLink + max-reduction of the error norm (the IEEE bits of a
		non-negative double are monotone in its value, so MAX on the float64 view is the max of the norms)"""
		import torch.distributed as dist
		stage, p = self._exchange
		dist.all_gather_into_tensor(stage, stage[2*self.lo:2*self.hi])      # in place: every rank's slice sits at its offset
		dist.all_reduce(p, op=dist.ReduceOp.MAX)

	def _run(self):
		E = self._engine
		while True:
			for _ in range(self.poll_every):
				E.attempt()
				if self.world > 1:
					self._exchange_anchors()
				E.control()
			done, t, status = E.poll()
			if done:
				return status

	def _grow(self, status):
		"""The reference's history is unbounded: a full ring is doubled (every rank takes the same decision) and the call goes on."""
		if status != _capi.STATUS_OVERFLOW or self._engine.ring_capacity >= self.max_ring_capacity:
			return False
		if self.verbose:
			print(f"Anchor ring full; growing to {2*self._engine.ring_capacity} anchors.")
		self._engine.grow_ring(2*self._engine.ring_capacity)
		return True

	def _raise(self, status):
		if status == _capi.STATUS_OK:
			return
		if status == _capi.STATUS_MIN_STEP:
			raise UnsuccessfulIntegration("Could not integrate with the given tolerance parameters (step size below min_step).")
		if status == _capi.STATUS_OVERFLOW:
			raise MemoryError("Anchor ring full and it could not be enlarged; pass a larger ring_capacity to jitcdde_network.")
		raise UnsuccessfulIntegration("A delay became shorter than the step, the step size collapsed or the state is not finite.")

	# ------------------------------------------------------------------ integration
	@property
	def t(self):
		self._initiate()
		return self._engine.poll()[1]

	@property
	def dt(self):
		self._initiate()
		return self._engine.get_dt()

	def integrate(self, target_time):
		"""as jitcdde.integrate (_jitcdde.py:807-836)"""
		self._initiate()
		if self.world == 1 or self._exchange == "peer":
			# (peer exchange: every rank runs the same loop in the runtime library; the GPUs meet on the device)
			out, status = self._engine.integrate(target_time)
			while self._grow(status):
				out, status = self._engine.resume()
			self._raise(status)
			return out
		self._engine.begin(MODE_ADAPTIVE, target_time)
		status = self._run()
		while self._grow(status):
			self._engine.begin(MODE_CONTINUE, 0.0)
			status = self._run()
		self._raise(status)
		return self._engine.recent_state(target_time)

	def adjust_diff(self, shift_ratio=1e-4):
		"""as jitcdde.adjust_diff (_jitcdde.py:863-878): a zero-amplitude jump at the last anchor, so that the integration starts
		with the derivative the differential equation gives there. Returns (minima, maxima) of every node over the jump interval."""
		self._initiate()
		if self.world > 1:
			if self._exchange != "peer":
				raise NotImplementedError("adjust_diff with several ranks needs the peer exchange (exchange='peer').")
			import torch.distributed as dist
			result = self._engine.adjust_diff(shift_ratio, barrier=dist.barrier)
			dist.barrier()      # nobody goes on (and overwrites the buffers) before everybody has read them
			return result
		return self._engine.adjust_diff(shift_ratio)

	def step_on_discontinuities(self, *, propagations=1, min_distance=1e-5, max_distinct_delays=64):
		"""as jitcdde.step_on_discontinuities (_jitcdde.py:935-999): adaptive steps that end exactly where the initial discontinuity
		arrives through the delays (and their sums, `propagations` times). Sensible for networks with FEW distinct delays (the
		number of steps grows like their number to the power of `propagations`); for heterogeneous delays use integrate_blindly or
		adjust_diff, as examples/kuramoto_network.py does."""
		from ._jitcdde import _propagate_delays
		self._initiate()
		if self.world > 1 and self._exchange != "peer":
			raise NotImplementedError("step_on_discontinuities with several ranks needs the peer exchange (exchange='peer').")
		assert min_distance > 0, "min_distance must be positive."
		assert isinstance(propagations, int), "Non-integer number of propagations."
		delays = [float(d) for d in np.unique(self._tau)]
		if len(delays) > max_distinct_delays:
			raise ValueError(f"{len(delays)} distinct delays: stepping on all discontinuities is not sensible; use integrate_blindly or adjust_diff.")
		delays.append(0.0)
		steps = _propagate_delays(delays, propagations, min_distance)
		steps.sort()
		steps.remove(0)
		if not steps:
			self.adjust_diff()
			return self._engine.recent_state(self.t, current_only=True)
		out = None
		for step in steps:
			out, status = self._engine.step_to(self._engine.poll()[1]+step)
			while self._grow(status):
				out, status = self._engine.resume()
			self._raise(status)
		return out

	def jump(self, amplitude, time, width=1e-5, forward=True):
		"""as jitcdde.jump (_jitcdde.py:935-974): adds `amplitude` (a number or one per node) to the state over the interval
		(time, time+width) — or (time−width, time) with forward=False —, for a time on the last segment of the history (usually
		the time the integration has reached). Returns (minima, maxima) of every node over the jump interval."""
		self._initiate()
		if self.world > 1:
			if self._exchange != "peer":
				raise NotImplementedError("jump with several ranks needs the peer exchange (exchange='peer').")
			import torch.distributed as dist
			result = self._engine.jump(amplitude, time, width, forward, barrier=dist.barrier)
			dist.barrier()
			return result
		return self._engine.jump(amplitude, time, width, forward)

	def integrate_blindly(self, target_time, step=None):
		"""as jitcdde.integrate_blindly (_jitcdde.py:880-933); with target_time == t it is adjust_diff, as there"""
		self._initiate()
		step = step or self.max_step
		assert step > 0, "step must be positive"
		if target_time == self._engine.poll()[1]:
			self.adjust_diff()
			return self._engine.recent_state(target_time, current_only=True)
		if self.world == 1 or self._exchange == "peer":
			out, status = self._engine.integrate_blindly(target_time, step)
			while self._grow(status):
				out, status = self._engine.resume()
			self._raise(status)
			return out
		total = target_time-self._engine.poll()[1]
		if total < 0:
			raise ValueError("Can’t integrate backwards in time.")
		number = 0
		if total > 0:
			if total < step:
				step = total
			number = round(total/step)
			self._engine.begin(MODE_BLIND, target_time, total/number, number)
			status = self._run()
			while self._grow(status):
				self._engine.begin(MODE_CONTINUE, 0.0)
				status = self._run()
			self._raise(status)
		return self._engine.recent_state(target_time, current_only=True)

	def counters(self):
		"""(accepted steps, rejected attempts, attempted steps)"""
		self._initiate()
		return self._engine.counters()
