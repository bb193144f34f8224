/*
 * jdn_kernels.cu — kernel image of the network engine, compiled per (local, coupling) function pair by
 * jitcdde_b200/_build.py with -DJDN_NPARAMS=… -DJDN_RHS_FILE="net_<key>.inc" for sm_100a.
 *
 * jdn_k_lookup: one thread per in-edge (10^7 threads for 10^6 nodes of in-degree 10); the anchor times live in
 * shared memory for the bracket search. jdn_k_attempt: one thread per owned node (3907 blocks of 256 for 10^6
 * nodes: 26 blocks per SM, several waves); the per-node error contributions are max-reduced per block (warp
 * shuffles, then shared memory) and merged with one atomicMax per block.
 */
#include "jdn_engine.cuh"

#ifndef JDN_BLOCK
#define JDN_BLOCK 256
#endif

extern "C" __device__ const jdn_image_desc_t jdn_image_desc = { JDN_IMAGE_MAGIC, JDDE_ABI_VERSION, JDN_NPARAMS, JDN_BLOCK };

/* lookup phase: one thread per edge of the owned nodes */
extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_lookup(jdn_problem const P, double * const scratch, long long const n_edges)
{
	extern __shared__ double s_times[];      /* cap doubles */
	jdn_control const & C = *P.ctrl;
	if (C.done)
		return;
	for (int k = threadIdx.x; k < P.cap; k += blockDim.x)
		s_times[k] = P.times[k];
	__syncthreads();
	jdn::View V = jdn::view_of(P);
	V.times = s_times;
	long long const e = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (e < n_edges)
	{
		int pws = 0;
		jdn::lookup_edge(P, V, C.h, e, scratch, pws);
		if (pws)
			P.ctrl->pws = 1;
	}
}

/* step phase: one thread per owned node */
extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_attempt(jdn_problem const P, double const * const scratch)
{
	__shared__ double s_max[JDN_BLOCK/32];
	jdn_control const & C = *P.ctrl;
	if (C.done)
		return;
	jdn::View const V = jdn::view_of(P);
	long long const i = P.lo + (long long)blockIdx.x*blockDim.x + threadIdx.x;
	double contribution = -1.0;
	if (i < P.hi)
	{
		contribution = jdn::attempt_node(P, V, C.h, i, scratch);
		if (!(contribution >= 0.0))      /* skipped (0/0) or NaN: ignored by the reference's `x > p` test */
			contribution = -1.0;
	}
	for (int o = 16; o > 0; o >>= 1)
		contribution = fmax(contribution, __shfl_down_sync(0xffffffffu, contribution, o));
	if ((threadIdx.x & 31) == 0)
		s_max[threadIdx.x >> 5] = contribution;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		double m = s_max[0];
		for (int w = 1; w < JDN_BLOCK/32; w++)
			m = fmax(m, s_max[w]);
		if (m > 0.0)
			atomicMax(&P.ctrl->p_bits, (unsigned long long)__double_as_longlong(m));
	}
}

extern "C" __global__ void jdn_k_control(jdn_problem const P)
{
	if (blockIdx.x == 0 && threadIdx.x == 0)
		jdn::control(P);
}

/* copy the accepted staged anchor into its ring slot */
extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_commit(jdn_problem const P)
{
	int const slot = P.ctrl->commit_slot;
	if (slot < 0)
		return;
	long long const i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i < P.n)
		reinterpret_cast<jdn::pair *>(P.hist)[(size_t)slot*P.n+i] = reinterpret_cast<jdn::pair const *>(P.stage)[i];
}

/* prepare a call: mode, target, step; blind: h and the number of steps are given by the host (_jitcdde.py:880-900) */
extern "C" __global__ void jdn_k_begin(jdn_problem const P, int const mode, double const target, double const h, long long const steps)
{
	if (blockIdx.x || threadIdx.x)
		return;
	jdn_control & C = *P.ctrl;
	if (C.status != JDDE_STATUS_OK)
	{
		C.done = 1;
		return;
	}
	C.mode = mode;
	C.target = target;
	C.commit_slot = -1;
	C.pws = 0;
	C.p_bits = 0;
	if (mode == JDN_MODE_BLIND)
	{
		C.h = h;
		C.remaining = steps;
		C.done = steps <= 0;
	}
	else
	{
		C.h = C.dt;
		C.done = !(C.t < target);
		if (C.done)
			jdn::forget(P, C);
	}
	C.last_start = C.t;
}

extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_recent_state(jdn_problem const P, double const t, int const current_only, double * const out)
{
	long long const i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= P.n)
		return;
	if (current_only)
		out[i] = P.hist[2*((size_t)(P.ctrl->current & (P.cap-1))*P.n + i)];      /* get_current_state */
	else
		out[i] = jdn::recent_state(P, t, i);
}

/* initial past (initiate_past_from_list, jitced_template.c:575-606): anchors [k][n] → rows of the ring */
extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_set_past(jdn_problem const P, int const n_anchors, double const * const times, double const * const states, double const * const diffs, long long const n_edges)
{
	long long const i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i < P.n)
		for (int k = 0; k < n_anchors; k++)
		{
			P.hist[2*((size_t)k*P.n+i)] = states[(size_t)k*P.n+i];
			P.hist[2*((size_t)k*P.n+i)+1] = diffs[(size_t)k*P.n+i];
		}
	for (long long e = i; e < n_edges; e += (long long)gridDim.x*blockDim.x)
		P.cursor[e] = n_anchors-2;      /* all cursors start at last_anchor->previous (:647-648) */
	if (i == 0)
	{
		jdn_control & C = *P.ctrl;
		for (int k = 0; k < n_anchors; k++)
			P.times[k] = times[k];
		C.first = 0;
		C.current = n_anchors-1;
		C.t = times[n_anchors-1];
		C.last_start = times[0];
		C.status = JDDE_STATUS_OK;
		C.done = 1;
		C.commit_slot = -1;
		C.accepted = C.rejected = C.attempts = 0;
	}
}
