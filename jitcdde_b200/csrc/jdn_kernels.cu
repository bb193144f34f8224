/*
 * jdn_kernels.cu — kernel image of the network engine, compiled per (local, coupling) function pair by
 * jitcdde_b200/_build.py with -DJDN_NPARAMS=… -DJDN_RHS_FILE="net_<key>.inc" for sm_100a.
 *
 * jdn_k_lookup: one thread per in-edge (10^7 threads for 10^6 nodes of in-degree 10); the anchor times live in
 * shared memory for the bracket search. jdn_k_attempt: one thread per owned node (3907 blocks of 256 for 10^6
 * nodes: 26 blocks per SM, several waves); the per-node error contributions are max-reduced per block (warp
 * shuffles, then shared memory) and merged with one atomicMax per block.
 */
#include <cooperative_groups.h>
#include "jdn_engine.cuh"

#ifndef JDN_BLOCK
#define JDN_BLOCK 256
#endif

#ifndef JDN_TILE_NODES
#define JDN_TILE_NODES 64
#endif
#ifndef JDN_TILE_EDGES
#define JDN_TILE_EDGES 1024
#endif
/* dynamic shared memory of jdn_k_run after the anchor times: values [3][JDN_TILE_EDGES], row offsets, stage states, block maxima, node of each edge */
#define JDN_RUN_SMEM_BYTES (3*JDN_TILE_EDGES*8 + (JDN_TILE_NODES+1)*8 + JDN_TILE_NODES*8 + 32*8 + JDN_TILE_EDGES)

extern "C" __device__ const jdn_image_desc_t jdn_image_desc = { JDN_IMAGE_MAGIC, JDDE_ABI_VERSION, JDN_NPARAMS, JDN_BLOCK, JDN_RUN_SMEM_BYTES };

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
	int const buffer = (int)(C.epoch & 1);
	if (i < P.hi)
	{
		jdn::pair staged;
		contribution = jdn::attempt_node(P, V, C.h, i, scratch, staged);
		if (!(contribution >= 0.0))      /* skipped (0/0) or NaN: ignored by the reference's `x > p` test */
			contribution = -1.0;
		if (P.world > 1)
		{
			/* fused exchange: the node's part of the tentative anchor goes straight into every rank's staging buffer
			 * (peer memory over NVLink; consecutive threads write consecutive 16-byte pairs) */
			for (int r = 0; r < P.world; r++)
				reinterpret_cast<jdn::pair *>(jdn_exchange_stage(P.peers[r], buffer, P.n))[i] = staged;
		}
		else
			reinterpret_cast<jdn::pair *>(P.stage)[i] = staged;
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
		{
			if (P.world > 1)
				for (int r = 0; r < P.world; r++)
					atomicMax_system(&P.peers[r]->p_bits[buffer], (unsigned long long)__double_as_longlong(m));
			else
				atomicMax(&P.ctrl->p_bits, (unsigned long long)__double_as_longlong(m));
		}
	}
}

/*
 * Completion of the fused exchange (several GPUs), one thread: tell every peer that this rank's part of the attempt has
 * been delivered — the step kernel that stored it has ended, so a release increment orders it after the data — and wait
 * until every peer has said the same. Then hand the merged error norm to the controller. The wait is bounded: a peer that
 * does not arrive within 20 s ends the integration with an error instead of hanging the device.
 */
extern "C" __global__ void jdn_k_exchange(jdn_problem const P)
{
	if (blockIdx.x || threadIdx.x || P.world <= 1)
		return;
	jdn_control & C = *P.ctrl;
	if (C.done)
		return;
	int const buffer = (int)(C.epoch & 1);
	if (C.pws)
		for (int r = 0; r < P.world; r++)
			atomicMax_system(&P.peers[r]->p_bits[buffer], JDN_PWS_SENTINEL);
	__threadfence_system();
	for (int r = 0; r < P.world; r++)
		if (r != P.rank)
			asm volatile("red.release.sys.global.add.u64 [%0], %1;" :: "l"(&P.peers[r]->arrived), "l"(1ull) : "memory");
	jdn_exchange * const own = P.peers[P.rank];
	unsigned long long const expected = (C.epoch+1)*(unsigned long long)(P.world-1);
	unsigned long long start, now, arrived;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
	for (;;)
	{
		asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(arrived) : "l"(&own->arrived) : "memory");
		if (arrived >= expected)
			break;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
		if (now-start > 20000000000ull)
		{
			C.exchange_failed = 1;
			C.status = JDDE_STATUS_NONFINITE;
			C.done = 1;
			return;
		}
		__nanosleep(200);
	}
	unsigned long long const bits = atomicExch_system(&own->p_bits[buffer], 0ull);
	if (bits == JDN_PWS_SENTINEL)
	{
		C.pws = 1;
		C.p_bits = 0;
	}
	else
		C.p_bits = bits;
	C.exchanged_buffer = buffer;
	C.epoch++;
}

/*
 * The persistent stepper: ONE cooperative launch runs up to `max_attempts` attempted steps instead of five launches
 * per attempt, and an attempted step is ONE pass over the network.
 *
 * A block takes tiles of JDN_TILE_NODES consecutive nodes. For a tile all threads first interpolate the delayed source
 * states of the tile's in-edges at the three stage times (one thread per edge: the scattered reads of the history, the
 * part that HBM bounds) — into SHARED memory, not into a scratch array in global memory; then, stage by stage, every
 * edge thread turns its value into its term weight·coupling(·, y_stage of its node) and one thread per node adds the
 * terms of its node in the order of the edge list (the reference's order) and forms the next stage state. Seven block
 * barriers per tile, no grid barrier between lookup and step phase, 48 bytes per edge less traffic and one pass less
 * over the edge arrays than the two-kernel version. (A tile with more than JDN_TILE_EDGES in-edges keeps its values in
 * the global scratch array.)
 *
 * A node's new (state, diff) pair goes straight into ring slot current+1 (the tentative anchor: no lookup can reach it,
 * every delay being at least a step), so an accepted step only moves `current`: no staging buffer, no commit pass on
 * one GPU. With several GPUs the pair is also stored into every other rank's staging buffer over NVLink — consecutive
 * nodes, full-width packets — while the rank still computes other tiles; the error norm is merged on the GPU first and
 * after a grid barrier thread 0 of block 0 passes the rank's maximum to every peer (one system-scope atomicMax each),
 * tells every peer that this rank's part has been delivered (release increment) and waits until all peers have said
 * the same — the only inter-GPU synchronisation of an attempt — and runs the controller. A second grid barrier
 * publishes its decision; after an accepted step the ranks copy the other ranks' nodes from staging into the ring.
 * Grid barriers (cooperative groups: release/acquire at GPU scope) are cumulative with the system-scope release/acquire
 * of the arrival counters. The controller state is read afresh after every barrier.
 */
/* -DJDN_PROFILE (JITCDDE_B200_NET_FLAGS): thread 0 of block 0 adds up the time it spends in each phase of an attempt and
 * prints the sums at the end of the launch (scripts/gpu_net_phases.sh) — a measuring aid, not compiled otherwise */
#if defined(JDN_PROFILE)
#define JDN_PHASE(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long now_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_)); jdn_phase_ns[k] += now_-jdn_phase_last; jdn_phase_last = now_; } } while (0)
#define JDN_PHASE_STATE unsigned long long jdn_phase_ns[8] = {0, 0, 0, 0, 0, 0, 0, 0}, jdn_phase_last = 0; int jdn_phase_attempts = 0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(jdn_phase_last))
#define JDN_PHASE_ARGS , jdn_phase_ns, jdn_phase_last
#define JDN_PHASE_PARAMS , unsigned long long * const jdn_phase_ns, unsigned long long & jdn_phase_last
#define JDN_PHASE_REPORT do { if (blockIdx.x == 0 && threadIdx.x == 0 && jdn_phase_attempts) printf("[jdn phases] rank %d attempts %d | tiles(own block) %.1f | grid barrier %.1f | push norm + arrival signals %.1f | wait for peers %.1f | controller %.1f | grid barrier %.1f | commit + grid barrier %.1f  (us per attempt)\n", P.rank, jdn_phase_attempts, jdn_phase_ns[0]*1e-3/jdn_phase_attempts, jdn_phase_ns[1]*1e-3/jdn_phase_attempts, jdn_phase_ns[2]*1e-3/jdn_phase_attempts, jdn_phase_ns[3]*1e-3/jdn_phase_attempts, jdn_phase_ns[4]*1e-3/jdn_phase_attempts, jdn_phase_ns[5]*1e-3/jdn_phase_attempts, jdn_phase_ns[6]*1e-3/jdn_phase_attempts); } while (0)
#else
#define JDN_PHASE(k) do { } while (0)
#define JDN_PHASE_STATE do { } while (0)
#define JDN_PHASE_ARGS
#define JDN_PHASE_PARAMS
#define JDN_PHASE_REPORT do { } while (0)
#endif

/* end of an attempt in the persistent steppers, thread 0 of block 0: exchange barrier between the GPUs, then the controller */
__device__ __forceinline__ void jdn_run_exchange_and_control(jdn_problem const & P, jdn_control & C, int const buffer JDN_PHASE_PARAMS)
{
	if (blockIdx.x != 0 || threadIdx.x != 0)
		return;
	bool arrived_all = true;
	if (P.world > 1)
	{
		unsigned long long const mine = C.pws ? JDN_PWS_SENTINEL : C.p_bits;
		if (mine)
			for (int r = 0; r < P.world; r++)      /* (reductions that return nothing: posted, not waited for one by one) */
				asm volatile("red.relaxed.sys.global.max.u64 [%0], %1;" :: "l"(&P.peers[r]->p_bits[buffer]), "l"(mine) : "memory");
		if (C.pws_bits)      /* rare: a delay shorter than the step */
			for (int r = 0; r < P.world; r++)
				atomicMax_system(&P.peers[r]->pws_bits[buffer], C.pws_bits);
		if (C.diff_bits)
			for (int r = 0; r < P.world; r++)
				atomicMax_system(&P.peers[r]->diff_bits[buffer], C.diff_bits);
		/* ONE system-scope fence orders everything this rank has stored (the grid barrier has made the other blocks' stores
		 * this thread's business) before the arrival increments, which can then be relaxed — a release increment per peer
		 * is a fence per peer, one after the other; the poll is relaxed, too, with one fence once all have arrived */
		__threadfence_system();
		for (int r = 0; r < P.world; r++)
			if (r != P.rank)
				asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" :: "l"(&P.peers[r]->arrived), "l"(1ull) : "memory");
		JDN_PHASE(2);
		jdn_exchange * const own = P.peers[P.rank];
		unsigned long long const expected = (C.epoch+1)*(unsigned long long)(P.world-1);
		unsigned long long start, now, arrived;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
		for (;;)
		{
			asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(arrived) : "l"(&own->arrived) : "memory");
			if (arrived >= expected)
			{
				__threadfence_system();
				break;
			}
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if (now-start > 20000000000ull)      /* a peer that does not arrive within 20 s ends the integration instead of hanging the device */
			{
				C.exchange_failed = 1;
				C.status = JDDE_STATUS_NONFINITE;
				C.done = 1;
				arrived_all = false;
				break;
			}
			__nanosleep(20);
		}
		if (arrived_all)
		{
			unsigned long long const bits = atomicExch_system(&own->p_bits[buffer], 0ull);
			if (bits == JDN_PWS_SENTINEL)
			{
				C.pws = 1;
				C.p_bits = 0;
			}
			else
				C.p_bits = bits;
			C.pws_bits = atomicExch_system(&own->pws_bits[buffer], 0ull);
			C.diff_bits = atomicExch_system(&own->diff_bits[buffer], 0ull);
			C.exchanged_buffer = buffer;
			C.epoch++;
		}
	}
	else
		C.epoch++;      /* (one GPU: the parity of the attempts selects the tentative buffer) */
	JDN_PHASE(3);
	C.next_tile = 0;      /* (all blocks are past their tiles: first grid barrier) */
	if (arrived_all)
		jdn::control(P);
	__threadfence();
	JDN_PHASE(4);
}

/* after an accepted step, several GPUs: the other ranks' nodes go from the staging buffer (written by the peers: read past L1)
 * into the ring slot that has just become `current` (all threads of the grid; a grid barrier follows) */
__device__ __forceinline__ void jdn_run_commit(jdn_problem const & P, jdn_control const & C, long long const owned)
{
	if (P.world <= 1 || C.commit_slot < 0)
		return;
	int const committed = C.commit_slot;
	double2 const * const staged = reinterpret_cast<double2 const *>(jdn_exchange_stage(P.peers[P.rank], C.exchanged_buffer, P.n));
	long long const others = P.n-owned;
	for (long long k = (long long)blockIdx.x*blockDim.x+threadIdx.x; k < others; k += (long long)gridDim.x*blockDim.x)
	{
		long long const i = k < P.lo ? k : k+owned;
		reinterpret_cast<double2 *>(P.hist)[jdn_pair_index(committed, i, P.n)] = __ldcg(staged+i);
	}
}

#if JDN_TILE_NODES > JDN_BLOCK-1 || JDN_TILE_NODES > 256
#error "a tile needs one thread per node (and one more for the row offsets) and its node indices fit a byte"
#endif

#ifndef JDN_RUN_MIN_BLOCKS
#define JDN_RUN_MIN_BLOCKS 4
#endif
/* Several GPUs: how a tile's new pairs reach the peers' staging buffers. 1: ONE bulk copy (cp.async.bulk shared → global, up to
 * 1 KB) per tile and peer, issued by one thread and completed asynchronously — the copy engine of the SM forms full NVLink
 * packets, the threads go on with the next tile. 0: every node thread stores its 16-byte pair to every peer (round 2 before
 * this: the overhead of the exchange grew in proportion to the bytes stored, ≈ 125 GB/s per GPU, profiles/configs_r02.md). */
#ifndef JDN_BULK_PUSH
#define JDN_BULK_PUSH 1
#endif

extern "C" __global__ void __launch_bounds__(JDN_BLOCK, JDN_RUN_MIN_BLOCKS)
jdn_k_run(jdn_problem const P, double * const scratch, long long const n_edges, int const max_attempts)
{
	extern __shared__ double s_times[];                /* cap doubles, then: */
	double * const s_value = s_times+P.cap;            /* per edge of the tile: delayed value, then term, of stage 0, 1, 2 */
	long long * const s_row = reinterpret_cast<long long *>(s_value+3*JDN_TILE_EDGES);     /* [JDN_TILE_NODES+1] */
	double * const s_y = reinterpret_cast<double *>(s_row+JDN_TILE_NODES+1);               /* stage state of the tile's nodes */
	double * const s_max = s_y+JDN_TILE_NODES;         /* [32] block reduction of the error norm */
	unsigned char * const s_node = reinterpret_cast<unsigned char *>(s_max+32);            /* node of the edge within the tile */
	namespace cg = cooperative_groups;
	cg::grid_group grid = cg::this_grid();
	jdn_control & C = *P.ctrl;
	jdn::pair * const own_hist = reinterpret_cast<jdn::pair *>(P.hist);
	long long const owned = P.hi-P.lo;
	int const tile_nodes = P.tile_nodes > 0 && P.tile_nodes < JDN_TILE_NODES ? P.tile_nodes : JDN_TILE_NODES;
	long long const n_tiles = (owned + tile_nodes-1)/tile_nodes;
	__shared__ long long s_next_tile[2];      /* the tile after the one under way (two slots: written while the other is still read) */
	__shared__ double s_tolerance[2];         /* atol, rtol */
#if JDN_BULK_PUSH
	__shared__ __align__(16) double2 s_out[JDN_TILE_NODES];      /* the tile's new (state, diff) pairs on their way to the peers */
#endif
	JDN_PHASE_STATE;
	for (int attempt = 0; attempt < max_attempts; attempt++)
	{
		if (C.done)
			break;      /* (every thread of every block sees the same value: written before the last grid barrier) */
		jdn::View V = jdn::view_of(P);
		double const h = C.h;
		int const slot = (V.current+1) & V.mask;
		int const buffer = (int)(C.epoch & 1);
		/* past within step (jdn_abi.h): the result also goes into tentative buffer `buffer` — with several GPUs that is this
		 * rank's own staging buffer —, where the next attempt finds it if this one is not accepted; in a repetition
		 * (pws_pass > 0) it is compared with the result before (check_new_y_diff, jitced_template.c:502-524) */
		bool const repetition = C.pws_pass > 0;
		for (int k = threadIdx.x; k < P.cap; k += blockDim.x)
			s_times[k] = P.times[k];
		if (threadIdx.x == 0)
		{
			s_tolerance[0] = C.P.atol;
			s_tolerance[1] = C.P.rtol;
		}
		__syncthreads();
		V.times = s_times;
		double block_contribution = -1.0;      /* (threads 0 … JDN_TILE_NODES−1: their nodes) */
		/* (what only the last stage or the rare past-within-step path needs — tolerances, stage times, the tentative buffer — is
		 * fetched where it is used, the tolerances from shared memory: at 64 registers per thread everything carried through
		 * the tile loop is a spill, and a load from global memory there would sit on the critical path of every tile) */

		/* a block starts with tile blockIdx.x; the tiles beyond the first gridDim.x are taken in the order in which the blocks
		 * get there (one atomic per tile, asked for while the tile before is computed): the blocks do not run at the same
		 * speed — 13 % of an attempt was waiting for the slowest one with a fixed assignment, profiles/configs_r02.md */
		int turn = 0;
		for (long long tile = blockIdx.x; tile < n_tiles; turn ^= 1)
		{
			long long const row0 = tile*tile_nodes;
			int const nn = (int)(owned-row0 < tile_nodes ? owned-row0 : tile_nodes);
			if (threadIdx.x <= nn)
				s_row[threadIdx.x] = P.row_ptr[row0+threadIdx.x];
			if (threadIdx.x == blockDim.x-1)
				s_next_tile[turn] = (long long)gridDim.x + atomicAdd(&P.ctrl->next_tile, 1);
#if JDN_BULK_PUSH
			if (P.world > 1 && threadIdx.x == 0)      /* the bulk copies of the tile before have read s_out */
				asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
			__syncthreads();
			long long const e0 = s_row[0];
			int const ne = (int)(s_row[nn]-e0);
			bool const in_shared = ne <= JDN_TILE_EDGES;
			double * const value = in_shared ? s_value : scratch+3*e0;      /* [3][ne] in shared memory, [ne][3] in the scratch array */
			int const edge_stride = in_shared ? 1 : 3, stage_stride = in_shared ? JDN_TILE_EDGES : 1;

			/* lookups: one thread per in-edge of the tile */
			for (int k = threadIdx.x; k < ne; k += blockDim.x)
			{
				double v1, v2, v3;
				double overlap = 0.0;
				jdn::lookup_edge_values(P, V, h, e0+k, v1, v2, v3, overlap);
				if (overlap > 0.0)      /* rare: a delay shorter than the step */
					atomicMax(&P.ctrl->pws_bits, (unsigned long long)__double_as_longlong(overlap));
				value[k*edge_stride] = v1;
				value[k*edge_stride+stage_stride] = v2;
				value[k*edge_stride+2*stage_stride] = v3;
				if (in_shared)
				{
					/* the node this edge leads to: the last row of the tile that starts at or before it */
					int lo = 0, hi = nn-1;
					while (lo < hi)
					{
						int const mid = (lo+hi+1) >> 1;
						if (s_row[mid]-e0 <= k)
							lo = mid;
						else
							hi = mid-1;
					}
					s_node[k] = (unsigned char)lo;
				}
			}
			/* one thread per node of the tile: current anchor, parameters, first stage state */
			long long const i = P.lo+row0+threadIdx.x;
			double y = 0.0, k_1 = 0.0, k_2 = 0.0, k_3 = 0.0, y_stage = 0.0;
			double par[JDN_NPARAMS > 0 ? JDN_NPARAMS : 1];
			if (threadIdx.x < nn)
			{
				jdn::pair const cur = own_hist[jdn_pair_index(V.current & V.mask, i, P.n)];
				y = cur.state;
				k_1 = cur.diff;
				for (int q = 0; q < JDN_NPARAMS; q++)
					par[q] = P.node_par[(size_t)q*P.n+i];
				y_stage = y + 0.5*h*k_1;
				s_y[threadIdx.x] = y_stage;
			}
			__syncthreads();
			for (int stage = 0; stage < 3; stage++)
			{
				/* every edge: its term of this stage (get_next_step's three evaluations of f, jitced_template.c:441-463) */
				if (in_shared)
					for (int k = threadIdx.x; k < ne; k += blockDim.x)
					{
						double const yd = value[k+stage*JDN_TILE_EDGES];
						value[k+stage*JDN_TILE_EDGES] = P.weight[e0+k]*JDN_COUPLING(yd, s_y[s_node[k]]);
					}
				__syncthreads();
				if (threadIdx.x < nn)
				{
					double const t_stage = V.t_cur + (stage == 0 ? 0.5*h : stage == 1 ? 0.75*h : h);
					double acc = JDN_LOCAL(y_stage, t_stage, par);
					long long const r0 = s_row[threadIdx.x]-e0, r1 = s_row[threadIdx.x+1]-e0;
					if (in_shared)
						for (long long k = r0; k < r1; k++)
							acc += value[k+stage*JDN_TILE_EDGES];
					else
						for (long long k = r0; k < r1; k++)
							acc += P.weight[e0+k]*JDN_COUPLING(value[3*k+stage], y_stage);
					if (stage == 0)
					{
						k_2 = acc;
						y_stage = y + 0.75*h*k_2;
					}
					else if (stage == 1)
					{
						k_3 = acc;
						y_stage = y + (h/9.) * (2*k_1+3*k_2+4*k_3);
					}
					else
					{
						double const k_4 = acc;
						double const error = fabs((5*k_1-6*k_2-8*k_3+9*k_4) * (h/72.));
						jdn::pair staged;
						staged.state = y_stage;
						staged.diff = k_4;
						/* into the tentative slot of this rank's ring, and — several GPUs — into the staging buffer of every
						 * other rank: consecutive nodes at consecutive 16-byte pairs, i.e. full-width NVLink packets (the ring
						 * keeps the pairs of four slots of a node together, 64 bytes from node to node: stores of single
						 * pairs into the peers' rings measured 0.40 efficiency on 8 GPUs, profiles/) */
						own_hist[jdn_pair_index(slot, i, P.n)] = staged;
						if (repetition)
						{
							double const difference = fabs(y_stage - jdn::last_tentative(P)[i].state);
							double const tolerance = C.P.pws_atol + fabs(C.P.pws_rtol*y_stage);
							if (!(tolerance >= difference))
								P.ctrl->diff_bits = 1;
						}
						if (P.keep_tentative)
							jdn::tentative_buffer(P, buffer)[i] = staged;
#if JDN_BULK_PUSH
						if (P.world > 1)
						{
							s_out[threadIdx.x] = make_double2(staged.state, staged.diff);
							asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      /* visible to the bulk copy below */
						}
#else
						for (int r = 0; r < P.world; r++)
							if (r != P.rank)
								reinterpret_cast<jdn::pair *>(jdn_exchange_stage(P.peers[r], buffer, P.n))[i] = staged;
#endif
						/* get_p, jitced_template.c:474-498: 0/0 is skipped, a NaN never wins the comparison */
						double const tolerance = s_tolerance[0] + s_tolerance[1]*fabs(y_stage);
						if (error != 0.0 || tolerance != 0.0)
						{
							double const c = error/tolerance;
							if (c >= 0.0)
								block_contribution = fmax(block_contribution, c);
						}
					}
					s_y[threadIdx.x] = y_stage;
				}
				__syncthreads();
			}
#if JDN_BULK_PUSH
			if (P.world > 1 && threadIdx.x == 0)
			{
				/* consecutive nodes at consecutive 16-byte pairs of every other rank's staging buffer */
				unsigned const source = (unsigned)__cvta_generic_to_shared(s_out);
				unsigned const bytes = (unsigned)nn*16u;
				for (int r = 0; r < P.world; r++)
					if (r != P.rank)
					{
						unsigned long long const target = (unsigned long long)__cvta_generic_to_global(
							reinterpret_cast<jdn::pair *>(jdn_exchange_stage(P.peers[r], buffer, P.n)) + (P.lo+row0));
						asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(target), "r"(source), "r"(bytes) : "memory");
					}
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			}
#endif
			tile = s_next_tile[turn];
		}
#if JDN_BULK_PUSH
		if (P.world > 1 && threadIdx.x == 0)      /* all of this block's pairs have been written at the peers */
			asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
		/* error norm of this block's nodes (threads 0 … JDN_TILE_NODES−1 hold them) */
		for (int o = 16; o > 0; o >>= 1)
			block_contribution = fmax(block_contribution, __shfl_down_sync(0xffffffffu, block_contribution, o));
		if ((threadIdx.x & 31) == 0)
			s_max[threadIdx.x >> 5] = block_contribution;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			for (int w = 1; w < (JDN_TILE_NODES+31)/32; w++)
				block_contribution = fmax(block_contribution, s_max[w]);
			/* merged on this GPU first (one atomic per block on a local word); thread 0 of block 0 passes the maximum of the
			 * rank on to the peers after the grid barrier — one remote atomic per peer instead of one per block and peer */
			if (block_contribution > 0.0)
				atomicMax(&P.ctrl->p_bits, (unsigned long long)__double_as_longlong(block_contribution));
		}
		JDN_PHASE(0);
		grid.sync();
		JDN_PHASE(1);

		jdn_run_exchange_and_control(P, C, buffer JDN_PHASE_ARGS);
		grid.sync();
		JDN_PHASE(5);
		jdn_run_commit(P, C, owned);
		if (P.world > 1 && C.commit_slot >= 0)
			grid.sync();
		JDN_PHASE(6);
#if defined(JDN_PROFILE)
		jdn_phase_attempts++;
#endif
	}
	JDN_PHASE_REPORT;
}

/*
 * The persistent stepper, warp tiles (a variant, JITCDDE_B200_NET_TILES=warp; jdn_k_run above takes a thread block per tile
 * of 64 nodes and needs seven block barriers per tile — 6.8 warps stalled at a barrier per issued instruction,
 * profiles/ — but this one needs more registers than it gets and measured 0.77 against 0.65 ms per attempt): a WARP takes a
 * tile of consecutive nodes with at most 32 in-edges together — lane k does the lookups of edge k (the three delayed
 * values stay in its registers) and, stage by stage, its term weight·coupling(·, y_stage of its node); lane j < nn is
 * also the thread of node j: it collects the terms of its node by shuffles in the order of the edge list (the
 * reference's order, bit-identical) and forms the next stage state, which the edge lanes fetch by shuffle. No block
 * barrier, no shared memory but the anchor times; the warps of an SM are at different points of their tiles and hide
 * each other's latencies. A node with more than 32 in-edges is a tile of its own: its lanes take the edges in rounds,
 * the delayed values go through the scratch array. Exchange between the GPUs, controller and grid barriers as above.
 */
#ifndef JDN_WARP_MIN_BLOCKS
#define JDN_WARP_MIN_BLOCKS 3
#endif

extern "C" __global__ void __launch_bounds__(JDN_BLOCK, JDN_WARP_MIN_BLOCKS)
jdn_k_run_warp(jdn_problem const P, double * const scratch, long long const n_edges, int const max_attempts)
{
	extern __shared__ double s_times[];                /* cap doubles */
	__shared__ double s_max[JDN_BLOCK/32];
	namespace cg = cooperative_groups;
	cg::grid_group grid = cg::this_grid();
	jdn_control & C = *P.ctrl;
	jdn::pair * const own_hist = reinterpret_cast<jdn::pair *>(P.hist);
	long long const owned = P.hi-P.lo;
	int const lane = threadIdx.x & 31;
	long long const warp = ((long long)blockIdx.x*blockDim.x + threadIdx.x) >> 5;
	long long const n_warps = ((long long)gridDim.x*blockDim.x) >> 5;
	unsigned const full = 0xffffffffu;
	JDN_PHASE_STATE;
	for (int attempt = 0; attempt < max_attempts; attempt++)
	{
		if (C.done)
			break;      /* (every thread of every block sees the same value: written before the last grid barrier) */
		jdn::View V = jdn::view_of(P);
		double const h = C.h;
		double const atol = C.P.atol, rtol = C.P.rtol;
		int const slot = (V.current+1) & V.mask;
		int const buffer = (int)(C.epoch & 1);
		for (int k = threadIdx.x; k < P.cap; k += blockDim.x)
			s_times[k] = P.times[k];
		__syncthreads();
		V.times = s_times;
		double const t1 = V.t_cur+0.5*h, t2 = V.t_cur+0.75*h, t3 = V.t_cur+h;
		int pws = 0;
		double contribution = -1.0;      /* (node lanes) */

		for (long long tile = warp; tile < P.n_warp_tiles; tile += n_warps)
		{
			int const row0 = P.tile_start[tile];
			int const nn = P.tile_start[tile+1]-row0;
			/* row offsets of the tile: lane j holds the first edge of node j (lane nn: the end) */
			long long const my_row = lane <= nn ? P.row_ptr[row0+lane] : 0;
			long long const e0 = __shfl_sync(full, my_row, 0);
			int const ne = (int)(__shfl_sync(full, my_row, nn)-e0);
			int const first = (int)(my_row-e0);                                   /* node lane: first edge of my node within the tile */
			int const degree = (int)(__shfl_down_sync(full, my_row, 1)-my_row);    /* node lane: in-degree of my node */
			long long const i = P.lo+row0+lane;                                    /* node lane: my node */
			bool const is_node = lane < nn;
			double y = 0.0, k_1 = 0.0, k_2 = 0.0, k_3 = 0.0, y_stage = 0.0;
			double par[JDN_NPARAMS > 0 ? JDN_NPARAMS : 1];
			if (is_node)
			{
				jdn::pair const cur = own_hist[jdn_pair_index(V.current & V.mask, i, P.n)];
				y = cur.state;
				k_1 = cur.diff;
				for (int q = 0; q < JDN_NPARAMS; q++)
					par[q] = P.node_par[(size_t)q*P.n+i];
				y_stage = y + 0.5*h*k_1;
			}
			bool const small = ne <= 32;
			double v1 = 0.0, v2 = 0.0, v3 = 0.0, weight = 0.0;
			int node = 0;
			int most = 0;       /* largest in-degree of the tile's nodes: how many terms the node lanes collect */
			if (small)
			{
				if (lane < ne)
				{
					jdn::lookup_edge_values(P, V, h, e0+lane, v1, v2, v3, pws);
					weight = P.weight[e0+lane];
				}
				/* the node this lane's edge leads to: the last node whose first edge is at or before it */
				for (int j = 1; j < nn; j++)
					node += __shfl_sync(full, first, j) <= lane;
				most = is_node ? degree : 0;
				for (int o = 16; o > 0; o >>= 1)
					most = max(most, __shfl_xor_sync(full, most, o));
			}
			else
			{
				/* one node with many in-edges: lookups in rounds of 32, values through the scratch array */
				for (long long e = e0+lane; e < e0+ne; e += 32)
					jdn::lookup_edge(P, V, h, e, scratch, pws);
				__syncwarp();
			}
			for (int stage = 0; stage < 3; stage++)
			{
				double const t_stage = stage == 0 ? t1 : stage == 1 ? t2 : t3;
				double acc = 0.0;
				if (small)
				{
					/* every edge lane: its term of this stage with the stage state of its node … */
					double const ys = __shfl_sync(full, y_stage, node);
					double const yd = stage == 0 ? v1 : stage == 1 ? v2 : v3;
					double const term = lane < ne ? weight*JDN_COUPLING(yd, ys) : 0.0;
					/* … every node lane: local part + the terms of its node in the order of the edge list */
					if (is_node)
						acc = JDN_LOCAL(y_stage, t_stage, par);
					for (int q = 0; q < most; q++)
					{
						double const tq = __shfl_sync(full, term, (first+q) & 31);
						if (is_node && q < degree)
							acc += tq;
					}
				}
				else if (lane == 0)
				{
					acc = JDN_LOCAL(y_stage, t_stage, par);
					for (long long e = e0; e < e0+ne; e++)
						acc += P.weight[e]*JDN_COUPLING(scratch[3*e+stage], y_stage);
				}
				if (is_node)
				{
					if (stage == 0)
					{
						k_2 = acc;
						y_stage = y + 0.75*h*k_2;
					}
					else if (stage == 1)
					{
						k_3 = acc;
						y_stage = y + (h/9.) * (2*k_1+3*k_2+4*k_3);
					}
					else
					{
						double const k_4 = acc;
						double const error = fabs((5*k_1-6*k_2-8*k_3+9*k_4) * (h/72.));
						jdn::pair staged;
						staged.state = y_stage;
						staged.diff = k_4;
						own_hist[jdn_pair_index(slot, i, P.n)] = staged;
						for (int r = 0; r < P.world; r++)
							if (r != P.rank)
								reinterpret_cast<jdn::pair *>(jdn_exchange_stage(P.peers[r], buffer, P.n))[i] = staged;
						/* get_p, jitced_template.c:474-498: 0/0 is skipped, a NaN never wins the comparison */
						double const tolerance = atol + rtol*fabs(y_stage);
						if (error != 0.0 || tolerance != 0.0)
						{
							double const c = error/tolerance;
							if (c >= 0.0)
								contribution = fmax(contribution, c);
						}
					}
				}
			}
		}
		if (pws)
			C.pws = 1;
		/* error norm of this block's nodes, merged on this GPU */
		for (int o = 16; o > 0; o >>= 1)
			contribution = fmax(contribution, __shfl_down_sync(full, contribution, o));
		if (lane == 0)
			s_max[threadIdx.x >> 5] = contribution;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			for (int w = 1; w < JDN_BLOCK/32; w++)
				contribution = fmax(contribution, s_max[w]);
			if (contribution > 0.0)
				atomicMax(&P.ctrl->p_bits, (unsigned long long)__double_as_longlong(contribution));
		}
		JDN_PHASE(0);
		grid.sync();
		JDN_PHASE(1);
		jdn_run_exchange_and_control(P, C, buffer JDN_PHASE_ARGS);
		grid.sync();
		JDN_PHASE(5);
		jdn_run_commit(P, C, owned);
		if (P.world > 1 && C.commit_slot >= 0)
			grid.sync();
		JDN_PHASE(6);
#if defined(JDN_PROFILE)
		jdn_phase_attempts++;
#endif
	}
	JDN_PHASE_REPORT;
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
	if (i >= P.n)
		return;
	if (P.world > 1)
	{
		/* written by the peers: read past L1 */
		double2 const * const staged = reinterpret_cast<double2 const *>(jdn_exchange_stage(P.peers[P.rank], P.ctrl->exchanged_buffer, P.n));
		double2 const value = __ldcg(staged+i);
		reinterpret_cast<double2 *>(P.hist)[jdn_pair_index(slot, i, P.n)] = value;
	}
	else
		reinterpret_cast<jdn::pair *>(P.hist)[jdn_pair_index(slot, i, P.n)] = reinterpret_cast<jdn::pair const *>(P.stage)[i];
}

/* prepare a call: mode, target, step; blind: h and the number of steps are given by the host (_jitcdde.py:880-900) */
extern "C" __global__ void jdn_k_begin(jdn_problem const P, int const mode, double const target, double const h, long long const steps)
{
	if (blockIdx.x || threadIdx.x)
		return;
	jdn_control & C = *P.ctrl;
	if (mode == JDN_MODE_CONTINUE)
	{
		jdn::continue_call(C);
		return;
	}
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
	C.pws_bits = C.diff_bits = 0;
	C.pws_pass = 0;
	if (mode == JDN_MODE_BLIND)
	{
		C.h = h;
		C.remaining = steps;
		C.done = steps <= 0;
	}
	else
	{
		C.h = jdn::next_length(C);      /* dt, or min(dt, target − t) for step_on_discontinuities */
		C.done = !(C.t < target);
		if (C.done)
			jdn::forget(P, C);
	}
	C.last_start = C.t;
}

/* jdde_net_grow_ring: live anchors into the new ring (node-parallel), anchor times by thread 0 */
extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_regrow(jdn_problem const P, double * const new_times, double * const new_hist, int const new_cap)
{
	long long const i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i < P.n)
		jdn::regrow_node(P, new_hist, new_cap, i);
	if (i == 0)
		for (int k = P.ctrl->first; k <= P.ctrl->current; k++)
			new_times[k & (new_cap-1)] = P.times[k & (P.cap-1)];
}

extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_recent_state(jdn_problem const P, double const t, int const current_only, double * const out)
{
	long long const i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= P.n)
		return;
	if (current_only)
		out[i] = P.hist[2*jdn_pair_index(P.ctrl->current & (P.cap-1), i, P.n)];      /* get_current_state */
	else
		out[i] = jdn::recent_state(P, t, i);
}

/* adjust_diff, phase 1 (jdn_engine.cuh): node-parallel over the owned nodes; the two anchors go into tentative buffers 0 and 1
 * of EVERY rank (several GPUs: the peers' staging buffers — the integration is at rest, nobody else uses them) */
extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_adjust_diff(jdn_problem const P, double const time, double const new_t, double const * const amplitude)
{
	long long const k = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (k >= P.hi-P.lo)
		return;
	long long const i = P.lo+k;
	jdn::View const V = jdn::view_of(P);
	jdn::pair truncated, fresh;
	jdn::adjust_diff_node(P, V, i, time, new_t, amplitude ? amplitude[i] : 0.0, truncated, fresh);
	if (P.world > 1)
		for (int r = 0; r < P.world; r++)
		{
			reinterpret_cast<jdn::pair *>(jdn_exchange_stage(P.peers[r], 0, P.n))[i] = truncated;
			reinterpret_cast<jdn::pair *>(jdn_exchange_stage(P.peers[r], 1, P.n))[i] = fresh;
		}
	else
	{
		jdn::tentative_buffer(P, 0)[i] = truncated;
		jdn::tentative_buffer(P, 1)[i] = fresh;
	}
}

/* phase 2: all nodes (every rank keeps the whole history), after the ranks have met; extrema to out[0…n) and out[n…2n) */
extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_adjust_diff_commit(jdn_problem const P, double const time, double const new_t, double * const out)
{
	long long const i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= P.n)
		return;
	/* (written by the peers: read past L1) */
	double2 const a = __ldcg(reinterpret_cast<double2 const *>(jdn::tentative_buffer(P, 0))+i);
	double2 const b = __ldcg(reinterpret_cast<double2 const *>(jdn::tentative_buffer(P, 1))+i);
	jdn::pair truncated, fresh;
	truncated.state = a.x; truncated.diff = a.y;
	fresh.state = b.x; fresh.diff = b.y;
	jdn::adjust_diff_commit_node(P, i, new_t-time, truncated, fresh, out[i], out[P.n+i]);
}

extern "C" __global__ void jdn_k_adjust_diff_control(jdn_problem const P, double const time, double const new_t)
{
	if (blockIdx.x == 0 && threadIdx.x == 0)
		jdn::adjust_diff_control(P, time, new_t);
}

/* initial past (initiate_past_from_list, jitced_template.c:575-606): anchors [k][n] → rows of the ring */
extern "C" __global__ void __launch_bounds__(JDN_BLOCK)
jdn_k_set_past(jdn_problem const P, int const n_anchors, double const * const times, double const * const states, double const * const diffs, long long const n_edges)
{
	long long const i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i < P.n)
		for (int k = 0; k < n_anchors; k++)
		{
			P.hist[2*jdn_pair_index(k, i, P.n)] = states[(size_t)k*P.n+i];
			P.hist[2*jdn_pair_index(k, i, P.n)+1] = diffs[(size_t)k*P.n+i];
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
		C.tentative = 0;
		C.pws_pass = 0;
		C.pws_bits = C.diff_bits = 0;
	}
}
