/*
 * jdde_kernels.cu — the per-DDE kernel image. Compiled once per right-hand side by
 * jitcdde_b200/_build.py:
 *
 *   nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 [-fmad=false]
 *        -DJD_N=… -DJD_NBASIC=… -DJD_NSITES=… -DJD_NPARAMS=… -DJD_RHS_FILE="rhs_<key>.inc"
 *
 * and loaded by the runtime library with cudaLibraryLoadData (jdde_load_rhs). This is the
 * counterpart of the reference rendering jitced_template.c and building a CPython extension per
 * DDE (/root/reference/jitcdde/_jitcdde.py:560-575).
 *
 * Every kernel maps one thread to one trajectory of the ensemble; the grid covers all members.
 * Block size JD_BLOCK (default 128): blocks are the scheduling granule that evens out the
 * different amounts of work of different members, 148 SMs × several resident blocks each.
 */
#ifndef JD_BLOCK
#define JD_BLOCK 128
#endif
#include "jdde_engine.cuh"
#ifndef JD_WIN_SMEM_BYTES
#define JD_WIN_SMEM_BYTES 0
#endif
#ifndef JD_MIN_BLOCKS
#define JD_MIN_BLOCKS 1
#endif
#ifndef JD_GROUP_BLOCK
#define JD_GROUP_BLOCK 64
#endif
#ifndef JD_GROUP_MIN_BLOCKS
#define JD_GROUP_MIN_BLOCKS 1
#endif

extern "C" __device__ const jdde_image_desc_t jdde_image_desc = {
	JDDE_IMAGE_MAGIC, JDDE_ABI_VERSION, JD_N, JD_NBASIC, JD_NSITES, JD_NPARAMS, JD_A, JD_BLOCK, JD_WIN_SMEM_BYTES,
#if defined(JD_GROUP)
	JD_GROUP, JD_GROUP_BLOCK, JD_GROUP_WINDOW ? (int)((JD_GROUP_BLOCK/32)*JD_GROUPS_PER_WARP*JD_GROUP_WIN_DOUBLES*sizeof(double)) : 0
#else
	0, 0, 0
#endif
};

#define JD_MEMBER_INDEX() ((long long)blockIdx.x*blockDim.x+threadIdx.x)

extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_set_past(jdde_problem const P, int const n_anchors, double const * const times, double const * const states, double const * const diffs, int const per_member)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_set_past(P, m, n_anchors, times, states, diffs, per_member);
}

extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_reset_controller(jdde_problem const P)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_reset_controller(P, m);
}

/* THE hot kernel: integrate / step_on_discontinuities for the whole ensemble */
extern "C" __global__ void __launch_bounds__(JD_BLOCK, JD_MIN_BLOCKS)
jdde_k_integrate(jdde_problem const P, double const * const targets, int const per_member, int const n_steps, int const resume, double * const out, int const n_out, long long const budget)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_integrate(P, m, targets, per_member, n_steps, resume, out, n_out, budget);
}

#if defined(JD_GROUP)
/* the same for Lyapunov systems with a group of JD_GROUP lanes per member (jdde_group.cuh): 32/JD_GROUP members per warp */
extern "C" __global__ void __launch_bounds__(JD_GROUP_BLOCK, JD_GROUP_MIN_BLOCKS)
jdde_k_integrate_group(jdde_problem const P, double const * const targets, int const per_member, int const n_steps, int const resume, double * const out, int const n_out, long long const budget)
{
	int const lane = threadIdx.x & 31;
	int const group = lane/JD_GROUP;
	if (group >= JD_GROUPS_PER_WARP)
		return;
	long long const warp = ((long long)blockIdx.x*blockDim.x+threadIdx.x) >> 5;
	long long const m = warp*JD_GROUPS_PER_WARP+group;
	extern __shared__ double jd_group_smem[];     /* JD_GROUP_WINDOW: staging area, JD_GROUP_WIN_DOUBLES per member of the block */
	if (m < P.n_members)
		jdde::group_integrate(P, m, lane-group*JD_GROUP, group*JD_GROUP, targets, per_member, n_steps, resume, out, n_out, budget,
			jd_group_smem + (size_t)((threadIdx.x >> 5)*JD_GROUPS_PER_WARP+group)*JD_GROUP_WIN_DOUBLES);
}
#endif

extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_integrate_blindly(jdde_problem const P, double const * const targets, int const per_member, double const step, double * const out, int const n_out, double * const lyaps, double * const total_time)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_integrate_blindly(P, m, targets, per_member, step, out, n_out, lyaps, total_time);
}

#if defined(JD_PROJECT)
/* restricted / transversal Lyapunov exponents: the projector after a call of integrate, and blind integration with the
 * projector after every step (one thread per member: cold paths) */
extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_project(jdde_problem const P, jdde_projector const R, double const delay, double * const norms)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_project(P, m, R, delay, norms);
}

extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_integrate_blindly_projected(jdde_problem const P, jdde_projector const R, double const * const targets, int const per_member, double const step, double * const out, int const n_out, double * const lyaps, double * const total_time)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_integrate_blindly(P, m, targets, per_member, step, out, n_out, lyaps, total_time, &R);
}
#endif

extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_jump(jdde_problem const P, int const mode, double const * const amplitude, int const amp_pm, double const * const time, int const time_pm, double const width, int const forward, double * const minima, double * const maxima)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_jump(P, m, mode, amplitude, amp_pm, time, time_pm, width, forward, minima, maxima);
}

extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_forget(jdde_problem const P)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_forget(P, m);
}

/* what = 0: t (get_t), 1: state of the last anchor (get_current_state), 2: dt */
extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_get(jdde_problem const P, int const what, double * const out)
{
	long long const m = JD_MEMBER_INDEX();
	if (m >= P.n_members)
		return;
	double const * const ring = P.ring + (size_t)m*P.ring_stride;
	if (what == 0)
		out[m] = ring[(size_t)(P.current[m] & (P.cap-1))*JD_A];
	else if (what == 1)
	{
		double const * const a = ring + (size_t)(P.last[m] & (P.cap-1))*JD_A;
		for (int i = 0; i < JD_N; i++)
			out[(size_t)m*JD_N+i] = a[1+i];
	}
	else
		out[m] = P.dt[m];
}

extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_clear_status(jdde_problem const P, int const code)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members && P.status[m] == code)
		P.status[m] = JDDE_STATUS_OK;
}

/* number of members whose status is not OK (so that callers need not copy the whole status array back) */
extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_count_failed(jdde_problem const P, unsigned long long * const out)
{
	long long const m = JD_MEMBER_INDEX();
	bool const failed = m < P.n_members && P.status[m] != JDDE_STATUS_OK;
	unsigned const ballot = __ballot_sync(0xffffffffu, failed);
	if ((threadIdx.x & 31) == 0 && ballot)
		atomicAdd(out, (unsigned long long)__popc(ballot));
}

/* sums of accepted / rejected / right-hand-side evaluations over the ensemble: warp shuffle
 * reduction, one atomic per warp */
extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_sum_counters(jdde_problem const P, unsigned long long * const out)
{
	long long const m = JD_MEMBER_INDEX();
	unsigned long long a = 0, r = 0, f = 0;
	if (m < P.n_members)
	{
		a = (unsigned long long)P.accepted[m];
		r = (unsigned long long)P.rejected[m];
		f = (unsigned long long)(3*P.attempts[m] + P.extra_rhs[m]);
	}
	for (int o = 16; o > 0; o >>= 1)
	{
		a += __shfl_down_sync(0xffffffffu, a, o);
		r += __shfl_down_sync(0xffffffffu, r, o);
		f += __shfl_down_sync(0xffffffffu, f, o);
	}
	if ((threadIdx.x & 31) == 0)
	{
		atomicAdd(out+0, a);
		atomicAdd(out+1, r);
		atomicAdd(out+2, f);
	}
}

/* copy every member's live anchors into a ring of another capacity (jdde_grow_ring) */
extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_regrow(jdde_problem const P, double * const new_ring, int const new_cap)
{
	long long const m = JD_MEMBER_INDEX();
	if (m >= P.n_members)
		return;
	double const * const src = P.ring + (size_t)m*P.ring_stride;
	double * const dst = new_ring + (size_t)m*jdde_ring_stride(new_cap, JD_A);
	int const first = P.first[m], last = P.last[m];
	for (int k = first; k <= last; k++)
	{
		double const * const a = src + (size_t)(k & (P.cap-1))*JD_A;
		double * const b = dst + (size_t)(k & (new_cap-1))*JD_A;
		for (int i = 0; i < JD_A; i++)
			b[i] = a[i];
	}
	if (P.status[m] == JDDE_STATUS_OVERFLOW)
		P.status[m] = JDDE_STATUS_OK;
}

#if JD_NBASIC != JD_N
extern "C" __global__ void __launch_bounds__(JD_BLOCK)
jdde_k_orthonormalise(jdde_problem const P, int const n_lyap, double const delay, double * const norms, double * const old_t, double * const lyaps, double * const delta_t)
{
	long long const m = JD_MEMBER_INDEX();
	if (m < P.n_members)
		jdde::member_orthonormalise(P, m, n_lyap, delay, norms, old_t, lyaps, delta_t);
}

/* one thread block (or cluster) per member (jdde_team.cuh); dynamic shared memory: blockDim.x + 2 + JD_TEAM_FLAG_DOUBLES +
 * (capacity|1)·(1 + 2·n_basic·n_lyap) doubles. A launch handles the members whose history length lies in [count_from, count_to]
 * (jdde_runtime.cu: launch_orthonormalise); the others are sorted out blockDim.x members at a time — one thread looks at one
 * member — so that a launch that has little to do costs microseconds, not a barrier per skipped member. */
extern "C" __global__ void __launch_bounds__(JD_TEAM_BLOCK)
jdde_k_orthonormalise_team(jdde_problem const P, int const n_lyap, double const delay, double * const norms, double * const old_t, double * const lyaps, double * const delta_t, int const capacity, int const count_from, int const count_to)
{
	extern __shared__ double jd_team_smem[];      /* [blockDim.x] scratch, [2] exchange slots of the cluster, the flags, then the staged anchors */
	unsigned crank, csize;
	asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
	asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
	jdde::Team T = { (int)threadIdx.x, (int)blockDim.x, jd_team_smem };
	T.crank = (int)crank;
	T.csize = (int)csize;
	unsigned char * const wanted = reinterpret_cast<unsigned char *>(jd_team_smem+blockDim.x+2);
	double * const staged = jd_team_smem+blockDim.x+2+JD_TEAM_FLAG_DOUBLES;
	/* the blocks of a cluster run the same members (and come to the same flags) */
	long long const n_teams = gridDim.x/csize, team = blockIdx.x/csize;
	for (long long base = team; base < P.n_members; base += n_teams*blockDim.x)
	{
		long long const candidate = base + (long long)threadIdx.x*n_teams;
		wanted[threadIdx.x] = candidate < P.n_members && jdde::team_has_work(P, candidate, old_t, count_from, count_to);
		__syncthreads();
		for (int k = 0; k < (int)blockDim.x; k++)
		{
			if (!wanted[k])
				continue;
			jdde::team_member_orthonormalise(T, P, base + (long long)k*n_teams, n_lyap, delay, norms, old_t, lyaps, delta_t, staged, capacity, count_from, count_to);
			__syncthreads();
			T.parity = 0;
			if (csize > 1)
				jdde::Team::cluster_sync();     /* nobody reads this block's exchange slots any more */
		}
		__syncthreads();
	}
}
#endif
