/*
 * jdde_group.cuh — the adaptive stepper for Lyapunov systems with a GROUP OF LANES per member
 * (kernel jdde_k_integrate_group): lane 0 of a group integrates the basic system, lane 1+i the i-th
 * separation function (tangent vector), JD_GROUP = 1 + n_lyap lanes in all, 32/JD_GROUP members per warp.
 *
 * Why: the extended system of jitcdde_lyap (/root/reference/jitcdde/_jitcdde.py:1046-1139) has
 * n = n_basic·(1+n_lyap) components (two Rösslers with 4 exponents: 30). One thread per member needs
 * 8 arrays of n doubles for a Bogacki–Shampine attempt — 480 registers at n = 30, i.e. local memory and
 * 4–8 resident warps per SM (profiles/configs_r01.md: ≈ 0.1 instructions per cycle and warp). The tangent
 * blocks are the SAME expressions with shifted indices, so the work splits evenly over 1+n_lyap lanes that
 * own n_basic components each and keep them in registers.
 *
 * What the lanes share:
 *   - the basic lane's stage argument y (needed by the Jacobian entries in the tangent blocks): broadcast by
 *     warp shuffle at the start of every right-hand-side evaluation;
 *   - the history: every lane of the group makes the same bracket search with its own copy of the cursors
 *     (same addresses → one transaction per warp), then interpolates the components it needs — the basic
 *     ones it reads are read by all lanes of the group from the same two anchors;
 *   - the error norm: max over the lanes (shuffle), after which every lane takes the same accept/reject and
 *     step-size decision with its own copy of the controller — no lane waits for a "leader";
 *   - the tentative anchor: every lane stores its own block, lane 0 the time; __syncwarp on the group's
 *     lanes orders these stores against the lookups of the next attempt (past within step) and the previous
 *     lookups against the store (replace_last_anchor, jitced_template.c:106-114).
 * Groups of one warp are NOT in lockstep with each other: every group runs its own adaptive loop; shuffles
 * and barriers name only the group's lanes.
 *
 * Arithmetic: per component exactly that of the one-thread-per-member engine (same generated statements,
 * same order), so the verification build is bit-identical to the reference core, too.
 *
 * The generated file (jitcdde_b200/_codegen.py, _generate_blocks) has the shape
 *     <lookups: JDDE_SEG s_k = JDDE_ANCHORS(k, …);>          — executed by every lane
 *     #ifdef JDDE_GROUP   if (JDDE_GROUP_IS_BASIC) { basic statements } else { tangent statements }
 *     #else               { basic } { tangent, JDDE_TOFF = n_basic } { tangent, JDDE_TOFF = 2·n_basic } …
 * with the tangent statements written once in the vocabulary JDDE_TY / JDDE_PAST_TY / JDDE_PAST_TDY /
 * JDDE_SET_TDY (own block, indices relative to it).
 *
 * Device only. The host twin (tests/host_twin) runs the one-thread engine; the group kernel is checked against the
 * oracle on the GPU (tests/test_gpu_parity.py).
 */
#ifndef JDDE_GROUP_CUH
#define JDDE_GROUP_CUH

#if defined(JD_GROUP) && defined(__CUDACC__)

#if JD_WINDOW || JD_SEG_BY_VALUE
#error "JD_GROUP needs the by-pointer anchor access (larger systems)"
#endif
#if JD_GROUP*JD_NBASIC != JD_N || JD_GROUP > 32
#error "JD_GROUP must be 1 + n_lyap with n = n_basic·(1+n_lyap)"
#endif

namespace jdde {

#define JD_GROUPS_PER_WARP (32/JD_GROUP)

/*
 * Shared-memory staging of the anchors a step can reach (JD_GROUP_WINDOW, the counterpart of JD_WINDOW=2 of the
 * one-thread engine): per member and lookup site a circular buffer of 4 whole anchors in shared memory holds the
 * anchors cursor−1 … cursor+2 (slot = index & 3). The group's lanes fill it together with cp.async — 16-byte
 * chunks dealt round-robin to the lanes, global → shared without passing through registers — one cursor move
 * ahead of their use, in either direction. A lookup then walks over times in shared memory and hands out
 * pointers into it; it waits (cp.async.wait_all + __syncwarp on the group) only when the cursor moves onto an
 * anchor whose copy may still be in flight, by which time the copy has had a whole step to arrive. Without it
 * every lookup is two to three DEPENDENT trips to DRAM (anchor times, then the data of the pair found): the
 * anchors around t−τ were written hundreds of steps ago (profiles/: 45 % of all stall samples).
 * Only accepted anchors (index <= current) are staged: they are immutable while the kernel runs. All lanes of a
 * group hold identical cursors and flags, so the control flow below is uniform within the group.
 */
#ifndef JD_GROUP_WINDOW
#define JD_GROUP_WINDOW 0
#endif
/* per member; + 16 bytes so that the areas of the members of a warp start in different banks (a multiple of 128 bytes
 * put equal offsets of all members into the same bank: 6-way conflicts on every access, profiles/) */
#define JD_GROUP_WIN_DOUBLES ((JD_NSITES > 0 ? JD_NSITES : 1)*4*JD_A + 2)

struct GWindow
{
	bool valid;
	bool fresh;        /* the copies of (cursor, cursor+1) have been requested but not yet awaited */
	bool prev_ok;
	bool next_ok;
};

struct GMember
{
	static constexpr int NL = JD_NBASIC;        /* components this lane owns */
	static constexpr bool PREFETCH = !JD_GROUP_WINDOW;
	double * ring;
	int mask;
	int first, last, current;
	int cursor[JD_NSITES > 0 ? JD_NSITES : 1];
	double par[JD_NPARAMS > 0 ? JD_NPARAMS : 1];

	double cur_t, cur_y[NL], cur_d[NL];
	double new_t, new_y[NL], new_d[NL];
	double err[NL];
	double pws;
	double last_start;
	double dt;
	bool pws_pending;
	long long index;
	int attempts, rejected, extra_rhs;
	int status;

	unsigned gmask;         /* the lanes of this group */
	int base;               /* lane of the basic system */
	int block;              /* 0: basic system, 1+i: separation function i */
	int off;                /* block·n_basic: first component of this lane */
	double last_pws, credit;    /* controller memory (cold), in registers for the launch */
	int count;
#if JD_GROUP_WINDOW
	GWindow win[JD_NSITES > 0 ? JD_NSITES : 1];
	double * win_base;      /* this member's staging area in shared memory */

	__device__ __forceinline__ double * win_slot(int const site, int const index) const { return win_base + (size_t)(site*4 + (index & 3))*JD_A; }
#endif

	__device__ __forceinline__ double * slot(int const index) const { return ring + (size_t)(index & mask)*JD_A; }
	__device__ __forceinline__ void invalidate_windows()
	{
#if JD_GROUP_WINDOW
		JD_UNROLL
		for (int s = 0; s < JD_NSITES; s++)
			win[s].valid = false;
#endif
	}

	/* the lines of an anchor this lane will read: its own state and diff blocks (lane 0: with the time) */
	__device__ __forceinline__ void prefetch(int const index) const
	{
		double const * const a = slot(index);
		asm volatile("prefetch.global.L1 [%0];" :: "l"(a+off));
		asm volatile("prefetch.global.L1 [%0];" :: "l"(a+off+NL));
		asm volatile("prefetch.global.L1 [%0];" :: "l"(a+1+JD_N+off));
		asm volatile("prefetch.global.L1 [%0];" :: "l"(a+JD_N+off+NL));
	}
};

__device__ __forceinline__ double & ctl_last_pws(GMember & M, jdde_problem const &) { return M.last_pws; }
__device__ __forceinline__ int & ctl_count(GMember & M, jdde_problem const &) { return M.count; }
__device__ __forceinline__ double & ctl_credit(GMember & M, jdde_problem const &) { return M.credit; }

#if JD_GROUP_WINDOW
/* this lane's share of the copy of one anchor: chunks block, block+JD_GROUP, … of JD_A/2 16-byte chunks */
__device__ __forceinline__ void gwin_fetch(GMember const & M, int const site, int const index)
{
	double const * const src = M.slot(index);
	double * const dst = M.win_slot(site, index);
	for (int c = M.block; c < JD_A/2; c += JD_GROUP)
	{
		unsigned const d = (unsigned)__cvta_generic_to_shared(dst+2*c);
		asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src+2*c) : "memory");
	}
}

/* every lane's copies have landed and are visible to the group */
__device__ __forceinline__ void gwin_wait(GMember const & M)
{
	asm volatile("cp.async.wait_all;" ::: "memory");
	__syncwarp(M.gmask);
}

__device__ __forceinline__ void fill_window(GMember & M, int const site)
{
	GWindow & W = M.win[site];
	int const ca = M.cursor[site];
	W.valid = ca+1 <= M.current;
	if (!W.valid)
		return;
	__syncwarp(M.gmask);          /* nobody reads the staging area any more (pointers handed out by earlier lookups) */
	gwin_fetch(M, site, ca);
	gwin_fetch(M, site, ca+1);
	W.fresh = true;
	W.next_ok = ca+2 <= M.current;
	if (W.next_ok)
		gwin_fetch(M, site, ca+2);
	W.prev_ok = ca-1 >= M.first;
	if (W.prev_ok)
		gwin_fetch(M, site, ca-1);
}

/* get_past_anchors (jitced_template.c:193-217) on the staged anchors: the same comparisons as walk(), on the same
 * values. Invariant: the copies of cursor and cursor+1 have landed (unless W.fresh); those of cursor−1 and cursor+2
 * may be in flight and are awaited when the cursor moves onto them. */
__device__ __forceinline__ Seg anchors(GMember & M, int const site, double const t)
{
	GWindow & W = M.win[site];
	Seg s;
	if (W.valid)
	{
		if (W.fresh)
		{
			gwin_wait(M);
			W.fresh = false;
		}
		int ca = M.cursor[site];
		double tv = M.win_slot(site, ca)[0];
		double tw = M.win_slot(site, ca+1)[0];
		bool hit = true;
		/* while (ca->time > t && ca->previous) ca = ca->previous; */
		while (tv > t && ca > M.first)
		{
			if (!W.prev_ok)
			{
				hit = false;
				break;
			}
			--ca;
			W.next_ok = true;
			W.prev_ok = ca-1 >= M.first;
			gwin_wait(M);         /* also: every lane is done with the slot that is overwritten next */
			tw = tv;
			tv = M.win_slot(site, ca)[0];
			if (W.prev_ok)
				gwin_fetch(M, site, ca-1);      /* into the slot of the old cursor+2 */
		}
		/* while (ca->next->time < t && ca->next->next) ca = ca->next; */
		while (hit && tw < t)
		{
			if (!W.next_ok)
			{
				hit = false;      /* next anchor is tentative or missing: let walk() decide */
				break;
			}
			++ca;
			W.prev_ok = true;
			W.next_ok = ca+2 <= M.current;
			gwin_wait(M);
			tv = tw;
			tw = M.win_slot(site, ca+1)[0];
			if (W.next_ok)
				gwin_fetch(M, site, ca+2);      /* into the slot of the old cursor−1 */
		}
		M.cursor[site] = ca;
		if (hit)
		{
			/* t <= time of an accepted anchor <= current.time: no past-within-step contribution */
			s.v = M.win_slot(site, ca);
			s.w = M.win_slot(site, ca+1);
			s.tv = tv;
			s.tw = tw;
			return s;
		}
	}
	walk(M, site, t, s.v, s.w, s.tv, s.tw);
	fill_window(M, site);
	return s;
}
#endif

__device__ __forceinline__ void load_member(jdde_problem const & P, long long const m, GMember & M)
{
	M.mask = P.cap-1;
	M.index = m;
	M.ring = P.ring + (size_t)m*P.ring_stride;
	M.first = P.first[m];
	M.last = P.last[m];
	M.current = P.current[m];
	JD_UNROLL
	for (int s = 0; s < JD_NSITES; s++)
		M.cursor[s] = P.cursor[(size_t)s*P.n_members+m];
	JD_UNROLL
	for (int k = 0; k < JD_NPARAMS; k++)
		M.par[k] = P.par[(size_t)k*P.n_members+m];
	double const * const c = M.slot(M.current);
	M.cur_t = c[0];
	JD_UNROLL
	for (int i = 0; i < GMember::NL; i++)
	{
		M.cur_y[i] = c[1+M.off+i];
		M.cur_d[i] = c[1+JD_N+M.off+i];
	}
	M.pws = 0.0;
	M.last_start = P.last_start[m];
	M.dt = P.dt[m];
	M.last_pws = P.last_pws[m];
	M.credit = P.credit[m];
	M.count = P.count[m];
	M.pws_pending = M.last_pws != 0.0;
	M.rejected = M.attempts = M.extra_rhs = 0;
	M.status = P.status[m];
#if JD_GROUP_WINDOW
	JD_UNROLL
	for (int s = 0; s < JD_NSITES; s++)
		fill_window(M, s);              /* issued now, needed at the first lookup */
#endif
}

/* all lanes hold the same controller state; the basic lane writes it back */
__device__ __forceinline__ void store_member(jdde_problem const & P, long long const m, GMember const & M)
{
	if (M.block != 0)
		return;
	P.first[m] = M.first;
	P.last[m] = M.last;
	P.current[m] = M.current;
	JD_UNROLL
	for (int s = 0; s < JD_NSITES; s++)
		P.cursor[(size_t)s*P.n_members+m] = M.cursor[s];
	P.last_start[m] = M.last_start;
	P.dt[m] = M.dt;
	P.last_pws[m] = M.last_pws;
	P.credit[m] = M.credit;
	P.count[m] = M.count;
	P.rejected[m] += M.rejected;
	P.attempts[m] += M.attempts;
	P.extra_rhs[m] += M.extra_rhs;
	P.status[m] = M.status;
}

/* eval_f for one lane: y and dY are this lane's block */
__device__ __forceinline__ void eval_f(GMember & M, double const t, double const * const y, double * const dY)
{
	double yb[JD_NBASIC];
	JD_UNROLL
	for (int k = 0; k < JD_NBASIC; k++)
		yb[k] = __shfl_sync(M.gmask, y[k], M.base);
#define JDDE_GROUP 1
#define JDDE_GROUP_IS_BASIC (M.block == 0)
#undef JDDE_Y
#define JDDE_Y(i) (yb[i])
#define JDDE_TY(k) (y[k])
#define JDDE_PAST_TY(tm, k, s) ::jdde::past_y(tm, M.off+(k), s)
#define JDDE_PAST_TDY(tm, k, s) ::jdde::past_dy(tm, M.off+(k), s)
#define JDDE_SET_TDY(c, value) (dY[c] = (value))
	#include JD_RHS_FILE
#undef JDDE_GROUP
#undef JDDE_GROUP_IS_BASIC
#undef JDDE_Y
#define JDDE_Y(i) (y[i])
#undef JDDE_TY
#undef JDDE_PAST_TY
#undef JDDE_PAST_TDY
#undef JDDE_SET_TDY
}

/* get_next_step, jitced_template.c:421-472 */
__device__ __forceinline__ void get_next_step(GMember & M, double const delta_t)
{
	constexpr int NL = GMember::NL;
	M.last_start = M.cur_t;
	M.pws = 0.0;
	M.attempts++;

	double argument[NL];
	double k_2[NL];
	JD_UNROLL
	for (int i = 0; i < NL; i++)
		argument[i] = M.cur_y[i] + 0.5*delta_t*M.cur_d[i];
	eval_f(M, M.cur_t+0.5*delta_t, argument, k_2);

	double k_3[NL];
	JD_UNROLL
	for (int i = 0; i < NL; i++)
		argument[i] = M.cur_y[i] + 0.75*delta_t*k_2[i];
	eval_f(M, M.cur_t+0.75*delta_t, argument, k_3);

	double ny[NL], nd[NL];
	JD_UNROLL
	for (int i = 0; i < NL; i++)
		ny[i] = M.cur_y[i] + divide_by<9>(delta_t) * (2*M.cur_d[i]+3*k_2[i]+4*k_3[i]);

	double const nt = M.cur_t + delta_t;
	eval_f(M, nt, ny, nd);

	JD_UNROLL
	for (int i = 0; i < NL; i++)
		M.err[i] = (5*M.cur_d[i]-6*k_2[i]-8*k_3[i]+9*nd[i]) * divide_by<72>(delta_t);

	/* every lookup of this attempt (which may have read the anchor that is replaced now) has been made */
	__syncwarp(M.gmask);
	if (M.last == M.current)
		M.last++;
	M.new_t = nt;
	double * const a = M.slot(M.last);
	if (M.block == 0)
	{
		a[0] = nt;
		for (int k = 1+2*JD_N; k < JD_A; k++)
			a[k] = 0.0;
	}
	JD_UNROLL
	for (int i = 0; i < NL; i++)
	{
		M.new_y[i] = ny[i];
		M.new_d[i] = nd[i];
		a[1+M.off+i] = ny[i];
		a[1+JD_N+M.off+i] = nd[i];
	}
	/* … and the new anchor is visible to the lookups of the group's other lanes */
	__syncwarp(M.gmask);
}

/* get_p, jitced_template.c:474-498: maximum over this lane's components, then over the group's lanes */
__device__ __forceinline__ double get_p(GMember const & M, double const atol, double const rtol)
{
	double const p = local_p<GMember::NL>(M.err, M.new_y, atol, rtol);
	double total = 0.0;
	JD_UNROLL
	for (int j = 0; j < JD_GROUP; j++)
	{
		double const other = __shfl_sync(M.gmask, p, M.base+j);
		if (other > total)
			total = other;
	}
	return total;
}

/* check_new_y_diff, jitced_template.c:502-524 */
__device__ __forceinline__ bool check_new_y_diff(GMember const & M, double const * const old_y, double const atol, double const rtol)
{
	bool result = true;
	JD_UNROLL
	for (int i = 0; i < GMember::NL; i++)
	{
		double const difference = fabs(M.new_y[i] - old_y[i]);
		double const tolerance = atol + fabs(rtol*M.new_y[i]);
		result &= (tolerance >= difference);
	}
	return __all_sync(M.gmask, result);
}

/* accept_step, jitced_template.c:526-532 */
__device__ __forceinline__ void accept_step(GMember & M)
{
	M.current = M.last;
	M.cur_t = M.new_t;
	JD_UNROLL
	for (int i = 0; i < GMember::NL; i++)
	{
		M.cur_y[i] = M.new_y[i];
		M.cur_d[i] = M.new_d[i];
	}
}

/* get_recent_state, jitced_template.c:287-320: every lane interpolates its own components */
__device__ __forceinline__ void get_recent_state(GMember const & M, double const t, double * const out, int const n_out)
{
	double const * const w = M.slot(M.last);
	double const * const v = M.slot(M.last-1);
	double const q = w[0]-v[0];
	double const x = (t - v[0])/q;
	JD_UNROLL
	for (int i = 0; i < GMember::NL; i++)
	{
		int const index = M.off+i;
		double const a = v[1+index];
		double const b = v[1+JD_N+index] * q;
		double const c = w[1+index];
		double const d = w[1+JD_N+index] * q;
		if (index < n_out)
			out[index] = (1-x) * ( (1-x) * (b*x + (a-c)*(2*x+1)) - d*x*x) + c;
	}
}

/* member_integrate (jdde_engine.cuh) for one lane of a group */
__device__ __forceinline__ void group_integrate(jdde_problem const & P, long long const m, int const lane_in_group, int const base_lane,
	double const * const targets, int const per_member, int const n_steps, int const resume, double * const out, int const n_out, long long const budget, double * const staging)
{
	GMember M;
#if JD_GROUP_WINDOW
	M.win_base = staging;
#else
	(void)staging;
#endif
	M.block = lane_in_group;
	M.off = lane_in_group*JD_NBASIC;
	M.base = base_lane;
	M.gmask = ((1u << JD_GROUP)-1u) << base_lane;
	load_member(P, m, M);
	if (M.status != JDDE_STATUS_OK)
		return;
	int const current_at_entry = M.current;
	double target;
	if (n_steps == 0)
	{
		target = per_member ? targets[m] : targets[0];
		step_until(M, P, target, false, budget);
	}
	else
	{
		int const resume_index = resume ? P.resume_index[m] : -1;
		target = resume ? P.resume_target[m] : M.cur_t;
		double const * const steps = per_member ? targets+(size_t)m*n_steps : targets;
		int s = resume ? resume_index : 0;
		/* the resume data have been read by every lane before the basic lane overwrites them below */
		__syncwarp(M.gmask);
		for (; s < n_steps; s++)
		{
			if (s != resume_index)
				target = M.cur_t + steps[s];
			step_until(M, P, target, true, budget);
			if (M.status != JDDE_STATUS_OK)
				break;
		}
		if (M.block == 0)
		{
			P.resume_index[m] = s;
			P.resume_target[m] = target;
		}
	}
	if (M.status == JDDE_STATUS_OK)
	{
		if (out)
			get_recent_state(M, target, out+(size_t)m*n_out, n_out);
		forget(M, P.max_delay);
	}
	if (M.block == 0)
		P.accepted[m] += M.current - current_at_entry;
	store_member(P, m, M);
}

} // namespace jdde

#endif
#endif
