/*
 * jdde_engine.cuh — the DDE integration engine: one trajectory ("member") per thread.
 *
 * Hand-written FP64 device code for the Shampine–Thompson method as implemented by
 * neurophysik/jitcdde: Bogacki–Shampine 3(2) stages with FSAL, cubic-Hermite history,
 * embedded error estimate, GSL-style step-size control, past-within-step iteration,
 * discontinuity stepping, jumps, dense output, forget, and Gram–Schmidt
 * orthonormalisation of separation functions for Lyapunov exponents.
 *
 * Reference behaviour reproduced here (file:line under /root/reference/jitcdde/):
 *   history + lookup      jitced_template.c:25-54, 56-114, 193-251
 *   step                  jitced_template.c:398-472
 *   error norm / accept   jitced_template.c:474-532
 *   forget / output       jitced_template.c:287-320, 534-561
 *   jumps                 jitced_template.c:253-285, 1089-1174
 *   Lyapunov              jitced_template.c:657-878
 *   controller            _jitcdde.py:621-1000 (runs INSIDE the kernel here; the reference
 *                         crosses Python↔C 5–7 times per step for it)
 *
 * Design for B200:
 *   - whole `integrate(target)` for an ensemble is ONE kernel launch; every thread runs its own
 *     adaptive loop (own step size, own accept/reject, no lockstep between trajectories);
 *   - the working set of a step — current anchor, tentative anchor, controller state, search
 *     cursors — lives in registers for the whole launch; HBM sees one anchor store per attempt
 *     and the anchor loads of the delayed lookups;
 *   - the history is a power-of-two ring per member in HBM (jdde_abi.h), addressed by logical
 *     indices; the bracket search keeps one cursor per lookup site and reproduces the reference's
 *     walk including its tie-breaks;
 *   - JD_WINDOW: staging of the anchors a step can reach. Per lookup site the anchors cursor−1 …
 *     cursor+2 are kept close to the thread and refilled one cursor move ahead of their use, so a
 *     lookup that stays in its segment or moves by one anchor never waits for L2 or HBM:
 *       2  in shared memory, filled by cp.async, handed out by value (anchors of up to 8 doubles);
 *       3  in shared memory, contiguous per thread, handed out as pointers (larger anchors);
 *       1  in registers (kept for comparison; measured slower);
 *       0  none: speculative vector loads (small anchors) or the plain walk with an L1 prefetch of
 *          the next anchor. Lyapunov systems use a group of lanes per member with its own staging
 *          (jdde_group.cuh). Only accepted anchors (index <= current) are staged: they are
 *          immutable while the kernel runs;
 *   - arithmetic keeps the reference's association order; with -fmad=false the results are
 *     bit-identical to the CPU oracle (tests/), with FMA contraction they agree to rounding.
 *
 * The functions are __host__ __device__ so that tests/ can compile the same engine with g++ and
 * compare it with the oracle without a GPU; the product only ever runs the CUDA build.
 */
#ifndef JDDE_ENGINE_CUH
#define JDDE_ENGINE_CUH

#include <math.h>
#include "jdde_math.h"
#include "jdde_abi.h"

#if !defined(JD_N) || !defined(JD_NSITES) || !defined(JD_NPARAMS) || !defined(JD_RHS_FILE)
#error "JD_N, JD_NSITES, JD_NPARAMS and JD_RHS_FILE must be defined by the generated translation unit"
#endif
#ifndef JD_NBASIC
#define JD_NBASIC JD_N
#endif

#define JD_A ((((1+2*JD_N))+3)&~3)
#ifndef JD_WINDOW
#define JD_WINDOW 0
#endif
/* Lookups hand anchor pairs around by value (registers) when anchors are small, by pointer otherwise */
#define JD_SEG_BY_VALUE (JD_WINDOW == 1 || JD_WINDOW == 2 || JD_WINDOW == 4 || JD_A <= 8)
#define JD_NORM_THRESHOLD (1e-30)     /* jitced_template.c:15 */

#if defined(__CUDACC__)
#  define JD_FN __host__ __device__ __forceinline__
#  define JD_UNROLL _Pragma("unroll")
#else
#  define JD_FN inline
#  define JD_UNROLL
#endif

namespace jdde {

/* 16-byte vector access to anchors (anchors are 32-byte aligned: LDG.128 / STG.128) */
struct alignas(16) d2 { double x, y; };

JD_FN void load_anchor(double const * const p, double (&a)[JD_A])
{
	JD_UNROLL
	for (int k = 0; k < JD_A; k += 2)
	{
		d2 const v = *reinterpret_cast<d2 const *>(p+k);
		a[k] = v.x;
		a[k+1] = v.y;
	}
}

JD_FN void store_anchor(double * const p, double const (&a)[JD_A])
{
	JD_UNROLL
	for (int k = 0; k < JD_A; k += 2)
	{
		d2 v;
		v.x = a[k];
		v.y = a[k+1];
		*reinterpret_cast<d2 *>(p+k) = v;
	}
}

/* pair of anchors bracketing a time (what `anchors(t)` returns in the reference: anchor *);
 * element 0 of an anchor is its time, 1..n the state, n+1..2n the derivative */
#if JD_SEG_BY_VALUE
struct Seg
{
	double v[JD_A];
	double w[JD_A];
};
JD_FN double seg_tv(Seg const & s) { return s.v[0]; }
JD_FN double seg_tw(Seg const & s) { return s.w[0]; }
#else
struct Seg
{
	const double * v;
	const double * w;
	double tv;
	double tw;
};
JD_FN double seg_tv(Seg const & s) { return s.tv; }
JD_FN double seg_tw(Seg const & s) { return s.tw; }
#endif

#if JD_WINDOW == 3
/* Shared-memory window for anchors too large for registers (device only; 8 < JD_A): as JD_WINDOW == 2 — per thread and
 * lookup site the anchors cursor−1 … cursor+2 in shared memory, filled by cp.async one cursor move ahead — but the
 * anchors lie CONTIGUOUSLY in the thread's private area and lookups hand out pointers into it (the by-pointer Seg), so
 * that the Hermite formulas read the components they need from shared memory instead of making two to three
 * dependent trips to L2/DRAM per lookup (the neutral ensemble: 20 of 33 cycles per issued instruction on the long
 * scoreboard, profiles/). The areas of consecutive threads are an odd number of 8-byte words apart: equal offsets of
 * the 32 lanes fall into different banks. */
#if JD_A <= 8
#error "JD_WINDOW == 3 is for anchors of more than 8 doubles"
#endif
struct Window
{
	bool valid;
	bool fresh;        /* the copies of (cursor, cursor+1) have been requested but not yet awaited */
	bool prev_ok;
	bool next_ok;
};
#ifndef JD_BLOCK
#define JD_BLOCK 128
#endif
#define JD_PWIN_DOUBLES ((JD_NSITES > 0 ? JD_NSITES : 1)*4*JD_A + 1)     /* per thread */
#define JD_WIN_SMEM_BYTES (JD_PWIN_DOUBLES*8*JD_BLOCK)
#elif JD_WINDOW == 2
/* Shared-memory window (device only): per thread and lookup site a private 4-anchor circular buffer in shared
 * memory holds the anchors cursor-1 … cursor+2 (slot = index & 3). Anchors are brought in with cp.async
 * (global → shared without passing through registers) one move ahead of their use, so a lookup that stays in
 * its segment or moves by one anchor in either direction reads shared memory only and never waits for L2 or
 * HBM; registers hold just three flags per site. Only accepted anchors (index <= current) are staged: they are
 * immutable while the kernel runs. This is the staging of `north_star`, without the register cost of JD_WINDOW=1. */
struct Window
{
	bool valid;
	bool fresh;        /* the copies of (cursor, cursor+1) have been requested but not yet awaited */
	bool prev_ok;
	bool next_ok;
};
#ifndef JD_BLOCK
#define JD_BLOCK 128
#endif
#define JD_WIN_SMEM_BYTES ((JD_NSITES > 0 ? JD_NSITES : 1)*4*JD_A*8*JD_BLOCK)
#elif JD_WINDOW == 4
/* Shared-memory window (device only): per thread and lookup site a private circular buffer of JD_WSLOTS = 8 anchors in
 * shared memory (slot = index & 7) holds the ACCEPTED anchors hi−7 … hi around the cursor; in the steady state
 * hi = cursor + JD_WAHEAD, i.e. cursor−3 … cursor+4. Anchors are brought in with cp.async (global → shared without
 * passing through registers), one commit group per anchor, always in increasing index order: a forward move of the
 * cursor asks for ONE new anchor, four moves ahead of its use, and waits (cp.async.wait_group 3) only for the anchor it
 * steps onto — whose copy was requested three moves earlier; the three younger copies stay in flight. Moving back (the
 * retry after a rejected step looks up earlier times) finds the three anchors behind the cursor still in place and
 * never waits. (The first version kept cursor−1 … cursor+2 and waited for ALL copies at every move: two moves within
 * one attempt, or one move back, paid a full memory round trip — 10 % of all stall samples, profiles/.)
 * Only accepted anchors (index <= current) are staged: they are immutable while the kernel runs.
 * Layout: an anchor's 1+2n doubles are (1+2n)/2 pairs of 16 bytes plus one single. Rows of 16-byte elements, one per
 * lane ([row][JD_BLOCK]: consecutive lanes at consecutive elements, conflict-free LDS.128): first the pair rows of all
 * slots, then rows that hold the singles of two neighbouring slots each — every address is the thread's base address
 * plus a function of the slot, nothing else to keep in registers. 24 bytes per anchor for n = 1 instead of the padded
 * 32, 192 bytes per thread. This is the staging of `north_star`. */
#define JD_WSLOTS 8
#define JD_WAHEAD 4
#define JD_WPAIRS ((1+2*JD_N)/2)
struct Window
{
	int hi;            /* highest staged index; all of hi−7 … hi are staged (requested in this order) */
	int sync;          /* highest index whose copy is known to have landed */
};
#ifndef JD_BLOCK
#define JD_BLOCK 128
#endif
#define JD_WIN_PAIR_ROWS ((JD_NSITES > 0 ? JD_NSITES : 1)*JD_WSLOTS*JD_WPAIRS)
#define JD_WIN_SMEM_BYTES ((JD_WIN_PAIR_ROWS + (JD_NSITES > 0 ? JD_NSITES : 1)*JD_WSLOTS/2)*16*JD_BLOCK)
#elif JD_WINDOW
/* register-resident window of one lookup site: the anchor pair (cursor, cursor+1) plus the prefetched
 * neighbours cursor-1 and cursor+2, so that the cursor can move one anchor in either direction (forward:
 * time advances; backward: retry with a smaller step after a rejection) without waiting for memory */
struct Window
{
	bool valid;
	bool prev_ok;
	bool next_ok;
	double p[JD_A];
	double v[JD_A];
	double w[JD_A];
	double n[JD_A];
};
#endif

/* everything one trajectory needs while a kernel runs; lives in registers / local memory */
struct Member
{
	static constexpr int NL = JD_N;             /* components this thread owns: all of them */
	static constexpr bool PREFETCH = true;
	double * ring;
	int mask;                                   /* ring capacity - 1 */
	int first, last, current;
	int cursor[JD_NSITES > 0 ? JD_NSITES : 1];
#if JD_WINDOW
	Window win[JD_NSITES > 0 ? JD_NSITES : 1];
#endif
#if JD_WINDOW == 4
	unsigned wbase;                             /* index of this thread's element in every row of the window (shared memory) */
#endif
	double par[JD_NPARAMS > 0 ? JD_NPARAMS : 1];

	double cur_t, cur_y[JD_N], cur_d[JD_N];     /* last accepted anchor (`current`) */
	double new_t, new_y[JD_N], new_d[JD_N];     /* tentative anchor (`last_anchor` while last != current) */
	double err[JD_N];
	double pws;                                 /* past_within_step */
	double last_start;                          /* last_actual_step_start */

	double dt;
	/* Past-within-step memory of the controller (last_pws, count, increase_credit; _jitcdde.py:727-741) is only
	 * touched when a delay is shorter than a step. It stays in the per-member arrays in global memory and is
	 * accessed by the cold paths directly; the hot loop only carries this flag (last_pws != False). */
	bool pws_pending;
	int index;                                  /* this member's index in the per-member arrays (jdde_create: n_members < 2^31) */

	/* Counters since load_member, added to the 64-bit totals by store_member. Only `attempts` is touched in
	 * the hot path; accepted steps are recovered from the advance of `current`. */
	int attempts, rejected, extra_rhs;
	int status;

	JD_FN double * slot(int const index) const { return ring + (size_t)(index & mask)*JD_A; }

	/* bring an anchor into L1 ahead of its use (walk) */
	JD_FN void prefetch(int const index) const
	{
#if defined(__CUDA_ARCH__)
		double const * const a = slot(index);
		JD_UNROLL
		for (int k = 0; k < JD_A; k += 16)
			asm volatile("prefetch.global.L1 [%0];" :: "l"(a+k));
#else
		(void)index;
#endif
	}

	JD_FN void invalidate_windows()
	{
#if JD_WINDOW
		JD_UNROLL
		for (int s = 0; s < JD_NSITES; s++)
#if JD_WINDOW == 4
			win[s].hi = -0x40000000;
#else
			win[s].valid = false;
#endif
#endif
	}
};

#if JD_WINDOW
JD_FN void fill_window(Member & M, int const site);
#endif
#if JD_WINDOW == 4
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void win_init(Member & M);
#else
inline void win_init(Member &) {}
#endif
#endif

/* Past-within-step memory of the controller (cold). One thread per member: the per-member arrays in global memory.
 * (A group of lanes per member keeps it in registers instead, jdde_group.cuh.) */
JD_FN double & ctl_last_pws(Member const & M, jdde_problem const & P) { return P.last_pws[M.index]; }
JD_FN int & ctl_count(Member const & M, jdde_problem const & P) { return P.count[M.index]; }
JD_FN double & ctl_credit(Member const & M, jdde_problem const & P) { return P.credit[M.index]; }

/* ------------------------------------------------------------------ load / store */

JD_FN void load_member(jdde_problem const & P, long long const m, Member & M)
{
	M.mask = P.cap-1;
	M.index = (int)m;
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
	double c[JD_A];
	load_anchor(M.slot(M.current), c);
	M.cur_t = c[0];
	JD_UNROLL
	for (int i = 0; i < JD_N; i++)
	{
		M.cur_y[i] = c[1+i];
		M.cur_d[i] = c[1+JD_N+i];
	}
#if JD_WINDOW == 4
	win_init(M);
#endif
#if JD_WINDOW
	JD_UNROLL
	for (int s = 0; s < JD_NSITES; s++)
		fill_window(M, s);              /* issued now, needed at the first lookup */
#endif
	M.pws = 0.0;
	M.last_start = P.last_start[m];
	M.dt = P.dt[m];
	M.pws_pending = P.last_pws[m] != 0.0;
	M.rejected = M.attempts = M.extra_rhs = 0;
	M.status = P.status[m];
}

JD_FN void store_member(jdde_problem const & P, long long const m, Member const & M)
{
	P.first[m] = M.first;
	P.last[m] = M.last;
	P.current[m] = M.current;
	JD_UNROLL
	for (int s = 0; s < JD_NSITES; s++)
		P.cursor[(size_t)s*P.n_members+m] = M.cursor[s];
	P.last_start[m] = M.last_start;
	P.dt[m] = M.dt;
	P.rejected[m] += M.rejected;
	P.attempts[m] += M.attempts;
	P.extra_rhs[m] += M.extra_rhs;
	P.status[m] = M.status;
}

/* ------------------------------------------------------------------ history lookup */

/* get_past_anchors, jitced_template.c:193-217. One cursor per lookup site; the walk's stopping
 * rules decide which segment is taken when t equals an anchor time.
 * This is the walk through memory; returns the index of the left anchor. */
template <class MT>
JD_FN int walk(MT & M, int const site, double const t, double const * & pa, double const * & pn, double & ta, double & tn)
{
	int ca = M.cursor[site];
	pa = M.slot(ca);
	ta = pa[0];
	while (ta > t && ca > M.first)
	{
		--ca;
		pa = M.slot(ca);
		ta = pa[0];
	}
	pn = M.slot(ca+1);
	tn = pn[0];
	while (tn < t && ca+2 <= M.last)
	{
		++ca;
		pa = pn;
		ta = tn;
		pn = M.slot(ca+1);
		tn = pn[0];
	}
	if (t > M.cur_t)
		M.pws = fmax(M.pws, t-M.cur_t);
	M.cursor[site] = ca;
	/* The cursor moves on by about one anchor per step: ask for the anchor after the bracketing pair now (L1
	 * prefetch, nothing waits for it), so that the lookup that reaches it one or two steps later does not pay
	 * the DRAM latency. Accepted anchors only: they are immutable. */
	if (MT::PREFETCH && ca+2 <= M.current)
		M.prefetch(ca+2);
	return ca;
}

#if JD_WINDOW == 3
#if defined(__CUDA_ARCH__)
extern __shared__ double jd_pwin_smem[];

__device__ __forceinline__ double * pwin_slot(int const site, int const index)
{
	return jd_pwin_smem + (size_t)threadIdx.x*JD_PWIN_DOUBLES + (size_t)(site*4 + (index & 3))*JD_A;
}

__device__ __forceinline__ void pwin_fetch(int const site, int const index, double const * const anchor)
{
	double * const dst = pwin_slot(site, index);
	JD_UNROLL
	for (int k = 0; k < 1+2*JD_N; k++)
	{
		unsigned const d = (unsigned)__cvta_generic_to_shared(dst+k);
		asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(anchor+k) : "memory");
	}
}

__device__ __forceinline__ void pwin_wait()
{
	asm volatile("cp.async.wait_all;" ::: "memory");
}
#else
/* host pass of nvcc: never executed */
inline double * pwin_slot(int, int) { return nullptr; }
inline void pwin_fetch(int, int, double const *) {}
inline void pwin_wait() {}
#endif

JD_FN void fill_window(Member & M, int const site)
{
	Window & W = M.win[site];
	int const ca = M.cursor[site];
	W.valid = ca+1 <= M.current;
	if (!W.valid)
		return;
	pwin_fetch(site, ca, M.slot(ca));
	pwin_fetch(site, ca+1, M.slot(ca+1));
	W.fresh = true;
	W.next_ok = ca+2 <= M.current;
	if (W.next_ok)
		pwin_fetch(site, ca+2, M.slot(ca+2));
	W.prev_ok = ca-1 >= M.first;
	if (W.prev_ok)
		pwin_fetch(site, ca-1, M.slot(ca-1));
}

/* the same comparisons as walk(), on staged times; invariant as for JD_WINDOW == 2 */
JD_FN Seg anchors(Member & M, int const site, double const t)
{
	Window & W = M.win[site];
	Seg s;
	if (W.valid)
	{
		if (W.fresh)
		{
			pwin_wait();
			W.fresh = false;
		}
		int ca = M.cursor[site];
		double tv = pwin_slot(site, ca)[0];
		double tw = pwin_slot(site, ca+1)[0];
		bool hit = true;
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
			pwin_wait();
			tw = tv;
			tv = pwin_slot(site, ca)[0];
			if (W.prev_ok)
				pwin_fetch(site, ca-1, M.slot(ca-1));
		}
		while (hit && tw < t)
		{
			if (!W.next_ok)
			{
				hit = false;
				break;
			}
			++ca;
			W.prev_ok = true;
			W.next_ok = ca+2 <= M.current;
			pwin_wait();
			tv = tw;
			tw = pwin_slot(site, ca+1)[0];
			if (W.next_ok)
				pwin_fetch(site, ca+2, M.slot(ca+2));
		}
		M.cursor[site] = ca;
		if (hit)
		{
			s.v = pwin_slot(site, ca);
			s.w = pwin_slot(site, ca+1);
			s.tv = tv;
			s.tw = tw;
			return s;
		}
	}
	walk(M, site, t, s.v, s.w, s.tv, s.tw);
	fill_window(M, site);
	return s;
}
#elif JD_WINDOW == 2
#if defined(__CUDA_ARCH__)
extern __shared__ d2 jd_win_smem[];

/* row r of this thread's column: rows are 16-byte elements, JD_BLOCK per row → conflict-free LDS.128 */
__device__ __forceinline__ d2 * win_row(int const site, int const index, int const half)
{
	return jd_win_smem + (size_t)((site*4 + (index & 3))*(JD_A/2) + half)*JD_BLOCK + threadIdx.x;
}

__device__ __forceinline__ void win_fetch(int const site, int const index, double const * const anchor)
{
	JD_UNROLL
	for (int h = 0; h < JD_A/2; h++)
	{
		unsigned const dst = (unsigned)__cvta_generic_to_shared(win_row(site, index, h));
		asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(anchor+2*h) : "memory");
	}
}

__device__ __forceinline__ void win_wait()
{
	asm volatile("cp.async.wait_all;" ::: "memory");
}

__device__ __forceinline__ double win_time(int const site, int const index)
{
	return win_row(site, index, 0)->x;
}

__device__ __forceinline__ void win_read(int const site, int const index, double (&a)[JD_A])
{
	JD_UNROLL
	for (int h = 0; h < JD_A/2; h++)
	{
		d2 const v = *win_row(site, index, h);
		a[2*h] = v.x;
		a[2*h+1] = v.y;
	}
}
#else
/* host pass of nvcc: never executed */
inline void win_fetch(int, int, double const *) {}
inline void win_wait() {}
inline double win_time(int, int) { return 0.0; }
inline void win_read(int, int, double (&)[JD_A]) {}
#endif

JD_FN void fill_window(Member & M, int const site)
{
	Window & W = M.win[site];
	int const ca = M.cursor[site];
	W.valid = ca+1 <= M.current;
	if (!W.valid)
		return;
	win_fetch(site, ca, M.slot(ca));
	win_fetch(site, ca+1, M.slot(ca+1));
	W.fresh = true;
	W.next_ok = ca+2 <= M.current;
	if (W.next_ok)
		win_fetch(site, ca+2, M.slot(ca+2));
	W.prev_ok = ca-1 >= M.first;
	if (W.prev_ok)
		win_fetch(site, ca-1, M.slot(ca-1));
}

/* Invariant: the copies of the slots of cursor and cursor+1 have landed (unless W.fresh); those of cursor-1 and
 * cursor+2 may still be in flight — they are awaited only when the cursor moves onto them, so a prefetch has
 * the time until the next move (about one integration step) to arrive. */
JD_FN Seg anchors(Member & M, int const site, double const t)
{
	Window & W = M.win[site];
	Seg s;
	if (W.valid)
	{
		if (W.fresh)
		{
			win_wait();
			W.fresh = false;
		}
		int ca = M.cursor[site];
		double tv = win_time(site, ca);
		double tw = win_time(site, ca+1);
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
			win_wait();
			tw = tv;
			tv = win_time(site, ca);
			if (W.prev_ok)
				win_fetch(site, ca-1, M.slot(ca-1));      /* overwrites the slot of the old cursor+2 */
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
			win_wait();
			tv = tw;
			tw = win_time(site, ca+1);
			if (W.next_ok)
				win_fetch(site, ca+2, M.slot(ca+2));      /* overwrites the slot of the old cursor-1 */
		}
		M.cursor[site] = ca;
		if (hit)
		{
			/* t <= time of an accepted anchor <= current.time: no past-within-step contribution */
			win_read(site, ca, s.v);
			win_read(site, ca+1, s.w);
			return s;
		}
	}
	double const * pa; double const * pn; double ta, tn;
	walk(M, site, t, pa, pn, ta, tn);
	load_anchor(pa, s.v);
	load_anchor(pn, s.w);
	fill_window(M, site);
	return s;
}
#elif JD_WINDOW == 4
#if defined(__CUDA_ARCH__)
extern __shared__ d2 jd_win_smem[];

__device__ __forceinline__ void win_init(Member & M)
{
	/* (through an opaque move: otherwise the compiler re-reads the thread index from its special register — a slow
	 * read — at every lookup instead of keeping this value in a register) */
	asm volatile("mov.u32 %0, %1;" : "=r"(M.wbase) : "r"(threadIdx.x));
}

/* this thread's element of pair row `row` of the slot of anchor `index` */
__device__ __forceinline__ d2 * win_pair_pointer(Member const & M, int const site, int const index, int const row)
{
	return jd_win_smem + (M.wbase + (unsigned)(((site*JD_WSLOTS + (index & (JD_WSLOTS-1)))*JD_WPAIRS + row)*JD_BLOCK));
}

/* … and the single, which shares a 16-byte element with the single of the neighbouring slot */
__device__ __forceinline__ double * win_single_pointer(Member const & M, int const site, int const index)
{
	int const slot = index & (JD_WSLOTS-1);
	return reinterpret_cast<double *>(jd_win_smem + (M.wbase + (unsigned)((JD_WIN_PAIR_ROWS + site*(JD_WSLOTS/2) + (slot >> 1))*JD_BLOCK))) + (slot & 1);
}

/* address of the window in the shared state space, as cp.async wants it: the symbol itself (an immediate), not a
 * conversion of a generic pointer (which reads a special register every time) */
__device__ __forceinline__ unsigned win_shared_base()
{
	unsigned base;
	asm("mov.u32 %0, _ZN4jdde11jd_win_smemE;" : "=r"(base));
	return base;
}

/* request the copy of one anchor: one commit group */
__device__ __forceinline__ void win_fetch(Member const & M, int const site, int const index)
{
	double const * const anchor = M.slot(index);
	int const slot = index & (JD_WSLOTS-1);
	unsigned const base = win_shared_base() + M.wbase*16u;
	JD_UNROLL
	for (int r = 0; r < JD_WPAIRS; r++)
		asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(base + (unsigned)(((site*JD_WSLOTS + slot)*JD_WPAIRS + r)*(JD_BLOCK*16))), "l"(anchor+2*r) : "memory");
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(base + (unsigned)((JD_WIN_PAIR_ROWS + site*(JD_WSLOTS/2) + (slot >> 1))*(JD_BLOCK*16) + (slot & 1)*8)), "l"(anchor+2*JD_WPAIRS) : "memory");
	asm volatile("cp.async.commit_group;" ::: "memory");
}

/* Make sure that the copy of anchor `need` has landed. Copies are requested in increasing index order and complete in
 * that order as far as wait_group is concerned: with hi − need >= 3 younger groups of this site (other sites only add
 * younger ones), "at most 3 groups pending" implies that the group of `need` is complete. (The "memory" clobbers keep
 * the compiler from moving reads of the window across the wait; the reads themselves are ordinary loads, free to be
 * scheduled ahead of their use.) */
__device__ __forceinline__ void win_sync(Window & W, int const need)
{
	if (W.sync < need)
	{
		if (W.hi-need >= JD_WAHEAD-1)
		{
			asm volatile("cp.async.wait_group 3;" ::: "memory");
			W.sync = W.hi-(JD_WAHEAD-1);
		}
		else
		{
			asm volatile("cp.async.wait_group 0;" ::: "memory");
			W.sync = W.hi;
		}
	}
}

__device__ __forceinline__ d2 win_pair(Member const & M, int const site, int const index, int const row)
{
	return *win_pair_pointer(M, site, index, row);
}

/* the rest of an anchor whose first pair (time, first component) has been read already */
__device__ __forceinline__ void win_read_rest(Member const & M, int const site, int const index, d2 const first, double (&a)[JD_A])
{
	a[0] = first.x;
	a[1] = first.y;
	JD_UNROLL
	for (int r = 1; r < JD_WPAIRS; r++)
	{
		d2 const v = win_pair(M, site, index, r);
		a[2*r] = v.x;
		a[2*r+1] = v.y;
	}
	a[2*JD_WPAIRS] = *win_single_pointer(M, site, index);
	JD_UNROLL
	for (int k = 2*JD_WPAIRS+1; k < JD_A; k++)
		a[k] = 0.0;
}
#else
/* host pass of nvcc: never executed */
inline void win_fetch(Member const &, int, int) {}
inline void win_sync(Window &, int) {}
inline d2 win_pair(Member const &, int, int, int) { return d2(); }
inline void win_read_rest(Member const &, int, int, d2, double (&)[JD_A]) {}
#endif

/* stage the anchors around the cursor of a site: hi−7 … hi with hi = min(cursor + JD_WAHEAD, current) */
JD_FN void fill_window(Member & M, int const site)
{
	Window & W = M.win[site];
	int const ca = M.cursor[site];
	int const hi = ca+JD_WAHEAD <= M.current ? ca+JD_WAHEAD : M.current;
	if (hi <= ca)
	{
		W.hi = -0x40000000;      /* nothing to stage: the segment of the cursor ends at a tentative anchor */
		return;
	}
	JD_UNROLL
	for (int k = JD_WSLOTS-1; k >= 0; k--)
		win_fetch(M, site, hi-k);       /* (indices below `first` are never looked at; their slots of the ring are this member's own memory) */
	W.hi = hi;
	W.sync = hi-JD_WSLOTS;
}

/* the same comparisons as walk(), on staged anchors */
JD_FN Seg anchors(Member & M, int const site, double const t)
{
	Window & W = M.win[site];
	Seg s;
	int ca = M.cursor[site];
	if (W.hi > ca && ca > W.hi-JD_WSLOTS)
	{
		if (W.hi < ca+JD_WAHEAD && W.hi < M.current)
			win_fetch(M, site, ++W.hi);      /* `current` has moved on since the window was filled */
		win_sync(W, ca+1);
		d2 v = win_pair(M, site, ca, 0);
		d2 w = win_pair(M, site, ca+1, 0);
		bool hit = true;
		/* while (ca->time > t && ca->previous) ca = ca->previous; */
		while (v.x > t && ca > M.first)
		{
			if (ca-1 <= W.hi-JD_WSLOTS)
			{
				hit = false;
				break;
			}
			--ca;
			w = v;
			v = win_pair(M, site, ca, 0);
		}
		/* while (ca->next->time < t && ca->next->next) ca = ca->next; */
		while (hit && w.x < t)
		{
			if (ca+2 > W.hi)
			{
				hit = false;      /* next anchor is tentative or missing: let walk() decide */
				break;
			}
			++ca;
			if (W.hi < ca+JD_WAHEAD && W.hi < M.current)
				win_fetch(M, site, ++W.hi);      /* overwrites the slot of anchor hi−8 <= cursor−4 */
			win_sync(W, ca+1);
			v = w;
			w = win_pair(M, site, ca+1, 0);
		}
		M.cursor[site] = ca;
		if (hit)
		{
			/* t <= time of an accepted anchor <= current.time: no past-within-step contribution */
			win_read_rest(M, site, ca, v, s.v);
			win_read_rest(M, site, ca+1, w, s.w);
			return s;
		}
	}
	double const * pa; double const * pn; double ta, tn;
	walk(M, site, t, pa, pn, ta, tn);
	load_anchor(pa, s.v);
	load_anchor(pn, s.w);
	fill_window(M, site);
	return s;
}
#elif JD_WINDOW
/* (re)load the window of a site around its cursor; only accepted anchors (index <= current) are cached */
JD_FN void fill_window(Member & M, int const site)
{
	Window & W = M.win[site];
	int const ca = M.cursor[site];
	W.valid = ca+1 <= M.current;
	if (!W.valid)
		return;
	load_anchor(M.slot(ca), W.v);
	load_anchor(M.slot(ca+1), W.w);
	W.next_ok = ca+2 <= M.current;
	if (W.next_ok)
		load_anchor(M.slot(ca+2), W.n);
	W.prev_ok = ca-1 >= M.first;
	if (W.prev_ok)
		load_anchor(M.slot(ca-1), W.p);
}

JD_FN Seg anchors(Member & M, int const site, double const t)
{
	Window & W = M.win[site];
	Seg s;
	/* Fast path: the same comparisons as in walk(), on cached anchors. They are accepted anchors, hence
	 * immutable while the kernel runs and not later than current.time. */
	if (W.valid)
	{
		bool hit = true;
		/* while (ca->time > t && ca->previous) ca = ca->previous; */
		while (W.v[0] > t && M.cursor[site] > M.first)
		{
			if (!W.prev_ok)
			{
				hit = false;
				break;
			}
			int const ca = --M.cursor[site];
			JD_UNROLL
			for (int k = 0; k < JD_A; k++)
			{
				W.n[k] = W.w[k];
				W.w[k] = W.v[k];
				W.v[k] = W.p[k];
			}
			W.next_ok = true;
			W.prev_ok = ca-1 >= M.first;
			if (W.prev_ok)
				load_anchor(M.slot(ca-1), W.p);    /* prefetch */
		}
		/* while (ca->next->time < t && ca->next->next) ca = ca->next; */
		while (hit && W.w[0] < t)
		{
			if (!W.next_ok)
			{
				hit = false;      /* next anchor is tentative or missing: let walk() decide */
				break;
			}
			int const ca = ++M.cursor[site];
			JD_UNROLL
			for (int k = 0; k < JD_A; k++)
			{
				W.p[k] = W.v[k];
				W.v[k] = W.w[k];
				W.w[k] = W.n[k];
			}
			W.prev_ok = true;
			W.next_ok = ca+2 <= M.current;
			if (W.next_ok)
				load_anchor(M.slot(ca+2), W.n);    /* prefetch: not needed before the cursor moves again */
		}
		if (hit)
		{
			/* t <= time of an accepted anchor <= current.time: no past-within-step contribution */
			JD_UNROLL
			for (int k = 0; k < JD_A; k++)
			{
				s.v[k] = W.v[k];
				s.w[k] = W.w[k];
			}
			return s;
		}
	}
	double const * pa; double const * pn; double ta, tn;
	walk(M, site, t, pa, pn, ta, tn);
	load_anchor(pa, s.v);
	load_anchor(pn, s.w);
	fill_window(M, site);
	return s;
}
#elif JD_SEG_BY_VALUE
/* Small anchors: the three anchors the cursor can reach without a second memory round trip (cursor,
 * cursor+1, cursor+2) are requested at once, as 16-byte vector loads, before the comparisons of the walk are
 * made on them. In the common cases — the cursor stays or moves on by one anchor — a lookup thus costs one
 * memory latency instead of three to four dependent ones; everything else continues with the generic walk. */
JD_FN Seg anchors(Member & M, int const site, double const t)
{
	Seg s;
	int ca = M.cursor[site];
	bool const third = ca+2 <= M.last;
	double a2[JD_A];
	load_anchor(M.slot(ca), s.v);
	load_anchor(M.slot(ca+1), s.w);
	load_anchor(M.slot(third ? ca+2 : ca+1), a2);
	if (s.v[0] > t && ca > M.first)
	{
		/* backwards (retry after a rejected step, state-dependent delays): generic walk */
		double const * pa; double const * pn; double ta, tn;
		walk(M, site, t, pa, pn, ta, tn);
		load_anchor(pa, s.v);
		load_anchor(pn, s.w);
		return s;
	}
	if (s.w[0] < t && third)
	{
		++ca;
		JD_UNROLL
		for (int k = 0; k < JD_A; k++)
		{
			s.v[k] = s.w[k];
			s.w[k] = a2[k];
		}
		while (s.w[0] < t && ca+2 <= M.last)
		{
			++ca;
			JD_UNROLL
			for (int k = 0; k < JD_A; k++)
				s.v[k] = s.w[k];
			load_anchor(M.slot(ca+1), s.w);
		}
	}
	if (t > M.cur_t)
		M.pws = fmax(M.pws, t-M.cur_t);
	M.cursor[site] = ca;
	return s;
}
#else
template <class MT>
JD_FN Seg anchors(MT & M, int const site, double const t)
{
	Seg s;
	walk(M, site, t, s.v, s.w, s.tv, s.tw);
	return s;
}
#endif

/* a/b where b is the length of a history segment (positive, normal). Production build on the device: hardware
 * reciprocal estimate (20 bits) + two Newton steps + one residual correction of the quotient — 9 branch-free
 * instructions instead of the ≈ 50 of the IEEE division (9 % of the Lyapunov stepper's instructions), within an
 * ulp of it. The verification build divides. */
JD_FN double segment_div(double const a, double const b)
{
#if defined(JD_FAST) && defined(__CUDA_ARCH__) && !defined(JD_EXACT_SEGMENT_DIV)
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
	r = fma(r, fma(-b, r, 1.0), r);
	r = fma(r, fma(-b, r, 1.0), r);
	double const q = a*r;
	return fma(fma(-q, b, a), r, q);
#else
	return a/b;
#endif
}

/* get_past_value, jitced_template.c:221-235 */
JD_FN double past_y(double const t, int const index, Seg const & s)
{
	double const q = seg_tw(s)-seg_tv(s);
	double const x = segment_div(t - seg_tv(s), q);
	double const a = s.v[1+index];
	double const b = s.v[1+JD_N+index] * q;
	double const c = s.w[1+index];
	double const d = s.w[1+JD_N+index] * q;
	return (1-x) * ( (1-x) * (b*x + (a-c)*(2*x+1)) - d*x*x) + c;
}

/* get_past_diff, jitced_template.c:237-251 */
JD_FN double past_dy(double const t, int const index, Seg const & s)
{
	double const q = seg_tw(s)-seg_tv(s);
	double const x = segment_div(t - seg_tv(s), q);
	double const a = s.v[1+index];
	double const b = s.v[1+JD_N+index] * q;
	double const c = s.w[1+index];
	double const d = s.w[1+JD_N+index] * q;
	return segment_div( (1-x)*(b-x*3*(2*(a-c)+b+d)) + d*x, q);
}

JD_FN Seg segment_at(Member const & M, int const vi)
{
	Seg s;
#if JD_SEG_BY_VALUE
	load_anchor(M.slot(vi), s.v);
	load_anchor(M.slot(vi+1), s.w);
#else
	s.v = M.slot(vi);
	s.w = M.slot(vi+1);
	s.tv = s.v[0];
	s.tw = s.w[0];
#endif
	return s;
}

/* ------------------------------------------------------------------ right-hand side */

/* macro vocabulary of the generated statements (jitced_template.c:382-391) */
#define JDDE_Y(i) (y[i])
#define JDDE_PAR(k) (M.par[k])
#define JDDE_SEG ::jdde::Seg const
#define JDDE_ANCHORS(k, tm) ::jdde::anchors(M, k, tm)
#define JDDE_PAST_Y(tm, i, s) ::jdde::past_y(tm, i, s)
#define JDDE_PAST_DY(tm, i, s) ::jdde::past_dy(tm, i, s)
#define JDDE_SET_DY(i, value) (dY[i] = (value))
#define JDDE_IPOW(x, e) jdde_ipow(x, e)

/* eval_f, jitced_template.c:398-419 */
JD_FN void eval_f(Member & M, double const t, double const * const y, double * const dY)
{
	#include JD_RHS_FILE
}

/* ------------------------------------------------------------------ one attempted step */

/* delta_t/9 and delta_t/72 of the Bogacki–Shampine formulas. The production build (JD_FAST) divides by these
 * constants with the compile-time reciprocal and two FMAs (Markstein's sequence: q = x·y, r = x − q·c exactly,
 * result = q + r·y), which returned the correctly rounded quotient for all of 4·10^8 random arguments tested
 * against the hardware division at a tenth of its instructions; the verification build divides. */
#if defined(JD_FAST) && defined(__CUDA_ARCH__)
template <int C>
JD_FN double divide_by(double const x)
{
	double const y = 1.0/C;
	double const q = x*y;
	double const r = fma(-q, (double)C, x);
	return fma(r, y, q);
}
#else
template <int C>
JD_FN double divide_by(double const x)
{
	return x/(double)C;
}
#endif

/* get_next_step, jitced_template.c:421-472 */
JD_FN void get_next_step(Member & M, double const delta_t)
{
	M.last_start = M.cur_t;
	M.pws = 0.0;
	M.attempts++;

	double argument[JD_N];
	double k_2[JD_N];
	JD_UNROLL
	for (int i = 0; i < JD_N; i++)
		argument[i] = M.cur_y[i] + 0.5*delta_t*M.cur_d[i];
	eval_f(M, M.cur_t+0.5*delta_t, argument, k_2);

	double k_3[JD_N];
	JD_UNROLL
	for (int i = 0; i < JD_N; i++)
		argument[i] = M.cur_y[i] + 0.75*delta_t*k_2[i];
	eval_f(M, M.cur_t+0.75*delta_t, argument, k_3);

	double ny[JD_N], nd[JD_N];
	JD_UNROLL
	for (int i = 0; i < JD_N; i++)
		ny[i] = M.cur_y[i] + divide_by<9>(delta_t) * (2*M.cur_d[i]+3*k_2[i]+4*k_3[i]);

	double const nt = M.cur_t + delta_t;
	eval_f(M, nt, ny, nd);

	JD_UNROLL
	for (int i = 0; i < JD_N; i++)
		M.err[i] = (5*M.cur_d[i]-6*k_2[i]-8*k_3[i]+9*nd[i]) * divide_by<72>(delta_t);

	/* append_anchor (:56-73) after an accepted step, replace_last_anchor (:106-114) on a retry. The state of
	 * the replaced anchor (`old_last`) is only needed by the past-within-step iteration: see try_single_step. */
	if (M.last == M.current)
		M.last++;
	M.new_t = nt;
	double a[JD_A];
	a[0] = nt;
	JD_UNROLL
	for (int i = 0; i < JD_N; i++)
	{
		M.new_y[i] = ny[i];
		M.new_d[i] = nd[i];
		a[1+i] = ny[i];
		a[1+JD_N+i] = nd[i];
	}
	JD_UNROLL
	for (int k = 1+2*JD_N; k < JD_A; k++)
		a[k] = 0.0;
	store_anchor(M.slot(M.last), a);
}

/* get_p, jitced_template.c:474-498: p = max_i |err_i| / (atol + rtol·|y_i|), skipping 0/0; a NaN never wins a
 * comparison. `local_p` is the part over the components one thread owns.
 * Production build with several components: the largest ratio is found by cross-multiplication
 * (e_i/τ_i > e_j/τ_j ⟺ e_i·τ_j > e_j·τ_i for τ >= 0) and only that one is divided out — one FP64 division per
 * attempt instead of one per component (divisions were 10 % of all instructions of the Lyapunov stepper,
 * profiles/). Near-ties may pick the other candidate, whose ratio then differs by rounding only. */
template <int NL>
JD_FN double local_p(double const (&err)[NL], double const (&new_y)[NL], double const atol, double const rtol)
{
#if defined(JD_FAST)
	if (NL > 1)
	{
		double best_e = 0.0, best_tol = 1.0;
		JD_UNROLL
		for (int i = 0; i < NL; i++)
		{
			double const error = fabs(err[i]);
			double const tolerance = atol+rtol*fabs(new_y[i]);
			/* (0/0 needs no test of its own here: 0·best_tol > best_e·0 is false) */
			if (error*best_tol > best_e*tolerance)
			{
				best_e = error;
				best_tol = tolerance;
			}
		}
		return best_e/best_tol;
	}
#endif
	double p = 0.0;
	JD_UNROLL
	for (int i = 0; i < NL; i++)
	{
		double const error = fabs(err[i]);
		double const tolerance = atol+rtol*fabs(new_y[i]);
		if (error != 0.0 || tolerance != 0.0)
		{
			double const x = error/tolerance;
			if (x > p)
				p = x;
		}
	}
	return p;
}

JD_FN double get_p(Member const & M, double const atol, double const rtol)
{
	return local_p<JD_N>(M.err, M.new_y, atol, rtol);
}

/* check_new_y_diff, jitced_template.c:502-524. `old_y` = state of the tentative anchor that the last
 * get_next_step replaced (old_last; it always exists where the controller calls this, _jitcdde.py:851-855). */
JD_FN bool check_new_y_diff(Member const & M, double const * const old_y, double const atol, double const rtol)
{
	bool result = true;
	JD_UNROLL
	for (int i = 0; i < JD_N; i++)
	{
		double const difference = fabs(M.new_y[i] - old_y[i]);
		double const tolerance = atol + fabs(rtol*M.new_y[i]);
		result &= (tolerance >= difference);
	}
	return result;
}

/* accept_step, jitced_template.c:526-532 */
JD_FN void accept_step(Member & M)
{
	M.current = M.last;
	M.cur_t = M.new_t;
	JD_UNROLL
	for (int i = 0; i < JD_N; i++)
	{
		M.cur_y[i] = M.new_y[i];
		M.cur_d[i] = M.new_d[i];
	}
}

/* forget, jitced_template.c:534-561: drop anchors while the next one is still older than the threshold.
 * The anchor times are inspected eight at a time (independent loads in flight together) instead of one
 * after the other; which anchor becomes the first one is unchanged. */
template <class MT>
JD_FN void forget(MT & M, double const delay)
{
	double const threshold = fmin(M.cur_t - delay, M.last_start);
	bool moved = false;
	for (;;)
	{
		double tm[8];
		JD_UNROLL
		for (int j = 0; j < 8; j++)
			tm[j] = M.slot(M.first+1+j)[0];      /* slots beyond `last` hold stale or unset times; they are masked below */
		int advance = 0;
		bool stop = false;
		JD_UNROLL
		for (int j = 0; j < 8; j++)
		{
			/* while (first+1 < last && time[first+1] < threshold) first++ */
			if (!stop && M.first+1+j < M.last && tm[j] < threshold)
				advance++;
			else
				stop = true;
		}
		M.first += advance;
		moved |= advance > 0;
		if (advance < 8)
			break;
	}
	if (moved)
	{
		JD_UNROLL
		for (int s = 0; s < JD_NSITES; s++)
			if (M.cursor[s] < M.first)
			{
				M.cursor[s] = M.first;
				M.invalidate_windows();
			}
	}
}

/* get_recent_state, jitced_template.c:287-320 */
JD_FN void get_recent_state(Member const & M, double const t, double * const out, int const n_out)
{
	double const * const w = M.slot(M.last);
	double const * const v = M.slot(M.last-1);
	double const q = w[0]-v[0];
	double const x = (t - v[0])/q;
	JD_UNROLL
	for (int index = 0; index < JD_N; index++)
	{
		double const a = v[1+index];
		double const b = v[1+JD_N+index] * q;
		double const c = w[1+index];
		double const d = w[1+JD_N+index] * q;
		if (index < n_out)
			out[index] = (1-x) * ( (1-x) * (b*x + (a-c)*(2*x+1)) - d*x*x) + c;
	}
}

/* ------------------------------------------------------------------ controller (_jitcdde.py) */

/* _jitcdde.py:21 */
JD_FN double sigmoid(double const x) { return (JDDE_TANH(x)+1)/2; }

/* _jitcdde.py:745-765: UnsuccessfulIntegration becomes a per-member status */
template <class MT>
JD_FN void control_for_min_step(MT & M, jdde_int_params const & P)
{
	if (M.dt < P.min_step)
		M.status = JDDE_STATUS_MIN_STEP;
}

/* _increase_chance (_jitcdde.py:767-773) and do_increase (:730-741: drawn, or the deterministic credit rule); cold: only
 * after a past-within-step event. Operates on the controller memory in the per-member arrays. */
template <class MT>
JD_FN bool pws_allows_increase(MT & M, jdde_problem const & P, double const new_dt)
{
	jdde_int_params const & I = P.P;
	double const q = new_dt/ctl_last_pws(M, P);
	double const is_explicit = sigmoid(1-q);
	double const far_from_explicit = sigmoid(q-I.pws_factor);
	double const few_iterations = sigmoid(1-ctl_count(M, P)/I.pws_factor);
	double const profile = is_explicit+far_from_explicit*few_iterations;
	double const chance = profile + I.pws_base_increase_chance*(1-profile);

	if (I.pws_fuzzy_seed)
	{
		/* pws_fuzzy_increase (_jitcdde.py:730-731): `rng.random() < chance`; the draw is a function of the seed, the member and
		 * the number of right-hand-side evaluations the member has made (3 per attempted step + those of jumps) */
		long long const evaluations = 3*(P.attempts[M.index]+M.attempts) + P.extra_rhs[M.index]+M.extra_rhs;
		return jdde_uniform((unsigned)I.pws_fuzzy_seed, (unsigned long long)M.index, (unsigned long long)evaluations) < chance;
	}

	double credit = ctl_credit(M, P) + chance;
	bool const increase = credit >= 0.9;
	if (increase)
		credit = 0.0;
	ctl_credit(M, P) = credit;
	return increase;
}

/* _adjust_step_size, _jitcdde.py:775-794 with q = 3 (:726). The two fractional powers use the libm-free root of
 * jdde_math.h so that CPU oracle and GPU take bit-identical step sizes. Decrease (p**(-1/3)) and increase (p**(-1/4))
 * share ONE instance of the root code, selected by a flag: in a warp of 32 trajectories some lane decreases and some
 * lane increases in nearly every attempt, so two branches cost the warp both of them one after the other. */
template <class MT>
JD_FN bool adjust_step_size(MT & M, jdde_problem const & PR)
{
	jdde_int_params const & P = PR.P;
	double const p = get_p(M, P.atol, P.rtol);
	bool const decrease = p > P.decrease_threshold;
	if (decrease || p <= P.increase_threshold)
	{
		double const root = jdde_inv_root(p, !decrease);
		if (decrease)
		{
			M.dt *= fmax(P.safety_factor*root, P.min_factor);
			control_for_min_step(M, P);
			return false;
		}
		double const factor = (p != 0.0) ? P.safety_factor*root : P.max_factor;
		double const new_dt = fmin(M.dt*fmin(factor, P.max_factor), P.max_step);
		if (!M.pws_pending || pws_allows_increase(M, PR, new_dt))
		{
			M.dt = new_dt;
			if (M.pws_pending)
			{
				ctl_count(M, PR) = 0;
				ctl_last_pws(M, PR) = 0.0;
				M.pws_pending = false;
			}
		}
	}
	return true;
}

/* try_single_step, _jitcdde.py:838-861. Written as one loop so that the step code is instantiated once:
 * pass 0 is the attempt with `length`; if a delayed lookup reached into the step (past within step),
 * passes 1..pws_max_iterations repeat the step with self.dt (sic, :853) until two successive results agree. */
template <class MT>
JD_FN bool try_single_step(MT & M, jdde_problem const & PR, double const length)
{
	jdde_int_params const & P = PR.P;
	double step = length;
	int count = 0;
	for (;;)
	{
		double old_y[MT::NL];
		if (count)
		{
			JD_UNROLL
			for (int i = 0; i < MT::NL; i++)
				old_y[i] = M.new_y[i];
		}
		get_next_step(M, step);
		if (count == 0)
		{
			if (M.pws == 0.0)
				break;
			ctl_last_pws(M, PR) = M.pws;
			M.pws_pending = true;
			/* if possible, shrink the step to make the integration explicit */
			if (M.dt > P.pws_factor*M.pws)
			{
				M.dt /= P.pws_factor;
				control_for_min_step(M, P);
				return false;
			}
		}
		else if (check_new_y_diff(M, old_y, P.pws_atol, P.pws_rtol))
			break;
		if (count == P.pws_max_iterations)
		{
			M.dt /= P.pws_factor;
			control_for_min_step(M, P);
			return false;
		}
		count++;
		ctl_count(M, PR) = count;
		step = M.dt;
	}
	return adjust_step_size(M, PR);
}

/* Room for one more anchor? The reference's list is unbounded; the ring is not. forget() first
 * (it never removes an anchor a valid lookup can reach), then report overflow. */
template <class MT>
JD_FN bool ensure_room(MT & M, double const max_delay)
{
	if (M.last+1-M.first <= M.mask)
		return true;
	forget(M, max_delay);
	if (M.last+1-M.first <= M.mask)
		return true;
	M.status = JDDE_STATUS_OVERFLOW;
	return false;
}

/* the `while t < target` loops of integrate (_jitcdde.py:830-832) and of
 * step_on_discontinuities (:988-993, steps capped at the target) */
/* One more rejected attempt. On the device the per-member total in global memory is incremented by a reduction that
 * returns nothing (RED: no register, nothing to wait for) instead of a counter carried through the hot loop. */
template <class MT>
JD_FN void count_rejection(MT & M, jdde_problem const &)
{
	M.rejected++;      /* (a group of lanes per member: every lane keeps the count, one of them stores it) */
}

JD_FN void count_rejection(Member & M, jdde_problem const & P)
{
#if defined(__CUDA_ARCH__)
	atomicAdd(reinterpret_cast<unsigned long long *>(P.rejected+M.index), 1ull);
#else
	P.rejected[M.index]++;
#endif
}

/* last_actual_step_start (jitced_template.c:48, set by get_next_step :424) after a series of attempts of the stepping
 * loop: the time of `current` if the last attempt left a tentative anchor behind (it was rejected), else the time of the
 * anchor before `current` (the accepted step started there). Recomputing it after the loop instead of carrying it
 * through the loop frees two registers of the hot loop (the assignments in get_next_step become dead stores there).
 * Valid where nothing but the stepper touches the history between load_member and store_member: the integrate kernels. */
template <class MT>
JD_FN void settle_last_start(MT & M, jdde_problem const & P)
{
	/* (attempts: calls of get_next_step since the member was loaded; none: the stored value stands) */
	M.last_start = M.attempts == 0 ? P.last_start[M.index] : M.last != M.current ? M.cur_t : M.slot(M.current-1)[0];
}

/* the `while t < target` loops of integrate (_jitcdde.py:830-832) and of
 * step_on_discontinuities (:988-993, steps capped at the target) */
template <class MT>
JD_FN void step_until(MT & M, jdde_problem const & P, double const target, bool const capped, long long const budget)
{
	int const attempt_budget = budget > 0x7fffffff ? 0x7fffffff : (int)budget;
	while (M.cur_t < target)
	{
		if (M.last+1-M.first > M.mask)
		{
			settle_last_start(M, P);
			if (!ensure_room(M, P.max_delay))
				break;
		}
		if (!(M.dt > 0.0))
		{
			/* a step size of zero/NaN would never reach the target: the reference would spin forever */
			M.status = JDDE_STATUS_NONFINITE;
			break;
		}
		if (M.attempts >= attempt_budget)
		{
			/* watchdog: a kernel must end; the caller clears the status and calls again to go on */
			M.status = JDDE_STATUS_BUDGET;
			break;
		}
		double const length = capped ? fmin(M.dt, target-M.cur_t) : M.dt;
		if (try_single_step(M, P, length))
			accept_step(M);
		else
		{
			count_rejection(M, P);
			if (M.status != JDDE_STATUS_OK)
				break;
		}
	}
	settle_last_start(M, P);
}

/* ------------------------------------------------------------------ jumps */

/* remove_last_anchor, jitced_template.c:88-104 */
JD_FN void remove_last_anchor(Member & M)
{
	M.invalidate_windows();
	M.last--;
	JD_UNROLL
	for (int s = 0; s < JD_NSITES; s++)
		if (M.cursor[s] == M.last)
			M.cursor[s] = M.last-1;
}

JD_FN void reload_current(Member & M)
{
	double const * const c = M.slot(M.current);
	M.cur_t = c[0];
	M.new_t = c[0];
	for (int i = 0; i < JD_N; i++)
	{
		M.cur_y[i] = M.new_y[i] = c[1+i];
		M.cur_d[i] = M.new_d[i] = c[1+JD_N+i];
	}
}

/* truncate_past, jitced_template.c:1089-1108 */
JD_FN void truncate_past(Member & M, double const time)
{
	while (M.last-1 > M.first && M.slot(M.last-1)[0] >= time)
		remove_last_anchor(M);

	M.invalidate_windows();
	Seg const left = segment_at(M, M.last-1);
	double ny[JD_N], nd[JD_N];
	for (int i = 0; i < JD_N; i++)
	{
		ny[i] = past_y (time, i, left);
		nd[i] = past_dy(time, i, left);
	}
	double * const dst = M.slot(M.last);
	dst[0] = time;
	for (int i = 0; i < JD_N; i++)
	{
		dst[1+i] = ny[i];
		dst[1+JD_N+i] = nd[i];
	}
	M.current = M.last;
	reload_current(M);
}

/* extrema, jitced_template.c:253-285 */
JD_FN void extrema(Seg const & s, int const index, double & minimum, double & maximum)
{
	double const q = seg_tw(s)-seg_tv(s);
	double const a = s.v[1+index];
	double const b = s.v[1+JD_N+index] * q;
	double const c = s.w[1+index];
	double const d = s.w[1+JD_N+index] * q;

	minimum = fmin(a,c);
	maximum = fmax(a,c);

	double const radicant = b*b + b*d + d*d + 3*(a-c)*(3*(a-c) + 2*(b+d));
	if (radicant >= 0)
	{
		double const A = 1/(2*a + b - 2*c + d);
		double const B = a + 2*b/3 - c + d/3;
		for (int sign = -1; sign <= 1; sign += 2)
		{
			double const x = (B+sign*sqrt(radicant)/3) * A;
			if (0 < x && x < 1)
			{
				double const value = (1-x) * ( (1-x) * (b*x + (a-c)*(2*x+1)) - d*x*x) + c;
				minimum = fmin(minimum,value);
				maximum = fmax(maximum,value);
			}
		}
	}
}

/* apply_jump, jitced_template.c:1122-1174 */
JD_FN void apply_jump(Member & M, jdde_problem const & P, double const * const change, double const time, double const width, double * const minima, double * const maxima)
{
	double const nt = time+width;

	int left = M.last-1;
	while (left > M.first && M.slot(left)[0] >= nt)
		left--;
	Seg const ls = segment_at(M, left);

	double ny[JD_N], nd[JD_N];
	for (int i = 0; i < JD_N; i++)
		ny[i] = past_y(nt, i, ls) + change[i];
	eval_f(M, nt, ny, nd);
	M.extra_rhs++;

	truncate_past(M, time);
	if (!ensure_room(M, P.max_delay))
		return;
	M.last++;
	double * const dst = M.slot(M.last);
	dst[0] = nt;
	for (int i = 0; i < JD_N; i++)
	{
		dst[1+i] = ny[i];
		dst[1+JD_N+i] = nd[i];
	}
	M.current = M.last;
	reload_current(M);

	Seg const js = segment_at(M, M.last-1);
	for (int i = 0; i < JD_N; i++)
		extrema(js, i, minima[i], maxima[i]);
}

/* adjust_diff, _jitcdde.py:863-878 (get_state forgets first, :367-371) */
JD_FN void adjust_diff(Member & M, jdde_problem const & P, double const shift_ratio, double * const minima, double * const maxima)
{
	forget(M, P.max_delay);
	double const t_last = M.slot(M.last)[0];
	double const width = shift_ratio*(t_last-M.slot(M.last-1)[0]);
	double zero[JD_N];
	for (int i = 0; i < JD_N; i++)
		zero[i] = 0.0;
	apply_jump(M, P, zero, t_last-width, width, minima, maxima);
}

/* ------------------------------------------------------------------ Lyapunov exponents */
#if JD_NBASIC != JD_N || defined(JD_PROJECT)

/* calculate_sp_matrix, jitced_template.c:657-674 */
JD_FN void sp_matrix(double const q, double m[4][4])
{
	double const cq = q/420.;
	double const cqq = cq*q;
	double const cqqq = cqq*q;
	m[0][0] = 156*cq;  m[0][1] = 22*cqq;  m[0][2] = 54*cq;   m[0][3] = -13*cqq;
	m[1][0] = 22*cqq;  m[1][1] = 4*cqqq;  m[1][2] = 13*cqq;  m[1][3] = -3*cqqq;
	m[2][0] = 54*cq;   m[2][1] = 13*cqq;  m[2][2] = 156*cq;  m[2][3] = -22*cqq;
	m[3][0] = -13*cqq; m[3][1] = -3*cqqq; m[3][2] = -22*cqq; m[3][3] = 4*cqqq;
}

/* calculate_partial_sp_matrix, jitced_template.c:676-709 */
JD_FN void partial_sp_matrix(double const q, double const z, double m[4][4])
{
	double const cq = q/420.;
	double const cqq = cq*q;
	double const cqqq = cqq*q;

	double const z3 = z*z*z;
	double const z4 = z*z3;
	double const z5 = z*z4;

	double const h_1 = (- 120*z*z - 350*z - 252) * z5 * cqq ;
	double const h_2 = (-  60*z*z - 140*z -  84) * z5 * cqqq;
	double const h_3 = (- 120*z*z - 420*z - 378) * z5 * cq  ;
	double const h_4 = (-  70*z*z - 168*z - 105) * z4 * cqqq;
	double const h_6 = (          - 105*z - 140) * z3 * cqq ;
	double const h_7 = (          - 210*z - 420) * z3 * cq  ;
	double const h_5 = (2*h_2 + 3*h_4)/q;
	double const h_8 = - h_5 + h_7*q - h_6 - 0.5*(z*q)*(z*q);

	m[0][0] = 2*h_3;     m[0][1] = h_1;     m[0][2] = h_7-2*h_3;       m[0][3] = h_5;
	m[1][0] = h_1;       m[1][1] = h_2;     m[1][2] = h_6-h_1;         m[1][3] = h_2+h_4;
	m[2][0] = h_7-2*h_3; m[2][1] = h_6-h_1; m[2][2] = 2*h_3-2*h_7-z*q; m[2][3] = h_8;
	m[3][0] = h_5;       m[3][1] = h_2+h_4; m[3][2] = h_8;             m[3][3] = h_2+(h_5+h_6-h_1)*q;
}

#endif
#if JD_NBASIC != JD_N
/* scalar_product, jitced_template.c:773-817 (norm_sq, :740-771, is the case begin_1 == begin_2).
 * Segment matrices (calculate_sp_matrices, :711-733) are recomputed from the anchor times
 * instead of being stored with every anchor: zero before threshold = current.time − delay,
 * partial for the segment containing it, full afterwards. */
JD_FN double scalar_product(Member const & M, double const threshold, int const begin_1, int const begin_2)
{
	double sum = 0;
	bool partial_done = false;
	for (int vi = M.first; vi < M.last; vi++)
	{
		double const * const v = M.slot(vi);
		double const * const w = M.slot(vi+1);
		double const q = w[0] - v[0];
		double m[4][4];
		if (!partial_done)
		{
			if (w[0] < threshold)
				continue;
			partial_sp_matrix(q, (threshold - w[0]) / q, m);
			partial_done = true;
		}
		else
			sp_matrix(q, m);
		double const * const vector_1[4] = { v+1+begin_1, v+1+JD_N+begin_1, w+1+begin_1, w+1+JD_N+begin_1 };
		double const * const vector_2[4] = { v+1+begin_2, v+1+JD_N+begin_2, w+1+begin_2, w+1+JD_N+begin_2 };
		double part = 0;
		for (int i = 0; i < 4; i++)
			for (int j = 0; j < 4; j++)
				for (int index = 0; index < JD_NBASIC; index++)
					part += m[i][j] * vector_1[i][index] * vector_2[j][index];
		sum += part;
	}
	return sum;
}

/*
 * One streaming pass over all anchors for orthonormalise: optionally  v[sub_i] -= sub_factor·v[sub_j]
 * (subtract_from_past, jitced_template.c:832-844) and/or  v[scale_i] *= scale_factor  (scale_past, :819-830) on every
 * anchor, and — on the values after these updates — the scalar product <v[acc_1], v[acc_2]> (:773-817). An anchor is
 * updated when the stream reaches it; the segment ending there then has both its anchors up to date, so every product
 * and every sum is formed from the same operands in the same order as by the reference's separate loops.
 * Fusing turns the 2+3+4+… full passes per call into about a third of them (the kernel is bound by this traffic).
 */
JD_FN double orthonormalise_pass(Member & M, double const threshold, int const sub_i, int const sub_j, double const sub_factor,
	int const scale_i, double const scale_factor, int const acc_1, int const acc_2)
{
	double sum = 0;
	bool partial_done = false;
	for (int k = M.first; k <= M.last; k++)
	{
		double * const w = M.slot(k);
		if (sub_i >= 0)
			for (int c = 0; c < JD_NBASIC; c++)
			{
				w[1+sub_i+c]      -= sub_factor*w[1+sub_j+c];
				w[1+JD_N+sub_i+c] -= sub_factor*w[1+JD_N+sub_j+c];
			}
		if (scale_i >= 0)
			for (int c = 0; c < JD_NBASIC; c++)
			{
				w[1+scale_i+c]      *= scale_factor;
				w[1+JD_N+scale_i+c] *= scale_factor;
			}
		if (acc_1 < 0 || k == M.first)
			continue;
		double const * const v = M.slot(k-1);
		double const q = w[0] - v[0];
		double m[4][4];
		if (!partial_done)
		{
			if (w[0] < threshold)
				continue;
			partial_sp_matrix(q, (threshold - w[0]) / q, m);
			partial_done = true;
		}
		else
			sp_matrix(q, m);
		double const * const vector_1[4] = { v+1+acc_1, v+1+JD_N+acc_1, w+1+acc_1, w+1+JD_N+acc_1 };
		double const * const vector_2[4] = { v+1+acc_2, v+1+JD_N+acc_2, w+1+acc_2, w+1+JD_N+acc_2 };
		double part = 0;
#if defined(JD_FAST)
		/* production build: Σ_ij m_ij (Σ_index v1_i·v2_j) — 16 independent short chains instead of one chain of
		 * 96 dependent additions. Another summation order than the reference's, i.e. other rounding: Lyapunov
		 * results are compared statistically for this build; the verification build keeps the order. */
		double row[4] = { 0.0, 0.0, 0.0, 0.0 };
		JD_UNROLL
		for (int i = 0; i < 4; i++)
		{
			JD_UNROLL
			for (int j = 0; j < 4; j++)
			{
				double dot = 0;
				JD_UNROLL
				for (int index = 0; index < JD_NBASIC; index++)
					dot += vector_1[i][index] * vector_2[j][index];
				row[i] += m[i][j]*dot;
			}
		}
		part = (row[0]+row[1])+(row[2]+row[3]);
#else
		for (int i = 0; i < 4; i++)
			for (int j = 0; j < 4; j++)
				for (int index = 0; index < JD_NBASIC; index++)
					part += m[i][j] * vector_1[i][index] * vector_2[j][index];
#endif
		sum += part;
	}
	return sum;
}

/* orthonormalise, jitced_template.c:846-878: modified Gram–Schmidt on the separation functions, in place over all
 * anchors; norms[i] = norm of vector i after orthogonalisation, before normalisation. The sequence of operations is
 * the reference's; consecutive loops over the anchors are fused (orthonormalise_pass). */
JD_FN void orthonormalise(Member & M, int const n_lyap, double const delay, double * const norms)
{
	double const threshold = M.slot(M.current)[0] - delay;
	int pending_scale = -1;           /* vector whose normalisation still has to be applied to the anchors */
	double pending_factor = 1.0;
	for (int i = 0; i < n_lyap; i++)
	{
		int const bi = (i+1)*JD_NBASIC;
		int sub_j = -1;
		double sub_factor = 0.0;
		for (int j = 0; j < i; j++)
		{
			int const bj = (j+1)*JD_NBASIC;
			/* sp = <v_i, v_j> after the previous subtraction; the scaling of v_{i-1} rides on the first pass */
			double const sp = orthonormalise_pass(M, threshold, sub_j >= 0 ? bi : -1, sub_j, sub_factor, pending_scale, pending_factor, bi, bj);
			pending_scale = -1;
			sub_j = bj;
			sub_factor = sp;
		}
		double const norm = sqrt(orthonormalise_pass(M, threshold, sub_j >= 0 ? bi : -1, sub_j, sub_factor, pending_scale, pending_factor, bi, bi));
		pending_scale = -1;
		if (norm > JD_NORM_THRESHOLD)
		{
			pending_scale = bi;
			pending_factor = 1./norm;
		}
		norms[i] = norm;
	}
	if (pending_scale >= 0)
		orthonormalise_pass(M, threshold, -1, -1, 0.0, pending_scale, pending_factor, -1, -1);
}
#endif

/* ------------------------------------------------------------------ restricted / transversal Lyapunov exponents */
#if defined(JD_PROJECT)

#if JD_NBASIC != JD_N
/* subtract_from_past (jitced_template.c:832-844) and scale_past (:819-830) as single passes */
JD_FN void subtract_from_past(Member & M, int const begin_1, int const begin_2, double const factor)
{
	orthonormalise_pass(M, 0.0, begin_1, begin_2, factor, -1, 0.0, -1, -1);
}

JD_FN void scale_past(Member & M, int const begin, double const factor)
{
	orthonormalise_pass(M, 0.0, -1, -1, 0.0, begin, factor, -1, -1);
}

/* remove_projections, jitced_template.c:885-971 (Python twin: past.py:75-127): for every anchor and every vector a
 * "dummy" separation function that is the vector at this anchor and zero elsewhere is orthonormalised against the
 * previous 2·n_vectors−… dummies and its projection removed from the separation function; finally the dummies are
 * cleared and the separation function is normalised. Returns its norm before normalisation.
 * `vectors`: [n_vectors][2][n_basic]. Dummy k lives in components (2+k)·n_basic …, the separation function in
 * n_basic … 2·n_basic−1. The index of a past dummy is formed as the reference forms it (unsigned difference widened
 * to a signed 64-bit integer before the modulo, :939). */
JD_FN double remove_projections(Member & M, double const delay, int const n_vectors, double const * const vectors)
{
	double const threshold = M.slot(M.current)[0] - delay;
	int const sep_func = JD_NBASIC;
	long long const d = 2*(long long)n_vectors;
	unsigned dummy_num = 0;
	long long len_dummies = 0;
	for (int ca = M.first; ca <= M.last; ca++)
	{
		for (int vi = 0; vi < n_vectors; vi++)
		{
			int const dummy = (2+(int)dummy_num)*JD_NBASIC;
			for (int j = 0; j < JD_NBASIC; j++)
			{
				for (int oa = M.first; oa <= M.last; oa++)
				{
					double * const o = M.slot(oa);
					o[1+dummy+j] = 0.0;
					o[1+JD_N+dummy+j] = 0.0;
				}
				double * const c = M.slot(ca);
				c[1+dummy+j] = vectors[((size_t)vi*2+0)*JD_NBASIC+j];
				c[1+JD_N+dummy+j] = vectors[((size_t)vi*2+1)*JD_NBASIC+j];
			}
			for (unsigned i = 0; i < len_dummies; i++)
			{
				int const past_dummy = (2+(int)((long long)(unsigned)(dummy_num-i-1u) % d))*JD_NBASIC;
				double const sp = scalar_product(M, threshold, dummy, past_dummy);
				subtract_from_past(M, dummy, past_dummy, sp);
			}
			double const norm = sqrt(scalar_product(M, threshold, dummy, dummy));
			if (norm > JD_NORM_THRESHOLD)
			{
				scale_past(M, dummy, 1./norm);
				double const sp = scalar_product(M, threshold, sep_func, dummy);
				subtract_from_past(M, sep_func, dummy, sp);
			}
			else
				scale_past(M, dummy, 0);
			len_dummies++;
			dummy_num = (unsigned)((dummy_num+1) % d);
		}
		if (len_dummies > n_vectors)
			len_dummies -= n_vectors;
	}
	for (int ca = M.first; ca <= M.last; ca++)
	{
		double * const c = M.slot(ca);
		for (int j = 2*JD_NBASIC; j < JD_N; j++)
			c[1+j] = c[1+JD_N+j] = 0.0;
	}
	double const norm = sqrt(scalar_product(M, threshold, sep_func, sep_func));
	scale_past(M, sep_func, 1./norm);
	return norm;
}
#endif

/* normalise_indices, jitced_template.c:1007-1083: norm of the components `indices` over the delay window
 * (norm_sq_tangent: per segment Σ_ij m_ij Σ_index vec_i[index]·vec_j[index]), normalisation if it is not tiny */
JD_FN double normalise_indices(Member & M, double const delay, int const n_indices, int const * const indices)
{
	double const threshold = M.slot(M.current)[0] - delay;
	double sum = 0;
	bool partial_done = false;
	for (int vi = M.first; vi < M.last; vi++)
	{
		double const * const v = M.slot(vi);
		double const * const w = M.slot(vi+1);
		double const q = w[0] - v[0];
		double m[4][4];
		if (!partial_done)
		{
			if (w[0] < threshold)
				continue;
			partial_sp_matrix(q, (threshold - w[0]) / q, m);
			partial_done = true;
		}
		else
			sp_matrix(q, m);
		double const * const vector[4] = { v+1, v+1+JD_N, w+1, w+1+JD_N };
		double part = 0;
		for (int i = 0; i < 4; i++)
			for (int j = 0; j < 4; j++)
				for (int k = 0; k < n_indices; k++)
					part += m[i][j] * vector[i][indices[k]] * vector[j][indices[k]];
		sum += part;
	}
	double const norm = sqrt(sum);
	if (norm > JD_NORM_THRESHOLD)
	{
		double const factor = 1./norm;
		for (int ca = M.first; ca <= M.last; ca++)
		{
			double * const c = M.slot(ca);
			for (int k = 0; k < n_indices; k++)
			{
				c[1+indices[k]] *= factor;
				c[1+JD_N+indices[k]] *= factor;
			}
		}
	}
	return norm;
}

/* what jitcdde_restricted_lyap.remove_projections (_jitcdde.py:1278-1284) or
 * jitcdde_transversal_lyap's DDE.normalise_indices (:1478) do after a step or a call of integrate; returns the norm */
JD_FN double project(Member & M, jdde_projector const & R, double const delay)
{
	M.invalidate_windows();
#if JD_NBASIC != JD_N
	if (R.mode == JDDE_PROJECT_RESTRICTED)
	{
		/* remove_state_component / remove_diff_component, jitced_template.c:973-1003 */
		for (int k = 0; k < R.n_state_components; k++)
			for (int ca = M.first; ca <= M.last; ca++)
				M.slot(ca)[1+JD_NBASIC+R.state_components[k]] = 0.0;
		for (int k = 0; k < R.n_diff_components; k++)
			for (int ca = M.first; ca <= M.last; ca++)
				M.slot(ca)[1+JD_N+JD_NBASIC+R.diff_components[k]] = 0.0;
		return remove_projections(M, delay, R.n_vectors, R.vectors);
	}
#endif
	return normalise_indices(M, delay, R.n_indices, R.indices);
}

JD_FN void member_project(jdde_problem const & P, long long const m, jdde_projector const & R, double const delay, double * const norms)
{
	Member M;
	load_member(P, m, M);
	if (M.status != JDDE_STATUS_OK)
		return;
	norms[m] = project(M, R, delay);
}
#endif

/* integrate_blindly, _jitcdde.py:880-933; with `lyaps` the variant of jitcdde_lyap (:1191-1209):
 * orthonormalise after every step and average log(norms)/dt */
JD_FN void integrate_blindly(Member & M, jdde_problem const & P, double const target, double step, double * const out, int const n_out, double * const lyaps, double * const total_time, jdde_projector const * const projector = nullptr)
{
	if (!(step > 0))
		step = P.P.max_step;
	double const total = target-M.cur_t;
	if (total_time)
		*total_time = total;
	/* number of local exponents averaged over the steps: n_lyap (jitcdde_lyap), 1 (restricted / transversal) */
	int n_exponents = 0;
#if JD_NBASIC != JD_N
	int const n_lyap = JD_N/JD_NBASIC-1;
	if (lyaps && !projector)
		n_exponents = n_lyap;
#endif
#if defined(JD_PROJECT)
	if (lyaps && projector)
		n_exponents = 1;
#endif
	double lyap_sum[JD_N/JD_NBASIC];
	for (int i = 0; i < n_exponents; i++)
		lyap_sum[i] = 0.0;
	if (total > 0)
	{
		if (total < step)
			step = total;
		long long const number = (long long) rint(total/step);   /* Python round(): ties to even */
		double const dt = total/number;
		for (long long k = 0; k < number && M.status == JDDE_STATUS_OK; k++)
		{
			if (!ensure_room(M, P.max_delay))
				break;
			get_next_step(M, dt);
			accept_step(M);
			P.accepted[M.index]++;
			forget(M, P.max_delay);
#if defined(JD_PROJECT)
			if (lyaps && projector)
			{
				/* _jitcdde.py:1335-1340, :1496-1501 */
				double const norm = project(M, *projector, P.max_delay);
				reload_current(M);
				lyap_sum[0] += log(norm)/dt;
			}
#endif
#if JD_NBASIC != JD_N
			if (lyaps && !projector)
			{
				double norms[JD_N/JD_NBASIC];
				orthonormalise(M, n_lyap, P.max_delay, norms);
				M.invalidate_windows();
				reload_current(M);
				for (int i = 0; i < n_lyap; i++)
					lyap_sum[i] += log(norms[i])/dt;
			}
#endif
		}
		for (int i = 0; i < n_exponents; i++)
			lyaps[i] = lyap_sum[i]/number;
	}
	else if (total == 0)
	{
		/* jitcdde.integrate_blindly calls adjust_diff here (_jitcdde.py:925-926); the variants of jitcdde_lyap (:1191-1209) and of
		 * the restricted / transversal classes just make no step */
		if (!lyaps)
		{
			double mi[JD_N], ma[JD_N];
			adjust_diff(M, P, 1e-4, mi, ma);
		}
		for (int i = 0; i < n_exponents; i++)
			lyaps[i] = 0.0;
	}
	else
		M.status = JDDE_STATUS_BACKWARDS;
	/* get_current_state, jitced_template.c:330-334: state of the last anchor */
	double const * const w = M.slot(M.last);
	for (int i = 0; i < n_out; i++)
		out[i] = w[1+i];
}

/* ------------------------------------------------------------------ per-member entry points
 * (one call = what one kernel thread does; the host twin in tests/ loops over members) */

/* initial past, dde_integrator_init + initiate_past_from_list, jitced_template.c:575-652 */
JD_FN void member_set_past(jdde_problem const & P, long long const m, int const n_anchors, double const * times, double const * states, double const * diffs, int const per_member)
{
	if (per_member)
	{
		times += (size_t)m*n_anchors;
		states += (size_t)m*n_anchors*JD_N;
		diffs += (size_t)m*n_anchors*JD_N;
	}
	double * const ring = P.ring + (size_t)m*P.ring_stride;
	for (int k = 0; k < n_anchors; k++)
	{
		double * const a = ring + (size_t)k*JD_A;
		a[0] = times[k];
		for (int i = 0; i < JD_N; i++)
		{
			a[1+i] = states[(size_t)k*JD_N+i];
			a[1+JD_N+i] = diffs[(size_t)k*JD_N+i];
		}
	}
	P.first[m] = 0;
	P.last[m] = n_anchors-1;
	P.current[m] = n_anchors-1;
	for (int s = 0; s < JD_NSITES; s++)
		P.cursor[(size_t)s*P.n_members+m] = n_anchors-2;
	P.last_start[m] = times[0];
	P.status[m] = JDDE_STATUS_OK;
	P.accepted[m] = P.rejected[m] = P.attempts[m] = P.extra_rhs[m] = 0;
}

/* controller reset of set_integration_parameters, _jitcdde.py:710-728 */
JD_FN void member_reset_controller(jdde_problem const & P, long long const m)
{
	P.dt[m] = P.P.first_step;
	P.last_pws[m] = 0.0;
	P.count[m] = 0;
	P.credit[m] = 0.0;
}

/* get_recent_state (jitced_template.c:287-320) at the moment a step is accepted: the segment (current, tentative anchor)
 * is in registers — same anchors, same formula as get_recent_state reads from the ring after accept_step */
JD_FN void recent_state_of_step(Member const & M, double const t, double * const out, int const n_out)
{
	double const q = M.new_t-M.cur_t;
	double const x = (t - M.cur_t)/q;
	JD_UNROLL
	for (int index = 0; index < JD_N; index++)
	{
		double const a = M.cur_y[index];
		double const b = M.cur_d[index] * q;
		double const c = M.new_y[index];
		double const d = M.new_d[index] * q;
		if (index < n_out)
			out[index] = (1-x) * ( (1-x) * (b*x + (a-c)*(2*x+1)) - d*x*x) + c;
	}
}

/* integrate (n_steps == 0: `targets` are absolute times), step_on_discontinuities
 * (n_steps > 0: `targets` are step lengths, applied cumulatively, _jitcdde.py:986-987) or a series of
 * integrate calls (n_steps < 0: `targets` are −n_steps absolute sample times, the user's sampling loop of
 * examples/mackey_glass.py:76-78 in one launch; results out[sample][member][component]).
 * `resume`: the call repeats an earlier one that stopped for some members (ring overflow, attempt
 * budget); members continue in the step and towards the target they had reached.
 *
 * ONE stepping loop serves all three (the stepper is instantiated once) and a member never leaves it before its last
 * target: the call is a sequence of stretches — integrate: one uncapped stretch to an absolute time; sampling: one
 * per sample time; step_on_discontinuities: capped stretches of given lengths. The step that reaches the end of a
 * stretch interpolates the output there from the two anchors it has in registers (what get_recent_state would read
 * back from the ring), moves on to the next stretch and is accepted; the lanes of a warp stay together in the loop
 * instead of taking turns with an epilogue per sample. forget() runs once, at the end: its threshold only grows, so
 * one call with the final threshold drops exactly the anchors that one call per sample would have dropped (and when
 * the ring runs full on the way, ensure_room forgets at once). */
JD_FN void member_integrate(jdde_problem const & P, long long const m, double const * const targets, int const per_member, int const n_steps, int const resume, double * const out, int const n_out, long long const budget)
{
	Member M;
	load_member(P, m, M);
	if (M.status != JDDE_STATUS_OK)
		return;
	bool const capped = n_steps > 0;
	int const n_stretches = capped ? n_steps : n_steps < 0 ? -n_steps : 1;
	double const * const steps = per_member ? targets+(size_t)m*n_stretches : targets;
	double * const result = out ? out+(size_t)m*n_out : nullptr;
	size_t const result_stride = (size_t)P.n_members*n_out;                 /* between the samples of one member */
	int s = n_steps != 0 && resume ? P.resume_index[m] : 0;
	double target = capped && resume ? P.resume_target[m] : 0.0;      /* (a member that had finished keeps its last target: it is the time of the output) */
	if (s < n_stretches && !(capped && resume))
		target = capped ? M.cur_t + steps[s] : steps[s];
	/* sample times that need no step (the reference extrapolates or interpolates on the last segment, _jitcdde.py:833-834) */
	while (!capped && s < n_stretches && !(M.cur_t < target))
	{
		if (result)
			get_recent_state(M, target, result+(size_t)s*result_stride, n_out);
		if (++s < n_stretches)
			target = steps[s];
	}
	while (s < n_stretches)
	{
		/* here: cur_t < target */
		if ((M.attempts & 31) == 31 || M.last+1-M.first > M.mask)
		{
			/* Drop what no lookup can reach any more, every 32 attempts — the lanes of a warp count their attempts
			 * alike, so they do it together — instead of each lane on its own when its ring is full (one lane in ten
			 * at every attempt, with the whole warp waiting: 17 % of the stall samples, profiles/). The reference forgets
			 * at the end of integrate; forgetting earlier with a smaller threshold drops a subset of the same anchors. */
			settle_last_start(M, P);
			forget(M, P.max_delay);
			if (M.last+1-M.first > M.mask)
			{
				M.status = JDDE_STATUS_OVERFLOW;
				break;
			}
		}
		if (!(M.dt > 0.0))
		{
			/* a step size of zero/NaN would never reach the target: the reference would spin forever */
			M.status = JDDE_STATUS_NONFINITE;
			break;
		}
		if (M.attempts >= (int)budget)      /* (the caller passes at most INT_MAX) */
		{
			/* watchdog: a kernel must end; the caller clears the status and calls again to go on */
			M.status = JDDE_STATUS_BUDGET;
			break;
		}
		double const length = capped ? fmin(M.dt, target-M.cur_t) : M.dt;
		if (try_single_step(M, P, length))
		{
			/* the `while t < target` of integrate (_jitcdde.py:830-832) / of step_on_discontinuities (:988-993) ends with this step? */
			while (!(M.new_t < target))
			{
				if (!capped && result)
					recent_state_of_step(M, target, result+(size_t)s*result_stride, n_out);
				if (++s == n_stretches)
					break;
				target = capped ? M.new_t + steps[s] : steps[s];
			}
			accept_step(M);
		}
		else
		{
			count_rejection(M, P);
			if (M.status != JDDE_STATUS_OK)
				break;
		}
	}
	settle_last_start(M, P);
	if (n_steps != 0)
	{
		P.resume_index[m] = s;
		P.resume_target[m] = target;
	}
	if (M.status == JDDE_STATUS_OK)
	{
		if (capped && result)
			get_recent_state(M, target, result, n_out);
		forget(M, P.max_delay);
	}
	P.accepted[m] += M.current - P.current[m];      /* every accepted step appends exactly one anchor */
	store_member(P, m, M);
}

JD_FN void member_integrate_blindly(jdde_problem const & P, long long const m, double const * const targets, int const per_member, double const step, double * const out, int const n_out, double * const lyaps, double * const total_time, jdde_projector const * const projector = nullptr)
{
	Member M;
	load_member(P, m, M);
	if (M.status != JDDE_STATUS_OK)
		return;
	double res[JD_N];
	double ly[JD_N/JD_NBASIC];
	double total;
	integrate_blindly(M, P, per_member ? targets[m] : targets[0], step, res, n_out, lyaps ? ly : nullptr, &total, projector);
	if (out)
		for (int i = 0; i < n_out; i++)
			out[(size_t)m*n_out+i] = res[i];
	int const n_exponents = projector ? 1 : JD_N/JD_NBASIC-1;
	if (lyaps)
		for (int i = 0; i < n_exponents; i++)
			lyaps[(size_t)m*n_exponents+i] = ly[i];
	if (total_time)
		total_time[m] = total;
	store_member(P, m, M);
}

/* mode 0: jump(amplitude, time, width, forward); mode 1: adjust_diff(shift_ratio = width) */
JD_FN void member_jump(jdde_problem const & P, long long const m, int const mode, double const * const amplitude, int const amp_pm, double const * const time, int const time_pm, double const width, int const forward, double * const minima, double * const maxima)
{
	Member M;
	load_member(P, m, M);
	if (M.status != JDDE_STATUS_OK)
		return;
	double mi[JD_N], ma[JD_N];
	if (mode == 1)
		adjust_diff(M, P, width, mi, ma);
	else
	{
		double change[JD_N];
		for (int i = 0; i < JD_N; i++)
			change[i] = amp_pm ? amplitude[(size_t)m*JD_N+i] : amplitude[i];
		double const tm = time_pm ? time[m] : time[0];
		apply_jump(M, P, change, forward ? tm : tm-width, width, mi, ma);
	}
	for (int i = 0; i < JD_N; i++)
	{
		if (minima) minima[(size_t)m*JD_N+i] = mi[i];
		if (maxima) maxima[(size_t)m*JD_N+i] = ma[i];
	}
	store_member(P, m, M);
}

JD_FN void member_forget(jdde_problem const & P, long long const m)
{
	Member M;
	load_member(P, m, M);
	forget(M, P.max_delay);
	store_member(P, m, M);
}

#if JD_NBASIC != JD_N
/* orthonormalise and, if old_t is given, the tail of jitcdde_lyap.integrate (_jitcdde.py:1162-1172) */
JD_FN void member_orthonormalise(jdde_problem const & P, long long const m, int const n_lyap, double const delay, double * const norms, double * const old_t, double * const lyaps, double * const delta_t)
{
	Member M;
	load_member(P, m, M);
	if (M.status != JDDE_STATUS_OK)
		return;
	double nrm[JD_N/JD_NBASIC];
	if (old_t)
	{
		/* A member that reached the target is orthonormalised once: its old_t is then marked (NaN), so that the
		 * repetition of the call for members that had stopped on the way (full ring, jdde_resume) leaves its
		 * exponents alone. */
		if (old_t[m] != old_t[m])
			return;
		double const dt = M.cur_t-old_t[m];
		old_t[m] = NAN;
		delta_t[m] = dt;
		if (dt != 0)
		{
			orthonormalise(M, n_lyap, delay, nrm);
			for (int i = 0; i < n_lyap; i++)
				lyaps[(size_t)m*n_lyap+i] = log(nrm[i]) / dt;
		}
		else
			for (int i = 0; i < n_lyap; i++)
				lyaps[(size_t)m*n_lyap+i] = 0.0;
	}
	else
	{
		orthonormalise(M, n_lyap, delay, nrm);
		for (int i = 0; i < n_lyap; i++)
			norms[(size_t)m*n_lyap+i] = nrm[i];
	}
}
#endif

} // namespace jdde

#include "jdde_team.cuh"
#include "jdde_group.cuh"

#endif
