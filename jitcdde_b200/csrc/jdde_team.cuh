/*
 * jdde_team.cuh — orthonormalisation of the separation functions by a TEAM of threads per member
 * (kernel jdde_k_orthonormalise_team: one thread block per member) instead of one thread per member.
 *
 * Why: modified Gram–Schmidt over the history (/root/reference/jitcdde/jitced_template.c:846-878) is 1+2+…+n_lyap
 * scalar-product passes over ALL anchors of a member (hundreds, 8·JD_A bytes each). One thread streaming its own ring
 * has one cache line in flight per load and every pass goes to DRAM again (profiles/ncu_full_r01_c3_jdde_k_orthonormalise.csv:
 * 2.7 MB of traffic per member, issue slots 1.8 % busy). Here the tangent part of a member's anchors
 * (2·n_basic·n_lyap doubles + the time per anchor) is read ONCE into shared memory, component-major so that threads
 * working on neighbouring anchors hit different banks, all passes run on chip, and the result is written back once.
 * A member whose live anchors do not fit (`capacity`) is streamed: every pass brings the anchors through shared memory
 * tile by tile (each tile with the last anchor of the previous one, for the segment between them).
 *
 * Arithmetic: every anchor update and every per-segment product is formed by ONE thread from the same operands in
 * the same order as by the reference's loops. What a team changes is the order in which the per-segment parts are
 * added up: the verification build (no JD_FAST) adds them in the reference's order (one thread, ascending segments:
 * bit-identical results), the production build reduces them as a tree.
 *
 * The code is written against the small `Team` interface below; the host twin (tests/host_twin) runs it with a
 * one-thread team, which checks indexing and arithmetic on the CPU.
 */
#ifndef JDDE_TEAM_CUH
#define JDDE_TEAM_CUH

#if JD_NBASIC != JD_N

namespace jdde {

#ifndef JD_TEAM_BLOCK
#define JD_TEAM_BLOCK 512
#endif
#define JD_TEAM_FLAG_DOUBLES JDDE_TEAM_FLAG_DOUBLES      /* one byte per thread of the largest block: "this member is for this launch" */
static_assert(JD_TEAM_BLOCK <= 8*JDDE_TEAM_FLAG_DOUBLES, "flag area of jdde_k_orthonormalise_team");

struct Team
{
	int rank, size;
	double * scratch;       /* shared by the team: at least `size` doubles, followed by the two exchange slots of the cluster */
	/* A thread-block CLUSTER per member (jdde_k_orthonormalise_team launched with a cluster dimension): every block
	 * of the cluster keeps its own contiguous part of the member's anchors resident in its shared memory; sums over
	 * segments are completed across the blocks through distributed shared memory (sum() below). 1: no cluster. */
	int crank = 0, csize = 1;
	mutable int parity = 0;  /* which of the two exchange slots the next cluster-wide sum uses */
	mutable int flip = 0;    /* which half of the scratch area the next block-wide sum uses (production build) */

#if defined(__CUDA_ARCH__)
	__device__ __forceinline__ static void cluster_sync()
	{
		asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
	}
	/* the exchange slot `which` of block `block_rank` of this cluster (distributed shared memory) */
	__device__ __forceinline__ double remote_slot(int const block_rank, int const which) const
	{
		unsigned const local = (unsigned)__cvta_generic_to_shared(scratch+size+which);
		unsigned remote;
		asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(block_rank));
		double value;
		asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(value) : "r"(remote) : "memory");
		return value;
	}
#endif

	/* body(e) for all e < count, then a barrier */
	template <class F>
	JD_FN void for_each(int const count, F && body) const
	{
#if defined(__CUDA_ARCH__)
		for (int e = rank; e < count; e += size)
			body(e);
		__syncthreads();
#else
		for (int e = 0; e < count; e++)
			body(e);
#endif
	}

	/* initial + Σ_{e<count} body(e), known to every thread afterwards */
	template <class F>
	JD_FN double sum(double const initial, int const count, F && body) const
	{
#if defined(__CUDA_ARCH__)
		if (csize > 1)
			return cluster_sum(initial, count, body);
		return block_sum(initial, count, body);
#else
		double total = initial;
		for (int e = 0; e < count; e++)
			total += body(e);
		return total;
#endif
	}

#if defined(__CUDA_ARCH__)
	/*
	 * Sum over the segments of all blocks of the cluster; every block calls this with its own segments.
	 * Production build: the blocks reduce their parts concurrently, publish them in their exchange slot, and after
	 * one cluster barrier everybody adds the csize parts in rank order. Verification build: the reference adds the
	 * segments in ascending order into one accumulator, so the running total is handed from block to block (block r
	 * continues from the slot of block r−1); the per-segment products are still formed concurrently.
	 * The two exchange slots alternate, so a slot is rewritten only after a later cluster barrier has shown that
	 * everybody has read it.
	 */
	template <class F>
	__device__ __forceinline__ double cluster_sum(double const initial, int const count, F && body) const
	{
		int const which = parity;
		parity ^= 1;
		double total;
#if defined(JD_FAST)
		/* every thread adds the blocks' parts itself, in rank order: one cluster barrier per sum, nothing to broadcast */
		total = block_sum(0.0, count, body);
		if (rank == 0)
			scratch[size+which] = total;
		cluster_sync();
		total = initial;
		for (int r = 0; r < csize; r++)
			total += remote_slot(r, which);
		return total;
#else
		for (int r = 0; r < csize; r++)
		{
			if (crank == r)
			{
				double running = initial;
				if (r > 0 && rank == 0)
					running = remote_slot(r-1, which);
				total = block_sum(running, count, body);      /* only thread 0's `running` enters the sum */
				if (rank == 0)
					scratch[size+which] = total;
			}
			cluster_sync();
		}
		if (rank == 0)
			scratch[0] = remote_slot(csize-1, which);
		__syncthreads();
		total = scratch[0];
		__syncthreads();
		return total;
#endif
	}

	template <class F>
	__device__ __forceinline__ double block_sum(double const initial, int const count, F && body) const
	{
		double total = 0;
#if defined(JD_FAST)
		/* tree within a warp, then every thread adds the warps' parts in warp order (the same value everywhere); the
		 * parts alternate between two halves of the scratch area, so one barrier per sum is enough */
		for (int e = rank; e < count; e += size)
			total += body(e);
		for (int offset = 16; offset > 0; offset >>= 1)
			total += __shfl_down_sync(0xffffffffu, total, offset);
		int const n_warps = (size+31) >> 5;
		double * const parts = scratch + (flip ? n_warps : 0);
		flip ^= 1;
		if ((rank & 31) == 0)
			parts[rank >> 5] = total;
		__syncthreads();
		total = initial;
		for (int w = 0; w < n_warps; w++)
			total += parts[w];
		return total;
#else
		/* the reference's order: ascending e, one accumulator */
		total = initial;
		for (int base = 0; base < count; base += size)
		{
			if (base+rank < count)
				scratch[rank] = body(base+rank);
			__syncthreads();
			if (rank == 0)
			{
				int const n = count-base < size ? count-base : size;
				for (int r = 0; r < n; r++)
					total += scratch[r];
			}
			__syncthreads();
		}
		if (rank == 0)
			scratch[0] = total;
		__syncthreads();
		total = scratch[0];
		__syncthreads();
		return total;
#endif
	}
#endif
};

/* anchors [first, first+count) of a member, staged in shared memory: times and the tangent rows
 * row = vector·2·n_basic + (0: state | n_basic: diff) + component */
struct Tangents
{
	double * ring;
	int mask, first, count;
	double * rows;          /* [2·n_basic·n_lyap + 1][ks]; the last row holds the anchor times */
	double * times;         /* = rows + 2·n_basic·n_lyap·ks */
	int ks;                 /* odd: rows of neighbouring index start in different banks */

	static constexpr int R = 2*JD_NBASIC;

	JD_FN double * slot(int const a) const { return ring + (size_t)((first+a) & mask)*JD_A; }
	JD_FN double time(int const a) const { return times[a]; }
	JD_FN double get(int const a, int const row) const { return rows[row*ks+a]; }
	JD_FN void set(int const a, int const row, double const x) const { rows[row*ks+a] = x; }

	/* body(a, r) for the anchors a >= from and the R rows r of one vector; threads run along the anchors (consecutive banks) */
	template <class F>
	JD_FN void sweep(Team const & T, int const from, F && body) const
	{
#if defined(__CUDA_ARCH__)
		/* element e = r·span + a of the R × span block, e = rank, rank + size, …: all threads are busy also when there are
		 * fewer anchors than threads; (r, a) are carried along instead of divided out */
		int const span = count-from;
		if (span > 0)
		{
			int r = 0, a = T.rank;
			while (a >= span)
			{
				a -= span;
				r++;
			}
			while (r < R)
			{
				body(from+a, r);
				a += T.size;
				while (a >= span)
				{
					a -= span;
					r++;
				}
			}
		}
		__syncthreads();
#else
		(void)T;
		for (int r = 0; r < R; r++)
			for (int a = from; a < count; a++)
				body(a, r);
#endif
	}
};

/*
 * Ring → shared memory (in = true: the tangent part of every anchor and its time) and back (in = false: the tangent
 * part of the anchors a >= from). One warp per anchor, lanes along the anchor's doubles: global accesses are contiguous
 * runs (the state part, then the diff part of the separation functions). Several anchors per lane are in flight
 * (round 2, profiles/ncu_r02_gram_schmidt.md: the cp.async loop of round 1 waited for each copy before it could reuse
 * the address registers of the next — 19 % of the kernel's stall samples).
 */
#ifndef JD_STAGE_DEPTH
#define JD_STAGE_DEPTH 8
#endif
JD_FN void stage(Team const & T, Tangents const & H, int const n_lyap, bool const in, int const from)
{
	constexpr int R = 2*JD_NBASIC;
	int const half = JD_NBASIC*n_lyap, per_anchor = 2*half;
#if defined(__CUDA_ARCH__)
	int const lane = T.rank & 31, warp = T.rank >> 5, n_warps = T.size >> 5;
	for (int c = lane; c < per_anchor+(in ? 1 : 0); c += 32)
#else
	(void)T;
	int const warp = 0, n_warps = 1;
	for (int c = 0; c < per_anchor+(in ? 1 : 0); c++)
#endif
	{
		int row = per_anchor, source = 0;                       /* the time */
		if (c < per_anchor)
		{
			int const part = c/half, rem = c-part*half, vector = rem/JD_NBASIC, component = rem-vector*JD_NBASIC;
			row = vector*R+part*JD_NBASIC+component;
			source = 1+part*JD_N+JD_NBASIC+rem;
		}
		double * const on_chip = H.rows + row*H.ks;
		/* JD_STAGE_DEPTH anchors per lane in flight: the loads of a batch are issued before the first store waits for its value */
		for (int a0 = from+warp; a0 < H.count; a0 += JD_STAGE_DEPTH*n_warps)
		{
			double value[JD_STAGE_DEPTH] = {};
			JD_UNROLL
			for (int u = 0; u < JD_STAGE_DEPTH; u++)
			{
				int const a = a0+u*n_warps;
				if (a < H.count)
					value[u] = in ? H.slot(a)[source] : on_chip[a];
			}
			JD_UNROLL
			for (int u = 0; u < JD_STAGE_DEPTH; u++)
			{
				int const a = a0+u*n_warps;
				if (a < H.count)
				{
					if (in)
						on_chip[a] = value[u];
					else
						H.slot(a)[source] = value[u];
				}
			}
		}
	}
#if defined(__CUDA_ARCH__)
	__syncthreads();
#endif
}

/*
 * Counterpart of orthonormalise_pass (jdde_engine.cuh) on the staged anchors; vectors are numbered from 0. Anchors
 * a >= update_from are updated (a tile's first anchor may be the already updated last one of the previous tile);
 * `leading` says that local anchor 0 is the member's first anchor; returns initial + the staged segments' parts.
 */
JD_FN double team_pass(Team const & T, Tangents const & H, double const threshold, int const sub_i, int const sub_j, double const sub_factor,
	int const scale_i, double const scale_factor, int const acc_1, int const acc_2, int const update_from, bool const leading, double const initial)
{
	constexpr int R = Tangents::R;
	if (sub_i >= 0)
		H.sweep(T, update_from, [&](int const a, int const r) { H.set(a, sub_i*R+r, H.get(a, sub_i*R+r) - sub_factor*H.get(a, sub_j*R+r)); });
	if (scale_i >= 0)
		H.sweep(T, update_from, [&](int const a, int const r) { H.set(a, scale_i*R+r, H.get(a, scale_i*R+r)*scale_factor); });
	if (acc_1 < 0)
		return initial;

	return T.sum(initial, H.count-1, [&](int const segment) -> double
	{
		int const a = segment+1;
		double const tw = H.time(a);
		if (tw < threshold)
			return 0.0;
		double const tv = H.time(a-1);
		double const q = tw - tv;
		double m[4][4];
		if ((leading && a == 1) || tv < threshold)      /* the first segment that reaches into the window */
			partial_sp_matrix(q, (threshold - tw) / q, m);
		else
			sp_matrix(q, m);
		int const b1 = acc_1*R, b2 = acc_2*R;
		/* vector_x[i]: i = 0 v.state, 1 v.diff, 2 w.state, 3 w.diff → (anchor a-1+(i>>1), row base + (i&1)·n_basic) */
		double part = 0;
#if defined(JD_FAST)
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
					dot += H.get(a-1+(i>>1), b1+(i&1)*JD_NBASIC+index) * H.get(a-1+(j>>1), b2+(j&1)*JD_NBASIC+index);
				row[i] += m[i][j]*dot;
			}
		}
		part = (row[0]+row[1])+(row[2]+row[3]);
#else
		for (int i = 0; i < 4; i++)
			for (int j = 0; j < 4; j++)
				for (int index = 0; index < JD_NBASIC; index++)
					part += m[i][j] * H.get(a-1+(i>>1), b1+(i&1)*JD_NBASIC+index) * H.get(a-1+(j>>1), b2+(j&1)*JD_NBASIC+index);
#endif
		return part;
	});
}

/* orthonormalise (jdde_engine.cuh): the sequence of passes, each carried out by pass(sub_i, sub_j, sub_factor, scale_i,
 * scale_factor, acc_1, acc_2) → scalar product; norms are known to every thread */
template <class Pass>
JD_FN void team_gram_schmidt(Pass && pass, int const n_lyap, double * const norms)
{
	int pending_scale = -1;
	double pending_factor = 1.0;
	for (int i = 0; i < n_lyap; i++)
	{
		int sub_j = -1;
		double sub_factor = 0.0;
		for (int j = 0; j < i; j++)
		{
			double const sp = pass(sub_j >= 0 ? i : -1, sub_j, sub_factor, pending_scale, pending_factor, i, j);
			pending_scale = -1;
			sub_j = j;
			sub_factor = sp;
		}
		double const norm = sqrt(pass(sub_j >= 0 ? i : -1, sub_j, sub_factor, pending_scale, pending_factor, i, i));
		pending_scale = -1;
		if (norm > JD_NORM_THRESHOLD)
		{
			pending_scale = i;
			pending_factor = 1./norm;
		}
		norms[i] = norm;
	}
	if (pending_scale >= 0)
		pass(-1, -1, 0.0, pending_scale, pending_factor, -1, -1);
}

/* is member m for the launch that takes the histories of count_from … count_to anchors? (the conditions under which
 * team_member_orthonormalise returns at once) */
JD_FN bool team_has_work(jdde_problem const & P, long long const m, double const * const old_t, int const count_from, int const count_to)
{
	if (P.status[m] != JDDE_STATUS_OK)
		return false;
	if (old_t && old_t[m] != old_t[m])
		return false;
	int const count = P.last[m]-P.first[m]+1;
	return count >= count_from && count <= count_to;
}

/*
 * member_orthonormalise (jdde_engine.cuh) for a team. `shared` = rows[2·n_basic·n_lyap + 1][ks] with ks = capacity | 1
 * (the team's scratch lies elsewhere); capacity >= 4 anchors.
 */
JD_FN void team_member_orthonormalise(Team const & T_given, jdde_problem const & P, long long const m, int const n_lyap, double const delay,
	double * const norms, double * const old_t, double * const lyaps, double * const delta_t, double * const shared, int const capacity,
	int const count_from = 0, int const count_to = 0x7fffffff)
{
	Team T = T_given;
	if (P.status[m] != JDDE_STATUS_OK)
		return;
	if (old_t && old_t[m] != old_t[m])      /* already orthonormalised by the call that is being resumed (member_orthonormalise) */
		return;
	constexpr int R = 2*JD_NBASIC;
	Tangents H;
	H.ring = P.ring + (size_t)m*P.ring_stride;
	H.mask = P.cap-1;
	H.ks = capacity | 1;
	H.rows = shared;
	H.times = shared + R*n_lyap*H.ks;
	int const first = P.first[m], count = P.last[m]-first+1;
	if (count < count_from || count > count_to)      /* this launch is for histories of another length (jdde_runtime.cu) */
		return;
	double const cur_t = H.ring[(size_t)(P.current[m] & H.mask)*JD_A];
	double const threshold = cur_t-delay;

	double dt = 0;
	if (old_t)
	{
		dt = cur_t-old_t[m];
		if (dt == 0)
		{
			T.for_each(n_lyap+1, [&](int const i)
			{
				if (i < n_lyap)
					lyaps[(size_t)m*n_lyap+i] = 0.0;
				else
					delta_t[m] = dt;
			});
			return;
		}
	}

	double nrm[JD_N/JD_NBASIC];
	bool const writer = T.crank == 0;       /* the block of a cluster that reports the results */
	if (T.csize > 1 && (count + T.csize - 1)/T.csize + 1 <= capacity)
	{
		/* resident in the cluster: block r keeps the anchors [lo, hi) plus a copy of anchor lo−1 (for the segment between
		 * the parts; it receives the same updates, computed from the same operands, and is not written back) */
		int const chunk = (count + T.csize - 1)/T.csize;
		int const lo = T.crank*chunk < count ? T.crank*chunk : count;
		int const hi = lo+chunk < count ? lo+chunk : count;
		int const halo = (lo > 0 && hi > lo) ? 1 : 0;
		H.first = first+lo-halo;
		H.count = hi-lo+halo;
		stage(T, H, n_lyap, true, 0);
		team_gram_schmidt([&](int const sub_i, int const sub_j, double const sub_factor, int const scale_i, double const scale_factor, int const acc_1, int const acc_2)
		{
			return team_pass(T, H, threshold, sub_i, sub_j, sub_factor, scale_i, scale_factor, acc_1, acc_2, 0, lo == 0, 0.0);
		}, n_lyap, nrm);
		stage(T, H, n_lyap, false, halo);
	}
	else if (T.csize > 1 && !writer)
		return;                             /* too long even for the cluster: its first block streams the member alone */
	else if (count <= capacity)
	{
		T.csize = 1;
		/* resident: one read, all passes on chip, one write */
		H.first = first;
		H.count = count;
		stage(T, H, n_lyap, true, 0);
		team_gram_schmidt([&](int const sub_i, int const sub_j, double const sub_factor, int const scale_i, double const scale_factor, int const acc_1, int const acc_2)
		{
			return team_pass(T, H, threshold, sub_i, sub_j, sub_factor, scale_i, scale_factor, acc_1, acc_2, 0, true, 0.0);
		}, n_lyap, nrm);
		stage(T, H, n_lyap, false, 0);
	}
	else
	{
		T.csize = 1;
		/* streamed: every pass goes through the anchors in tiles [lo, hi); a tile is staged together with anchor lo-1 */
		team_gram_schmidt([&](int const sub_i, int const sub_j, double const sub_factor, int const scale_i, double const scale_factor, int const acc_1, int const acc_2)
		{
			double sum = 0.0;
			for (int lo = 0; lo < count; )
			{
				int const halo = lo > 0 ? 1 : 0;
				int const hi = count-lo <= capacity-halo ? count : lo+capacity-halo;
				H.first = first+lo-halo;
				H.count = hi-lo+halo;
				stage(T, H, n_lyap, true, 0);
				sum = team_pass(T, H, threshold, sub_i, sub_j, sub_factor, scale_i, scale_factor, acc_1, acc_2, halo, lo == 0, sum);
				if (sub_i >= 0 || scale_i >= 0)
					stage(T, H, n_lyap, false, halo);
				lo = hi;
			}
			return sum;
		}, n_lyap, nrm);
	}

	if (!writer)
		return;
	T.for_each(n_lyap+1, [&](int const i)
	{
		if (i < n_lyap)
		{
			if (old_t)
				lyaps[(size_t)m*n_lyap+i] = log(nrm[i]) / dt;
			else
				norms[(size_t)m*n_lyap+i] = nrm[i];
		}
		else if (old_t)
		{
			delta_t[m] = dt;
			old_t[m] = NAN;
		}
	});
}

} // namespace jdde

#endif
#endif
