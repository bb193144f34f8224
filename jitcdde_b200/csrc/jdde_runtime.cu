/*
 * jdde_runtime.cu — libjitcdde_b200.so: the C ABI of include/jitcdde_b200.h.
 *
 * Host-side runtime of the engine: owns the device memory of one ensemble (layout: jdde_abi.h),
 * loads the per-DDE kernel image (jdde_kernels.cu compiled for sm_100a) with
 * cudaLibraryLoadData and launches its kernels on one CUDA stream. It contains no device code
 * and no numerical code of its own: all arithmetic of the hot path is in jdde_engine.cuh.
 *
 * It replaces the CPython glue of the reference's generated module
 * (/root/reference/jitcdde/jitced_template.c:1178-1268) and the module build/load plumbing of
 * jitcxde_common used at /root/reference/jitcdde/_jitcdde.py:560-588.
 *
 * There is deliberately no CPU fallback: every entry point needs a CUDA device.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <new>

#include "jdde_abi.h"

namespace {

thread_local std::string g_create_error;

struct Kernels
{
	cudaKernel_t set_past = nullptr, reset_controller = nullptr, integrate = nullptr, integrate_blindly = nullptr,
		jump = nullptr, forget = nullptr, get = nullptr, clear_status = nullptr, sum_counters = nullptr, count_failed = nullptr,
		regrow = nullptr, orthonormalise = nullptr, orthonormalise_team = nullptr, integrate_group = nullptr, project = nullptr, integrate_blindly_projected = nullptr;
};

} // namespace

struct jdde_handle
{
	jdde_desc desc{};
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	cudaLibrary_t lib = nullptr;
	Kernels k;
	int block = 128;
	int smem_bytes = 0;      /* dynamic shared memory per block (anchor staging window of the image) */
	int group_lanes = 0, group_block = 0, group_smem_bytes = 0;   /* jdde_k_integrate_group: lanes per member, block size (0: one thread per member) */
	/* pipelined sampling (jdde_integrate_async): second result buffer, copy stream, events per buffer */
	cudaStream_t copy_stream = nullptr;
	double * d_out_alt = nullptr;
	double * d_samples[2] = { nullptr, nullptr };  size_t samples_count[2] = { 0, 0 };   /* results of jdde_integrate_samples that cannot go straight to the host */
	cudaEvent_t kernel_done[2] = { nullptr, nullptr }, copy_done[2] = { nullptr, nullptr };
	bool copy_pending[2] = { false, false };
	bool result_direct[2] = { false, false }, result_queued[2] = { false, false };
	int out_slot = 0;
	bool queueing = false;    /* inside jdde_integrate_async: the only entry point that does not wait for pending copies */
	int team_smem_max = 0;   /* dynamic shared memory the team kernel may use (0: not yet asked) */
	int sm_count = 0;
	jdde_problem P{};
	bool have_past = false, have_params = false, have_int_params = false;
	long long budget = 1ll << 24;
	int64_t launches = 0;
	std::string err;

	/* device scratch */
	double * d_in = nullptr;  size_t in_bytes = 0;      /* staged inputs */
	double * d_out = nullptr;                            /* [N][n] result of the last integration */
	double * d_out2 = nullptr;                           /* [N][n] second result (maxima, lyaps) */
	double * d_aux = nullptr;                            /* [N] (old t, delta t) */
	double * d_aux2 = nullptr;                           /* [N] */
	unsigned long long * d_sums = nullptr;               /* [3] */
	/* last integration call, for jdde_resume */
	int last_kind = 0;                                   /* 0 none, 1 integrate, 2 step_on_discontinuities, 3 integrate_lyap, 4 integrate_samples */
	int last_per_member = 0, last_n_steps = 0;
	std::vector<double> last_targets;
	std::vector<void *> allocations;
	/* restricted / transversal Lyapunov exponents (jdde_set_projector): device copies of the arrays */
	jdde_projector projector{};
	void * projector_memory = nullptr;
};

namespace {

int fail(jdde_handle * h, int code, const std::string & msg)
{
	if (h)
		h->err = msg;
	else
		g_create_error = msg;
	return code;
}

#define JD_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(h, JDDE_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

template <class T>
int dev_alloc(jdde_handle * h, T ** p, size_t count)
{
	void * q = nullptr;
	cudaError_t e = cudaMalloc(&q, count*sizeof(T) ? count*sizeof(T) : 1);
	if (e != cudaSuccess)
		return fail(h, e == cudaErrorMemoryAllocation ? JDDE_E_NOMEM : JDDE_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
	h->allocations.push_back(q);
	*p = static_cast<T *>(q);
	return JDDE_OK;
}

int ensure_stage(jdde_handle * h, size_t doubles)
{
	size_t const need = doubles*sizeof(double);
	if (need <= h->in_bytes)
		return JDDE_OK;
	if (h->d_in)
	{
		JD_CUDA(h, cudaStreamSynchronize(h->stream));
		JD_CUDA(h, cudaFree(h->d_in));
		h->d_in = nullptr;
		h->in_bytes = 0;
	}
	JD_CUDA(h, cudaMalloc((void **)&h->d_in, need*2));
	h->in_bytes = need*2;
	return JDDE_OK;
}

/* host → device copy of an input into the staging buffer (same stream as the kernels: ordered) */
int stage_in(jdde_handle * h, const double * host, size_t count, double ** dev, size_t offset_doubles = 0)
{
	int rc = ensure_stage(h, offset_doubles+count);
	if (rc) return rc;
	*dev = h->d_in+offset_doubles;
	JD_CUDA(h, cudaMemcpyAsync(*dev, host, count*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	return JDDE_OK;
}

/*
 * Device-side alias of a caller's result buffer if it is page-locked host memory the GPU can address (cudaHostAlloc /
 * jdde_host_alloc, registered memory): the kernel then writes its results straight into it over PCIe — posted
 * writes spread over the kernel's run time, no device-side result buffer, no copy afterwards. nullptr otherwise.
 */
double * direct_alias(double * host)
{
	if (!host)
		return nullptr;
	cudaPointerAttributes attributes{};
	if (cudaPointerGetAttributes(&attributes, host) != cudaSuccess)
	{
		(void)cudaGetLastError();
		return nullptr;
	}
	const char * const choice = getenv("JITCDDE_B200_DIRECT");
	if (choice && choice[0] == '0')
		return nullptr;
	if (attributes.type == cudaMemoryTypeHost && attributes.devicePointer)
		return static_cast<double *>(attributes.devicePointer);
	return nullptr;
}

int launch(jdde_handle * h, cudaKernel_t kernel, void ** args)
{
	if (!kernel)
		return fail(h, JDDE_E_INVALID, "kernel image not loaded (jdde_load_rhs) or kernel missing from the image");
	long long const N = h->desc.n_members;
	unsigned const grid = (unsigned)((N + h->block - 1)/h->block);
	JD_CUDA(h, cudaLaunchKernel((const void *)kernel, dim3(grid), dim3(h->block), args, (size_t)h->smem_bytes, h->stream));
	h->launches++;
	return JDDE_OK;
}

/* integrate / step_on_discontinuities: Lyapunov images bring a kernel with a group of lanes per member
 * (csrc/jdde_group.cuh; 32/lanes members per warp), all others one thread per member */
int launch_integrate(jdde_handle * h, void ** args)
{
	if (!h->group_lanes)
		return launch(h, h->k.integrate, args);
	long long const N = h->desc.n_members;
	long long const per_block = (long long)(h->group_block/32)*(32/h->group_lanes);
	unsigned const grid = (unsigned)((N + per_block - 1)/per_block);
	JD_CUDA(h, cudaLaunchKernel((const void *)h->k.integrate_group, dim3(grid), dim3(h->group_block), args, (size_t)h->group_smem_bytes, h->stream));
	h->launches++;
	return JDDE_OK;
}

/*
 * Gram–Schmidt over the history with one thread block per member (csrc/jdde_team.cuh). The tangent part of a member's
 * anchors is staged in shared memory: up to the ring capacity if that fits into the opt-in maximum of the device
 * (227 KB on sm_100), else as many anchors as fit (longer histories are streamed through in tiles, pass by pass).
 * JITCDDE_B200_TEAM=0 selects the one-thread-per-member kernel.
 */
int launch_orthonormalise(jdde_handle * h, int n_lyap, double delay, double * norms, double * old_t, double * lyaps, double * delta_t)
{
	const char * const choice = getenv("JITCDDE_B200_TEAM");
	if (!h->k.orthonormalise_team || (choice && choice[0] == '0'))
	{
		void * args[] = { &h->P, &n_lyap, &delay, &norms, &old_t, &lyaps, &delta_t };
		return launch(h, h->k.orthonormalise, args);
	}
	int const per_anchor = 1 + 2*h->desc.n_basic*n_lyap;
	int const cap = h->P.cap;
	int block = ((cap/2 + 31)/32)*32;
	block = block < 64 ? 64 : block > 512 ? 512 : block;
	if (h->team_smem_max == 0)
	{
		int optin = 0;
		if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->desc.device) != cudaSuccess || optin < 48*1024)
			optin = 48*1024;
		if (optin > 48*1024 && cudaFuncSetAttribute((const void *)h->k.orthonormalise_team, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) != cudaSuccess)
		{
			(void)cudaGetLastError();
			optin = 48*1024;
		}
		h->team_smem_max = optin;
		if (cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->desc.device) != cudaSuccess)
			h->sm_count = 148;
	}
	/* JITCDDE_B200_TEAM_SMEM (bytes per block) and JITCDDE_B200_TEAM_THREADS trade the size of a block for the number of
	 * blocks per SM: smaller blocks let one member's barriers and staging overlap another member's arithmetic, and the
	 * anchors of a member then span more blocks of a cluster (profiles/ncu_r02_gram_schmidt.md) */
	const char * const smem_choice = getenv("JITCDDE_B200_TEAM_SMEM");
	const char * const threads_choice = getenv("JITCDDE_B200_TEAM_THREADS");
	int smem_limit = h->team_smem_max;
	if (smem_choice && atoi(smem_choice) >= 16*1024 && atoi(smem_choice) < smem_limit)
		smem_limit = atoi(smem_choice);
	int team_threads = 512;
	if (threads_choice && atoi(threads_choice) >= 64 && atoi(threads_choice) <= 512)
		team_threads = (atoi(threads_choice)/32)*32;
	if (block > team_threads)
		block = team_threads;
	long long const room = smem_limit/(long long)sizeof(double) - 512 - 2 - JDDE_TEAM_FLAG_DOUBLES;     /* doubles left for the staged anchors */
	int capacity = cap;
	if ((long long)(capacity | 1)*per_anchor > room)
		capacity = (int)(room/per_anchor) - 1;
	if (capacity < 4)       /* anchors too large for the staging area: one thread per member */
	{
		void * targs[] = { &h->P, &n_lyap, &delay, &norms, &old_t, &lyaps, &delta_t };
		return launch(h, h->k.orthonormalise, targs);
	}
	/*
	 * One launch per history length class. Members whose live anchors fit one block's staging area (`capacity`) are
	 * handled by a block each. If the ring can hold more, further launches with a cluster of 2, 4, 8 blocks per member
	 * take the members with up to 2, 4, 8 × (capacity − 1) live anchors — each block keeps its part resident, sums
	 * are completed through distributed shared memory (csrc/jdde_team.cuh) — and the last launch also streams
	 * anything longer. A launch skips the members of the other classes at once. JITCDDE_B200_CLUSTER limits the
	 * cluster size (1: no clusters, longer histories are streamed by one block).
	 */
	const char * const cluster_choice = getenv("JITCDDE_B200_CLUSTER");
	int const cluster_max = cluster_choice ? atoi(cluster_choice) : 8;
	long long const N = h->desc.n_members;
	int count_from = 0;
	for (int cluster = 1; count_from <= cap; cluster *= 2)
	{
		long long const fits = cluster == 1 ? capacity : (long long)cluster*(capacity-1);
		bool const final_launch = fits >= cap || 2*cluster > cluster_max;
		int count_to = final_launch ? 0x7fffffff : (int)fits;
		int const threads = cluster == 1 ? block : team_threads;
		size_t const smem = sizeof(double)*((size_t)threads + 2 + JDDE_TEAM_FLAG_DOUBLES + (size_t)(capacity | 1)*per_anchor);
		/* cluster launches: one wave of clusters looping over the members (blocks of 200 KB start slowly) */
		long long per_sm = h->team_smem_max/(long long)(smem+1024);      /* resident blocks per SM: shared memory, 128 registers per thread */
		if (per_sm > 65536/(128*threads))
			per_sm = 65536/(128*threads);
		if (per_sm < 1)
			per_sm = 1;
		long long const most = cluster == 1 ? (long long)h->sm_count*32 : (long long)h->sm_count*per_sm/cluster;
		unsigned const grid = (unsigned)(N < most ? N : most)*cluster;
		void * args[] = { &h->P, &n_lyap, &delay, &norms, &old_t, &lyaps, &delta_t, &capacity, &count_from, &count_to };
		if (getenv("JITCDDE_B200_DEBUG"))      /* launch shape and kernel attributes on stderr */
		{
			cudaFuncAttributes fa{};
			cudaError_t const e = cudaFuncGetAttributes(&fa, (const void *)h->k.orthonormalise_team);
			fprintf(stderr, "[jdde] orthonormalise_team: grid %u block %d cluster %d smem %zu capacity %d (ring %d) live anchors %d..%d smem_max %d | attributes: %s maxThreads %d regs %d local %zu maxDynamic %d static %zu\n",
				grid, threads, cluster, smem, capacity, cap, count_from, count_to, h->team_smem_max, cudaGetErrorString(e), fa.maxThreadsPerBlock, fa.numRegs, fa.localSizeBytes, fa.maxDynamicSharedSizeBytes, fa.sharedSizeBytes);
			(void)cudaGetLastError();
		}
		if (cluster > 1)
		{
			cudaLaunchConfig_t config{};
			config.gridDim = dim3(grid);
			config.blockDim = dim3(threads);
			config.dynamicSmemBytes = smem;
			config.stream = h->stream;
			cudaLaunchAttribute attribute{};
			attribute.id = cudaLaunchAttributeClusterDimension;
			attribute.val.clusterDim.x = cluster;
			attribute.val.clusterDim.y = 1;
			attribute.val.clusterDim.z = 1;
			config.attrs = &attribute;
			config.numAttrs = 1;
			JD_CUDA(h, cudaLaunchKernelExC(&config, (const void *)h->k.orthonormalise_team, args));
		}
		else
			JD_CUDA(h, cudaLaunchKernel((const void *)h->k.orthonormalise_team, dim3(grid), dim3(threads), args, smem, h->stream));
		h->launches++;
		if (final_launch)
			break;
		count_from = count_to+1;
	}
	return JDDE_OK;
}

int ready(jdde_handle * h)
{
	if (!h)
		return JDDE_E_INVALID;
	if (cudaSetDevice(h->desc.device) != cudaSuccess)
		return fail(h, JDDE_E_CUDA, "cudaSetDevice failed");
	if (!h->queueing && (h->copy_pending[0] || h->copy_pending[1]))
	{
		/* results of jdde_integrate_async still on their way: every other call waits for them first, so that it
		 * may reuse the result buffers */
		JD_CUDA(h, cudaStreamSynchronize(h->copy_stream));
		h->copy_pending[0] = h->copy_pending[1] = false;
	}
	return JDDE_OK;
}

int ready_to_integrate(jdde_handle * h)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!h->lib) return fail(h, JDDE_E_INVALID, "no kernel image loaded (jdde_load_rhs)");
	if (!h->have_past) return fail(h, JDDE_E_INVALID, "no past set (jdde_set_past)");
	if (!h->have_int_params) return fail(h, JDDE_E_INVALID, "integration parameters not set (jdde_set_integration_parameters)");
	if (h->desc.n_params && !h->have_params) return fail(h, JDDE_E_INVALID, "control parameters not set (jdde_set_parameters)");
	return JDDE_OK;
}

/* copy results to the caller and synchronise, if the caller wants anything */
int finish(jdde_handle * h, double * out, const double * d_src, size_t count, int32_t * status)
{
	bool sync = false;
	if (out)
	{
		JD_CUDA(h, cudaMemcpyAsync(out, d_src, count*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
		sync = true;
	}
	if (status)
	{
		JD_CUDA(h, cudaMemcpyAsync(status, h->P.status, (size_t)h->desc.n_members*sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
		sync = true;
	}
	if (sync)
		JD_CUDA(h, cudaStreamSynchronize(h->stream));
	return JDDE_OK;
}

} // namespace

extern "C" {

int jdde_create(const jdde_desc * desc, jdde_handle ** out)
{
	if (!desc || !out)
		return fail(nullptr, JDDE_E_INVALID, "null argument");
	*out = nullptr;
	if (desc->n < 1 || desc->n_basic < 1 || desc->n % desc->n_basic || desc->n_sites < 0 || desc->n_params < 0 || desc->n_members < 1)
		return fail(nullptr, JDDE_E_INVALID, "invalid descriptor");
	if (desc->ring_capacity < 4 || (desc->ring_capacity & (desc->ring_capacity-1)))
		return fail(nullptr, JDDE_E_INVALID, "ring_capacity must be a power of two >= 4");
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count < 1)
		return fail(nullptr, JDDE_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this engine has no CPU fallback)");
	if (desc->device < 0 || desc->device >= count)
		return fail(nullptr, JDDE_E_INVALID, "device ordinal out of range");

	jdde_handle * h = new (std::nothrow) jdde_handle;
	if (!h)
		return fail(nullptr, JDDE_E_NOMEM, "out of host memory");
	h->desc = *desc;
	auto bail = [&](int rc) { g_create_error = h->err; jdde_destroy(h); return rc; };
	if (cudaSetDevice(desc->device) != cudaSuccess)
		return bail(fail(h, JDDE_E_CUDA, "cudaSetDevice failed"));
	if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
		return bail(fail(h, JDDE_E_CUDA, "cudaStreamCreate failed"));
	h->own_stream = true;

	size_t const N = (size_t)desc->n_members;
	int const A = jdde_anchor_stride(desc->n);
	jdde_problem & P = h->P;
	P.n_members = desc->n_members;
	P.cap = desc->ring_capacity;
	P.anchor_stride = A;
	P.ring_stride = jdde_ring_stride(P.cap, A);
	int rc;
	if ((rc = dev_alloc(h, &P.ring, N*(size_t)P.ring_stride))) return bail(rc);
	if ((rc = dev_alloc(h, &P.par, N*(size_t)(desc->n_params ? desc->n_params : 1)))) return bail(rc);
	if ((rc = dev_alloc(h, &P.dt, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.last_pws, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.credit, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.last_start, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.first, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.last, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.current, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.count, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.status, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.cursor, N*(size_t)(desc->n_sites ? desc->n_sites : 1)))) return bail(rc);
	if ((rc = dev_alloc(h, &P.accepted, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.rejected, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.attempts, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.extra_rhs, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.resume_index, N))) return bail(rc);
	if ((rc = dev_alloc(h, &P.resume_target, N))) return bail(rc);
	if ((rc = dev_alloc(h, &h->d_out, N*(size_t)desc->n))) return bail(rc);
	if ((rc = dev_alloc(h, &h->d_out2, N*(size_t)desc->n))) return bail(rc);
	if ((rc = dev_alloc(h, &h->d_aux, N))) return bail(rc);
	if ((rc = dev_alloc(h, &h->d_aux2, N))) return bail(rc);
	if ((rc = dev_alloc(h, &h->d_sums, 3))) return bail(rc);
	if (cudaMemsetAsync(P.status, 0, N*sizeof(int), h->stream) != cudaSuccess)
		return bail(fail(h, JDDE_E_CUDA, "cudaMemset failed"));
	*out = h;
	return JDDE_OK;
}

int jdde_destroy(jdde_handle * h)
{
	if (!h)
		return JDDE_OK;
	cudaSetDevice(h->desc.device);
	if (h->stream)
		cudaStreamSynchronize(h->stream);
	if (h->copy_stream)
	{
		cudaStreamSynchronize(h->copy_stream);
		cudaStreamDestroy(h->copy_stream);
	}
	for (int s = 0; s < 2; s++)
	{
		if (h->kernel_done[s]) cudaEventDestroy(h->kernel_done[s]);
		if (h->copy_done[s]) cudaEventDestroy(h->copy_done[s]);
	}
	for (void * p : h->allocations)
		cudaFree(p);
	if (h->d_in)
		cudaFree(h->d_in);
	for (double * p : h->d_samples)
		if (p)
			cudaFree(p);
	if (h->projector_memory)
		cudaFree(h->projector_memory);
	if (h->lib)
		cudaLibraryUnload(h->lib);
	if (h->own_stream && h->stream)
		cudaStreamDestroy(h->stream);
	delete h;
	return JDDE_OK;
}

const char * jdde_last_error(const jdde_handle * h)
{
	return h ? h->err.c_str() : g_create_error.c_str();
}

int jdde_load_rhs(jdde_handle * h, const void * image, size_t len)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!image || !len)
		return fail(h, JDDE_E_INVALID, "empty image");
	if (h->lib)
	{
		cudaLibraryUnload(h->lib);
		h->lib = nullptr;
		h->k = Kernels();
	}
	JD_CUDA(h, cudaLibraryLoadData(&h->lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0));

	void * dptr = nullptr;
	size_t bytes = 0;
	if (cudaLibraryGetGlobal(&dptr, &bytes, h->lib, "jdde_image_desc") != cudaSuccess || bytes != sizeof(jdde_image_desc_t))
		return fail(h, JDDE_E_IMAGE, "image has no jdde_image_desc");
	jdde_image_desc_t d;
	JD_CUDA(h, cudaMemcpy(&d, dptr, sizeof(d), cudaMemcpyDeviceToHost));
	if (d.magic != JDDE_IMAGE_MAGIC || d.abi != JDDE_ABI_VERSION)
		return fail(h, JDDE_E_IMAGE, "image built for another ABI version");
	if (d.n != h->desc.n || d.n_basic != h->desc.n_basic || d.n_sites != h->desc.n_sites || d.n_params != h->desc.n_params || d.anchor_stride != h->P.anchor_stride)
	{
		char buf[256];
		snprintf(buf, sizeof buf, "image (n=%d, n_basic=%d, sites=%d, params=%d) does not match descriptor (n=%d, n_basic=%d, sites=%d, params=%d)",
			d.n, d.n_basic, d.n_sites, d.n_params, h->desc.n, h->desc.n_basic, h->desc.n_sites, h->desc.n_params);
		return fail(h, JDDE_E_IMAGE, buf);
	}
	h->block = d.block;
	h->smem_bytes = d.smem_bytes;
	h->group_lanes = d.group_lanes;
	h->group_block = d.group_block;
	h->group_smem_bytes = d.group_smem_bytes;
	h->team_smem_max = 0;

	struct { const char * name; cudaKernel_t * slot; bool required; } const table[] = {
		{"jdde_k_set_past", &h->k.set_past, true},
		{"jdde_k_reset_controller", &h->k.reset_controller, true},
		{"jdde_k_integrate", &h->k.integrate, true},
		{"jdde_k_integrate_blindly", &h->k.integrate_blindly, true},
		{"jdde_k_jump", &h->k.jump, true},
		{"jdde_k_forget", &h->k.forget, true},
		{"jdde_k_get", &h->k.get, true},
		{"jdde_k_clear_status", &h->k.clear_status, true},
		{"jdde_k_sum_counters", &h->k.sum_counters, true},
		{"jdde_k_count_failed", &h->k.count_failed, true},
		{"jdde_k_regrow", &h->k.regrow, true},
		{"jdde_k_orthonormalise", &h->k.orthonormalise, false},
		{"jdde_k_orthonormalise_team", &h->k.orthonormalise_team, false},
		{"jdde_k_integrate_group", &h->k.integrate_group, false},
		{"jdde_k_project", &h->k.project, false},
		{"jdde_k_integrate_blindly_projected", &h->k.integrate_blindly_projected, false},
	};
	for (auto const & entry : table)
	{
		cudaError_t e = cudaLibraryGetKernel(entry.slot, h->lib, entry.name);
		if (e != cudaSuccess)
		{
			*entry.slot = nullptr;
			(void)cudaGetLastError();
			if (entry.required)
				return fail(h, JDDE_E_IMAGE, std::string("image lacks kernel ") + entry.name);
		}
	}
	if (h->smem_bytes > 48*1024)
		JD_CUDA(h, cudaFuncSetAttribute((const void *)h->k.integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
	/* (no preferred carve-out: the driver sizes shared memory for the resident blocks and leaves the rest to L1, which
	 * the register spills of the stepper live in — forcing the maximum carve-out cost 25 % of the throughput, profiles/) */
	if (getenv("JITCDDE_B200_CARVEOUT") && h->smem_bytes > 0)
	{
		if (cudaFuncSetAttribute((const void *)h->k.integrate, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("JITCDDE_B200_CARVEOUT"))) != cudaSuccess)
			(void)cudaGetLastError();
	}
	if (h->group_lanes && !h->k.integrate_group)
		return fail(h, JDDE_E_IMAGE, "image announces a lane group per member but lacks jdde_k_integrate_group");
	if (h->group_lanes && h->group_smem_bytes > 48*1024)
		JD_CUDA(h, cudaFuncSetAttribute((const void *)h->k.integrate_group, cudaFuncAttributeMaxDynamicSharedMemorySize, h->group_smem_bytes));
	return JDDE_OK;
}

int jdde_set_stream(jdde_handle * h, void * cuda_stream)
{
	int rc = ready(h);
	if (rc) return rc;
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	if (h->own_stream)
	{
		cudaStreamDestroy(h->stream);
		h->own_stream = false;
	}
	h->stream = static_cast<cudaStream_t>(cuda_stream);
	return JDDE_OK;
}

int jdde_synchronize(jdde_handle * h)
{
	int rc = ready(h);
	if (rc) return rc;
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	if (h->copy_stream)
	{
		JD_CUDA(h, cudaStreamSynchronize(h->copy_stream));
		h->copy_pending[0] = h->copy_pending[1] = false;
	}
	return JDDE_OK;
}

int jdde_set_past(jdde_handle * h, int32_t n_anchors, const double * times, const double * states, const double * diffs, int32_t per_member)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!h->lib) return fail(h, JDDE_E_INVALID, "no kernel image loaded (jdde_load_rhs)");
	if (n_anchors < 2)
		return fail(h, JDDE_E_INVALID, "need at least two anchors (reference: _jitcdde.py:582)");
	if (n_anchors+1 > h->P.cap)
		return fail(h, JDDE_E_INVALID, "initial past does not fit into the ring; increase ring_capacity");
	if (!times || !states || !diffs)
		return fail(h, JDDE_E_INVALID, "null argument");
	size_t const members = per_member ? (size_t)h->desc.n_members : 1;
	size_t const nt = members*(size_t)n_anchors, ns = nt*(size_t)h->desc.n;
	/* one staging buffer: [times | states | diffs] */
	if ((rc = ensure_stage(h, nt+2*ns))) return rc;
	double * d_times, * d_states, * d_diffs;
	if ((rc = stage_in(h, times, nt, &d_times, 0))) return rc;
	if ((rc = stage_in(h, states, ns, &d_states, nt))) return rc;
	if ((rc = stage_in(h, diffs, ns, &d_diffs, nt+ns))) return rc;
	int pm = per_member ? 1 : 0;
	int na = n_anchors;
	void * args[] = { &h->P, &na, &d_times, &d_states, &d_diffs, &pm };
	if ((rc = launch(h, h->k.set_past, args))) return rc;
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	h->have_past = true;
	return JDDE_OK;
}

int jdde_set_parameters(jdde_handle * h, const double * params, int32_t per_member)
{
	int rc = ready(h);
	if (rc) return rc;
	int const np = h->desc.n_params;
	if (np == 0)
		return JDDE_OK;
	if (!params)
		return fail(h, JDDE_E_INVALID, "null argument");
	size_t const N = (size_t)h->desc.n_members;
	/* transpose to [n_params][n_members] on the host: coalesced loads on the device */
	std::vector<double> soa((size_t)np*N);
	for (size_t m = 0; m < N; m++)
		for (int k = 0; k < np; k++)
			soa[(size_t)k*N+m] = per_member ? params[m*np+k] : params[k];
	JD_CUDA(h, cudaMemcpyAsync(h->P.par, soa.data(), soa.size()*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	h->have_params = true;
	return JDDE_OK;
}

int jdde_set_integration_parameters(jdde_handle * h, const jdde_int_params * p, double max_delay)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!h->lib) return fail(h, JDDE_E_INVALID, "no kernel image loaded (jdde_load_rhs)");
	if (!p)
		return fail(h, JDDE_E_INVALID, "null argument");
	/* the assertions of the reference, _jitcdde.py:695-708 */
	if (!(p->decrease_threshold >= 1.0) || !(p->increase_threshold <= 1.0) || !(p->max_factor >= 1.0) || !(p->min_factor <= 1.0)
		|| !(p->safety_factor <= 1.0) || !(p->atol >= 0.0) || !(p->rtol >= 0.0) || !(p->pws_atol >= 0.0) || !(p->pws_rtol >= 0.0)
		|| !(p->pws_max_iterations > 0) || !(p->pws_factor >= 2) || !(p->pws_base_increase_chance >= 0) || !(max_delay >= 0.0))
		return fail(h, JDDE_E_INVALID, "integration parameter out of range (see _jitcdde.py:695-708)");
	h->P.P = *p;
	/* clipping of first_step / min_step, _jitcdde.py:688-693 */
	if (h->P.P.first_step > h->P.P.max_step)
		h->P.P.first_step = h->P.P.max_step;
	if (h->P.P.min_step > h->P.P.first_step)
		h->P.P.min_step = h->P.P.first_step;
	h->P.max_delay = max_delay;
	void * args[] = { &h->P };
	if ((rc = launch(h, h->k.reset_controller, args))) return rc;
	h->have_int_params = true;
	return JDDE_OK;
}

/* device buffer for the results of a sampling call: [n_samples][N][n] */
static int samples_buffer(jdde_handle * h, int slot, size_t count, double ** out)
{
	if (h->samples_count[slot] < count)
	{
		if (h->d_samples[slot])
		{
			JD_CUDA(h, cudaStreamSynchronize(h->stream));
			if (h->copy_stream)
				JD_CUDA(h, cudaStreamSynchronize(h->copy_stream));
			JD_CUDA(h, cudaFree(h->d_samples[slot]));
			h->d_samples[slot] = nullptr;
			h->samples_count[slot] = 0;
		}
		cudaError_t e = cudaMalloc((void **)&h->d_samples[slot], count*sizeof(double));
		if (e != cudaSuccess)
			return fail(h, e == cudaErrorMemoryAllocation ? JDDE_E_NOMEM : JDDE_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
		h->samples_count[slot] = count;
	}
	*out = h->d_samples[slot];
	return JDDE_OK;
}

/* n_steps == 0: integrate; > 0: step_on_discontinuities; < 0: −n_steps samples (results [samples][N][n_out]) */
static int run_integrate(jdde_handle * h, const double * targets, size_t count, int per_member, int n_steps, int resume, double * out_state, int n_out, int32_t * status, bool pipelined = false)
{
	int rc;
	double * d_targets;
	if (!targets)
		return fail(h, JDDE_E_INVALID, "null argument");
	if (n_steps < 0 && h->group_lanes)
		return fail(h, JDDE_E_INVALID, "sampling in one launch is not available for images with a group of lanes per member");
	size_t const n_results = n_steps < 0 ? (size_t)-n_steps : 1;
	size_t const result_count = n_results*(size_t)h->desc.n_members*(size_t)n_out;
	/* A single sample per launch is written by the kernel straight into page-locked host memory. Several samples per launch
	 * are not: lanes reach their samples at different times, so the 8-byte results would cross PCIe one by one; they go to
	 * a device buffer and the copy engine brings them over while the next launch runs. */
	double * const direct = n_steps == 0 ? direct_alias(out_state) : nullptr;
	if (pipelined)
	{
		if (!out_state)
			return fail(h, JDDE_E_INVALID, "null argument");
		if (!h->kernel_done[0])
			for (int s = 0; s < 2; s++)
			{
				JD_CUDA(h, cudaEventCreateWithFlags(&h->kernel_done[s], cudaEventDisableTiming));
				JD_CUDA(h, cudaEventCreateWithFlags(&h->copy_done[s], cudaEventDisableTiming));
			}
		if (!direct && !h->copy_stream)
		{
			JD_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
			if ((rc = dev_alloc(h, &h->d_out_alt, (size_t)h->desc.n_members*h->desc.n))) return rc;
		}
		int const slot = h->out_slot ^= 1;
		if (!direct && h->copy_pending[slot])       /* the copy that still reads this buffer (two calls ago) */
			JD_CUDA(h, cudaStreamWaitEvent(h->stream, h->copy_done[slot], 0));
		double * d_result = direct ? direct : slot ? h->d_out_alt : h->d_out;
		if (!direct && n_results > 1 && (rc = samples_buffer(h, slot, result_count, &d_result))) return rc;
		h->last_targets.assign(targets, targets+count);
		h->last_per_member = per_member;
		h->last_n_steps = n_steps;
		if ((rc = stage_in(h, targets, count, &d_targets))) return rc;
		int pm = per_member ? 1 : 0;
		long long budget = h->budget;
		void * args[] = { &h->P, &d_targets, &pm, &n_steps, &resume, &d_result, &n_out, &budget };
		if ((rc = launch_integrate(h, args))) return rc;
		JD_CUDA(h, cudaEventRecord(h->kernel_done[slot], h->stream));
		h->result_direct[slot] = direct != nullptr;
		h->result_queued[slot] = true;
		if (direct)
			return JDDE_OK;
		JD_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->kernel_done[slot], 0));
		JD_CUDA(h, cudaMemcpyAsync(out_state, d_result, result_count*sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
		JD_CUDA(h, cudaEventRecord(h->copy_done[slot], h->copy_stream));
		h->copy_pending[slot] = true;
		return JDDE_OK;
	}
	if (!resume)
	{
		h->last_targets.assign(targets, targets+count);
		h->last_per_member = per_member;
		h->last_n_steps = n_steps;
	}
	if ((rc = stage_in(h, targets, count, &d_targets))) return rc;
	int pm = per_member ? 1 : 0;
	long long budget = h->budget;
	double * d_result = direct ? direct : h->d_out;
	if (!direct && n_results > 1 && (rc = samples_buffer(h, 0, result_count, &d_result))) return rc;
	void * args[] = { &h->P, &d_targets, &pm, &n_steps, &resume, &d_result, &n_out, &budget };
	if ((rc = launch_integrate(h, args))) return rc;
	if (direct)
	{
		/* the states are already on their way into out_state; they have arrived when the kernel has ended */
		if ((rc = finish(h, nullptr, nullptr, 0, status))) return rc;
		if (!status)
			JD_CUDA(h, cudaStreamSynchronize(h->stream));
		return JDDE_OK;
	}
	return finish(h, out_state, d_result, result_count, status);
}

int jdde_integrate(jdde_handle * h, const double * target, int32_t per_member, double * out_state, int32_t * status)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	h->last_kind = 1;
	return run_integrate(h, target, per_member ? (size_t)h->desc.n_members : 1, per_member, 0, 0, out_state, h->desc.n, status);
}

int jdde_integrate_t(jdde_handle * h, const double * target, int32_t per_member, double * out_state, int32_t * status, double * t_after)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (!t_after)
		return fail(h, JDDE_E_INVALID, "null argument");
	h->last_kind = 1;
	/* the kernel, the times of the members (jdde_k_get) and the copies are queued; one synchronisation at the end */
	if ((rc = run_integrate(h, target, per_member ? (size_t)h->desc.n_members : 1, per_member, 0, 0, nullptr, h->desc.n, nullptr))) return rc;
	int what = 0;
	void * gargs[] = { &h->P, &what, &h->d_aux };
	if ((rc = launch(h, h->k.get, gargs))) return rc;
	JD_CUDA(h, cudaMemcpyAsync(t_after, h->d_aux, (size_t)h->desc.n_members*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	return finish(h, out_state, h->d_out, (size_t)h->desc.n_members*h->desc.n, status);
}

int jdde_integrate_async(jdde_handle * h, const double * target, int32_t per_member, double * out_state)
{
	if (!h)
		return JDDE_E_INVALID;
	h->queueing = true;
	int rc = ready_to_integrate(h);
	h->queueing = false;
	if (rc) return rc;
	h->last_kind = 0;      /* nothing to resume */
	return run_integrate(h, target, per_member ? (size_t)h->desc.n_members : 1, per_member, 0, 0, out_state, h->desc.n, nullptr, true);
}

int jdde_wait_result(jdde_handle * h, int32_t age)
{
	if (!h)
		return JDDE_E_INVALID;
	if (age < 0 || age > 1)
		return fail(h, JDDE_E_INVALID, "age must be 0 or 1 (two results can be in flight)");
	int const slot = h->out_slot ^ age;
	if (!h->result_queued[slot])
		return JDDE_OK;
	if (h->result_direct[slot])
		JD_CUDA(h, cudaEventSynchronize(h->kernel_done[slot]));
	else if (h->copy_pending[slot])
		JD_CUDA(h, cudaEventSynchronize(h->copy_done[slot]));
	return JDDE_OK;
}

int jdde_integrate_samples(jdde_handle * h, const double * targets, int32_t n_samples, int32_t per_member, double * out_states, int32_t * status)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (n_samples < 1)
		return fail(h, JDDE_E_INVALID, "n_samples must be positive");
	h->last_kind = 4;
	return run_integrate(h, targets, (per_member ? (size_t)h->desc.n_members : 1)*(size_t)n_samples, per_member, -n_samples, 0, out_states, h->desc.n, status);
}

int jdde_integrate_samples_async(jdde_handle * h, const double * targets, int32_t n_samples, int32_t per_member, double * out_states)
{
	if (!h)
		return JDDE_E_INVALID;
	h->queueing = true;
	int rc = ready_to_integrate(h);
	h->queueing = false;
	if (rc) return rc;
	if (n_samples < 1)
		return fail(h, JDDE_E_INVALID, "n_samples must be positive");
	h->last_kind = 0;      /* nothing to resume */
	return run_integrate(h, targets, (per_member ? (size_t)h->desc.n_members : 1)*(size_t)n_samples, per_member, -n_samples, 0, out_states, h->desc.n, nullptr, true);
}

int jdde_step_on_discontinuities(jdde_handle * h, int32_t n_steps, const double * steps, int32_t per_member, double * out_state, int32_t * status)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (n_steps < 1)
		return fail(h, JDDE_E_INVALID, "n_steps must be positive (for an ODE the reference calls adjust_diff instead, _jitcdde.py:998-1000)");
	h->last_kind = 2;
	return run_integrate(h, steps, (per_member ? (size_t)h->desc.n_members : 1)*(size_t)n_steps, per_member, n_steps, 0, out_state, h->desc.n, status);
}

static int run_blind(jdde_handle * h, const double * target, int per_member, double step, int n_out, bool lyap, double * out_state, double * lyaps, double * total_time, int32_t * status)
{
	int rc;
	if (!target)
		return fail(h, JDDE_E_INVALID, "null argument");
	size_t const N = (size_t)h->desc.n_members;
	double * d_targets;
	if ((rc = stage_in(h, target, per_member ? N : 1, &d_targets))) return rc;
	int pm = per_member ? 1 : 0;
	double * d_lyaps = lyap ? h->d_out2 : nullptr;
	double * d_total = lyap ? h->d_aux2 : nullptr;
	void * args[] = { &h->P, &d_targets, &pm, &step, &h->d_out, &n_out, &d_lyaps, &d_total };
	if ((rc = launch(h, h->k.integrate_blindly, args))) return rc;
	if (lyap && lyaps)
		JD_CUDA(h, cudaMemcpyAsync(lyaps, h->d_out2, N*(size_t)(h->desc.n/h->desc.n_basic-1)*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (lyap && total_time)
		JD_CUDA(h, cudaMemcpyAsync(total_time, h->d_aux2, N*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	return finish(h, out_state, h->d_out, N*n_out, status);
}

int jdde_integrate_blindly(jdde_handle * h, const double * target, int32_t per_member, double step, double * out_state, int32_t * status)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	return run_blind(h, target, per_member, step, h->desc.n, false, out_state, nullptr, nullptr, status);
}

int jdde_integrate_blindly_lyap(jdde_handle * h, const double * target, int32_t per_member, double step, double * out_state, double * lyaps, double * total_time, int32_t * status)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (h->desc.n == h->desc.n_basic)
		return fail(h, JDDE_E_INVALID, "integrate_blindly_lyap needs n_basic < n");
	return run_blind(h, target, per_member, step, h->desc.n_basic, true, out_state, lyaps, total_time, status);
}

int jdde_set_max_delay(jdde_handle * h, double max_delay)
{
	if (!h || !(max_delay >= 0.0))
		return fail(h, JDDE_E_INVALID, "max_delay must be non-negative");
	h->P.max_delay = max_delay;
	return JDDE_OK;
}

static int run_jump(jdde_handle * h, int mode, const double * amplitude, int amp_pm, const double * time, int time_pm, double width, int forward, double * minima, double * maxima)
{
	int rc;
	size_t const N = (size_t)h->desc.n_members;
	int const n = h->desc.n;
	double * d_amp = nullptr, * d_time = nullptr;
	if (mode == 0)
	{
		if (!amplitude || !time)
			return fail(h, JDDE_E_INVALID, "null argument");
		size_t const na = (amp_pm ? N : 1)*(size_t)n, nt = time_pm ? N : 1;
		std::vector<double> buf(na+nt);
		memcpy(buf.data(), amplitude, na*sizeof(double));
		memcpy(buf.data()+na, time, nt*sizeof(double));
		if ((rc = stage_in(h, buf.data(), na+nt, &d_amp))) return rc;
		JD_CUDA(h, cudaStreamSynchronize(h->stream));   /* buf is a temporary */
		d_time = d_amp+na;
	}
	void * args[] = { &h->P, &mode, &d_amp, &amp_pm, &d_time, &time_pm, &width, &forward, &h->d_out, &h->d_out2 };
	if ((rc = launch(h, h->k.jump, args))) return rc;
	if (minima)
		JD_CUDA(h, cudaMemcpyAsync(minima, h->d_out, N*n*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (maxima)
		JD_CUDA(h, cudaMemcpyAsync(maxima, h->d_out2, N*n*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (minima || maxima)
		JD_CUDA(h, cudaStreamSynchronize(h->stream));
	return JDDE_OK;
}

int jdde_jump(jdde_handle * h, const double * amplitude, int32_t amplitude_per_member, const double * time, int32_t time_per_member, double width, int32_t forward, double * minima, double * maxima)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (!(width >= 0))
		return fail(h, JDDE_E_INVALID, "width must be non-negative (_jitcdde.py:1035)");
	return run_jump(h, 0, amplitude, amplitude_per_member ? 1 : 0, time, time_per_member ? 1 : 0, width, forward ? 1 : 0, minima, maxima);
}

int jdde_adjust_diff(jdde_handle * h, double shift_ratio, double * minima, double * maxima)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	return run_jump(h, 1, nullptr, 0, nullptr, 0, shift_ratio, 0, minima, maxima);
}

int jdde_orthonormalise(jdde_handle * h, int32_t n_lyap, double delay, double * norms)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (h->desc.n == h->desc.n_basic || n_lyap < 0 || (n_lyap+1)*h->desc.n_basic > h->desc.n)
		return fail(h, JDDE_E_INVALID, "orthonormalise needs n_basic < n and (n_lyap+1)·n_basic <= n");
	if ((rc = launch_orthonormalise(h, n_lyap, delay, h->d_out2, nullptr, nullptr, nullptr))) return rc;
	return finish(h, norms, h->d_out2, (size_t)h->desc.n_members*n_lyap, nullptr);
}

static int run_lyap(jdde_handle * h, const double * target, size_t count, int per_member, int resume, double * out_state, double * lyaps, double * delta_t, int32_t * status)
{
	int rc;
	int const nb = h->desc.n_basic;
	int n_lyap = h->desc.n/nb-1;
	size_t const N = (size_t)h->desc.n_members;
	if (!resume)
	{
		/* old_t = get_t() (_jitcdde.py:1162); a resumed call keeps the one of the original call */
		int what = 0;
		void * gargs[] = { &h->P, &what, &h->d_aux };
		if ((rc = launch(h, h->k.get, gargs))) return rc;
	}
	if ((rc = run_integrate(h, target, count, per_member, 0, resume, nullptr, nb, nullptr))) return rc;
	if ((rc = launch_orthonormalise(h, n_lyap, h->P.max_delay, h->d_out2, h->d_aux, h->d_out2, h->d_aux2))) return rc;
	if (lyaps)
		JD_CUDA(h, cudaMemcpyAsync(lyaps, h->d_out2, N*n_lyap*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (delta_t)
		JD_CUDA(h, cudaMemcpyAsync(delta_t, h->d_aux2, N*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	return finish(h, out_state, h->d_out, N*nb, status);
}

int jdde_integrate_lyap(jdde_handle * h, const double * target, int32_t per_member, double * out_state, double * lyaps, double * delta_t, int32_t * status)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (h->desc.n == h->desc.n_basic)
		return fail(h, JDDE_E_INVALID, "integrate_lyap needs n_basic < n");
	h->last_kind = 3;
	return run_lyap(h, target, per_member ? (size_t)h->desc.n_members : 1, per_member, 0, out_state, lyaps, delta_t, status);
}

int jdde_set_projector(jdde_handle * h, int32_t mode, const double * vectors, int32_t n_vectors,
	const int32_t * state_components, int32_t n_state, const int32_t * diff_components, int32_t n_diff, const int32_t * indices, int32_t n_indices)
{
	int rc = ready(h);
	if (rc) return rc;
	int const n = h->desc.n, nb = h->desc.n_basic;
	if (mode == JDDE_PROJECT_RESTRICTED)
	{
		if (n_vectors < 0 || n_state < 0 || n_diff < 0 || n != nb*(2+2*n_vectors) || (n_vectors && !vectors) || (n_state && !state_components) || (n_diff && !diff_components))
			return fail(h, JDDE_E_INVALID, "restricted projector needs n = n_basic·(2 + 2·n_vectors) (_jitcdde.py:1270-1274)");
		for (int k = 0; k < n_state; k++) if (state_components[k] < 0 || state_components[k] >= nb) return fail(h, JDDE_E_INVALID, "state component out of range");
		for (int k = 0; k < n_diff; k++) if (diff_components[k] < 0 || diff_components[k] >= nb) return fail(h, JDDE_E_INVALID, "diff component out of range");
		n_indices = 0;
	}
	else if (mode == JDDE_PROJECT_TRANSVERSAL)
	{
		if (n_indices < 1 || !indices)
			return fail(h, JDDE_E_INVALID, "transversal projector needs tangent indices");
		for (int k = 0; k < n_indices; k++) if (indices[k] < 0 || indices[k] >= n) return fail(h, JDDE_E_INVALID, "tangent index out of range");
		n_vectors = n_state = n_diff = 0;
	}
	else
		return fail(h, JDDE_E_INVALID, "unknown projector mode");
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	if (h->projector_memory)
	{
		cudaFree(h->projector_memory);
		h->projector_memory = nullptr;
	}
	size_t const vector_bytes = (size_t)n_vectors*2*nb*sizeof(double);
	size_t const int_count = (size_t)n_state+n_diff+n_indices;
	std::vector<char> host(vector_bytes + int_count*sizeof(int32_t) + 16);
	if (vector_bytes) memcpy(host.data(), vectors, vector_bytes);
	int32_t * ints = (int32_t *)(host.data()+vector_bytes);
	for (int k = 0; k < n_state; k++) ints[k] = state_components[k];
	for (int k = 0; k < n_diff; k++) ints[n_state+k] = diff_components[k];
	for (int k = 0; k < n_indices; k++) ints[n_state+n_diff+k] = indices[k];
	JD_CUDA(h, cudaMalloc(&h->projector_memory, host.size()));
	JD_CUDA(h, cudaMemcpy(h->projector_memory, host.data(), host.size(), cudaMemcpyHostToDevice));
	jdde_projector & R = h->projector;
	R.mode = mode;
	R.n_vectors = n_vectors;
	R.vectors = (const double *)h->projector_memory;
	const int * const device_ints = (const int *)((const char *)h->projector_memory+vector_bytes);
	R.n_state_components = n_state;
	R.state_components = device_ints;
	R.n_diff_components = n_diff;
	R.diff_components = device_ints+n_state;
	R.n_indices = n_indices;
	R.indices = device_ints+n_state+n_diff;
	return JDDE_OK;
}

int jdde_project(jdde_handle * h, double delay, double * norms)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (!h->projector.mode) return fail(h, JDDE_E_INVALID, "no projector set (jdde_set_projector)");
	if (!h->k.project) return fail(h, JDDE_E_IMAGE, "image was built without JD_PROJECT");
	void * args[] = { &h->P, &h->projector, &delay, &h->d_aux2 };
	if ((rc = launch(h, h->k.project, args))) return rc;
	return finish(h, norms, h->d_aux2, (size_t)h->desc.n_members, nullptr);
}

int jdde_integrate_blindly_projected(jdde_handle * h, const double * target, int32_t per_member, double step, double * out_state, double * lyap, double * total_time, int32_t * status)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	if (!h->projector.mode) return fail(h, JDDE_E_INVALID, "no projector set (jdde_set_projector)");
	if (!h->k.integrate_blindly_projected) return fail(h, JDDE_E_IMAGE, "image was built without JD_PROJECT");
	if (!target) return fail(h, JDDE_E_INVALID, "null argument");
	size_t const N = (size_t)h->desc.n_members;
	double * d_targets;
	if ((rc = stage_in(h, target, per_member ? N : 1, &d_targets))) return rc;
	int pm = per_member ? 1 : 0;
	int n_out = h->desc.n;
	void * args[] = { &h->P, &h->projector, &d_targets, &pm, &step, &h->d_out, &n_out, &h->d_aux, &h->d_aux2 };
	if ((rc = launch(h, h->k.integrate_blindly_projected, args))) return rc;
	if (lyap)
		JD_CUDA(h, cudaMemcpyAsync(lyap, h->d_aux, N*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	if (total_time)
		JD_CUDA(h, cudaMemcpyAsync(total_time, h->d_aux2, N*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	return finish(h, out_state, h->d_out, N*n_out, status);
}

int jdde_resume(jdde_handle * h, double * out_state, double * lyaps, double * delta_t, int32_t * status)
{
	int rc = ready_to_integrate(h);
	if (rc) return rc;
	std::vector<double> const targets = h->last_targets;
	switch (h->last_kind)
	{
		case 1:
		case 2:
		case 4:      /* sampling: out_state is [samples][N][n]; samples written before the stop are written again only for members that had not reached them */
			return run_integrate(h, targets.data(), targets.size(), h->last_per_member, h->last_n_steps, 1, out_state, h->desc.n, status);
		case 3:
			return run_lyap(h, targets.data(), targets.size(), h->last_per_member, 1, out_state, lyaps, delta_t, status);
		default:
			return fail(h, JDDE_E_INVALID, "nothing to resume");
	}
}

static int run_get(jdde_handle * h, int what, double * out, size_t per_member_count)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!h->have_past) return fail(h, JDDE_E_INVALID, "no past set (jdde_set_past)");
	if (!out) return fail(h, JDDE_E_INVALID, "null argument");
	void * args[] = { &h->P, &what, &h->d_out2 };
	if ((rc = launch(h, h->k.get, args))) return rc;
	return finish(h, out, h->d_out2, (size_t)h->desc.n_members*per_member_count, nullptr);
}

int jdde_get_t(jdde_handle * h, double * t) { return run_get(h, 0, t, 1); }
int jdde_get_current_state(jdde_handle * h, double * state) { return run_get(h, 1, state, h ? h->desc.n : 0); }
int jdde_get_dt(jdde_handle * h, double * dt) { return run_get(h, 2, dt, 1); }

int jdde_get_status(jdde_handle * h, int32_t * status)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!status) return fail(h, JDDE_E_INVALID, "null argument");
	return finish(h, nullptr, nullptr, 0, status);
}

int jdde_count_failed(jdde_handle * h, int64_t * count)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!count) return fail(h, JDDE_E_INVALID, "null argument");
	JD_CUDA(h, cudaMemsetAsync(h->d_sums, 0, sizeof(unsigned long long), h->stream));
	void * args[] = { &h->P, &h->d_sums };
	if ((rc = launch(h, h->k.count_failed, args))) return rc;
	unsigned long long n = 0;
	JD_CUDA(h, cudaMemcpyAsync(&n, h->d_sums, sizeof n, cudaMemcpyDeviceToHost, h->stream));
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	*count = (int64_t)n;
	return JDDE_OK;
}

int jdde_get_counters(jdde_handle * h, int64_t * accepted, int64_t * rejected, int64_t * rhs_evals)
{
	int rc = ready(h);
	if (rc) return rc;
	size_t const N = (size_t)h->desc.n_members;
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	if (accepted)
		JD_CUDA(h, cudaMemcpy(accepted, h->P.accepted, N*sizeof(int64_t), cudaMemcpyDeviceToHost));
	if (rejected)
		JD_CUDA(h, cudaMemcpy(rejected, h->P.rejected, N*sizeof(int64_t), cudaMemcpyDeviceToHost));
	if (rhs_evals)
	{
		std::vector<long long> attempts(N), extra(N);
		JD_CUDA(h, cudaMemcpy(attempts.data(), h->P.attempts, N*sizeof(long long), cudaMemcpyDeviceToHost));
		JD_CUDA(h, cudaMemcpy(extra.data(), h->P.extra_rhs, N*sizeof(long long), cudaMemcpyDeviceToHost));
		for (size_t m = 0; m < N; m++)
			rhs_evals[m] = 3*attempts[m]+extra[m];
	}
	return JDDE_OK;
}

int jdde_sum_counters(jdde_handle * h, int64_t * accepted, int64_t * rejected, int64_t * rhs_evals)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!h->have_past) return fail(h, JDDE_E_INVALID, "no past set (jdde_set_past)");
	JD_CUDA(h, cudaMemsetAsync(h->d_sums, 0, 3*sizeof(unsigned long long), h->stream));
	void * args[] = { &h->P, &h->d_sums };
	if ((rc = launch(h, h->k.sum_counters, args))) return rc;
	unsigned long long sums[3];
	JD_CUDA(h, cudaMemcpyAsync(sums, h->d_sums, sizeof sums, cudaMemcpyDeviceToHost, h->stream));
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	if (accepted) *accepted = (int64_t)sums[0];
	if (rejected) *rejected = (int64_t)sums[1];
	if (rhs_evals) *rhs_evals = (int64_t)sums[2];
	return JDDE_OK;
}

int jdde_get_state(jdde_handle * h, int64_t member, int32_t max_anchors, int32_t * n_anchors, double * times, double * states, double * diffs)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!h->have_past) return fail(h, JDDE_E_INVALID, "no past set (jdde_set_past)");
	if (member < 0 || member >= h->desc.n_members || !n_anchors)
		return fail(h, JDDE_E_INVALID, "member out of range or null argument");
	/* get_state() forgets first (_jitcdde.py:368) */
	if (h->have_int_params)
	{
		void * args[] = { &h->P };
		if ((rc = launch(h, h->k.forget, args))) return rc;
	}
	int first, last;
	JD_CUDA(h, cudaMemcpyAsync(&first, h->P.first+member, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	JD_CUDA(h, cudaMemcpyAsync(&last, h->P.last+member, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	int const count = last-first+1;
	*n_anchors = count;
	if (count > max_anchors || !times || !states || !diffs)
		return JDDE_OK;
	int const A = h->P.anchor_stride, cap = h->P.cap, n = h->desc.n;
	std::vector<double> ring((size_t)cap*A);
	JD_CUDA(h, cudaMemcpy(ring.data(), h->P.ring+(size_t)member*h->P.ring_stride, ring.size()*sizeof(double), cudaMemcpyDeviceToHost));
	for (int k = 0; k < count; k++)
	{
		const double * a = ring.data()+(size_t)((first+k) & (cap-1))*A;
		times[k] = a[0];
		memcpy(states+(size_t)k*n, a+1, n*sizeof(double));
		memcpy(diffs+(size_t)k*n, a+1+n, n*sizeof(double));
	}
	return JDDE_OK;
}

int jdde_max_live_anchors(jdde_handle * h, int32_t * count)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!h->have_past) return fail(h, JDDE_E_INVALID, "no past set (jdde_set_past)");
	if (!count) return fail(h, JDDE_E_INVALID, "null argument");
	size_t const N = (size_t)h->desc.n_members;
	std::vector<int> first(N), last(N);
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	JD_CUDA(h, cudaMemcpy(first.data(), h->P.first, N*sizeof(int), cudaMemcpyDeviceToHost));
	JD_CUDA(h, cudaMemcpy(last.data(), h->P.last, N*sizeof(int), cudaMemcpyDeviceToHost));
	int most = 0;
	for (size_t m = 0; m < N; m++)
		if (last[m]-first[m]+1 > most)
			most = last[m]-first[m]+1;
	*count = most;
	return JDDE_OK;
}

int jdde_grow_ring(jdde_handle * h, int32_t new_capacity)
{
	int rc = ready(h);
	if (rc) return rc;
	if (!h->lib) return fail(h, JDDE_E_INVALID, "no kernel image loaded (jdde_load_rhs)");
	if (new_capacity <= h->P.cap || (new_capacity & (new_capacity-1)))
		return fail(h, JDDE_E_INVALID, "new capacity must be a larger power of two");
	size_t const N = (size_t)h->desc.n_members;
	double * new_ring = nullptr;
	cudaError_t e = cudaMalloc((void **)&new_ring, N*(size_t)jdde_ring_stride(new_capacity, h->P.anchor_stride)*sizeof(double));
	if (e != cudaSuccess)
		return fail(h, JDDE_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
	if (h->have_past)
	{
		int nc = new_capacity;
		void * args[] = { &h->P, &new_ring, &nc };
		if ((rc = launch(h, h->k.regrow, args))) { cudaFree(new_ring); return rc; }
	}
	JD_CUDA(h, cudaStreamSynchronize(h->stream));
	for (auto & p : h->allocations)
		if (p == h->P.ring)
			p = new_ring;
	cudaFree(h->P.ring);
	h->P.ring = new_ring;
	h->P.cap = new_capacity;
	h->P.ring_stride = jdde_ring_stride(new_capacity, h->P.anchor_stride);
	h->desc.ring_capacity = new_capacity;
	return JDDE_OK;
}

int jdde_clear_status(jdde_handle * h, int32_t code)
{
	int rc = ready(h);
	if (rc) return rc;
	int c = code;
	void * args[] = { &h->P, &c };
	return launch(h, h->k.clear_status, args);
}

int jdde_set_attempt_budget(jdde_handle * h, int64_t attempts)
{
	if (!h || attempts < 1)
		return fail(h, JDDE_E_INVALID, "budget must be positive");
	h->budget = attempts > 0x7fffffff ? 0x7fffffff : attempts;      /* the kernels count attempts per call in 32 bits */
	return JDDE_OK;
}

int jdde_host_alloc(size_t bytes, void ** out)
{
	if (!out)
		return JDDE_E_INVALID;
	*out = nullptr;
	cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
	if (e != cudaSuccess)
		return fail(nullptr, JDDE_E_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
	return JDDE_OK;
}

int jdde_host_free(void * p)
{
	return cudaFreeHost(p) == cudaSuccess ? JDDE_OK : JDDE_E_CUDA;
}

int jdde_measure_fp64_peak(int32_t device, const void * image, size_t len, double * flops)
{
	if (!image || !len || !flops)
		return fail(nullptr, JDDE_E_INVALID, "null argument");
	JD_CUDA(nullptr, cudaSetDevice(device));
	cudaLibrary_t lib;
	JD_CUDA(nullptr, cudaLibraryLoadData(&lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0));
	cudaKernel_t kernel;
	JD_CUDA(nullptr, cudaLibraryGetKernel(&kernel, lib, "jdde_k_fp64_peak"));
	cudaDeviceProp prop;
	JD_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
	int const block = 256, grid = prop.multiProcessorCount*16, iterations = 4096;
	double * out = nullptr;
	JD_CUDA(nullptr, cudaMalloc((void **)&out, (size_t)grid*block*sizeof(double)));
	cudaEvent_t e0, e1;
	JD_CUDA(nullptr, cudaEventCreate(&e0));
	JD_CUDA(nullptr, cudaEventCreate(&e1));
	int it = iterations; double a = 1.0000001, b = 1e-9;
	void * args[] = { &out, &it, &a, &b };
	float best = 1e30f;
	for (int r = 0; r < 6; r++)
	{
		JD_CUDA(nullptr, cudaEventRecord(e0, 0));
		JD_CUDA(nullptr, cudaLaunchKernel((const void *)kernel, dim3(grid), dim3(block), args, 0, 0));
		JD_CUDA(nullptr, cudaEventRecord(e1, 0));
		JD_CUDA(nullptr, cudaEventSynchronize(e1));
		float ms = 0;
		JD_CUDA(nullptr, cudaEventElapsedTime(&ms, e0, e1));
		if (r > 0 && ms < best)
			best = ms;
	}
	*flops = 2.0*8.0*iterations*(double)grid*block/(best*1e-3);
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out); cudaLibraryUnload(lib);
	return JDDE_OK;
}

int64_t jdde_launch_count(const jdde_handle * h) { return h ? h->launches : 0; }
uint64_t jdde_device_result(const jdde_handle * h) { return h ? (uint64_t)(uintptr_t)h->d_out : 0; }

} // extern "C"
