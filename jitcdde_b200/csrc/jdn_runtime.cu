/*
 * jdn_runtime.cu — host side of the network engine (jdde_net_* in include/jitcdde_b200.h); part of
 * libjitcdde_b200.so. Owns the device memory of one system (layout: jdn_abi.h), loads the kernel image and runs the
 * integration calls: by default as cooperative launches of the persistent stepper jdn_k_run (hundreds of attempted steps
 * per launch, exchange between the GPUs included); the per-attempt kernels  lookup → attempt → [exchange] → control →
 * commit  remain for callers that exchange the anchors themselves (NCCL) and for comparison (JITCDDE_B200_NET_PERSISTENT=0).
 * The controller runs on the device, the host polls a flag per launch. No numerical code here.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <new>

#include "jdn_abi.h"

#define JDN_MAX_RING_CAPACITY 16384

struct jdde_net
{
	jdde_net_desc desc{};
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	cudaLibrary_t lib = nullptr;
	cudaKernel_t k_lookup = nullptr, k_attempt = nullptr, k_control = nullptr, k_commit = nullptr, k_begin = nullptr, k_recent = nullptr, k_set_past = nullptr, k_exchange = nullptr, k_regrow = nullptr, k_run = nullptr, k_run_warp = nullptr, k_adjust = nullptr, k_adjust_commit = nullptr, k_adjust_control = nullptr;
	/* peer exchange: this rank's buffer (mapped by the peers through CUDA IPC), the peers' buffers mapped here */
	jdn_exchange * exchange = nullptr;
	std::vector<void *> mapped_peers;
	jdn_exchange ** d_peers = nullptr;
	int peer_rank = 0, peer_world = 1;
	bool persistent = true;                   /* jdn_k_run (one cooperative launch per batch of attempts) instead of five kernels per attempt */
	int run_grid = 0;                         /* co-resident blocks of jdn_k_run */
	double min_tau = 0.0, max_step = 1e300;
	double adjust_time = 0.0, adjust_new_t = 0.0;      /* between jdde_net_adjust_diff_begin and _end */
	bool adjusting = false;   /* smallest delay of this rank's in-edges; max_step of the integration parameters */
	int run_smem_bytes = 0;                   /* its dynamic shared memory beyond the anchor times (image descriptor) */
	bool warp_tiles = false;                  /* jdn_k_run_warp (a warp per tile of <= 32 in-edges) instead of jdn_k_run (a block per tile of 64 nodes) */
	int * d_tile_start = nullptr;
	/* `poll_every` attempts (lookup, step, [exchange], control, commit: 4–5 launches each) captured once as a CUDA graph and
	 * replayed: small networks are bound by launch gaps, not by the kernels (10^5 nodes: 45 µs of kernels per 70 µs attempt) */
	cudaGraphExec_t batch = nullptr;
	int64_t launches_per_batch = 0;
	int block = 256;
	jdn_problem P{};
	double * scratch = nullptr, * d_out = nullptr, * d_par = nullptr;
	long long * d_row_ptr = nullptr; int * d_src = nullptr; double * d_tau = nullptr, * d_weight = nullptr;
	jdn_control host_ctrl{};
	bool have_graph = false, have_past = false, have_int = false;
	int64_t launches = 0;
	int poll_every = 8;
	std::string err;
	std::vector<void *> allocations;
};

namespace {
thread_local std::string g_net_error;

int nfail(jdde_net * h, int code, const std::string & msg) { (h ? h->err : g_net_error) = msg; return code; }

#define JN_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return nfail(h, JDDE_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

template <class T> int nalloc(jdde_net * h, T ** p, size_t count)
{
	void * q = nullptr;
	cudaError_t e = cudaMalloc(&q, count*sizeof(T) ? count*sizeof(T) : 1);
	if (e != cudaSuccess)
		return nfail(h, e == cudaErrorMemoryAllocation ? JDDE_E_NOMEM : JDDE_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
	h->allocations.push_back(q);
	*p = static_cast<T *>(q);
	return JDDE_OK;
}

int nready(jdde_net * h)
{
	if (!h) return JDDE_E_INVALID;
	if (cudaSetDevice(h->desc.device) != cudaSuccess) return nfail(h, JDDE_E_CUDA, "cudaSetDevice failed");
	return JDDE_OK;
}

int nlaunch(jdde_net * h, cudaKernel_t k, long long threads, int block, size_t smem, void ** args)
{
	if (!k) return nfail(h, JDDE_E_INVALID, "kernel image not loaded (jdde_net_load_image)");
	unsigned const grid = (unsigned)((threads+block-1)/block);
	JN_CUDA(h, cudaLaunchKernel((const void *)k, dim3(grid ? grid : 1), dim3(block), args, smem, h->stream));
	h->launches++;
	return JDDE_OK;
}

void drop_batch(jdde_net * h)
{
	if (h->batch)
	{
		cudaGraphExecDestroy(h->batch);
		h->batch = nullptr;
	}
}

int nready_run(jdde_net * h)
{
	int rc = nready(h);
	if (rc) return rc;
	if (!h->lib) return nfail(h, JDDE_E_INVALID, "no kernel image loaded (jdde_net_load_image)");
	if (!h->have_graph) return nfail(h, JDDE_E_INVALID, "no graph set (jdde_net_set_graph)");
	if (!h->have_past) return nfail(h, JDDE_E_INVALID, "no past set (jdde_net_set_past)");
	if (!h->have_int) return nfail(h, JDDE_E_INVALID, "integration parameters not set (jdde_net_set_integration_parameters)");
	return JDDE_OK;
}
} // namespace

extern "C" {

int jdde_net_create(const jdde_net_desc * d, jdde_net ** out)
{
	if (!d || !out) return nfail(nullptr, JDDE_E_INVALID, "null argument");
	*out = nullptr;
	if (d->n < 1 || d->lo < 0 || d->hi > d->n || d->lo >= d->hi || d->n_edges < 0 || d->n_node_params < 0)
		return nfail(nullptr, JDDE_E_INVALID, "invalid descriptor");
	if (d->ring_capacity < 4 || (d->ring_capacity & (d->ring_capacity-1)))
		return nfail(nullptr, JDDE_E_INVALID, "ring_capacity must be a power of two >= 4");
	if (d->ring_capacity > JDN_MAX_RING_CAPACITY)
		return nfail(nullptr, JDDE_E_INVALID, "ring_capacity above 16384 (the anchor times of the ring are staged in shared memory: 128 KB)");
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count < 1)
		return nfail(nullptr, JDDE_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this engine has no CPU fallback)");
	if (d->device < 0 || d->device >= count) return nfail(nullptr, JDDE_E_INVALID, "device ordinal out of range");
	jdde_net * h = new (std::nothrow) jdde_net;
	if (!h) return nfail(nullptr, JDDE_E_NOMEM, "out of host memory");
	h->desc = *d;
	auto bail = [&](int rc) { g_net_error = h->err; jdde_net_destroy(h); return rc; };
	if (cudaSetDevice(d->device) != cudaSuccess) return bail(nfail(h, JDDE_E_CUDA, "cudaSetDevice failed"));
	if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(nfail(h, JDDE_E_CUDA, "cudaStreamCreate failed"));
	h->own_stream = true;
	size_t const n = (size_t)d->n, cap = (size_t)d->ring_capacity, E = (size_t)d->n_edges, owned = (size_t)(d->hi-d->lo);
	jdn_problem & P = h->P;
	P.n = d->n; P.lo = d->lo; P.hi = d->hi; P.cap = d->ring_capacity; P.n_node_params = d->n_node_params;
	int rc;
	if ((rc = nalloc(h, &P.ctrl, 1))) return bail(rc);
	if ((rc = nalloc(h, &P.times, cap))) return bail(rc);
	if ((rc = nalloc(h, &P.hist, 2*cap*n))) return bail(rc);
	if ((rc = nalloc(h, &P.stage, 2*n))) return bail(rc);
	/* the two tentative buffers of the past-within-step iteration (jdn_abi.h); several GPUs use the staging buffers of the exchange */
	if ((rc = nalloc(h, &P.tentative[0], 2*n))) return bail(rc);
	if ((rc = nalloc(h, &P.tentative[1], 2*n))) return bail(rc);
	if ((rc = nalloc(h, &h->d_par, n*(size_t)(d->n_node_params ? d->n_node_params : 1)))) return bail(rc);
	if ((rc = nalloc(h, &h->d_row_ptr, owned+1))) return bail(rc);
	if ((rc = nalloc(h, &h->d_src, E))) return bail(rc);
	if ((rc = nalloc(h, &h->d_tau, E))) return bail(rc);
	if ((rc = nalloc(h, &h->d_weight, E))) return bail(rc);
	if ((rc = nalloc(h, &P.cursor, E))) return bail(rc);
	if ((rc = nalloc(h, &h->scratch, 3*E))) return bail(rc);
	if ((rc = nalloc(h, &h->d_out, 2*n))) return bail(rc);      /* states of all nodes; minima and maxima of adjust_diff */
	P.node_par = h->d_par; P.row_ptr = h->d_row_ptr; P.src = h->d_src; P.tau = h->d_tau; P.weight = h->d_weight;
	if (cudaMemsetAsync(P.ctrl, 0, sizeof(jdn_control), h->stream) != cudaSuccess) return bail(nfail(h, JDDE_E_CUDA, "cudaMemset failed"));
	*out = h;
	return JDDE_OK;
}

int jdde_net_destroy(jdde_net * h)
{
	if (!h) return JDDE_OK;
	cudaSetDevice(h->desc.device);
	if (h->stream) cudaStreamSynchronize(h->stream);
	if (h->batch) cudaGraphExecDestroy(h->batch);
	for (void * p : h->mapped_peers) cudaIpcCloseMemHandle(p);
	for (void * p : h->allocations) cudaFree(p);
	if (h->lib) cudaLibraryUnload(h->lib);
	if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
	delete h;
	return JDDE_OK;
}

const char * jdde_net_last_error(const jdde_net * h) { return h ? h->err.c_str() : g_net_error.c_str(); }

int jdde_net_load_image(jdde_net * h, const void * image, size_t len)
{
	int rc = nready(h);
	if (rc) return rc;
	if (!image || !len) return nfail(h, JDDE_E_INVALID, "empty image");
	drop_batch(h);
	if (h->lib) { cudaLibraryUnload(h->lib); h->lib = nullptr; }
	JN_CUDA(h, cudaLibraryLoadData(&h->lib, image, nullptr, nullptr, 0, nullptr, nullptr, 0));
	void * dptr = nullptr; size_t bytes = 0;
	if (cudaLibraryGetGlobal(&dptr, &bytes, h->lib, "jdn_image_desc") != cudaSuccess || bytes != sizeof(jdn_image_desc_t))
		return nfail(h, JDDE_E_IMAGE, "image has no jdn_image_desc");
	jdn_image_desc_t d;
	JN_CUDA(h, cudaMemcpy(&d, dptr, sizeof d, cudaMemcpyDeviceToHost));
	if (d.magic != JDN_IMAGE_MAGIC || d.abi != JDDE_ABI_VERSION) return nfail(h, JDDE_E_IMAGE, "image built for another ABI version");
	if (d.n_node_params != h->desc.n_node_params) return nfail(h, JDDE_E_IMAGE, "image does not match the descriptor (node parameters)");
	h->block = d.block;
	h->run_smem_bytes = d.run_smem_bytes;
	struct { const char * name; cudaKernel_t * slot; } const table[] = {
		{"jdn_k_lookup", &h->k_lookup}, {"jdn_k_attempt", &h->k_attempt}, {"jdn_k_control", &h->k_control}, {"jdn_k_commit", &h->k_commit},
		{"jdn_k_begin", &h->k_begin}, {"jdn_k_recent_state", &h->k_recent}, {"jdn_k_set_past", &h->k_set_past},
	};
	for (auto const & entry : table)
		if (cudaLibraryGetKernel(entry.slot, h->lib, entry.name) != cudaSuccess)
			return nfail(h, JDDE_E_IMAGE, std::string("image lacks kernel ") + entry.name);
	if (cudaLibraryGetKernel(&h->k_regrow, h->lib, "jdn_k_regrow") != cudaSuccess)
		return nfail(h, JDDE_E_IMAGE, "image lacks kernel jdn_k_regrow");
	if (cudaLibraryGetKernel(&h->k_exchange, h->lib, "jdn_k_exchange") != cudaSuccess)
	{
		h->k_exchange = nullptr;
		(void)cudaGetLastError();
	}
	if (cudaLibraryGetKernel(&h->k_run, h->lib, "jdn_k_run") != cudaSuccess)
	{
		h->k_run = nullptr;
		(void)cudaGetLastError();
	}
	if (cudaLibraryGetKernel(&h->k_run_warp, h->lib, "jdn_k_run_warp") != cudaSuccess)
	{
		h->k_run_warp = nullptr;
		(void)cudaGetLastError();
	}
	if (cudaLibraryGetKernel(&h->k_adjust, h->lib, "jdn_k_adjust_diff") != cudaSuccess
		|| cudaLibraryGetKernel(&h->k_adjust_commit, h->lib, "jdn_k_adjust_diff_commit") != cudaSuccess
		|| cudaLibraryGetKernel(&h->k_adjust_control, h->lib, "jdn_k_adjust_diff_control") != cudaSuccess)
		return nfail(h, JDDE_E_IMAGE, "image lacks the adjust_diff kernels");
	h->run_grid = 0;
	h->P.tile_nodes = 0;
	const char * const choice = getenv("JITCDDE_B200_NET_PERSISTENT");
	h->persistent = h->k_run && !(choice && choice[0] == '0');
	const char * const tiles = getenv("JITCDDE_B200_NET_TILES");
	h->warp_tiles = h->k_run_warp && tiles && tiles[0] == 'w';      /* "warp": the warp-per-tile variant (measured slower: its registers spill, profiles/configs_r02.md) */
	if (h->k_run_warp && (size_t)h->P.cap*sizeof(double) > 48*1024)
		JN_CUDA(h, cudaFuncSetAttribute((const void *)h->k_run_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)h->P.cap*sizeof(double))));
	/* the bracket search keeps the anchor times in shared memory: 8 bytes per ring slot */
	if ((size_t)h->P.cap*sizeof(double) > 48*1024)
		JN_CUDA(h, cudaFuncSetAttribute((const void *)h->k_lookup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)h->P.cap*sizeof(double))));
	if (h->k_run && (size_t)h->P.cap*sizeof(double)+h->run_smem_bytes > 48*1024)
		JN_CUDA(h, cudaFuncSetAttribute((const void *)h->k_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)h->P.cap*sizeof(double)+h->run_smem_bytes)));
	return JDDE_OK;
}

int jdde_net_set_stream(jdde_net * h, void * s)
{
	int rc = nready(h);
	if (rc) return rc;
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	drop_batch(h);
	if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
	h->stream = static_cast<cudaStream_t>(s);
	return JDDE_OK;
}

int jdde_net_set_graph(jdde_net * h, const int64_t * row_ptr, const int32_t * src, const double * tau, const double * weight)
{
	int rc = nready(h);
	if (rc) return rc;
	if (!row_ptr || (h->desc.n_edges && (!src || !tau || !weight))) return nfail(h, JDDE_E_INVALID, "null argument");
	size_t const owned = (size_t)(h->desc.hi-h->desc.lo), E = (size_t)h->desc.n_edges;
	if (row_ptr[0] != 0 || (size_t)row_ptr[owned] != E) return nfail(h, JDDE_E_INVALID, "row_ptr does not match n_edges");
	h->min_tau = E ? tau[0] : 1e300;
	for (size_t e = 0; e < E; e++)
		if (tau[e] < h->min_tau) h->min_tau = tau[e];
	for (size_t e = 0; e < E; e++)
		if (src[e] < 0 || src[e] >= h->desc.n || !(tau[e] > 0.0))
			return nfail(h, JDDE_E_INVALID, "edge with source out of range or non-positive delay");
	JN_CUDA(h, cudaMemcpyAsync(h->d_row_ptr, row_ptr, (owned+1)*sizeof(long long), cudaMemcpyHostToDevice, h->stream));
	JN_CUDA(h, cudaMemcpyAsync(h->d_src, src, E*sizeof(int), cudaMemcpyHostToDevice, h->stream));
	JN_CUDA(h, cudaMemcpyAsync(h->d_tau, tau, E*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	JN_CUDA(h, cudaMemcpyAsync(h->d_weight, weight, E*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	/* warp tiles of the persistent stepper: consecutive nodes while they are at most 32 and have at most 32 in-edges together;
	 * a node with more in-edges is a tile of its own */
	{
		std::vector<int> tiles;
		tiles.push_back(0);
		size_t row = 0;
		while (row < owned)
		{
			size_t end = row+1;
			int64_t edges = row_ptr[row+1]-row_ptr[row];
			while (end < owned && end-row < 32 && edges+(row_ptr[end+1]-row_ptr[end]) <= 32)
			{
				edges += row_ptr[end+1]-row_ptr[end];
				end++;
			}
			tiles.push_back((int)end);
			row = end;
		}
		if (h->d_tile_start)
		{
			for (auto it = h->allocations.begin(); it != h->allocations.end(); ++it)
				if (*it == h->d_tile_start) { h->allocations.erase(it); break; }
			cudaFree(h->d_tile_start);
			h->d_tile_start = nullptr;
		}
		int rc2 = nalloc(h, &h->d_tile_start, tiles.size());
		if (rc2) return rc2;
		JN_CUDA(h, cudaMemcpy(h->d_tile_start, tiles.data(), tiles.size()*sizeof(int), cudaMemcpyHostToDevice));
		h->P.tile_start = h->d_tile_start;
		h->P.n_warp_tiles = (long long)tiles.size()-1;
		h->run_grid = 0;
		h->P.tile_nodes = 0;
	}
	h->have_graph = true;
	return JDDE_OK;
}

int jdde_net_set_node_parameters(jdde_net * h, const double * par)
{
	int rc = nready(h);
	if (rc) return rc;
	if (!h->desc.n_node_params) return JDDE_OK;
	if (!par) return nfail(h, JDDE_E_INVALID, "null argument");
	JN_CUDA(h, cudaMemcpyAsync(h->d_par, par, (size_t)h->desc.n*h->desc.n_node_params*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	return JDDE_OK;
}

int jdde_net_set_past(jdde_net * h, int32_t n_anchors, const double * times, const double * states, const double * diffs)
{
	int rc = nready(h);
	if (rc) return rc;
	if (!h->lib) return nfail(h, JDDE_E_INVALID, "no kernel image loaded (jdde_net_load_image)");
	if (n_anchors < 2) return nfail(h, JDDE_E_INVALID, "need at least two anchors (reference: _jitcdde.py:582)");
	if (n_anchors+1 > h->P.cap) return nfail(h, JDDE_E_INVALID, "initial past does not fit into the ring; increase ring_capacity");
	if (!times || !states || !diffs) return nfail(h, JDDE_E_INVALID, "null argument");
	size_t const n = (size_t)h->desc.n, k = (size_t)n_anchors;
	double * d_t = nullptr, * d_s = nullptr, * d_d = nullptr;
	JN_CUDA(h, cudaMalloc((void **)&d_t, k*sizeof(double)));
	JN_CUDA(h, cudaMalloc((void **)&d_s, k*n*sizeof(double)));
	JN_CUDA(h, cudaMalloc((void **)&d_d, k*n*sizeof(double)));
	JN_CUDA(h, cudaMemcpyAsync(d_t, times, k*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	JN_CUDA(h, cudaMemcpyAsync(d_s, states, k*n*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	JN_CUDA(h, cudaMemcpyAsync(d_d, diffs, k*n*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	int na = n_anchors; long long E = h->desc.n_edges;
	void * args[] = { &h->P, &na, &d_t, &d_s, &d_d, &E };
	rc = nlaunch(h, h->k_set_past, (long long)n, h->block, 0, args);
	cudaStreamSynchronize(h->stream);
	cudaFree(d_t); cudaFree(d_s); cudaFree(d_d);
	if (rc) return rc;
	h->have_past = true;
	return JDDE_OK;
}

int jdde_net_set_integration_parameters(jdde_net * h, const jdde_int_params * p, double max_delay)
{
	int rc = nready(h);
	if (rc) return rc;
	if (!p) return nfail(h, JDDE_E_INVALID, "null argument");
	if (!(p->decrease_threshold >= 1.0) || !(p->increase_threshold <= 1.0) || !(p->max_factor >= 1.0) || !(p->min_factor <= 1.0)
		|| !(p->safety_factor <= 1.0) || !(p->atol >= 0.0) || !(p->rtol >= 0.0) || !(max_delay >= 0.0))
		return nfail(h, JDDE_E_INVALID, "integration parameter out of range (see _jitcdde.py:695-708)");
	jdde_int_params q = *p;
	if (q.first_step > q.max_step) q.first_step = q.max_step;
	if (q.min_step > q.first_step) q.min_step = q.first_step;
	/* controller fields are set in place: P, dt (= first_step, _jitcdde.py:712), max_delay */
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	jdn_control c;
	JN_CUDA(h, cudaMemcpy(&c, h->P.ctrl, sizeof c, cudaMemcpyDeviceToHost));
	c.P = q; c.dt = q.first_step; c.max_delay = max_delay; c.cap = h->P.cap;
	h->max_step = q.max_step;
	c.last_pws = 0.0; c.count = 0; c.credit = 0.0;      /* _jitcdde.py:727-729 */
	JN_CUDA(h, cudaMemcpy(h->P.ctrl, &c, sizeof c, cudaMemcpyHostToDevice));
	h->have_int = true;
	return JDDE_OK;
}

int jdde_net_begin(jdde_net * h, int32_t mode, double target, double step, int64_t steps)
{
	int rc = nready_run(h);
	if (rc) return rc;
	int m = mode; long long s = steps;
	void * args[] = { &h->P, &m, &target, &step, &s };
	return nlaunch(h, h->k_begin, 1, 1, 0, args);
}

int jdde_net_attempt(jdde_net * h)
{
	int rc = nready_run(h);
	if (rc) return rc;
	h->P.iterate = 0;      /* the two-kernel attempt does not iterate on a past within the step: it stops the integration */
	long long E = h->desc.n_edges;
	void * largs[] = { &h->P, &h->scratch, &E };
	if (E && (rc = nlaunch(h, h->k_lookup, E, h->block, (size_t)h->P.cap*sizeof(double), largs))) return rc;
	void * args[] = { &h->P, &h->scratch };
	if ((rc = nlaunch(h, h->k_attempt, h->desc.hi-h->desc.lo, h->block, 0, args))) return rc;
	if (h->P.world > 1)
	{
		/* the step kernel has stored this rank's part into every rank's staging buffer: complete the exchange */
		void * xargs[] = { &h->P };
		return nlaunch(h, h->k_exchange, 1, 1, 0, xargs);
	}
	return JDDE_OK;
}

int jdde_net_control(jdde_net * h)
{
	int rc = nready_run(h);
	if (rc) return rc;
	void * args[] = { &h->P };
	if ((rc = nlaunch(h, h->k_control, 1, 1, 0, args))) return rc;
	return nlaunch(h, h->k_commit, h->desc.n, h->block, 0, args);
}

int jdde_net_poll(jdde_net * h, int32_t * done, double * t, int32_t * status)
{
	int rc = nready(h);
	if (rc) return rc;
	JN_CUDA(h, cudaMemcpyAsync(&h->host_ctrl, h->P.ctrl, sizeof(jdn_control), cudaMemcpyDeviceToHost, h->stream));
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	if (done) *done = h->host_ctrl.done;
	if (t) *t = h->host_ctrl.t;
	if (status) *status = h->host_ctrl.status;
	return JDDE_OK;
}

int jdde_net_recent_state(jdde_net * h, double t, int32_t current_only, double * out)
{
	int rc = nready_run(h);
	if (rc) return rc;
	int co = current_only;
	void * args[] = { &h->P, &t, &co, &h->d_out };
	if ((rc = nlaunch(h, h->k_recent, h->desc.n, h->block, 0, args))) return rc;
	if (out)
	{
		JN_CUDA(h, cudaMemcpyAsync(out, h->d_out, (size_t)h->desc.n*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
		JN_CUDA(h, cudaStreamSynchronize(h->stream));
	}
	return JDDE_OK;
}

/* poll_every × (attempt, control) as one graph launch; kernels return at once when the controller has set `done` */
static int launch_batch(jdde_net * h)
{
	int rc;
	const char * const choice = getenv("JITCDDE_B200_NET_GRAPH");
	if (choice && choice[0] == '0')
	{
		for (int k = 0; k < h->poll_every; k++)
		{
			if ((rc = jdde_net_attempt(h))) return rc;
			if ((rc = jdde_net_control(h))) return rc;
		}
		return JDDE_OK;
	}
	if (!h->batch)
	{
		if ((rc = nready_run(h))) return rc;
		int64_t const before = h->launches;
		JN_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
		rc = JDDE_OK;
		for (int k = 0; k < h->poll_every && !rc; k++)
		{
			rc = jdde_net_attempt(h);
			if (!rc) rc = jdde_net_control(h);
		}
		cudaGraph_t graph = nullptr;
		cudaError_t const e = cudaStreamEndCapture(h->stream, &graph);
		h->launches = before;
		if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
		if (e != cudaSuccess) return nfail(h, JDDE_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
		cudaError_t const ei = cudaGraphInstantiate(&h->batch, graph, 0);
		cudaGraphDestroy(graph);
		if (ei != cudaSuccess) { h->batch = nullptr; return nfail(h, JDDE_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei)); }
		h->launches_per_batch = 0;
		long long const per_attempt = (h->desc.n_edges ? 1 : 0) + 1 + (h->P.world > 1 ? 1 : 0) + 2;
		h->launches_per_batch = per_attempt*h->poll_every;
	}
	JN_CUDA(h, cudaGraphLaunch(h->batch, h->stream));
	h->launches += h->launches_per_batch;
	return JDDE_OK;
}

/* one cooperative launch of the persistent stepper: up to `attempts` attempted steps on the device */
static int launch_run(jdde_net * h, int attempts)
{
	int rc;
	if ((rc = nready_run(h))) return rc;
	cudaKernel_t const kernel = h->warp_tiles ? h->k_run_warp : h->k_run;
	size_t const smem = (size_t)h->P.cap*sizeof(double)+(h->warp_tiles ? 0 : (size_t)h->run_smem_bytes);
	if (!h->run_grid)
	{
		int per_sm = 0, sms = 0;
		JN_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kernel, h->block, smem));
		JN_CUDA(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->desc.device));
		if (per_sm < 1) return nfail(h, JDDE_E_CUDA, "the persistent stepper does not fit an SM");
		long long needed;
		if (h->warp_tiles)
			needed = (h->P.n_warp_tiles*32 + h->block - 1)/h->block;
		else
		{
			long long const work = h->desc.n_edges > h->desc.hi-h->desc.lo ? h->desc.n_edges : h->desc.hi-h->desc.lo;
			needed = (work + h->block - 1)/h->block;
		}
		long long const resident = (long long)per_sm*sms;
		h->run_grid = (int)(needed < resident ? (needed ? needed : 1) : resident);
		const char * const cap = getenv("JITCDDE_B200_NET_GRID");      /* experiments: fewer blocks make the grid barriers cheaper */
		if (cap && atoi(cap) > 0 && atoi(cap) < h->run_grid)
			h->run_grid = atoi(cap);
	}
	if (!h->P.tile_nodes)
	{
		/* nodes per block tile (at most the image's JDN_TILE_NODES = 64): the in-edges of a tile just below a multiple of the
		 * block size, so that no round of the one-thread-per-edge loops runs half empty (in-degree 10, 256 threads: 51 nodes —
		 * measured 0.083 ms per attempt against 0.088 with 64 and 0.090–0.102 with 25–38 nodes at 125 000 nodes,
		 * profiles/configs_r02.md), and at least two tiles per block for small networks */
		long long const owned = h->desc.hi-h->desc.lo;
		double const degree = owned > 0 ? (double)h->desc.n_edges/(double)owned : 0.0;
		long long nodes = owned/(2ll*h->run_grid);
		nodes = nodes > 64 ? 64 : nodes < 16 ? 16 : nodes;
		if (degree > 0)
		{
			long long const rounds = (long long)((double)nodes*degree/h->block);
			if (rounds >= 1)
			{
				long long const aligned = (long long)((double)rounds*h->block/degree);
				if (aligned >= 16 && aligned < nodes)
					nodes = aligned;
			}
		}
		const char * const choice = getenv("JITCDDE_B200_NET_TILE_NODES");
		if (choice && atoi(choice) > 0)
			nodes = atoi(choice);
		h->P.tile_nodes = (int)nodes;
	}
	h->P.iterate = h->warp_tiles ? 0 : 1;      /* block tiles handle a past within the step; warp tiles stop the integration */
	/* an adaptive step can only reach into a delay if max_step exceeds the smallest delay of this rank's in-edges (blind steps
	 * are all accepted: no tentative anchor outlives them); otherwise no lookup ever reads a tentative anchor and the
	 * stepper does not store it */
	h->P.keep_tentative = h->P.iterate && h->max_step > h->min_tau;
	long long E = h->desc.n_edges;
	void * args[] = { &h->P, &h->scratch, &E, &attempts };
	cudaLaunchConfig_t config{};
	config.gridDim = dim3((unsigned)h->run_grid);
	config.blockDim = dim3((unsigned)h->block);
	config.dynamicSmemBytes = smem;
	config.stream = h->stream;
	cudaLaunchAttribute attribute{};
	attribute.id = cudaLaunchAttributeCooperative;
	attribute.val.cooperative = 1;
	config.attrs = &attribute;
	config.numAttrs = 1;
	JN_CUDA(h, cudaLaunchKernelExC(&config, (const void *)kernel, args));
	h->launches++;
	return JDDE_OK;
}

static int run_until_done(jdde_net * h)
{
	for (;;)
	{
		int rc;
		if (h->persistent)
		{
			if ((rc = launch_run(h, h->poll_every*32))) return rc;
		}
		else if ((rc = launch_batch(h))) return rc;
		int32_t done = 0;
		if ((rc = jdde_net_poll(h, &done, nullptr, nullptr))) return rc;
		if (done) return JDDE_OK;
	}
}

/* single-rank convenience: the whole jitcdde.integrate (_jitcdde.py:807-836) */
int jdde_net_integrate(jdde_net * h, double target, double * out_state, int32_t * status)
{
	int rc = jdde_net_begin(h, JDN_MODE_ADAPTIVE, target, 0.0, 0);
	if (rc) return rc;
	if ((rc = run_until_done(h))) return rc;
	if (status) *status = h->host_ctrl.status;
	if (h->host_ctrl.status == JDDE_STATUS_OK)
		return jdde_net_recent_state(h, target, 0, out_state);
	return JDDE_OK;
}

/* one step of step_on_discontinuities (_jitcdde.py:985-997): adaptive steps, each capped at the distance to the target */
int jdde_net_step_to(jdde_net * h, double target, double * out_state, int32_t * status)
{
	int rc = jdde_net_begin(h, JDN_MODE_CAPPED, target, 0.0, 0);
	if (rc) return rc;
	if ((rc = run_until_done(h))) return rc;
	if (status) *status = h->host_ctrl.status;
	if (h->host_ctrl.status == JDDE_STATUS_OK)
		return jdde_net_recent_state(h, target, 0, out_state);
	return JDDE_OK;
}

/* integrate_blindly (_jitcdde.py:880-933); step <= 0 means max_step */
int jdde_net_integrate_blindly(jdde_net * h, double target, double step, double * out_state, int32_t * status)
{
	int rc = nready_run(h);
	if (rc) return rc;
	if ((rc = jdde_net_poll(h, nullptr, nullptr, nullptr))) return rc;
	if (!(step > 0)) step = h->host_ctrl.P.max_step;
	double const total = target-h->host_ctrl.t;
	if (total < 0) return nfail(h, JDDE_E_INVALID, "Can't integrate backwards in time.");
	long long number = 0; double dt = 0;
	if (total > 0)
	{
		if (total < step) step = total;
		number = (long long) rint(total/step);
		dt = total/number;
	}
	if ((rc = jdde_net_begin(h, JDN_MODE_BLIND, target, dt, number))) return rc;
	if (number > 0 && (rc = run_until_done(h))) return rc;
	if ((rc = jdde_net_poll(h, nullptr, nullptr, nullptr))) return rc;
	if (status) *status = h->host_ctrl.status;
	return jdde_net_recent_state(h, target, 1, out_state);
}

/* the reference's history is an unbounded list: when the ring is full (status JDDE_STATUS_OVERFLOW) it is enlarged … */
int jdde_net_grow_ring(jdde_net * h, int32_t new_capacity)
{
	int rc = nready_run(h);
	if (rc) return rc;
	if (new_capacity <= h->P.cap || (new_capacity & (new_capacity-1)))
		return nfail(h, JDDE_E_INVALID, "new capacity must be a larger power of two");
	if (new_capacity > JDN_MAX_RING_CAPACITY)
		return nfail(h, JDDE_E_NOMEM, "ring capacity above 16384 (the anchor times of the ring are staged in shared memory)");
	if ((size_t)new_capacity*sizeof(double) > 48*1024)
		JN_CUDA(h, cudaFuncSetAttribute((const void *)h->k_lookup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)new_capacity*sizeof(double))));
	if (h->k_run && (size_t)new_capacity*sizeof(double)+h->run_smem_bytes > 48*1024)
		JN_CUDA(h, cudaFuncSetAttribute((const void *)h->k_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)new_capacity*sizeof(double)+h->run_smem_bytes)));
	if (h->k_run_warp && (size_t)new_capacity*sizeof(double) > 48*1024)
		JN_CUDA(h, cudaFuncSetAttribute((const void *)h->k_run_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)new_capacity*sizeof(double))));
	h->run_grid = 0;
	h->P.tile_nodes = 0;
	double * new_times = nullptr, * new_hist = nullptr;
	if ((rc = nalloc(h, &new_times, (size_t)new_capacity))) return rc;
	if ((rc = nalloc(h, &new_hist, 2*(size_t)new_capacity*(size_t)h->desc.n))) return rc;
	int cap = new_capacity;
	void * args[] = { &h->P, &new_times, &new_hist, &cap };
	if ((rc = nlaunch(h, h->k_regrow, h->desc.n, h->block, 0, args))) return rc;
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	for (double * old : { h->P.times, h->P.hist })
	{
		cudaFree(old);
		for (auto it = h->allocations.begin(); it != h->allocations.end(); ++it)
			if (*it == old) { h->allocations.erase(it); break; }
	}
	h->P.times = new_times;
	h->P.hist = new_hist;
	h->P.cap = new_capacity;
	h->desc.ring_capacity = new_capacity;
	JN_CUDA(h, cudaMemcpy(&h->P.ctrl->cap, &cap, sizeof cap, cudaMemcpyHostToDevice));
	drop_batch(h);
	return JDDE_OK;
}

/*
 * adjust_diff (_jitcdde.py:863-878) in two halves, so that several ranks can meet in between (phase 1 stores this rank's
 * nodes into every rank's buffers, phase 2 reads all nodes): begin → [barrier of the ranks] → end. One rank: jdde_net_adjust_diff.
 */
/* apply_jump(amplitude, time, width) (jitced_template.c:1122-1174) for a `time` on the last segment of the history; phase 1 */
static int jump_begin(jdde_net * h, const double * amplitude, double time, double width, double t_last, double t_before)
{
	int rc;
	if (!(time > t_before) || !(time <= t_last) || !(width > 0))
		return nfail(h, JDDE_E_INVALID, "jump: the time must lie on the last segment of the history (after the last anchor but one, not after the last) and the width must be positive");
	h->adjust_time = time;
	h->adjust_new_t = time+width;                           /* apply_jump: new->time = time+width, jitced_template.c:1141 */
	double * d_amplitude = nullptr;
	if (amplitude)
	{
		d_amplitude = h->d_out;                             /* (free until phase 2 writes the extrema) */
		JN_CUDA(h, cudaMemcpyAsync(d_amplitude, amplitude, (size_t)h->desc.n*sizeof(double), cudaMemcpyHostToDevice, h->stream));
	}
	void * args[] = { &h->P, &h->adjust_time, &h->adjust_new_t, &d_amplitude };
	if ((rc = nlaunch(h, h->k_adjust, h->desc.hi-h->desc.lo, h->block, 0, args))) return rc;
	JN_CUDA(h, cudaStreamSynchronize(h->stream));          /* (the peers may read what this rank stored once the ranks have met) */
	h->adjusting = true;
	return JDDE_OK;
}

static int last_two_times(jdde_net * h, double * t_last, double * t_before)
{
	int rc = nready_run(h);
	if (rc) return rc;
	if ((rc = jdde_net_poll(h, nullptr, nullptr, nullptr))) return rc;
	jdn_control const & c = h->host_ctrl;
	if (c.status != JDDE_STATUS_OK) return nfail(h, JDDE_E_INVALID, "the integration has failed; set a new past first");
	if (c.current+2-c.first > h->P.cap) return nfail(h, JDDE_E_NOMEM, "anchor ring full; increase ring_capacity");
	JN_CUDA(h, cudaMemcpyAsync(t_before, h->P.times+((c.current-1) & (h->P.cap-1)), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	*t_last = c.t;
	return JDDE_OK;
}

int jdde_net_adjust_diff_begin(jdde_net * h, double shift_ratio)
{
	double t_last = 0.0, t_before = 0.0;
	int rc = last_two_times(h, &t_last, &t_before);
	if (rc) return rc;
	double const width = shift_ratio*(t_last-t_before);      /* _jitcdde.py:875-878 and jump(…, forward=False), :958-960 */
	return jump_begin(h, nullptr, t_last-width, width, t_last, t_before);
}

/* jump (_jitcdde.py:935-974): amplitude [n]; forward != 0: the jump interval is (time, time+width), else (time−width, time) */
int jdde_net_jump_begin(jdde_net * h, const double * amplitude, double time, double width, int32_t forward)
{
	if (!amplitude) return nfail(h, JDDE_E_INVALID, "null argument");
	double t_last = 0.0, t_before = 0.0;
	int rc = last_two_times(h, &t_last, &t_before);
	if (rc) return rc;
	return jump_begin(h, amplitude, forward ? time : time-width, width, t_last, t_before);
}

int jdde_net_adjust_diff_end(jdde_net * h, double * minima, double * maxima)
{
	int rc = nready_run(h);
	if (rc) return rc;
	if (!h->adjusting) return nfail(h, JDDE_E_INVALID, "jdde_net_adjust_diff_end without jdde_net_adjust_diff_begin");
	h->adjusting = false;
	void * args[] = { &h->P, &h->adjust_time, &h->adjust_new_t, &h->d_out };
	if ((rc = nlaunch(h, h->k_adjust_commit, h->desc.n, h->block, 0, args))) return rc;
	void * cargs[] = { &h->P, &h->adjust_time, &h->adjust_new_t };
	if ((rc = nlaunch(h, h->k_adjust_control, 1, 1, 0, cargs))) return rc;
	size_t const bytes = (size_t)h->desc.n*sizeof(double);
	if (minima) JN_CUDA(h, cudaMemcpyAsync(minima, h->d_out, bytes, cudaMemcpyDeviceToHost, h->stream));
	if (maxima) JN_CUDA(h, cudaMemcpyAsync(maxima, h->d_out+h->desc.n, bytes, cudaMemcpyDeviceToHost, h->stream));
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	return JDDE_OK;
}

int jdde_net_jump(jdde_net * h, const double * amplitude, double time, double width, int32_t forward, double * minima, double * maxima)
{
	if (h && h->P.world > 1) return nfail(h, JDDE_E_INVALID, "several ranks: jdde_net_jump_begin, a barrier of the ranks, jdde_net_adjust_diff_end");
	int rc = jdde_net_jump_begin(h, amplitude, time, width, forward);
	if (rc) return rc;
	return jdde_net_adjust_diff_end(h, minima, maxima);
}

int jdde_net_adjust_diff(jdde_net * h, double shift_ratio, double * minima, double * maxima)
{
	if (h && h->P.world > 1) return nfail(h, JDDE_E_INVALID, "several ranks: jdde_net_adjust_diff_begin, a barrier of the ranks, jdde_net_adjust_diff_end");
	int rc = jdde_net_adjust_diff_begin(h, shift_ratio);
	if (rc) return rc;
	return jdde_net_adjust_diff_end(h, minima, maxima);
}

/* … and the call goes on where it stopped (same target, same remaining steps) */
int jdde_net_resume(jdde_net * h, double * out_state, int32_t * status)
{
	int rc = jdde_net_begin(h, JDN_MODE_CONTINUE, 0.0, 0.0, 0);
	if (rc) return rc;
	if ((rc = run_until_done(h))) return rc;
	if (status) *status = h->host_ctrl.status;
	if (h->host_ctrl.mode == JDN_MODE_BLIND)
		return jdde_net_recent_state(h, h->host_ctrl.target, 1, out_state);
	if (h->host_ctrl.status == JDDE_STATUS_OK)
		return jdde_net_recent_state(h, h->host_ctrl.target, 0, out_state);
	return JDDE_OK;
}

int jdde_net_get_counters(jdde_net * h, int64_t * accepted, int64_t * rejected, int64_t * attempts)
{
	int rc = jdde_net_poll(h, nullptr, nullptr, nullptr);
	if (rc) return rc;
	if (accepted) *accepted = h->host_ctrl.accepted;
	if (rejected) *rejected = h->host_ctrl.rejected;
	if (attempts) *attempts = h->host_ctrl.attempts;
	return JDDE_OK;
}

int jdde_net_get_dt(jdde_net * h, double * dt)
{
	int rc = jdde_net_poll(h, nullptr, nullptr, nullptr);
	if (rc) return rc;
	if (dt) *dt = h->host_ctrl.dt;
	return JDDE_OK;
}

/*
 * get_state (_jitcdde.py:357-371; the reference forgets first, then hands out all anchors): the live anchors first … current in
 * order — checkpoint / resume, switching integrators. A first call with times == nullptr only reports their number. No kernel:
 * the groups of four ring slots that hold live anchors are copied to the host one at a time and unpacked there.
 */
int jdde_net_get_history(jdde_net * h, int32_t * count, double * times, double * states, double * diffs, int32_t capacity)
{
	int rc = nready_run(h);
	if (rc) return rc;
	if ((rc = jdde_net_poll(h, nullptr, nullptr, nullptr))) return rc;
	jdn_control const & c = h->host_ctrl;
	int const live = c.current-c.first+1;
	if (count) *count = live;
	if (!times) return JDDE_OK;
	if (!states || !diffs) return nfail(h, JDDE_E_INVALID, "null argument");
	if (capacity < live) return nfail(h, JDDE_E_INVALID, "room for fewer anchors than there are (ask for their number first)");
	size_t const n = (size_t)h->desc.n;
	int const mask = h->P.cap-1;
	int const group_slots = 1 << JDN_ROW_GROUP_LOG2;
	std::vector<double> ring_times((size_t)h->P.cap), group(2*n*(size_t)group_slots);
	JN_CUDA(h, cudaMemcpyAsync(ring_times.data(), h->P.times, ring_times.size()*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	int loaded = -1;
	for (int k = 0; k < live; k++)
	{
		int const slot = (c.first+k) & mask;
		times[k] = ring_times[(size_t)slot];
		if ((slot >> JDN_ROW_GROUP_LOG2) != loaded)
		{
			loaded = slot >> JDN_ROW_GROUP_LOG2;
			JN_CUDA(h, cudaMemcpyAsync(group.data(), h->P.hist + 2*jdn_pair_index(loaded << JDN_ROW_GROUP_LOG2, 0, (long long)n), group.size()*sizeof(double), cudaMemcpyDeviceToHost, h->stream));
			JN_CUDA(h, cudaStreamSynchronize(h->stream));
		}
		size_t const within = (size_t)(slot & (group_slots-1));
		for (size_t i = 0; i < n; i++)
		{
			states[(size_t)k*n+i] = group[2*(i*group_slots+within)];
			diffs[(size_t)k*n+i] = group[2*(i*group_slots+within)+1];
		}
	}
	return JDDE_OK;
}

int jdde_net_exchange_pointers(jdde_net * h, uint64_t * stage, uint64_t * p_bits)
{
	if (!h) return JDDE_E_INVALID;
	if (stage) *stage = (uint64_t)(uintptr_t)h->P.stage;
	if (p_bits) *p_bits = (uint64_t)(uintptr_t)&h->P.ctrl->p_bits;
	return JDDE_OK;
}

int jdde_net_exchange_handle(jdde_net * h, void * handle, size_t handle_bytes)
{
	int rc = nready(h);
	if (rc) return rc;
	if (!handle || handle_bytes < sizeof(cudaIpcMemHandle_t)) return nfail(h, JDDE_E_INVALID, "handle buffer too small (needs 64 bytes, 128 with the history replica)");
	if (!h->exchange)
	{
		size_t const bytes = sizeof(jdn_exchange) + 2*2*(size_t)h->desc.n*sizeof(double);
		void * q = nullptr;
		cudaError_t e = cudaMalloc(&q, bytes);
		if (e != cudaSuccess) return nfail(h, e == cudaErrorMemoryAllocation ? JDDE_E_NOMEM : JDDE_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
		h->allocations.push_back(q);
		JN_CUDA(h, cudaMemset(q, 0, bytes));
		h->exchange = static_cast<jdn_exchange *>(q);
	}
	cudaIpcMemHandle_t ipc;
	JN_CUDA(h, cudaIpcGetMemHandle(&ipc, h->exchange));
	memset(handle, 0, handle_bytes);
	memcpy(handle, &ipc, sizeof ipc);
	return JDDE_OK;
}

int jdde_net_connect_peers(jdde_net * h, int32_t rank, int32_t world, const void * handles, size_t handle_bytes)
{
	int rc = nready(h);
	if (rc) return rc;
	if (world < 2 || rank < 0 || rank >= world || !handles || handle_bytes < sizeof(cudaIpcMemHandle_t))
		return nfail(h, JDDE_E_INVALID, "invalid rank, world or handles");
	if (!h->exchange) return nfail(h, JDDE_E_INVALID, "call jdde_net_exchange_handle first");
	if (!h->k_exchange) return nfail(h, JDDE_E_IMAGE, "image lacks kernel jdn_k_exchange");
	if (h->P.world > 1) return nfail(h, JDDE_E_INVALID, "peers already connected");
	JN_CUDA(h, cudaStreamSynchronize(h->stream));
	{
		std::vector<jdn_exchange *> peers((size_t)world, nullptr);
		for (int r = 0; r < world; r++)
		{
			if (r == rank)
			{
				peers[r] = h->exchange;
				continue;
			}
			cudaIpcMemHandle_t ipc;
			memcpy(&ipc, (const char *)handles + (size_t)r*handle_bytes, sizeof ipc);
			void * q = nullptr;
			JN_CUDA(h, cudaIpcOpenMemHandle(&q, ipc, cudaIpcMemLazyEnablePeerAccess));
			h->mapped_peers.push_back(q);
			peers[r] = static_cast<jdn_exchange *>(q);
		}
		if ((rc = nalloc(h, &h->d_peers, (size_t)world))) return rc;
		JN_CUDA(h, cudaMemcpy(h->d_peers, peers.data(), (size_t)world*sizeof(jdn_exchange *), cudaMemcpyHostToDevice));
	}
	drop_batch(h);
	h->P.peers = h->d_peers;
	h->P.rank = rank;
	h->P.world = world;
	return JDDE_OK;
}

int64_t jdde_net_launch_count(const jdde_net * h) { return h ? h->launches : 0; }

} // extern "C"
