/*
 * jdde_peak.cu — FP64 FMA throughput microbenchmark (the pool's MEASURED_PEAKS.json has no FP64 entry; SURVEY.md §0
 * asks the builder to measure it). 8 independent DFMA chains per thread, 4096 iterations, every SM full:
 * FLOP/s = 2 · 8 · iterations · threads / time. Used by bench.py for the FP64 fraction of the integrate kernel.
 */
extern "C" __global__ void __launch_bounds__(256)
jdde_k_fp64_peak(double * const out, int const iterations, double const a, double const b)
{
	double x0 = threadIdx.x, x1 = x0+1, x2 = x0+2, x3 = x0+3, x4 = x0+4, x5 = x0+5, x6 = x0+6, x7 = x0+7;
	for (int i = 0; i < iterations; i++)
	{
		x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
		x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
	}
	double const s = ((x0+x1)+(x2+x3))+((x4+x5)+(x6+x7));
	if (s == 12345.678)      /* never true: keeps the chains alive */
		out[blockIdx.x*blockDim.x+threadIdx.x] = s;
}
