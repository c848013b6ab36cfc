// k_ubench.cu -- measures the FP32 roofline K3 is reported against, on the device the context owns.
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the scoring kernels are FP32-FMA-pipe bound (SURVEY.md
// 8d), so bench.py measures that pipe itself, in the same process and under the same clocks as the timed region:
// mode 0 = dependent chains of packed FFMA2 (the pipe's peak: 2 flop per lane per FMA), mode 1 = the same with K3's side
// work per 19 FFMA2 (2 MUFU.RCP, FMNMX / sign-bit count / FADD2) -- the ceiling of K3's instruction mix.
#include "gcr_internal.cuh"

#define UB_ITERS 2048
#define UB_CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k_ubench_fp32(float *out, float a, float b)
{
	float2 v[UB_CHAINS];
	unsigned cnt = 0;
#pragma unroll
	for (int c = 0; c < UB_CHAINS; ++c) v[c] = make_float2(threadIdx.x * 1e-3f + c, c * 0.5f);
	const float2 A = make_float2(a, a), B = make_float2(b, b);
	for (int it = 0; it < UB_ITERS; ++it)
	{
#pragma unroll
		for (int c = 0; c < UB_CHAINS; ++c)
		{
#pragma unroll
			for (int j = 0; j < 9; ++j) { v[c] = __ffma2_rn(v[c], A, B); v[c] = __ffma2_rn(v[c], B, A); }
			if (MODE == 1)
			{
				float rx, ry;
				asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(v[c].x));
				asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(v[c].y));
				v[c] = __ffma2_rn(make_float2(rx, ry), A, B);
				cnt += __float_as_uint(v[c].x) >> 31;
				cnt += __float_as_uint(v[c].y) >> 31;
				v[c] = __fadd2_rn(v[c], make_float2(fminf(v[c].x, 0.f), fminf(v[c].y, 0.f)));
			}
		}
	}
	float s = 0;
#pragma unroll
	for (int c = 0; c < UB_CHAINS; ++c) s += v[c].x + v[c].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + cnt;
}

extern "C" gcr_status gcr_ubench_fp32(gcr_ctx *ctx, int mode, double min_ms, double *tflops_out, double *ms_out)
{
	if (!ctx || !tflops_out || (mode != 0 && mode != 1))
		return GCR_ERR_ARG;
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const int blocks = 4 * ctx->sm_count;
	gcr_status st = gcr_ensure_scratch(ctx, (size_t)blocks * 256 * sizeof(float));
	if (st)
		return st;
	cudaEvent_t e0, e1;
	GCR_CHECK_CUDA(ctx, cudaEventCreate(&e0));
	GCR_CHECK_CUDA(ctx, cudaEventCreate(&e1));
	auto launch = [&]() {
		if (mode == 0) k_ubench_fp32<0><<<blocks, 256, 0, ctx->stream>>>((float *)ctx->d_scratch, 1.0001f, 0.5f);
		else k_ubench_fp32<1><<<blocks, 256, 0, ctx->stream>>>((float *)ctx->d_scratch, 1.0001f, 0.5f);
	};
	launch(); // warm-up
	++ctx->launches;
	double total_ms = 0.0, best = 0.0;
	long launches = 0;
	// FMA lanes per launch: mode 0: 18 FFMA2 x 2 lanes per chain-iteration; mode 1: 19 FFMA2 (+ 1 FADD2 not counted)
	const double fma_lanes = (double)blocks * 256 * UB_ITERS * UB_CHAINS * (mode == 0 ? 36.0 : 38.0);
	do
	{
		GCR_CHECK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
		launch();
		GCR_CHECK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
		GCR_CHECK_CUDA(ctx, cudaEventSynchronize(e1));
		float ms = 0.f;
		GCR_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
		++ctx->launches;
		++launches;
		total_ms += ms;
		const double tf = 2.0 * fma_lanes / (ms * 1e-3) / 1e12;
		if (tf > best) best = tf;
	} while (total_ms < min_ms);
	GCR_CHECK_CUDA(ctx, cudaGetLastError());
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	// sustained figure over the whole interval (>= min_ms); the best single launch is the burst figure
	tflops_out[0] = 2.0 * fma_lanes * launches / (total_ms * 1e-3) / 1e12;
	tflops_out[1] = best;
	if (ms_out)
		*ms_out = total_ms;
	return GCR_OK;
}
