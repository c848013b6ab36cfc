// tools/ubench_fma.cu -- what is the FP32 issue ceiling of one B200 SM for the instruction mixes K3 uses?
// Measures FMA lanes / clk / SM for: scalar FFMA, packed FFMA2, FFMA2 + MUFU.RCP at K3's ratio, and
// FFMA2 + MUFU + the ALU side work (FMNMX, LEA.HI) at K3's ratio.  Build: nvcc -arch=sm_100a -O3 -o ubench_fma ...
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b)
{
	float2 v[CHAINS];
	unsigned cnt = 0;
	for (int c = 0; c < CHAINS; ++c) v[c] = make_float2(threadIdx.x * 1e-3f + c, c * 0.5f);
	const float2 A = make_float2(a, a), B = make_float2(b, b);
	for (int it = 0; it < ITERS; ++it)
	{
#pragma unroll
		for (int c = 0; c < CHAINS; ++c)
		{
			if (MODE == 0)
			{ // 18 scalar FFMA x 2 halves
#pragma unroll
				for (int j = 0; j < 9; ++j) { v[c].x = fmaf(v[c].x, a, b); v[c].y = fmaf(v[c].y, a, b); v[c].x = fmaf(v[c].x, b, a); v[c].y = fmaf(v[c].y, b, a); }
			}
			else
			{
#pragma unroll
				for (int j = 0; j < 9; ++j) { v[c] = __ffma2_rn(v[c], A, B); v[c] = __ffma2_rn(v[c], B, A); }
				if (MODE >= 2)
				{
					float rx, ry;
					asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(v[c].x));
					asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(v[c].y));
					v[c] = __ffma2_rn(make_float2(rx, ry), A, B);
				}
				if (MODE >= 3)
				{
					cnt += __float_as_uint(v[c].x) >> 31;
					cnt += __float_as_uint(v[c].y) >> 31;
					v[c] = __fadd2_rn(v[c], make_float2(fminf(v[c].x, 0.f), fminf(v[c].y, 0.f)));
				}
			}
		}
	}
	float s = 0;
	for (int c = 0; c < CHAINS; ++c) s += v[c].x + v[c].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + cnt;
}

template <int MODE>
void run(const char *name, int blocks_per_sm, int sms, double fma_per_iter_chain)
{
	float *out;
	cudaMalloc(&out, sizeof(float) * 256 * blocks_per_sm * sms);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k<MODE><<<blocks_per_sm * sms, 256>>>(out, 1.0001f, 0.5f);
	cudaEventRecord(e0);
	k<MODE><<<blocks_per_sm * sms, 256>>>(out, 1.0001f, 0.5f);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double fma = (double)blocks_per_sm * sms * 256 * ITERS * CHAINS * fma_per_iter_chain;
	printf("%-34s blocks/SM %d: %.3f ms  %.2f TFLOP/s (2 flop per FMA lane)  %.1f FMA lanes/clk/SM @1965 MHz\n", name, blocks_per_sm, ms,
	       2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.965e9);
	cudaFree(out);
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount;
	for (int b = 2; b <= 8; b *= 2)
	{
		run<0>("scalar FFMA", b, sms, 36);
		run<1>("packed FFMA2", b, sms, 36);
		run<2>("FFMA2 + 2 MUFU / 19 FFMA2", b, sms, 38);
		run<3>("FFMA2 + MUFU + FMNMX/LEA/FADD2", b, sms, 40);
	}
	return 0;
}
