// k_refit_pnp.cu -- K6 for the absolute-pose problem: the non-minimal fit of DefaultPnPEstimator =
// PerspectiveNPointEstimator<P3PSolver, PnPBundleAdjustment> (types.h:135-138;
// estimators/solver_pnp_bundle_adjustment.h:108-228) -- an initial pose from the whole index set, refined by refine_pnp =
// lm_pnp_impl over CameraJacobianAccumulator, identity camera, TruncatedLoss(1.0), <= 25 iterations
// (relative_pose/bundle.cpp:23-99, 558-600; relative_pose/jacobian_impl.h:8-160).
//
// Stated deviations (the reference has no defined behaviour here): the initial pose is the normalised DLT
// (>= 6 points; smaller sets return no model) instead of OpenCV's EPnP, a dependency outside the tree; the accumulator
// sums over the index set with weights by position (all ones when absent) -- as written it loops over all rows of the
// matrix while indexing sample[i] and multiplies by a weight pointer its callers pass as nullptr.  Statement by
// statement the arithmetic of oracle/gcr_oracle_pba.cpp, which the tests pin against cv2's EPnP + scipy.
//
// One CTA per index set: moments / cost / Jacobian passes block-parallel with a fixed reduction order, the 12 x 12
// inverse iteration, the polar decomposition and the 6 x 6 Cholesky solves on thread 0.
#include "gcr_internal.cuh"

#include <math.h>

#define PNP_THREADS 256
#define PNP_WARPS (PNP_THREADS / 32)

namespace
{
struct PnpRefitArgs
{
	const double *pts; // [n_points][5] of the pair
	const uint32_t *sets, *offsets, *sizes;
	const double *weights; // by position inside a set (sets concatenated like the index sets), or by point, or null
	int weight_mode;
	double *models; // [count][12]
	uint8_t *ok;
	int max_iterations;
};
struct Pose { double q[4], t[3]; };

__device__ void mul3(const double *A, const double *B, double *C)
{
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c)
			C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
__device__ void rotmat_to_quat(const double *R, double *q)
{
	// Eigen::Quaterniond(R) (QuaternionBase::operator=(MatrixBase), quaternionbase_assign_impl<Other, 3, 3>) + normalize();
	// the three cases of its largest-diagonal branch written out (no runtime-indexed local array)
	const double trace = R[0] + R[4] + R[8];
	double w, x, y, z;
	if (trace > 0.0)
	{
		double t = sqrt(trace + 1.0);
		w = 0.5 * t;
		t = 0.5 / t;
		x = (R[7] - R[5]) * t;
		y = (R[2] - R[6]) * t;
		z = (R[3] - R[1]) * t;
	}
	else
	{
		int i = 0;
		if (R[4] > R[0]) i = 1;
		if (R[8] > (i == 1 ? R[4] : R[0])) i = 2;
		if (i == 0) // j = 1, k = 2
		{
			double t = sqrt(R[0] - R[4] - R[8] + 1.0);
			x = 0.5 * t;
			t = 0.5 / t;
			w = (R[7] - R[5]) * t;
			y = (R[3] + R[1]) * t;
			z = (R[6] + R[2]) * t;
		}
		else if (i == 1) // j = 2, k = 0
		{
			double t = sqrt(R[4] - R[8] - R[0] + 1.0);
			y = 0.5 * t;
			t = 0.5 / t;
			w = (R[2] - R[6]) * t;
			z = (R[7] + R[5]) * t;
			x = (R[1] + R[3]) * t;
		}
		else // j = 0, k = 1
		{
			double t = sqrt(R[8] - R[0] - R[4] + 1.0);
			z = 0.5 * t;
			t = 0.5 / t;
			w = (R[3] - R[1]) * t;
			x = (R[2] + R[6]) * t;
			y = (R[5] + R[7]) * t;
		}
	}
	const double nrm = sqrt(w * w + x * x + y * y + z * z);
	q[0] = w / nrm; q[1] = x / nrm; q[2] = y / nrm; q[3] = z / nrm;
}
__device__ void quat_to_rotmat(const double *q, double *R)
{
	const double w = q[0], x = q[1], y = q[2], z = q[3];
	const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
	const double twx = tx * w, twy = ty * w, twz = tz * w;
	const double txx = tx * x, txy = ty * x, txz = tz * x;
	const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
	R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
	R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
	R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}
__device__ void inv3(const double *M, double *R)
{
	const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
	const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
	const double inv = 1.0 / det;
	R[0] = c00 * inv; R[1] = (M[2] * M[7] - M[1] * M[8]) * inv; R[2] = (M[1] * M[5] - M[2] * M[4]) * inv;
	R[3] = c01 * inv; R[4] = (M[0] * M[8] - M[2] * M[6]) * inv; R[5] = (M[2] * M[3] - M[0] * M[5]) * inv;
	R[6] = c02 * inv; R[7] = (M[1] * M[6] - M[0] * M[7]) * inv; R[8] = (M[0] * M[4] - M[1] * M[3]) * inv;
}
// x = -(L L^T)^-1 b from the lower triangle of A (N x N row-major)
template <int N>
__device__ __forceinline__ void llt_solve_neg(const double *A, const double *b, double *x)
{
	double L[N * N];
	for (int i = 0; i < N * N; ++i) L[i] = 0.0;
	for (int j = 0; j < N; ++j)
	{
		double d = A[N * j + j];
		for (int k = 0; k < j; ++k) d -= L[N * j + k] * L[N * j + k];
		L[N * j + j] = sqrt(d);
		for (int i = j + 1; i < N; ++i)
		{
			double v = A[N * i + j];
			for (int k = 0; k < j; ++k) v -= L[N * i + k] * L[N * j + k];
			L[N * i + j] = v / L[N * j + j];
		}
	}
	double y[N];
	for (int i = 0; i < N; ++i)
	{
		double v = b[i];
		for (int k = 0; k < i; ++k) v -= L[N * i + k] * y[k];
		y[i] = v / L[N * i + i];
	}
	for (int i = N - 1; i >= 0; --i)
	{
		double v = y[i];
		for (int k = i + 1; k < N; ++k) v -= L[N * k + i] * x[k];
		x[i] = v / L[N * i + i];
	}
	for (int i = 0; i < N; ++i) x[i] = -x[i];
}
template <int NV>
__device__ void block_sum(double *v, double *red, double *tot)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int j = 0; j < NV; ++j)
	{
		double x = v[j];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
			x += __shfl_xor_sync(0xffffffffu, x, o);
		if (lane == 0)
			red[warp * NV + j] = x;
	}
	__syncthreads();
	if (threadIdx.x < NV)
	{
		double s = 0.0;
		for (int w = 0; w < PNP_WARPS; ++w)
			s += red[w * NV + threadIdx.x];
		tot[threadIdx.x] = s;
	}
	__syncthreads();
}
__device__ __forceinline__ double weight_of(const PnpRefitArgs &a, const uint32_t *idx, size_t set_offset, int k)
{
	if (!a.weights)
		return 1.0;
	(void)set_offset;
	return a.weight_mode == GCR_WEIGHT_BY_POINT ? a.weights[idx[k]] : a.weights[k]; // positional: entry k of the set, as k_refit*
}
// CameraJacobianAccumulator::residual (jacobian_impl.h:26-59), identity camera, TruncatedLoss(1.0); R, t in shared memory
__device__ double cost_pass(const PnpRefitArgs &a, const uint32_t *idx, size_t set_offset, int n, const double *R, const double *t,
                            double *red, double *tot)
{
	double c[1] = {0.0};
	for (int k = threadIdx.x; k < n; k += PNP_THREADS)
	{
		const double *p = a.pts + 5 * (size_t)idx[k];
		const double X0 = p[2], X1 = p[3], X2 = p[4];
		const double Z0 = (R[0] * X0 + R[1] * X1 + R[2] * X2) + t[0], Z1 = (R[3] * X0 + R[4] * X1 + R[5] * X2) + t[1],
		             Z2 = (R[6] * X0 + R[7] * X1 + R[8] * X2) + t[2];
		if (Z2 < 0)
			continue;
		const double inv_z = 1.0 / Z2;
		const double r0 = Z0 * inv_z - p[0], r1 = Z1 * inv_z - p[1];
		const double r_squared = r0 * r0 + r1 * r1;
		const double loss = r_squared < 1.0 ? r_squared : 1.0;
		c[0] += weight_of(a, idx, set_offset, k) * loss;
	}
	block_sum<1>(c, red, tot);
	return tot[0];
}
} // namespace

__global__ void __launch_bounds__(PNP_THREADS) k_refit_pnp(const PnpRefitArgs a)
{
	__shared__ double red[PNP_WARPS * 40], tot[40];
	__shared__ Pose s_pose, s_new;
	__shared__ double s_R[9], s_Rn[9], s_JtJ[36], s_Jtr[6], s_c[3], s_scale, s_M[144];
	__shared__ int s_state; // 0 = failed, 1 = iterate, 2 = converged
	__shared__ double s_cost, s_lambda;
	__shared__ int s_recompute;
	const int set = blockIdx.x, tid = threadIdx.x;
	const size_t set_offset = a.offsets[set];
	const uint32_t *idx = a.sets + set_offset;
	const int n = (int)a.sizes[set];
	double *out = a.models + (size_t)set * 12;
	if (n < 6)
	{
		if (tid < 12) out[tid] = 0.0;
		if (tid == 0) a.ok[set] = 0;
		return;
	}
	// ---- normalised DLT: centroid, mean distance, the four 4 x 4 moment blocks (10 unique entries each) ----
	{
		double c[3] = {0.0, 0.0, 0.0};
		for (int k = tid; k < n; k += PNP_THREADS)
		{
			const double *p = a.pts + 5 * (size_t)idx[k];
			c[0] += p[2]; c[1] += p[3]; c[2] += p[4];
		}
		block_sum<3>(c, red, tot);
		if (tid < 3) s_c[tid] = tot[tid] / n;
		__syncthreads();
		double sp[1] = {0.0};
		for (int k = tid; k < n; k += PNP_THREADS)
		{
			const double *p = a.pts + 5 * (size_t)idx[k];
			const double d0 = p[2] - s_c[0], d1 = p[3] - s_c[1], d2 = p[4] - s_c[2];
			sp[0] += sqrt((d0 * d0 + d1 * d1) + d2 * d2);
		}
		block_sum<1>(sp, red, tot);
		if (tid == 0)
		{
			const double spread = tot[0] / n;
			s_scale = spread > 0.0 ? sqrt(3.0) / spread : 0.0;
		}
		__syncthreads();
	}
	if (!(s_scale > 0.0))
	{
		if (tid < 12) out[tid] = 0.0;
		if (tid == 0) a.ok[set] = 0;
		return;
	}
	{
		double acc[40];
#pragma unroll
		for (int j = 0; j < 40; ++j) acc[j] = 0.0;
		const double s = s_scale;
		for (int k = tid; k < n; k += PNP_THREADS)
		{
			const double *p = a.pts + 5 * (size_t)idx[k];
			const double w = weight_of(a, idx, set_offset, k);
			const double Xh[4] = {s * (p[2] - s_c[0]), s * (p[3] - s_c[1]), s * (p[4] - s_c[2]), 1.0};
			const double u = p[0], v = p[1], q = u * u + v * v;
			int e = 0;
#pragma unroll
			for (int i = 0; i < 4; ++i)
#pragma unroll
				for (int j = i; j < 4; ++j)
				{
					const double ab = w * (Xh[i] * Xh[j]);
					acc[e] += ab;
					acc[10 + e] += u * ab;
					acc[20 + e] += v * ab;
					acc[30 + e] += q * ab;
					++e;
				}
		}
		block_sum<40>(acc, red, tot);
	}
	if (tid == 0)
	{
		s_state = 0;
		double S[4][16];
		int e = 0;
		for (int i = 0; i < 4; ++i)
			for (int j = i; j < 4; ++j)
			{
				for (int b = 0; b < 4; ++b)
					S[b][4 * i + j] = S[b][4 * j + i] = tot[10 * b + e];
				++e;
			}
		// the 12 x 12 normal matrix lives in shared memory and is factored ONCE, in place (a 144-double local array per
		// solve overflowed the thread's stack frame: the factor read back zeros)
		double *M = s_M;
		for (int r = 0; r < 4; ++r)
			for (int c = 0; c < 4; ++c)
			{
				M[12 * r + c] = S[0][4 * r + c];
				M[12 * r + 4 + c] = 0.0;
				M[12 * r + 8 + c] = -S[1][4 * r + c];
				M[12 * (4 + r) + c] = 0.0;
				M[12 * (4 + r) + 4 + c] = S[0][4 * r + c];
				M[12 * (4 + r) + 8 + c] = -S[2][4 * r + c];
				M[12 * (8 + r) + c] = -S[1][4 * r + c];
				M[12 * (8 + r) + 4 + c] = -S[2][4 * r + c];
				M[12 * (8 + r) + 8 + c] = S[3][4 * r + c];
			}
		double tr = 0.0;
		for (int i = 0; i < 12; ++i) tr += M[13 * i];
		for (int i = 0; i < 12; ++i) M[13 * i] += 1e-14 * tr;
		// Cholesky M = L L^T, L over the lower triangle (the same operations, in the same order, as the oracle's llt)
		for (int j = 0; j < 12; ++j)
		{
			double d = M[12 * j + j];
			for (int k = 0; k < j; ++k) d -= M[12 * j + k] * M[12 * j + k];
			M[12 * j + j] = sqrt(d);
			for (int i = j + 1; i < 12; ++i)
			{
				double v = M[12 * i + j];
				for (int k = 0; k < j; ++k) v -= M[12 * i + k] * M[12 * j + k];
				M[12 * i + j] = v / M[12 * j + j];
			}
		}
		double x[12], y[12];
		for (int i = 0; i < 12; ++i) x[i] = 1.0;
		bool good = true;
		for (int it = 0; it < 8 && good; ++it)
		{
			// x <- -(L L^T)^-1 x, normalised (inverse iteration; the sign alternates as in the oracle's llt_solve_neg)
			for (int i = 0; i < 12; ++i)
			{
				double v = x[i];
				for (int k = 0; k < i; ++k) v -= M[12 * i + k] * y[k];
				y[i] = v / M[12 * i + i];
			}
			for (int i = 11; i >= 0; --i)
			{
				double v = y[i];
				for (int k = i + 1; k < 12; ++k) v -= M[12 * k + i] * x[k];
				x[i] = v / M[12 * i + i];
			}
			double nn = 0.0;
			for (int i = 0; i < 12; ++i) { x[i] = -x[i]; nn += x[i] * x[i]; }
			nn = sqrt(nn);
			if (!(nn > 0.0) || !isfinite(nn))
				good = false;
			else
				for (int i = 0; i < 12; ++i) x[i] = x[i] / nn;
		}
		if (good)
		{
			const double s = s_scale;
			double P[12];
			for (int r = 0; r < 3; ++r)
			{
				const double *pr = x + 4 * r;
				P[4 * r + 0] = pr[0] * s; P[4 * r + 1] = pr[1] * s; P[4 * r + 2] = pr[2] * s;
				P[4 * r + 3] = pr[3] - s * (pr[0] * s_c[0] + pr[1] * s_c[1] + pr[2] * s_c[2]);
			}
			double A[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
			double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
			if (det < 0)
			{
				for (int i = 0; i < 12; ++i) P[i] = -P[i];
				for (int i = 0; i < 9; ++i) A[i] = -A[i];
				det = -det;
			}
			if (det > 0.0 && isfinite(det))
			{
				const double scale = cbrt(det);
				double R[9];
				for (int i = 0; i < 9; ++i) R[i] = A[i] / scale;
				for (int it = 0; it < 10; ++it)
				{
					double Ri[9];
					inv3(R, Ri);
					for (int r = 0; r < 3; ++r)
						for (int c = 0; c < 3; ++c)
							A[3 * r + c] = 0.5 * (R[3 * r + c] + Ri[3 * c + r]);
					for (int i = 0; i < 9; ++i) R[i] = A[i];
				}
				bool finite = true;
				for (int i = 0; i < 9; ++i)
					if (!isfinite(R[i])) finite = false;
				if (finite)
				{
					rotmat_to_quat(R, s_pose.q);
					s_pose.t[0] = P[3] / scale; s_pose.t[1] = P[7] / scale; s_pose.t[2] = P[11] / scale;
					quat_to_rotmat(s_pose.q, s_R);
					s_state = 1;
					s_lambda = 1e-3;
					s_recompute = 1;
				}
			}
		}
	}
	__syncthreads();
	if (s_state == 0)
	{
		if (tid < 12) out[tid] = 0.0;
		if (tid == 0) a.ok[set] = 0;
		return;
	}
	// ---- lm_pnp_impl (bundle.cpp:23-99) ----
	{
		const double c0 = cost_pass(a, idx, set_offset, n, s_R, s_pose.t, red, tot);
		if (tid == 0) s_cost = c0;
		__syncthreads();
	}
	for (int iter = 0; iter < a.max_iterations; ++iter)
	{
		if (s_recompute)
		{
			// accumulate (jacobian_impl.h:63-159): 21 lower-triangle entries of JtJ + 6 of Jtr
			double acc[27];
#pragma unroll
			for (int j = 0; j < 27; ++j) acc[j] = 0.0;
			const double *R = s_R, *t = s_pose.t;
			for (int k = tid; k < n; k += PNP_THREADS)
			{
				const double *p = a.pts + 5 * (size_t)idx[k];
				const double X0 = p[2], X1 = p[3], X2 = p[4];
				const double Z0 = (R[0] * X0 + R[1] * X1 + R[2] * X2) + t[0], Z1 = (R[3] * X0 + R[4] * X1 + R[5] * X2) + t[1],
				             Z2 = (R[6] * X0 + R[7] * X1 + R[8] * X2) + t[2];
				const double z0 = Z0 / Z2, z1 = Z1 / Z2;
				if (Z2 < 0)
					continue;
				double r0 = z0 - p[0], r1 = z1 - p[1];
				const double r_squared = r0 * r0 + r1 * r1;
				const double weight = weight_of(a, idx, set_offset, k) * (r_squared < 1.0 ? 1.0 : 0.0);
				if (weight == 0.0)
					continue;
				const double iz = 1.0 / Z2;
				const double A0[3] = {1.0 * iz, 0.0 * iz, -z0 * iz}, A1[3] = {0.0 * iz, 1.0 * iz, -z1 * iz};
				double dZ0[3], dZ1[3];
#pragma unroll
				for (int c = 0; c < 3; ++c)
				{
					dZ0[c] = A0[0] * R[c] + A0[1] * R[3 + c] + A0[2] * R[6 + c];
					dZ1[c] = A1[0] * R[c] + A1[1] * R[3 + c] + A1[2] * R[6 + c];
				}
				const double d00 = weight * (dZ0[0] * dZ0[0] + dZ1[0] * dZ1[0]), d10 = weight * (dZ0[1] * dZ0[0] + dZ1[1] * dZ1[0]),
				             d11 = weight * (dZ0[1] * dZ0[1] + dZ1[1] * dZ1[1]), d20 = weight * (dZ0[2] * dZ0[0] + dZ1[2] * dZ1[0]),
				             d21 = weight * (dZ0[2] * dZ0[1] + dZ1[2] * dZ1[1]), d22 = weight * (dZ0[2] * dZ0[2] + dZ1[2] * dZ1[2]);
				// row-major lower triangle: (0,0) (1,0) (1,1) (2,0) (2,1) (2,2) (3,0) ... (5,5)
				acc[0] += X2 * (X2 * d11 - X1 * d21) + X1 * (X1 * d22 - X2 * d21);
				acc[1] += -X2 * (X2 * d10 - X0 * d21) - X1 * (X0 * d22 - X2 * d20);
				acc[2] += X2 * (X2 * d00 - X0 * d20) + X0 * (X0 * d22 - X2 * d20);
				acc[3] += X1 * (X0 * d21 - X1 * d20) - X2 * (X0 * d11 - X1 * d10);
				acc[4] += -X2 * (X1 * d00 - X0 * d10) - X0 * (X0 * d21 - X1 * d20);
				acc[5] += X1 * (X1 * d00 - X0 * d10) + X0 * (X0 * d11 - X1 * d10);
				acc[6] += X1 * d20 - X2 * d10;
				acc[7] += X2 * d00 - X0 * d20;
				acc[8] += X0 * d10 - X1 * d00;
				acc[9] += d00;
				acc[10] += X1 * d21 - X2 * d11;
				acc[11] += X2 * d10 - X0 * d21;
				acc[12] += X0 * d11 - X1 * d10;
				acc[13] += d10;
				acc[14] += d11;
				acc[15] += X1 * d22 - X2 * d21;
				acc[16] += X2 * d20 - X0 * d22;
				acc[17] += X0 * d21 - X1 * d20;
				acc[18] += d20;
				acc[19] += d21;
				acc[20] += d22;
				r0 *= weight;
				r1 *= weight;
				acc[21] += (r0 * (X1 * dZ0[2] - X2 * dZ0[1]) + r1 * (X1 * dZ1[2] - X2 * dZ1[1]));
				acc[22] += (-r0 * (X0 * dZ0[2] - X2 * dZ0[0]) - r1 * (X0 * dZ1[2] - X2 * dZ1[0]));
				acc[23] += (r0 * (X0 * dZ0[1] - X1 * dZ0[0]) + r1 * (X0 * dZ1[1] - X1 * dZ1[0]));
				acc[24] += (dZ0[0] * r0 + dZ1[0] * r1);
				acc[25] += (dZ0[1] * r0 + dZ1[1] * r1);
				acc[26] += (dZ0[2] * r0 + dZ1[2] * r1);
			}
			block_sum<27>(acc, red, tot);
			if (tid == 0)
			{
				int q = 0;
				for (int i = 0; i < 6; ++i)
					for (int j = 0; j <= i; ++j)
						s_JtJ[6 * i + j] = tot[q++];
				double g = 0.0;
				for (int i = 0; i < 6; ++i)
				{
					s_Jtr[i] = tot[21 + i];
					g += s_Jtr[i] * s_Jtr[i];
				}
				if (sqrt(g) < 1e-8)
					s_state = 2;
			}
			__syncthreads();
			if (s_state == 2)
				break;
		}
		if (tid == 0)
		{
			for (int k = 0; k < 6; ++k) s_JtJ[6 * k + k] += s_lambda;
			double sol[6];
			llt_solve_neg<6>(s_JtJ, s_Jtr, sol);
			double sn = 0.0;
			for (int i = 0; i < 6; ++i) sn += sol[i] * sol[i];
			if (sqrt(sn) < 1e-8)
				s_state = 2;
			else
			{
				const double theta = sqrt(sol[0] * sol[0] + sol[1] * sol[1] + sol[2] * sol[2]);
				const double w[3] = {sol[0] / theta, sol[1] / theta, sol[2] / theta};
				const double sa = sin(theta), cb = cos(theta);
				const double sw[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
				double sw2[9], M[9], RM[9], Rn[9];
				mul3(sw, sw, sw2);
				for (int i = 0; i < 9; ++i) M[i] = sa * sw[i] + (1 - cb) * sw2[i];
				mul3(s_R, M, RM);
				for (int i = 0; i < 9; ++i) Rn[i] = s_R[i] + RM[i];
				rotmat_to_quat(Rn, s_new.q);
				for (int i = 0; i < 3; ++i)
					s_new.t[i] = s_pose.t[i] + (s_R[3 * i] * sol[3] + s_R[3 * i + 1] * sol[4] + s_R[3 * i + 2] * sol[5]);
				quat_to_rotmat(s_new.q, s_Rn);
#ifdef GCR_DEBUG_PNP
				printf("iter %d theta %g sol %g %g %g %g %g %g q %.6f %.6f %.6f %.6f -> %.6f %.6f %.6f %.6f  Rn0 %.6f R0 %.6f t %g -> %g\n", iter, theta, sol[0], sol[1], sol[2], sol[3], sol[4], sol[5],
				       s_pose.q[0], s_pose.q[1], s_pose.q[2], s_pose.q[3], s_new.q[0], s_new.q[1], s_new.q[2], s_new.q[3], s_Rn[0], s_R[0], s_pose.t[2], s_new.t[2]);
#endif
			}
		}
		__syncthreads();
		if (s_state == 2)
			break;
		const double cost_new = cost_pass(a, idx, set_offset, n, s_Rn, s_new.t, red, tot);
#ifdef GCR_DEBUG_PNP
		if (tid == 0) printf("set %d iter %d cost %.6g new %.6g lambda %g\n", set, iter, s_cost, cost_new, s_lambda);
#endif
		if (tid == 0)
		{
			if (cost_new < s_cost)
			{
				s_pose = s_new;
				for (int i = 0; i < 9; ++i) s_R[i] = s_Rn[i];
				s_lambda /= 10;
				s_cost = cost_new;
				s_recompute = 1;
			}
			else
			{
				for (int k = 0; k < 6; ++k) s_JtJ[6 * k + k] -= s_lambda;
				s_lambda *= 10;
				s_recompute = 0;
			}
		}
		__syncthreads();
	}
	if (tid < 12)
	{
		const int r = tid / 4, c = tid % 4;
		out[tid] = c < 3 ? s_R[3 * r + c] : s_pose.t[r];
	}
	if (tid == 0)
	{
		bool finite = true;
		for (int i = 0; i < 9; ++i)
			if (!isfinite(s_R[i])) finite = false;
		for (int i = 0; i < 3; ++i)
			if (!isfinite(s_pose.t[i])) finite = false;
		a.ok[set] = finite ? 1 : 0;
	}
}

// called by gcr_refit for GCR_PNP (k_refit.cu)
gcr_status gcr_refit_pnp_launch(gcr_ctx *ctx, int pair, const uint32_t *d_index_sets, const uint32_t *d_set_offsets,
                                const uint32_t *d_set_sizes, int count, const double *d_weights, int weight_mode, double *d_models,
                                uint8_t *d_ok)
{
	PnpRefitArgs a;
	a.pts = ctx->d_pts64 + (size_t)pair * ctx->n * 5;
	a.sets = d_index_sets;
	a.offsets = d_set_offsets;
	a.sizes = d_set_sizes;
	a.weights = d_weights;
	a.weight_mode = weight_mode;
	a.models = d_models;
	a.ok = d_ok;
	a.max_iterations = 25; // PnPBundleAdjustment's default (solver_pnp_bundle_adjustment.h:56)
	k_refit_pnp<<<count, PNP_THREADS, 0, ctx->stream>>>(a);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}
