// k_refit_eba.cu -- K6: the relative-pose bundle adjustment behind findEssentialMatrix_'s post-run polish
// (py/gcransac_python.cpp:2766-2821): EssentialMatrixBundleAdjustmentSolver::estimateModel
// (estimators/solver_essential_matrix_bundle_adjustment.h:113-186) on one index set --
//   initial models (the Stewenius non-minimal fit, gcr_refit on GCR_E: up to 10)
//   -> motion_from_essential + cheirality on the set's first point (relative_pose/essential.cpp:44-117: up to 4 poses)
//   -> refine_relpose: lm_5dof_impl over RelativePoseJacobianAccumulator, TruncatedLoss(1.0), <= 25 iterations
//      (relative_pose/bundle.cpp:176-250, 412-452; relative_pose/jacobian_impl.h:307-460)
//   -> essential_from_motion (essential.cpp:34-39).
// One CTA per (initial model, pose) candidate: the per-point cost / Jacobian passes are block-parallel with a fixed
// reduction order (deterministic), the 5 x 5 Cholesky solve and the pose update run on thread 0.  FP64, -fmad=false: the
// arithmetic of oracle/gcr_oracle_eba.cpp statement by statement (the sums over the points associate differently).
// Weights are read BY POSITION in the set, in the accumulator (as the reference does) and in the initialiser (where the
// reference reads by point index, out of bounds: see the oracle's header for the deviation).
#include "gcr_internal.cuh"

#include <math.h>

#define EBA_THREADS 256
#define EBA_WARPS (EBA_THREADS / 32)
#define EBA_INIT 10
#define EBA_POSES 4

namespace
{
struct EbaArgs
{
	const double *pts;     // [n_points][4] of the pair
	const uint32_t *set;   // [n]
	int n;
	const double *weights; // [n] by position, or null
	const double *init;    // [EBA_INIT][9]
	const uint8_t *init_ok;
	double *models;        // [EBA_INIT][EBA_POSES][9]
	uint8_t *ok;           // [EBA_INIT][EBA_POSES]
	int max_iterations;
};

struct Pose { double q[4], t[3]; };

__device__ void mul3(const double *A, const double *B, double *C)
{
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c)
			C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
__device__ void cross3(const double *a, const double *b, double *c)
{
	c[0] = a[1] * b[2] - a[2] * b[1];
	c[1] = a[2] * b[0] - a[0] * b[2];
	c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ double sqn(const double *v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }

// Eigen::Quaterniond(R) + normalize (CameraPose::rotmat_to_quat, relative_pose/bundle.h:35-41); q = (w, x, y, z)
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
__device__ void essential_from_pose(const Pose &p, double *E)
{
	double R[9];
	quat_to_rotmat(p.q, R);
	const double tx[9] = {0.0, -p.t[2], p.t[1], p.t[2], 0.0, -p.t[0], -p.t[1], p.t[0], 0.0};
	mul3(tx, R, E);
}
__device__ bool cheirality(const Pose &p, const double *x1, const double *x2)
{
	double R[9];
	quat_to_rotmat(p.q, R);
	const double Rx1[3] = {R[0] * x1[0] + R[1] * x1[1] + R[2] * x1[2], R[3] * x1[0] + R[4] * x1[1] + R[5] * x1[2],
	                       R[6] * x1[0] + R[7] * x1[1] + R[8] * x1[2]};
	const double a = -(Rx1[0] * x2[0] + Rx1[1] * x2[1] + Rx1[2] * x2[2]);
	const double b1 = -(Rx1[0] * p.t[0] + Rx1[1] * p.t[1] + Rx1[2] * p.t[2]);
	const double b2 = x2[0] * p.t[0] + x2[1] * p.t[1] + x2[2] * p.t[2];
	const double lambda1 = b1 - a * b2, lambda2 = -a * b1 + b2;
	const double min_depth = 0.0 * (1 - a * a);
	return lambda1 > min_depth && lambda2 > min_depth;
}
__device__ void colv(const double *M, int c, double *v) { v[0] = M[c]; v[1] = M[3 + c]; v[2] = M[6 + c]; }

// motion_from_essential (essential.cpp:63-117): the poses that pass the cheirality test, in the reference's order
__device__ int motion_from_essential(const double *E, const double *x1, const double *x2, Pose *out)
{
	double e0[3], e1[3], e2[3], u12[3], u13[3], u23[3];
	colv(E, 0, e0); colv(E, 1, e1); colv(E, 2, e2);
	cross3(e0, e1, u12); cross3(e0, e2, u13); cross3(e1, e2, u23);
	const double n12 = sqn(u12), n13 = sqn(u13), n23 = sqn(u23);
	const double *ecol, *u;
	double nn;
	if (n12 > n13)
	{
		if (n12 > n23) { ecol = e0; u = u12; nn = n12; }
		else { ecol = e1; u = u23; nn = n23; }
	}
	else
	{
		if (n13 > n23) { ecol = e0; u = u13; nn = n13; }
		else { ecol = e1; u = u23; nn = n23; }
	}
	double c0[3], c1[3], c2[3];
	{
		const double ne = sqrt(sqn(ecol)), s = sqrt(nn);
		for (int i = 0; i < 3; ++i) { c1[i] = ecol[i] / ne; c2[i] = u[i] / s; }
	}
	cross3(c2, c1, c0);
	for (int i = 0; i < 3; ++i) c0[i] = -c0[i];
	double UW[9], Vt[9], r0[3], r1[3], r2[3];
	for (int i = 0; i < 3; ++i) { UW[3 * i] = c0[i]; UW[3 * i + 1] = c1[i]; UW[3 * i + 2] = c2[i]; }
	for (int c = 0; c < 3; ++c)
	{
		r0[c] = c1[0] * E[c] + c1[1] * E[3 + c] + c1[2] * E[6 + c];
		r1[c] = -(c0[0] * E[c] + c0[1] * E[3 + c] + c0[2] * E[6 + c]);
	}
	double nr = sqrt(sqn(r0));
	for (int c = 0; c < 3; ++c) r0[c] /= nr;
	const double d = r0[0] * r1[0] + r0[1] * r1[1] + r0[2] * r1[2];
	for (int c = 0; c < 3; ++c) r1[c] -= d * r0[c];
	nr = sqrt(sqn(r1));
	for (int c = 0; c < 3; ++c) r1[c] /= nr;
	cross3(r0, r1, r2);
	for (int c = 0; c < 3; ++c) { Vt[c] = r0[c]; Vt[3 + c] = r1[c]; Vt[6 + c] = r2[c]; }
	int count = 0;
	Pose pose;
	double R[9];
	mul3(UW, Vt, R);
	rotmat_to_quat(R, pose.q);
	for (int i = 0; i < 3; ++i) pose.t[i] = c2[i];
	if (cheirality(pose, x1, x2)) out[count++] = pose;
	for (int i = 0; i < 3; ++i) pose.t[i] = -pose.t[i];
	if (cheirality(pose, x1, x2)) out[count++] = pose;
	for (int r = 0; r < 3; ++r) { UW[3 * r] = -UW[3 * r]; UW[3 * r + 1] = -UW[3 * r + 1]; }
	mul3(UW, Vt, R);
	rotmat_to_quat(R, pose.q);
	if (cheirality(pose, x1, x2)) out[count++] = pose;
	for (int i = 0; i < 3; ++i) pose.t[i] = -pose.t[i];
	if (cheirality(pose, x1, x2)) out[count++] = pose;
	return count;
}

// deterministic block sum of NV doubles per thread: warp shuffles, then the warps in order; result in tot[] (shared)
template <int NV>
__device__ void block_sum(double *v, double *red /* [EBA_WARPS][NV] */, double *tot /* [NV] */)
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
		for (int w = 0; w < EBA_WARPS; ++w)
			s += red[w * NV + threadIdx.x];
		tot[threadIdx.x] = s;
	}
	__syncthreads();
}

// RelativePoseJacobianAccumulator::residual (jacobian_impl.h:321-346) with TruncatedLoss(1.0); E in shared memory
__device__ double cost_pass(const EbaArgs &a, const double *E, double *red, double *tot)
{
	double c[1] = {0.0};
	for (int k = threadIdx.x; k < a.n; k += EBA_THREADS)
	{
		const double *p = a.pts + 4 * (size_t)a.set[k];
		const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
		const double Ex0 = E[0] * x1 + E[1] * y1 + E[2], Ex1 = E[3] * x1 + E[4] * y1 + E[5], Ex2 = E[6] * x1 + E[7] * y1 + E[8];
		const double C = x2 * Ex0 + y2 * Ex1 + Ex2;
		const double Etx0 = E[0] * x2 + E[3] * y2 + E[6], Etx1 = E[1] * x2 + E[4] * y2 + E[7];
		const double nJc_sq = (Ex0 * Ex0 + Ex1 * Ex1) + (Etx0 * Etx0 + Etx1 * Etx1);
		const double r2 = (C * C) / nJc_sq;
		const double loss = r2 < 1.0 ? r2 : 1.0;
		c[0] += a.weights ? a.weights[k] * loss : loss;
	}
	block_sum<1>(c, red, tot);
	return tot[0];
}
} // namespace

__global__ void __launch_bounds__(EBA_THREADS) k_refit_e_ba(const EbaArgs a)
{
	__shared__ double red[EBA_WARPS * 20], tot[20];
	__shared__ Pose s_pose, s_new;
	__shared__ double s_E[9], s_dR[27], s_dt[18], s_b0[3], s_b1[3], s_JtJ[25], s_Jtr[5];
	__shared__ int s_state; // 0 = no pose for this slot, 1 = iterate, 2 = done
	__shared__ double s_cost, s_lambda;
	__shared__ int s_recompute;
	const int m = blockIdx.x / EBA_POSES, slot = blockIdx.x % EBA_POSES, tid = threadIdx.x;
	double *out = a.models + (size_t)blockIdx.x * 9;
	if (tid == 0)
	{
		s_state = 0;
		if (a.n >= 6 && a.init_ok[m]) // sampleSize() of the BA solver is 6 (:84-87)
		{
			const double *p = a.pts + 4 * (size_t)a.set[0];
			double x1[3] = {p[0], p[1], 1.0}, x2[3] = {p[2], p[3], 1.0};
			const double n1 = sqrt(sqn(x1)), n2 = sqrt(sqn(x2));
			for (int i = 0; i < 3; ++i) { x1[i] /= n1; x2[i] /= n2; }
			Pose poses[EBA_POSES];
			const int np = motion_from_essential(a.init + 9 * m, x1, x2, poses);
			if (slot < np)
			{
				s_pose = poses[slot];
				s_state = 1;
				s_lambda = 1e-3;
				s_recompute = 1;
				essential_from_pose(s_pose, s_E);
			}
		}
	}
	__syncthreads();
	if (s_state == 0)
	{
		if (tid < 9) out[tid] = 0.0;
		if (tid == 0) a.ok[blockIdx.x] = 0;
		return;
	}
	{
		const double c0 = cost_pass(a, s_E, red, tot);
		if (tid == 0) s_cost = c0;
		__syncthreads();
	}
	for (int iter = 0; iter < a.max_iterations; ++iter)
	{
		if (s_recompute)
		{
			if (tid == 0)
			{
				// tangent basis of the translation, dE / d(rotation), dE / d(translation) (jacobian_impl.h:349-394)
				const double *t = s_pose.t;
				int axis;
				if (fabs(t[0]) < fabs(t[1])) axis = fabs(t[0]) < fabs(t[2]) ? 0 : 2;
				else axis = fabs(t[1]) < fabs(t[2]) ? 1 : 2;
				double e[3] = {0.0, 0.0, 0.0};
				e[axis] = 1.0;
				cross3(t, e, s_b0);
				double nn = sqrt(sqn(s_b0));
				for (int i = 0; i < 3; ++i) s_b0[i] /= nn;
				cross3(s_b0, t, s_b1);
				nn = sqrt(sqn(s_b1));
				for (int i = 0; i < 3; ++i) s_b1[i] /= nn;
				double R[9], Ec[3][3];
				quat_to_rotmat(s_pose.q, R);
				for (int c = 0; c < 3; ++c) colv(s_E, c, Ec[c]);
				for (int i = 0; i < 3; ++i)
				{
					s_dR[3 * i] = 0.0;              s_dR[3 * i + 1] = -Ec[2][i];      s_dR[3 * i + 2] = Ec[1][i];
					s_dR[3 * (3 + i)] = Ec[2][i];   s_dR[3 * (3 + i) + 1] = 0.0;      s_dR[3 * (3 + i) + 2] = -Ec[0][i];
					s_dR[3 * (6 + i)] = -Ec[1][i];  s_dR[3 * (6 + i) + 1] = Ec[0][i]; s_dR[3 * (6 + i) + 2] = 0.0;
				}
				for (int c = 0; c < 3; ++c)
				{
					double Rc[3], v0[3], v1[3];
					colv(R, c, Rc);
					cross3(s_b0, Rc, v0);
					cross3(s_b1, Rc, v1);
					for (int i = 0; i < 3; ++i) { s_dt[2 * (3 * c + i)] = v0[i]; s_dt[2 * (3 * c + i) + 1] = v1[i]; }
				}
			}
			__syncthreads();
			// accumulate (jacobian_impl.h:396-458): 15 lower-triangle entries of JtJ + 5 of Jtr per thread, block sum
			double acc[20];
#pragma unroll
			for (int j = 0; j < 20; ++j) acc[j] = 0.0;
			const double *E = s_E;
			for (int k = tid; k < a.n; k += EBA_THREADS)
			{
				const double *p = a.pts + 4 * (size_t)a.set[k];
				const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
				const double Ex0 = E[0] * x1 + E[1] * y1 + E[2], Ex1 = E[3] * x1 + E[4] * y1 + E[5], Ex2 = E[6] * x1 + E[7] * y1 + E[8];
				const double C = x2 * Ex0 + y2 * Ex1 + Ex2;
				const double JC0 = E[0] * x2 + E[3] * y2 + E[6], JC1 = E[1] * x2 + E[4] * y2 + E[7], JC2 = Ex0, JC3 = Ex1;
				const double nJ_C = sqrt(JC0 * JC0 + JC1 * JC1 + JC2 * JC2 + JC3 * JC3);
				const double inv_nJ_C = 1.0 / nJ_C;
				const double r = C * inv_nJ_C;
				double weight = ((r * r < 1.0) ? 1.0 : 0.0) / a.n;
				if (a.weights) weight = a.weights[k] * weight;
				if (weight == 0.0)
					continue;
				double dF[9] = {x1 * x2, x1 * y2, x1, y1 * x2, y1 * y2, y1, x2, y2, 1.0};
				const double s = C * inv_nJ_C * inv_nJ_C;
				dF[0] -= s * (JC2 * x1 + JC0 * x2);
				dF[1] -= s * (JC3 * x1 + JC0 * y2);
				dF[2] -= s * (JC0);
				dF[3] -= s * (JC2 * y1 + JC1 * x2);
				dF[4] -= s * (JC3 * y1 + JC1 * y2);
				dF[5] -= s * (JC1);
				dF[6] -= s * (JC2);
				dF[7] -= s * (JC3);
				double J[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
				for (int i = 0; i < 9; ++i)
				{
					const double f = dF[i] * inv_nJ_C;
					J[0] += f * s_dR[3 * i]; J[1] += f * s_dR[3 * i + 1]; J[2] += f * s_dR[3 * i + 2];
					J[3] += f * s_dt[2 * i]; J[4] += f * s_dt[2 * i + 1];
				}
				int q = 0;
#pragma unroll
				for (int i = 0; i < 5; ++i)
#pragma unroll
					for (int j = 0; j <= i; ++j)
						acc[q++] += weight * (J[i] * J[j]);
#pragma unroll
				for (int i = 0; i < 5; ++i)
					acc[15 + i] += weight * C * inv_nJ_C * J[i];
			}
			block_sum<20>(acc, red, tot);
			if (tid == 0)
			{
				int q = 0;
				for (int i = 0; i < 5; ++i)
					for (int j = 0; j <= i; ++j)
						s_JtJ[5 * i + j] = tot[q++];
				double g = 0.0;
				for (int i = 0; i < 5; ++i)
				{
					s_Jtr[i] = tot[15 + i];
					g += s_Jtr[i] * s_Jtr[i];
				}
				if (sqrt(g) < 1e-8) // gradient_tol
					s_state = 2;
			}
			__syncthreads();
			if (s_state == 2)
				break;
		}
		if (tid == 0)
		{
			for (int k = 0; k < 5; ++k) s_JtJ[5 * k + k] += s_lambda;
			// sol = -(L L^T)^-1 Jtr from the lower triangle (selfadjointView<Lower>().llt().solve)
			double L[25], y[5], sol[5];
			for (int i = 0; i < 25; ++i) L[i] = 0.0;
			for (int j = 0; j < 5; ++j)
			{
				double d = s_JtJ[5 * j + j];
				for (int k = 0; k < j; ++k) d -= L[5 * j + k] * L[5 * j + k];
				L[5 * j + j] = sqrt(d);
				for (int i = j + 1; i < 5; ++i)
				{
					double v = s_JtJ[5 * i + j];
					for (int k = 0; k < j; ++k) v -= L[5 * i + k] * L[5 * j + k];
					L[5 * i + j] = v / L[5 * j + j];
				}
			}
			for (int i = 0; i < 5; ++i)
			{
				double v = s_Jtr[i];
				for (int k = 0; k < i; ++k) v -= L[5 * i + k] * y[k];
				y[i] = v / L[5 * i + i];
			}
			for (int i = 4; i >= 0; --i)
			{
				double v = y[i];
				for (int k = i + 1; k < 5; ++k) v -= L[5 * k + i] * sol[k];
				sol[i] = v / L[5 * i + i];
			}
			double sn = 0.0;
			for (int i = 0; i < 5; ++i) { sol[i] = -sol[i]; sn += sol[i] * sol[i]; }
			if (sqrt(sn) < 1e-8) // step_tol
				s_state = 2;
			else
			{
				const double theta = sqrt(sol[0] * sol[0] + sol[1] * sol[1] + sol[2] * sol[2]);
				const double w[3] = {sol[0] / theta, sol[1] / theta, sol[2] / theta};
				const double sa = sin(theta), cb = cos(theta);
				const double sw[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
				double sw2[9], M[9], R[9], RM[9], Rn[9];
				mul3(sw, sw, sw2);
				for (int i = 0; i < 9; ++i) M[i] = sa * sw[i] + (1 - cb) * sw2[i];
				quat_to_rotmat(s_pose.q, R);
				mul3(R, M, RM);
				for (int i = 0; i < 9; ++i) Rn[i] = R[i] + RM[i];
				rotmat_to_quat(Rn, s_new.q);
				for (int i = 0; i < 3; ++i)
					s_new.t[i] = s_pose.t[i] + (s_b0[i] * sol[3] + s_b1[i] * sol[4]);
				essential_from_pose(s_new, red); // the candidate's E: red[0..8] is free between block sums
			}
		}
		__syncthreads();
		if (s_state == 2)
			break;
		__shared__ double s_En[9];
		if (tid < 9) s_En[tid] = red[tid];
		__syncthreads();
		const double cost_new = cost_pass(a, s_En, red, tot);
		if (tid == 0)
		{
			if (cost_new < s_cost)
			{
				s_pose = s_new;
				for (int i = 0; i < 9; ++i) s_E[i] = s_En[i];
				s_lambda /= 10;
				s_cost = cost_new;
				s_recompute = 1;
			}
			else
			{
				for (int k = 0; k < 5; ++k) s_JtJ[5 * k + k] -= s_lambda;
				s_lambda *= 10;
				s_recompute = 0;
			}
		}
		__syncthreads();
	}
	if (tid < 9) out[tid] = s_E[tid];
	if (tid == 0)
	{
		bool finite = true;
		for (int i = 0; i < 9; ++i)
			if (!isfinite(s_E[i])) finite = false;
		a.ok[blockIdx.x] = finite ? 1 : 0;
	}
}

extern "C" int gcr_max_e_ba_models(void) { return EBA_INIT * EBA_POSES; }

extern "C" gcr_status gcr_refit_e_ba(gcr_ctx *ctx, int pair, const uint32_t *d_index_set, uint32_t set_size, const double *d_weights,
                                     double *d_models, uint8_t *d_ok)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0 || ctx->cols != 4)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_refit_e_ba: no 4-column points uploaded");
	if (pair < 0 || pair >= ctx->n_pairs || !d_index_set || !d_models || !d_ok || set_size == 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit_e_ba: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	// initial models: the estimator's non-minimal fit on the set (Stewenius; weights by position)
	// scratch: offsets | sizes | init models | init ok
	gcr_status st = gcr_ensure_scratch(ctx, 4096);
	if (st)
		return st;
	char *base = (char *)ctx->d_scratch;
	uint32_t *d_off = (uint32_t *)base, *d_size = (uint32_t *)(base + 256);
	double *d_init = (double *)(base + 512);
	uint8_t *d_init_ok = (uint8_t *)(base + 512 + EBA_INIT * 9 * sizeof(double));
	const uint32_t h_meta[2] = {0u, set_size};
	st = gcr_memcpy_h2d(ctx, d_off, &h_meta[0], sizeof(uint32_t));
	if (!st) st = gcr_memcpy_h2d(ctx, d_size, &h_meta[1], sizeof(uint32_t));
	if (!st)
		st = gcr_refit(ctx, GCR_E, pair, d_index_set, d_off, d_size, 1, d_weights, d_weights ? GCR_WEIGHT_POSITIONAL : GCR_WEIGHT_NONE,
		               d_init, d_init_ok);
	if (st)
		return st;
	EbaArgs a;
	a.pts = ctx->d_pts64 + (size_t)pair * ctx->n * 4;
	a.set = d_index_set;
	a.n = (int)set_size;
	a.weights = d_weights;
	a.init = d_init;
	a.init_ok = d_init_ok;
	a.models = d_models;
	a.ok = d_ok;
	a.max_iterations = 25; // EssentialMatrixBundleAdjustmentSolver's default (:58)
	k_refit_e_ba<<<EBA_INIT * EBA_POSES, EBA_THREADS, 0, ctx->stream>>>(a);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}
