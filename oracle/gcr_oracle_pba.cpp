// oracle/gcr_oracle_pba.cpp -- TEST INFRASTRUCTURE ONLY (see gcr_oracle.h).
// The non-minimal absolute-pose fit behind DefaultPnPEstimator = PerspectiveNPointEstimator<P3PSolver, PnPBundleAdjustment>
// (types.h:135-138; estimators/solver_pnp_bundle_adjustment.h:108-228): an initial pose from the whole index set, then
// refine_pnp = lm_pnp_impl over CameraJacobianAccumulator with the identity camera and TruncatedLoss(1.0), <= 25 iterations
// (relative_pose/bundle.cpp:23-99, 558-600; relative_pose/jacobian_impl.h:8-160).
//
// Two stated deviations, because the reference has no defined behaviour to match here:
//  * its initial pose is OpenCV's cv::solvePnP(SOLVEPNP_EPNP) (:150-158), a dependency outside the tree.  Stated
//    replacement: the normalised DLT (Hartley & Zisserman alg. 7.1 on K-normalised image points) -- smallest eigenvector of
//    the 12 x 12 normal matrix by inverse iteration, P = s [R | t] split by a Newton polar decomposition -- for sets of at
//    least six points; smaller sets return no model.  Both initialisers land in the basin of the same reprojection minimum
//    on the inlier sets GC-RANSAC refits, and the LM below decides the result (tests pin it against cv2's EPnP + scipy).
//  * the accumulator as written loops over ALL rows of the matrix while indexing sample[i], multiplies by weights[i] where
//    the callers pass nullptr, and is constructed with the loss where its camera belongs (jacobian_impl.h:12-18, 32-37, 56;
//    bundle.cpp:589-594).  Restated with the evident intent: the sum runs over the index set, a missing weight vector is
//    all ones, weights (when given) are by position.
#include "gcr_oracle.h"
#include "linalg.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace
{
	// Eigen::Quaterniond(R).normalized() / toRotationMatrix(): see gcr_oracle_eba.cpp (same statements)
	void rotmat_to_quat(const double *R, double *q)
	{
		double t = R[0] + R[4] + R[8];
		double w, v[3];
		if (t > 0.0)
		{
			t = std::sqrt(t + 1.0);
			w = 0.5 * t;
			t = 0.5 / t;
			v[0] = (R[7] - R[5]) * t; v[1] = (R[2] - R[6]) * t; v[2] = (R[3] - R[1]) * t;
		}
		else
		{
			int i = 0;
			if (R[4] > R[0]) i = 1;
			if (R[8] > R[4 * i]) i = 2;
			const int j = (i + 1) % 3, k = (j + 1) % 3;
			t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
			v[i] = 0.5 * t;
			t = 0.5 / t;
			w = (R[3 * k + j] - R[3 * j + k]) * t;
			v[j] = (R[3 * j + i] + R[3 * i + j]) * t;
			v[k] = (R[3 * k + i] + R[3 * i + k]) * t;
		}
		const double nrm = std::sqrt(w * w + v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
		q[0] = w / nrm; q[1] = v[0] / nrm; q[2] = v[1] / nrm; q[3] = v[2] / nrm;
	}
	void quat_to_rotmat(const double *q, double *R)
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
	struct Pose { double q[4], t[3]; };

	// CameraJacobianAccumulator::residual (jacobian_impl.h:26-59), identity camera, TruncatedLoss(1.0)
	double pnp_cost(const double *pts, const uint32_t *sample, int n, const double *weights, const Pose &pose)
	{
		double R[9];
		quat_to_rotmat(pose.q, R);
		double cost = 0.0;
		for (int k = 0; k < n; ++k)
		{
			const double *p = pts + 5 * (size_t)sample[k];
			const double X0 = p[2], X1 = p[3], X2 = p[4];
			const double Z0 = (R[0] * X0 + R[1] * X1 + R[2] * X2) + pose.t[0], Z1 = (R[3] * X0 + R[4] * X1 + R[5] * X2) + pose.t[1],
			             Z2 = (R[6] * X0 + R[7] * X1 + R[8] * X2) + pose.t[2];
			if (Z2 < 0)
				continue;
			const double inv_z = 1.0 / Z2;
			const double r0 = Z0 * inv_z - p[0], r1 = Z1 * inv_z - p[1];
			const double r_squared = r0 * r0 + r1 * r1;
			const double loss = r_squared < 1.0 ? r_squared : 1.0;
			cost += (weights ? weights[k] : 1.0) * loss;
		}
		return cost;
	}
	// CameraJacobianAccumulator::accumulate (jacobian_impl.h:63-159): lower triangle of JtJ (6 x 6 row-major), Jtr
	void pnp_accumulate(const double *pts, const uint32_t *sample, int n, const double *weights, const Pose &pose, double *JtJ, double *Jtr)
	{
		double R[9];
		quat_to_rotmat(pose.q, R);
		for (int k = 0; k < n; ++k)
		{
			const double *p = pts + 5 * (size_t)sample[k];
			const double X0 = p[2], X1 = p[3], X2 = p[4];
			const double Z[3] = {(R[0] * X0 + R[1] * X1 + R[2] * X2) + pose.t[0], (R[3] * X0 + R[4] * X1 + R[5] * X2) + pose.t[1],
			                     (R[6] * X0 + R[7] * X1 + R[8] * X2) + pose.t[2]};
			const double z[2] = {Z[0] / Z[2], Z[1] / Z[2]};
			if (Z[2] < 0)
				continue;
			double r[2] = {z[0] - p[0], z[1] - p[1]};
			const double r_squared = r[0] * r[0] + r[1] * r[1];
			const double weight = (weights ? weights[k] : 1.0) * (r_squared < 1.0 ? 1.0 : 0.0);
			if (weight == 0.0)
				continue;
			// dZ = [I | -z] / Z2 * R  (2 x 3)
			const double iz = 1.0 / Z[2];
			const double A[2][3] = {{1.0 * iz, 0.0 * iz, -z[0] * iz}, {0.0 * iz, 1.0 * iz, -z[1] * iz}};
			double dZ[2][3];
			for (int a = 0; a < 2; ++a)
				for (int c = 0; c < 3; ++c)
					dZ[a][c] = A[a][0] * R[c] + A[a][1] * R[3 + c] + A[a][2] * R[6 + c];
			auto cd = [&](int i, int j) { return dZ[0][i] * dZ[0][j] + dZ[1][i] * dZ[1][j]; };
			const double d00 = weight * cd(0, 0), d10 = weight * cd(1, 0), d11 = weight * cd(1, 1), d20 = weight * cd(2, 0),
			             d21 = weight * cd(2, 1), d22 = weight * cd(2, 2);
			JtJ[0] += X2 * (X2 * d11 - X1 * d21) + X1 * (X1 * d22 - X2 * d21);
			JtJ[6] += -X2 * (X2 * d10 - X0 * d21) - X1 * (X0 * d22 - X2 * d20);
			JtJ[12] += X1 * (X0 * d21 - X1 * d20) - X2 * (X0 * d11 - X1 * d10);
			JtJ[18] += X1 * d20 - X2 * d10;
			JtJ[24] += X1 * d21 - X2 * d11;
			JtJ[30] += X1 * d22 - X2 * d21;
			JtJ[7] += X2 * (X2 * d00 - X0 * d20) + X0 * (X0 * d22 - X2 * d20);
			JtJ[13] += -X2 * (X1 * d00 - X0 * d10) - X0 * (X0 * d21 - X1 * d20);
			JtJ[19] += X2 * d00 - X0 * d20;
			JtJ[25] += X2 * d10 - X0 * d21;
			JtJ[31] += X2 * d20 - X0 * d22;
			JtJ[14] += X1 * (X1 * d00 - X0 * d10) + X0 * (X0 * d11 - X1 * d10);
			JtJ[20] += X0 * d10 - X1 * d00;
			JtJ[26] += X0 * d11 - X1 * d10;
			JtJ[32] += X0 * d21 - X1 * d20;
			JtJ[21] += d00;
			JtJ[27] += d10;
			JtJ[33] += d20;
			JtJ[28] += d11;
			JtJ[34] += d21;
			JtJ[35] += d22;
			r[0] *= weight;
			r[1] *= weight;
			Jtr[0] += (r[0] * (X1 * dZ[0][2] - X2 * dZ[0][1]) + r[1] * (X1 * dZ[1][2] - X2 * dZ[1][1]));
			Jtr[1] += (-r[0] * (X0 * dZ[0][2] - X2 * dZ[0][0]) - r[1] * (X0 * dZ[1][2] - X2 * dZ[1][0]));
			Jtr[2] += (r[0] * (X0 * dZ[0][1] - X1 * dZ[0][0]) + r[1] * (X0 * dZ[1][1] - X1 * dZ[1][0]));
			Jtr[3] += (dZ[0][0] * r[0] + dZ[1][0] * r[1]);
			Jtr[4] += (dZ[0][1] * r[0] + dZ[1][1] * r[1]);
			Jtr[5] += (dZ[0][2] * r[0] + dZ[1][2] * r[1]);
		}
	}
	template <int N>
	void llt_solve_neg(const double *A, const double *b, double *x)
	{
		double L[N * N] = {0};
		for (int j = 0; j < N; ++j)
		{
			double d = A[N * j + j];
			for (int k = 0; k < j; ++k) d -= L[N * j + k] * L[N * j + k];
			L[N * j + j] = std::sqrt(d);
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
	// lm_pnp_impl (bundle.cpp:23-99)
	int refine_pnp(const double *pts, const uint32_t *sample, int n, const double *weights, Pose *pose, int max_iterations)
	{
		double JtJ[36], Jtr[6];
		double lambda = 1e-3;
		double cost = pnp_cost(pts, sample, n, weights, *pose);
		bool recompute_jac = true;
		int iter;
		for (iter = 0; iter < max_iterations; ++iter)
		{
			if (recompute_jac)
			{
				for (int i = 0; i < 36; ++i) JtJ[i] = 0.0;
				for (int i = 0; i < 6; ++i) Jtr[i] = 0.0;
				pnp_accumulate(pts, sample, n, weights, *pose, JtJ, Jtr);
				double g = 0.0;
				for (int i = 0; i < 6; ++i) g += Jtr[i] * Jtr[i];
				if (std::sqrt(g) < 1e-8)
					break;
			}
			for (int k = 0; k < 6; ++k) JtJ[6 * k + k] += lambda;
			double sol[6];
			llt_solve_neg<6>(JtJ, Jtr, sol);
			double sn = 0.0;
			for (int i = 0; i < 6; ++i) sn += sol[i] * sol[i];
			if (std::sqrt(sn) < 1e-8)
				break;
			const double theta = std::sqrt(sol[0] * sol[0] + sol[1] * sol[1] + sol[2] * sol[2]);
			const double w[3] = {sol[0] / theta, sol[1] / theta, sol[2] / theta};
			const double a = std::sin(theta), b = std::cos(theta);
			const double sw[9] = {0.0, -w[2], w[1], w[2], 0.0, -w[0], -w[1], w[0], 0.0};
			double sw2[9], M[9], R[9], RM[9], Rn[9];
			gcro::mul3(sw, sw, sw2);
			for (int i = 0; i < 9; ++i) M[i] = a * sw[i] + (1 - b) * sw2[i];
			quat_to_rotmat(pose->q, R);
			gcro::mul3(R, M, RM);
			for (int i = 0; i < 9; ++i) Rn[i] = R[i] + RM[i];
			Pose pn;
			rotmat_to_quat(Rn, pn.q);
			for (int i = 0; i < 3; ++i)
				pn.t[i] = pose->t[i] + (R[3 * i] * sol[3] + R[3 * i + 1] * sol[4] + R[3 * i + 2] * sol[5]);
			const double cost_new = pnp_cost(pts, sample, n, weights, pn);
#ifdef GCR_DEBUG_PNP
			printf("oracle iter %d cost %.6g new %.6g lambda %g\n", iter, cost, cost_new, lambda);
#endif
			if (cost_new < cost)
			{
				*pose = pn;
				lambda /= 10;
				cost = cost_new;
				recompute_jac = true;
			}
			else
			{
				for (int k = 0; k < 6; ++k) JtJ[6 * k + k] -= lambda;
				lambda *= 10;
				recompute_jac = false;
			}
		}
		return iter;
	}

	// The stated initialiser: normalised DLT.  Rows (Xh, 0, -u Xh), (0, Xh, -v Xh) with Xh = (s (X - c), 1); the normal
	// matrix is assembled from four 4 x 4 moment blocks; its smallest eigenvector by inverse iteration on M + eps I
	// (Cholesky, 8 iterations from the all-ones vector); P = P~ T; sign so that det(P[:, :3]) > 0; P[:, :3] / cbrt(det)
	// -> nearest rotation by the Newton polar iteration R <- (R + R^-T) / 2 (10 steps); t = P[:, 3] / cbrt(det).
	bool dlt_pose(const double *pts, const uint32_t *sample, int n, const double *weights, Pose *pose)
	{
		double c[3] = {0, 0, 0};
		for (int k = 0; k < n; ++k)
			for (int d = 0; d < 3; ++d) c[d] += pts[5 * (size_t)sample[k] + 2 + d];
		for (int d = 0; d < 3; ++d) c[d] /= n;
		double spread = 0.0;
		for (int k = 0; k < n; ++k)
		{
			double dd = 0.0;
			for (int d = 0; d < 3; ++d)
			{
				const double v = pts[5 * (size_t)sample[k] + 2 + d] - c[d];
				dd += v * v;
			}
			spread += std::sqrt(dd);
		}
		spread /= n;
		if (!(spread > 0.0))
			return false;
		const double s = std::sqrt(3.0) / spread;
		double S0[16] = {0}, Su[16] = {0}, Sv[16] = {0}, Sq[16] = {0};
		for (int k = 0; k < n; ++k)
		{
			const double *p = pts + 5 * (size_t)sample[k];
			const double w = weights ? weights[k] : 1.0;
			const double Xh[4] = {s * (p[2] - c[0]), s * (p[3] - c[1]), s * (p[4] - c[2]), 1.0};
			const double u = p[0], v = p[1];
			for (int a = 0; a < 4; ++a)
				for (int b = 0; b < 4; ++b)
				{
					const double ab = w * (Xh[a] * Xh[b]);
					S0[4 * a + b] += ab;
					Su[4 * a + b] += u * ab;
					Sv[4 * a + b] += v * ab;
					Sq[4 * a + b] += (u * u + v * v) * ab;
				}
		}
		double M[144];
		for (int a = 0; a < 4; ++a)
			for (int b = 0; b < 4; ++b)
			{
				M[12 * a + b] = S0[4 * a + b];
				M[12 * a + 4 + b] = 0.0;
				M[12 * a + 8 + b] = -Su[4 * a + b];
				M[12 * (4 + a) + b] = 0.0;
				M[12 * (4 + a) + 4 + b] = S0[4 * a + b];
				M[12 * (4 + a) + 8 + b] = -Sv[4 * a + b];
				M[12 * (8 + a) + b] = -Su[4 * a + b];
				M[12 * (8 + a) + 4 + b] = -Sv[4 * a + b];
				M[12 * (8 + a) + 8 + b] = Sq[4 * a + b];
			}
		double tr = 0.0;
		for (int i = 0; i < 12; ++i) tr += M[13 * i];
		for (int i = 0; i < 12; ++i) M[13 * i] += 1e-14 * tr;
		double x[12], b[12], y[12];
		for (int i = 0; i < 12; ++i) x[i] = 1.0;
		for (int it = 0; it < 8; ++it)
		{
			for (int i = 0; i < 12; ++i) b[i] = x[i];
			llt_solve_neg<12>(M, b, y);
			double nn = 0.0;
			for (int i = 0; i < 12; ++i) nn += y[i] * y[i];
			nn = std::sqrt(nn);
			if (!(nn > 0.0) || !std::isfinite(nn))
				return false;
			for (int i = 0; i < 12; ++i) x[i] = y[i] / nn;
		}
		// P = P~ T, T = [s I | -s c; 0 0 0 1]
		double P[12];
		for (int r = 0; r < 3; ++r)
		{
			const double *pr = x + 4 * r;
			P[4 * r + 0] = pr[0] * s; P[4 * r + 1] = pr[1] * s; P[4 * r + 2] = pr[2] * s;
			P[4 * r + 3] = pr[3] - s * (pr[0] * c[0] + pr[1] * c[1] + pr[2] * c[2]);
		}
		double A[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
		double det = gcro::det3(A);
		if (det < 0)
		{
			for (int i = 0; i < 12; ++i) P[i] = -P[i];
			for (int i = 0; i < 9; ++i) A[i] = -A[i];
			det = -det;
		}
		if (!(det > 0.0) || !std::isfinite(det))
			return false;
		const double scale = std::cbrt(det);
		double R[9];
		for (int i = 0; i < 9; ++i) R[i] = A[i] / scale;
		for (int it = 0; it < 10; ++it)
		{
			double Ri[9], Rit[9];
			gcro::inv3(R, Ri);
			gcro::transpose3(Ri, Rit);
			for (int i = 0; i < 9; ++i) R[i] = 0.5 * (R[i] + Rit[i]);
		}
		for (int i = 0; i < 9; ++i)
			if (!std::isfinite(R[i])) return false;
		rotmat_to_quat(R, pose->q);
		pose->t[0] = P[3] / scale; pose->t[1] = P[7] / scale; pose->t[2] = P[11] / scale;
		return true;
	}
}

extern "C" {

// refine_pnp from a given pose (test hook): P12 = [R | t] row-major in / out; returns the LM iteration count
int gcro_refine_pnp_lm(const double *pts5, const uint32_t *sample, int n, const double *weights, double *P12)
{
	Pose pose;
	const double R[9] = {P12[0], P12[1], P12[2], P12[4], P12[5], P12[6], P12[8], P12[9], P12[10]};
	rotmat_to_quat(R, pose.q);
	pose.t[0] = P12[3]; pose.t[1] = P12[7]; pose.t[2] = P12[11];
	const int it = refine_pnp(pts5, sample, n, weights, &pose, 25);
	double Rn[9];
	quat_to_rotmat(pose.q, Rn);
	for (int r = 0; r < 3; ++r)
	{
		P12[4 * r] = Rn[3 * r]; P12[4 * r + 1] = Rn[3 * r + 1]; P12[4 * r + 2] = Rn[3 * r + 2];
		P12[4 * r + 3] = pose.t[r];
	}
	return it;
}

// PnPBundleAdjustment::estimateModel with the stated initialiser: DLT (n >= 6) + LM.  weights: by position, or null.
int gcro_refit_pnp(const double *pts5, const uint32_t *sample, int n, const double *weights, double *P12)
{
	if (n < 6)
		return 0;
	Pose pose;
	if (!dlt_pose(pts5, sample, n, weights, &pose))
		return 0;
	refine_pnp(pts5, sample, n, weights, &pose, 25);
	double R[9];
	quat_to_rotmat(pose.q, R);
	for (int r = 0; r < 3; ++r)
	{
		P12[4 * r] = R[3 * r]; P12[4 * r + 1] = R[3 * r + 1]; P12[4 * r + 2] = R[3 * r + 2];
		P12[4 * r + 3] = pose.t[r];
	}
	for (int i = 0; i < 12; ++i)
		if (!std::isfinite(P12[i])) return 0;
	return 1;
}

}
