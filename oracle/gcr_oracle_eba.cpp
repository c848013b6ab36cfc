// oracle/gcr_oracle_eba.cpp -- TEST INFRASTRUCTURE ONLY (see gcr_oracle.h).
// The post-run relative-pose bundle adjustment of findEssentialMatrix_ (py/gcransac_python.cpp:2766-2821):
// EssentialMatrixBundleAdjustmentSolver::estimateModel (estimators/solver_essential_matrix_bundle_adjustment.h:113-186)
//   = Stewenius initialisation on the inliers -> motion_from_essential + cheirality on the first inlier
//     (relative_pose/essential.cpp:44-117) -> refine_relpose = lm_5dof_impl over RelativePoseJacobianAccumulator with
//     TruncatedLoss(1.0), <= 25 iterations (relative_pose/bundle.cpp:176-250, 412-452; jacobian_impl.h:307-460)
//     -> essential_from_motion.
// Pinned against an independent scipy least-squares solve of the same cost in tests/test_oracle_solvers.py and against the
// reference's own source (solver_essential_matrix_bundle_adjustment.h + relative_pose/*.cpp compiled verbatim over
// oracle/shim_solvers, tests/test_reference_solver_sources.py: same model sets to 1e-9).  Restated with plain arrays in the
// reference's operation order.
//
// One deliberate deviation.  The reference fills its weight vector BY POSITION in the inlier list
// (py/gcransac_python.cpp:2776-2786) and the LM accumulator reads it by position (jacobian_impl.h:342, 416), but the
// Stewenius initialiser reads weights[point index] (solver_essential_matrix_five_point_stewenius.h:143) -- past the end
// of the vector whenever an inlier's index exceeds the inlier count.  Restated with the evident intent: the weight of
// inlier k is weights[k] in both places.
#include "gcr_oracle.h"
#include "linalg.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace
{
	// Eigen::Quaterniond(R) (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other, 3, 3>), then normalize():
	// CameraPose::rotmat_to_quat (relative_pose/bundle.h:35-41).  q = (w, x, y, z); R row-major.
	void rotmat_to_quat(const double *R, double *q)
	{
		double t = R[0] + R[4] + R[8];
		double w, v[3];
		if (t > 0.0)
		{
			t = std::sqrt(t + 1.0);
			w = 0.5 * t;
			t = 0.5 / t;
			v[0] = (R[7] - R[5]) * t;
			v[1] = (R[2] - R[6]) * t;
			v[2] = (R[3] - R[1]) * t;
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
	// Eigen's QuaternionBase::toRotationMatrix
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

	// essential_from_motion (essential.cpp:34-39): E = [t]_x R
	void essential_from_pose(const Pose &p, double *E)
	{
		double R[9];
		quat_to_rotmat(p.q, R);
		const double tx[9] = {0.0, -p.t[2], p.t[1], p.t[2], 0.0, -p.t[0], -p.t[1], p.t[0], 0.0};
		gcro::mul3(tx, R, E);
	}
	// check_cheirality (essential.cpp:42-60), min_depth = 0
	bool cheirality(const Pose &p, const double *x1, const double *x2)
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
	void col(const double *M, int c, double *v) { v[0] = M[c]; v[1] = M[3 + c]; v[2] = M[6 + c]; }
	void set_col(double *M, int c, const double *v) { M[c] = v[0]; M[3 + c] = v[1]; M[6 + c] = v[2]; }
	double sqn(const double *v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }

	// motion_from_essential (essential.cpp:63-117): up to four poses, the ones that pass the cheirality test, in its order
	int motion_from_essential(const double *E, const double *x1, const double *x2, Pose *out)
	{
		double e0[3], e1[3], e2[3], u12[3], u13[3], u23[3];
		col(E, 0, e0); col(E, 1, e1); col(E, 2, e2);
		gcro::cross3(e0, e1, u12); gcro::cross3(e0, e2, u13); gcro::cross3(e1, e2, u23);
		const double n12 = sqn(u12), n13 = sqn(u13), n23 = sqn(u23);
		double UW[9], Vt[9], c1[3], c2[3];
		auto pick = [&](const double *ecol, const double *u, double nn) {
			const double ne = std::sqrt(sqn(ecol)), s = std::sqrt(nn);
			for (int i = 0; i < 3; ++i) { c1[i] = ecol[i] / ne; c2[i] = u[i] / s; }
		};
		if (n12 > n13)
		{
			if (n12 > n23) pick(e0, u12, n12);
			else pick(e1, u23, n23);
		}
		else
		{
			if (n13 > n23) pick(e0, u13, n13);
			else pick(e1, u23, n23);
		}
		double c0[3];
		gcro::cross3(c2, c1, c0);
		for (int i = 0; i < 3; ++i) c0[i] = -c0[i];
		set_col(UW, 0, c0); set_col(UW, 1, c1); set_col(UW, 2, c2);
		double r0[3], r1[3], r2[3];
		for (int c = 0; c < 3; ++c)
		{
			r0[c] = c1[0] * E[c] + c1[1] * E[3 + c] + c1[2] * E[6 + c];
			r1[c] = -(c0[0] * E[c] + c0[1] * E[3 + c] + c0[2] * E[6 + c]);
		}
		double nr = std::sqrt(sqn(r0));
		for (int c = 0; c < 3; ++c) r0[c] /= nr;
		const double d = r0[0] * r1[0] + r0[1] * r1[1] + r0[2] * r1[2];
		for (int c = 0; c < 3; ++c) r1[c] -= d * r0[c];
		nr = std::sqrt(sqn(r1));
		for (int c = 0; c < 3; ++c) r1[c] /= nr;
		gcro::cross3(r0, r1, r2);
		for (int c = 0; c < 3; ++c) { Vt[c] = r0[c]; Vt[3 + c] = r1[c]; Vt[6 + c] = r2[c]; }
		int count = 0;
		Pose pose;
		double R[9];
		gcro::mul3(UW, Vt, R);
		rotmat_to_quat(R, pose.q);
		for (int i = 0; i < 3; ++i) pose.t[i] = c2[i];
		if (cheirality(pose, x1, x2)) out[count++] = pose;
		for (int i = 0; i < 3; ++i) pose.t[i] = -pose.t[i];
		if (cheirality(pose, x1, x2)) out[count++] = pose;
		for (int r = 0; r < 3; ++r) { UW[3 * r] = -UW[3 * r]; UW[3 * r + 1] = -UW[3 * r + 1]; }
		gcro::mul3(UW, Vt, R);
		rotmat_to_quat(R, pose.q);
		if (cheirality(pose, x1, x2)) out[count++] = pose;
		for (int i = 0; i < 3; ++i) pose.t[i] = -pose.t[i];
		if (cheirality(pose, x1, x2)) out[count++] = pose;
		return count;
	}

	// RelativePoseJacobianAccumulator::residual (jacobian_impl.h:321-346), TruncatedLoss(1.0)
	double relpose_cost(const double *pts, const uint32_t *sample, int n, const double *weights, const Pose &pose)
	{
		double E[9];
		essential_from_pose(pose, E);
		double cost = 0.0;
		for (int k = 0; k < n; ++k)
		{
			const double *p = pts + 4 * (size_t)sample[k];
			const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
			const double Ex[3] = {E[0] * x1 + E[1] * y1 + E[2], E[3] * x1 + E[4] * y1 + E[5], E[6] * x1 + E[7] * y1 + E[8]};
			const double C = x2 * Ex[0] + y2 * Ex[1] + Ex[2];
			const double Etx0 = E[0] * x2 + E[3] * y2 + E[6], Etx1 = E[1] * x2 + E[4] * y2 + E[7];
			const double nJc_sq = (Ex[0] * Ex[0] + Ex[1] * Ex[1]) + (Etx0 * Etx0 + Etx1 * Etx1);
			const double r2 = (C * C) / nJc_sq;
			const double loss = r2 < 1.0 ? r2 : 1.0;
			cost += weights ? weights[k] * loss : loss;
		}
		return cost;
	}
	void cross_unit(const double *t, int axis, double *out)
	{
		double e[3] = {0, 0, 0};
		e[axis] = 1.0;
		gcro::cross3(t, e, out);
		const double nn = std::sqrt(sqn(out));
		for (int i = 0; i < 3; ++i) out[i] /= nn;
	}
	// RelativePoseJacobianAccumulator::accumulate (jacobian_impl.h:348-459): lower triangle of JtJ (5 x 5, row-major), Jtr,
	// tangent basis (3 x 2, columns b0 b1)
	void relpose_accumulate(const double *pts, const uint32_t *sample, int n, const double *weights, const Pose &pose, double *JtJ,
	                        double *Jtr, double *b0, double *b1)
	{
		const double *t = pose.t;
		if (std::fabs(t[0]) < std::fabs(t[1]))
			cross_unit(t, std::fabs(t[0]) < std::fabs(t[2]) ? 0 : 2, b0);
		else
			cross_unit(t, std::fabs(t[1]) < std::fabs(t[2]) ? 1 : 2, b0);
		gcro::cross3(b0, t, b1);
		{
			const double nn = std::sqrt(sqn(b1));
			for (int i = 0; i < 3; ++i) b1[i] /= nn;
		}
		double E[9], R[9];
		essential_from_pose(pose, E);
		quat_to_rotmat(pose.q, R);
		// dR (9 x 3): vec(E) is row-major here -- block (3 r .. 3 r + 2, k) of the reference = row r of E, see :376-385
		double dR[9][3], dt[9][2];
		double Ec[3][3];
		for (int c = 0; c < 3; ++c) col(E, c, Ec[c]);
		for (int i = 0; i < 3; ++i)
		{
			dR[i][0] = 0.0;        dR[i][1] = -Ec[2][i];   dR[i][2] = Ec[1][i];
			dR[3 + i][0] = Ec[2][i]; dR[3 + i][1] = 0.0;     dR[3 + i][2] = -Ec[0][i];
			dR[6 + i][0] = -Ec[1][i]; dR[6 + i][1] = Ec[0][i]; dR[6 + i][2] = 0.0;
		}
		for (int c = 0; c < 3; ++c)
		{
			double Rc[3], x0[3], x1v[3];
			col(R, c, Rc);
			gcro::cross3(b0, Rc, x0);
			gcro::cross3(b1, Rc, x1v);
			for (int i = 0; i < 3; ++i) { dt[3 * c + i][0] = x0[i]; dt[3 * c + i][1] = x1v[i]; }
		}
		for (int k = 0; k < n; ++k)
		{
			const double *p = pts + 4 * (size_t)sample[k];
			const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
			const double Ex[3] = {E[0] * x1 + E[1] * y1 + E[2], E[3] * x1 + E[4] * y1 + E[5], E[6] * x1 + E[7] * y1 + E[8]};
			const double C = x2 * Ex[0] + y2 * Ex[1] + Ex[2];
			const double JC[4] = {E[0] * x2 + E[3] * y2 + E[6], E[1] * x2 + E[4] * y2 + E[7], Ex[0], Ex[1]};
			const double nJ_C = std::sqrt(JC[0] * JC[0] + JC[1] * JC[1] + JC[2] * JC[2] + JC[3] * JC[3]);
			const double inv_nJ_C = 1.0 / nJ_C;
			const double r = C * inv_nJ_C;
			double weight = ((r * r < 1.0) ? 1.0 : 0.0) / n;
			if (weights) weight = weights[k] * weight;
			if (weight == 0.0)
				continue;
			// dF in the reference's vec order (:422-433): (x1 x2, x1 y2, x1, y1 x2, y1 y2, y1, x2, y2, 1) = column-major vec(E)
			double dF[9] = {x1 * x2, x1 * y2, x1, y1 * x2, y1 * y2, y1, x2, y2, 1.0};
			const double s = C * inv_nJ_C * inv_nJ_C;
			dF[0] -= s * (JC[2] * x1 + JC[0] * x2);
			dF[1] -= s * (JC[3] * x1 + JC[0] * y2);
			dF[2] -= s * (JC[0]);
			dF[3] -= s * (JC[2] * y1 + JC[1] * x2);
			dF[4] -= s * (JC[3] * y1 + JC[1] * y2);
			dF[5] -= s * (JC[1]);
			dF[6] -= s * (JC[2]);
			dF[7] -= s * (JC[3]);
			for (int i = 0; i < 9; ++i) dF[i] *= inv_nJ_C;
			double J[5] = {0, 0, 0, 0, 0};
			for (int i = 0; i < 9; ++i)
			{
				J[0] += dF[i] * dR[i][0]; J[1] += dF[i] * dR[i][1]; J[2] += dF[i] * dR[i][2];
				J[3] += dF[i] * dt[i][0]; J[4] += dF[i] * dt[i][1];
			}
			for (int i = 0; i < 5; ++i)
				Jtr[i] += weight * C * inv_nJ_C * J[i];
			for (int i = 0; i < 5; ++i)
				for (int j = 0; j <= i; ++j)
					JtJ[5 * i + j] += weight * (J[i] * J[j]);
		}
	}
	// -(L L^T)^-1 b from the lower triangle (selfadjointView<Lower>().llt().solve)
	void llt_solve5_neg(const double *A, const double *b, double *x)
	{
		double L[25] = {0};
		for (int j = 0; j < 5; ++j)
		{
			double d = A[5 * j + j];
			for (int k = 0; k < j; ++k) d -= L[5 * j + k] * L[5 * j + k];
			L[5 * j + j] = std::sqrt(d);
			for (int i = j + 1; i < 5; ++i)
			{
				double v = A[5 * i + j];
				for (int k = 0; k < j; ++k) v -= L[5 * i + k] * L[5 * j + k];
				L[5 * i + j] = v / L[5 * j + j];
			}
		}
		double y[5];
		for (int i = 0; i < 5; ++i)
		{
			double v = b[i];
			for (int k = 0; k < i; ++k) v -= L[5 * i + k] * y[k];
			y[i] = v / L[5 * i + i];
		}
		for (int i = 4; i >= 0; --i)
		{
			double v = y[i];
			for (int k = i + 1; k < 5; ++k) v -= L[5 * k + i] * x[k];
			x[i] = v / L[5 * i + i];
		}
		for (int i = 0; i < 5; ++i) x[i] = -x[i];
	}
	// lm_5dof_impl (bundle.cpp:176-250)
	int refine_relpose(const double *pts, const uint32_t *sample, int n, const double *weights, Pose *pose, int max_iterations)
	{
		double JtJ[25], Jtr[5], b0[3], b1[3];
		double lambda = 1e-3;
		double cost = relpose_cost(pts, sample, n, weights, *pose);
		bool recompute_jac = true;
		int iter;
		for (iter = 0; iter < max_iterations; ++iter)
		{
			if (recompute_jac)
			{
				for (int i = 0; i < 25; ++i) JtJ[i] = 0.0;
				for (int i = 0; i < 5; ++i) Jtr[i] = 0.0;
				relpose_accumulate(pts, sample, n, weights, *pose, JtJ, Jtr, b0, b1);
				double g = 0.0;
				for (int i = 0; i < 5; ++i) g += Jtr[i] * Jtr[i];
				if (std::sqrt(g) < 1e-8)
					break;
			}
			for (int k = 0; k < 5; ++k) JtJ[5 * k + k] += lambda;
			double sol[5];
			llt_solve5_neg(JtJ, Jtr, sol);
			double sn = 0.0;
			for (int i = 0; i < 5; ++i) sn += sol[i] * sol[i];
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
				pn.t[i] = pose->t[i] + (b0[i] * sol[3] + b1[i] * sol[4]);
			const double cost_new = relpose_cost(pts, sample, n, weights, pn);
			if (cost_new < cost)
			{
				*pose = pn;
				lambda /= 10;
				cost = cost_new;
				recompute_jac = true;
			}
			else
			{
				for (int k = 0; k < 5; ++k) JtJ[5 * k + k] -= lambda;
				lambda *= 10;
				recompute_jac = false;
			}
		}
		return iter;
	}
}

extern "C" {

// LM refinement of one pose from an essential matrix + the cheirality point (test hook): E9 in, refined E9 out per
// surviving pose; returns their number (<= 4)
int gcro_refine_e_lm(const double *pts, const uint32_t *sample, int n, const double *weights, const double *E9, double *out36)
{
	const double *p = pts + 4 * (size_t)sample[0];
	double x1[3] = {p[0], p[1], 1.0}, x2[3] = {p[2], p[3], 1.0};
	const double n1 = std::sqrt(sqn(x1)), n2 = std::sqrt(sqn(x2));
	for (int i = 0; i < 3; ++i) { x1[i] /= n1; x2[i] /= n2; }
	Pose poses[4];
	const int np = motion_from_essential(E9, x1, x2, poses);
	for (int q = 0; q < np; ++q)
	{
		refine_relpose(pts, sample, n, weights, &poses[q], 25);
		essential_from_pose(poses[q], out36 + 9 * q);
	}
	return np;
}

// EssentialMatrixBundleAdjustmentSolver::estimateModel on the inliers (the post-run polish of findEssentialMatrix_):
// weights[k] belongs to inliers[k]; n_points = rows of pts.  Initial models: the estimator's non-minimal fit
// (gcro_refit_e: Stewenius, unit Frobenius norm, E[8] >= 0).  out: up to 40 models; returns their number.
int gcro_polish_e_ba(const double *pts, int n_points, const uint32_t *inliers, int n, const double *weights, double *out360)
{
	if (n < 6) // sampleSize() of the BA solver
		return 0;
	std::vector<double> by_point;
	if (weights)
	{
		by_point.assign((size_t)n_points, 0.0);
		for (int k = 0; k < n; ++k)
			by_point[inliers[k]] = weights[k];
	}
	double init[90];
	const int n_init = gcro_refit_e(pts, inliers, n, weights ? by_point.data() : nullptr, init);
	int count = 0;
	for (int m = 0; m < n_init; ++m)
		count += gcro_refine_e_lm(pts, inliers, n, weights, init + 9 * m, out360 + 9 * count);
	return count;
}

}
