// Host build of the DEVICE solver headers (graph-cut-ransac_b200/csrc/solve_f7.cuh, solve_e5_group.cuh, solve_e_stewenius.cuh,
// solve_p3p.cuh): the CUDA
// qualifiers are defined away, the ten lanes of a 5-point group become ten threads and the group barrier a
// pthread barrier.  Lets the CPU-only test tier check the kernels' arithmetic against the oracle without a GPU
// (tests/test_device_solvers_on_host.py).  Test infrastructure, not part of the product.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#define __device__
#define __constant__
#define __forceinline__ inline
using std::isfinite;
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
#include "solve_f7.cuh"
#include "solve_e5_group.cuh"
#include "solve_e_stewenius.cuh"
#include "solve_p3p.cuh"

namespace
{
	struct Pt4
	{
		const double *P;
		const uint32_t *s;
		double operator()(int i, int c) const { return P[(size_t)s[i] * 4 + c]; }
	};
	struct LaneSync
	{
		pthread_barrier_t *b;
		unsigned *scratch;
		int li;
		void operator()() const { pthread_barrier_wait(b); }
		bool any(bool x) const { return x; } // one group: its lanes agree
		unsigned ballot(bool x) const
		{
			scratch[li] = x ? 1u : 0u;
			pthread_barrier_wait(b);
			unsigned m = 0;
			for (int i = 0; i < 10; ++i) m |= scratch[i] << i;
			pthread_barrier_wait(b);
			return m;
		}
	};
	struct Job
	{
		int li;
		double *S;
		pthread_barrier_t *b;
		unsigned *scratch;
		double *out;
		unsigned char *valid;
		int ret;
	};
	void *lane_main(void *a)
	{
		Job *j = (Job *)a;
		LaneSync s{j->b, j->scratch, j->li};
		j->ret = gcr_e5g::solve_group(j->li, j->S, s, j->out, j->valid);
		return nullptr;
	}
}

// 7-point: models[3][9], valid[3] exactly as k_solve_f7 writes them
extern "C" int host_f7(const double *pts, const uint32_t *sidx, double *models, unsigned char *valid)
{
	double smem[63];
	Pt4 pt{pts, sidx};
	gcr_f7::SmemMat A{smem, 1};
	double f1[9], f2[9], roots[3];
	const int nr = gcr_f7::solve(pt, A, f1, f2, roots);
	for (int m = 0; m < 3; ++m)
	{
		double F[9];
		bool ok = m < nr && gcr_f7::model(f1, f2, roots[m], F);
		if (ok) ok = gcr_f7::orientation_valid(F, pt, 7);
		for (int j = 0; j < 9; ++j) models[m * 9 + j] = ok ? F[j] : 0.0;
		valid[m] = ok;
	}
	return nr;
}

// 5-point, ten lanes: pts5 = the five sample points (x0 y0 x1 y1); models[10][9], valid[10] as k_solve_e5 writes them
extern "C" int host_e5_group(const double *pts5, double *models, unsigned char *valid)
{
	double S[gcr_e5g::SAMPLE_DOUBLES];
	for (int i = 0; i < gcr_e5g::SAMPLE_DOUBLES; ++i) S[i] = NAN; // uninitialised shared memory
	memcpy(S + gcr_e5g::OFF_PT, pts5, 20 * sizeof(double));
	unsigned scratch[10];
	pthread_barrier_t b;
	pthread_barrier_init(&b, nullptr, 10);
	pthread_t th[10];
	Job jobs[10];
	for (int l = 0; l < 10; ++l)
	{
		jobs[l] = {l, S, &b, scratch, models, valid, 0};
		pthread_create(&th[l], nullptr, lane_main, &jobs[l]);
	}
	for (int l = 0; l < 10; ++l) pthread_join(th[l], nullptr);
	pthread_barrier_destroy(&b);
	for (int l = 1; l < 10; ++l)
		if (jobs[l].ret != jobs[0].ret) return -100; // the lanes must agree
	return jobs[0].ret;
}

// Stewenius non-minimal solver from a 4-D null-space basis N[4][9] (what k_refit_e hands it): E_out[10][9] un-normalised,
// returns the number of models
extern "C" int host_stewenius_from_basis(const double *N36, double *E90)
{
	double N[4][9];
	for (int b = 0; b < 4; ++b)
		for (int j = 0; j < 9; ++j)
			N[b][j] = N36[9 * b + j];
	return gcr_e_nm::solve_from_basis(N, E90);
}

// P3P (Lambda-Twist) on three 2D-3D matches, rows u v X Y Z: P48 = up to four 3 x 4 poses, returns their number
extern "C" int host_p3p(const double *pts5, const uint32_t *sidx, double *P48)
{
	struct Pt5
	{
		const double *P;
		const uint32_t *s;
		double operator()(int i, int c) const { return P[(size_t)s[i] * 5 + c]; }
	};
	const Pt5 pt{pts5, sidx};
	return gcr_p3p::solve(pt, P48);
}
