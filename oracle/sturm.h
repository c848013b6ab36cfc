// oracle/sturm.h -- TEST INFRASTRUCTURE ONLY.  Real roots of a real polynomial of degree <= 10 by Sturm-sequence
// isolation + safeguarded Newton, ascending order.
//
// Stands in for poselib::sturm::bisect_sturm<10> (maths/sturm_polynomial_solver.h:223-264 of the reference), which the
// Nister 5-point solver calls (estimators/solver_essential_matrix_five_point_nister.h:574).  Same mathematical
// contract (all real roots inside the Cauchy bound, each isolated by sign-change counts of a Sturm chain, then
// polished), written independently: explicit Euclidean remainder chain, explicit interval stack instead of recursion,
// bisection-safeguarded Newton instead of Ridders + Newton.  Differences from the reference are confined to
// ill-conditioned inputs (clustered / multiple roots), where the reference gives up at recursion depth 30.
#pragma once
#include <cmath>

namespace gcro_sturm
{
	constexpr int MAXD = 10;

	struct Chain
	{
		double p[MAXD + 1][MAXD + 1]; // p[k][i]: coefficient of x^i of the k-th polynomial
		int deg[MAXD + 1];
		int len;
	};

	inline double horner(const double *c, int deg, double x)
	{
		double v = c[deg];
		for (int i = deg - 1; i >= 0; --i)
			v = v * x + c[i];
		return v;
	}

	// p0 = f, p1 = f', p_{k+1} = -rem(p_{k-1}, p_k), every remainder scaled to unit max-magnitude (positive factor)
	inline void build(const double *f, int n, Chain &ch)
	{
		for (int i = 0; i <= n; ++i)
			ch.p[0][i] = f[i];
		ch.deg[0] = n;
		for (int i = 0; i < n; ++i)
			ch.p[1][i] = f[i + 1] * (double)(i + 1);
		ch.deg[1] = n - 1;
		ch.len = 2;
		while (ch.deg[ch.len - 1] > 0 && ch.len <= MAXD)
		{
			const double *a = ch.p[ch.len - 2];
			const double *b = ch.p[ch.len - 1];
			const int da = ch.deg[ch.len - 2], db = ch.deg[ch.len - 1];
			double r[MAXD + 1];
			for (int i = 0; i <= da; ++i)
				r[i] = a[i];
			for (int k = da; k >= db; --k) // long division, remainder left in r[0..db-1]
			{
				const double q = r[k] / b[db];
				for (int i = 0; i <= db; ++i)
					r[k - db + i] -= q * b[i];
				r[k] = 0.0;
			}
			int dr = db - 1;
			while (dr >= 0 && r[dr] == 0.0)
				--dr;
			if (dr < 0)
				break; // exact common factor: chain ends
			double mx = 0.0;
			for (int i = 0; i <= dr; ++i)
				mx = std::fmax(mx, std::fabs(r[i]));
			const double s = -1.0 / mx;
			for (int i = 0; i <= dr; ++i)
				ch.p[ch.len][i] = r[i] * s;
			ch.deg[ch.len] = dr;
			++ch.len;
		}
	}

	inline int sign_changes(const Chain &ch, double x)
	{
		int changes = 0, last = 0;
		for (int k = 0; k < ch.len; ++k)
		{
			const double v = horner(ch.p[k], ch.deg[k], x);
			const int s = v > 0.0 ? 1 : (v < 0.0 ? -1 : 0);
			if (s != 0)
			{
				if (last != 0 && s != last)
					++changes;
				last = s;
			}
		}
		return changes;
	}

	// one simple root of f in [lo, hi] with f(lo) f(hi) <= 0: Newton with bisection safeguard
	inline double polish(const double *f, const double *df, int n, double lo, double hi)
	{
		const double fl = horner(f, n, lo), fh = horner(f, n, hi);
		if (fl == 0.0) return lo;
		if (fh == 0.0) return hi;
		double xl, xh;
		if (fl < 0.0) { xl = lo; xh = hi; }
		else { xl = hi; xh = lo; }
		double rts = 0.5 * (lo + hi);
		double dxold = std::fabs(hi - lo), dx = dxold;
		double v = horner(f, n, rts), d = horner(df, n - 1, rts);
		for (int it = 0; it < 200; ++it)
		{
			if ((((rts - xh) * d - v) * ((rts - xl) * d - v) > 0.0) || (std::fabs(2.0 * v) > std::fabs(dxold * d)))
			{
				dxold = dx;
				dx = 0.5 * (xh - xl);
				rts = xl + dx;
				if (xl == rts) return rts;
			}
			else
			{
				dxold = dx;
				dx = v / d;
				const double t = rts;
				rts -= dx;
				if (t == rts) return rts;
			}
			if (std::fabs(dx) <= 4.440892098500626e-16 * std::fabs(rts)) return rts;
			v = horner(f, n, rts);
			d = horner(df, n - 1, rts);
			if (v < 0.0) xl = rts;
			else xh = rts;
		}
		return rts;
	}

	// coeffs ascending (c[0] + c[1] x + ... + c[n] x^n), c[n] != 0; roots ascending; returns their number
	inline int real_roots(const double *coeffs, int n, double *roots)
	{
		if (n < 1 || n > MAXD || coeffs[n] == 0.0)
			return 0;
		double f[MAXD + 1], df[MAXD + 1];
		const double inv = 1.0 / coeffs[n];
		double mx = 0.0;
		for (int i = 0; i < n; ++i)
		{
			f[i] = coeffs[i] * inv;
			mx = std::fmax(mx, std::fabs(f[i]));
		}
		f[n] = 1.0;
		for (int i = 0; i < n; ++i)
			df[i] = f[i + 1] * (double)(i + 1);
		const double R = 1.0 + mx; // Cauchy bound
		Chain ch;
		build(f, n, ch);
		struct Iv { double a, b; int sa, sb, depth; };
		Iv stack[160];
		int top = 0, n_roots = 0;
		const int s_lo = sign_changes(ch, -R), s_hi = sign_changes(ch, R);
		if (s_lo - s_hi <= 0)
			return 0;
		stack[top++] = {-R, R, s_lo, s_hi, 0};
		while (top > 0 && n_roots < n)
		{
			const Iv iv = stack[--top];
			const int cnt = iv.sa - iv.sb;
			if (cnt <= 0)
				continue;
			if (cnt == 1)
			{
				const double fa = horner(f, n, iv.a), fb = horner(f, n, iv.b);
				if ((fa < 0.0) != (fb < 0.0) || fa == 0.0 || fb == 0.0)
				{
					roots[n_roots++] = polish(f, df, n, iv.a, iv.b);
					continue;
				}
				// no sign change although the chain reports one root (rounding in the chain): keep bisecting
			}
			if (iv.depth >= 60 || !(iv.b - iv.a > 1e-15 * (std::fabs(iv.a) + std::fabs(iv.b))))
				continue; // clustered / multiple roots beyond double resolution: dropped (the reference stops at depth 30)
			const double c = 0.5 * (iv.a + iv.b);
			const int sc = sign_changes(ch, c);
			// push the right half first so the left half is processed first (ascending output)
			stack[top++] = {c, iv.b, sc, iv.sb, iv.depth + 1};
			stack[top++] = {iv.a, c, iv.sa, sc, iv.depth + 1};
		}
		return n_roots;
	}
}
