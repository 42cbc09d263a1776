// tbk_tile.cuh -- per-mesh iterative sigma-clipped statistics (generic CTA-per-tile version).
//
// Restates astropy 5.1 SigmaClip(sigma=3, maxiters=5, cenfunc='median', stdfunc='std') as called
// by photutils 1.3.0 Background2D on one 64x64 mesh, followed by the nan-aware
// median / mean / std(ddof=0) of the surviving pixels (SExtractorBackground inputs); see
// photometry/backgrounds.py:105-106, 200-205 and SURVEY.md section 8a.
#pragma once
#include "tbk_common.cuh"

struct TileSmem {
	SelectSmem sel;
	RedSmem red;
};

// Thread -> pixel mapping inside a 64x64 tile for 256 threads x 16 pixels:
// element e = 4*j + q  (j = 0..3, q = 0..3) sits at row (tid/16 + 16*j), column (tid%16)*4 + q.
__device__ __forceinline__ int tile_lrow(int tid, int j) { return (tid >> 4) + 16 * j; }
__device__ __forceinline__ int tile_lcol(int tid) { return (tid & 15) << 2; }

// Sigma-clipped statistics of the valid elements of v[] (valid = bit e of ``valid``).
template <typename T>
__device__ TileStat tile_sigma_clip(const T (&v)[TBK_VPT], unsigned valid, TileSmem& sm)
{
	TileStat out;
	out.mean = out.med = out.std = nan_d();
	out.nfin = 0; out.pad = 0;

	// tight range of the valid data
	int cnt = 0; double mn = INFINITY, mx = -INFINITY;
#pragma unroll
	for (int e = 0; e < TBK_VPT; ++e) {
		if (valid >> e & 1u) { double d = (double)v[e]; ++cnt; mn = fmin(mn, d); mx = fmax(mx, d); }
	}
	block_sum_min_max(sm.red, cnt, mn, mx);
	if (cnt == 0) return out;

	double lo_run = -INFINITY, hi_run = INFINITY;   // running intersection = the clip buffer
	double lo_last = -INFINITY, hi_last = INFINITY; // most recent bounds
	bool nested_last = true;                        // last bounds lie inside the previous buffer range
	bool converged = false;
	double pivot = mn, med = mn;
	int n_prev = -1;

	for (int it = 0; it < 5; ++it) {
		// mean = sum / n, std = sqrt(sum((mean - v)^2) / n): the two passes of _fast_sigma_clip.c
		int n = 0; double s1 = 0.0, s2 = 0.0;
#pragma unroll
		for (int e = 0; e < TBK_VPT; ++e) {
			const double d = (double)v[e];
			if ((valid >> e & 1u) && d >= lo_run && d <= hi_run) { ++n; s1 += d - pivot; }
		}
		block_sum3(sm.red, n, s1, s2);
		if (it > 0 && n == n_prev) { converged = true; break; }
		const double mean_it = pivot + s1 / (double)n;
		int nd = 0; double ss = 0.0, dz = 0.0;
#pragma unroll
		for (int e = 0; e < TBK_VPT; ++e) {
			double d = (double)v[e];
			if ((valid >> e & 1u) && d >= lo_run && d <= hi_run) { d -= mean_it; ss += d * d; }
		}
		block_sum3(sm.red, nd, ss, dz);
		const double sd = sqrt(ss / (double)n);

		auto each = [&](auto f) {
#pragma unroll
			for (int e = 0; e < TBK_VPT; ++e) {
				double d = (double)v[e];
				if ((valid >> e & 1u) && d >= lo_run && d <= hi_run) f(d);
			}
		};
		const double a = fmax(lo_run, mn), b = fmin(hi_run, mx);
		const double q1 = block_select(sm.sel, sm.red, each, (n - 1) >> 1, a, b);
		double q2 = q1;
		if ((n & 1) == 0) {
			int cle = 0; double nxt = INFINITY, dummy = -INFINITY;
			each([&](double d) { if (d <= q1) ++cle; else nxt = fmin(nxt, d); });
			block_sum_min_max(sm.red, cle, nxt, dummy);
			q2 = (cle > (n >> 1)) ? q1 : nxt;
		}
		med = 0.5 * (q1 + q2);
		lo_last = med - 3.0 * sd;
		hi_last = med + 3.0 * sd;
		nested_last = (lo_last >= lo_run) && (hi_last <= hi_run);
		lo_run = fmax(lo_run, lo_last);
		hi_run = fmin(hi_run, hi_last);
		n_prev = n;
		pivot = med;
	}

	// Final set: the ORIGINAL valid values inside the last bounds (not the running intersection).
	int nf = 0; double s1 = 0.0, dummy = 0.0;
#pragma unroll
	for (int e = 0; e < TBK_VPT; ++e) {
		double d = (double)v[e];
		if ((valid >> e & 1u) && d >= lo_last && d <= hi_last) { ++nf; s1 += d - pivot; }
	}
	block_sum3(sm.red, nf, s1, dummy);
	out.nfin = nf;
	if (nf == 0) return out;
	const double mean = pivot + s1 / (double)nf;
	int ndum = 0; double ss = 0.0; dummy = 0.0;
#pragma unroll
	for (int e = 0; e < TBK_VPT; ++e) {
		double d = (double)v[e];
		if ((valid >> e & 1u) && d >= lo_last && d <= hi_last) { d -= mean; ss += d * d; }
	}
	block_sum3(sm.red, ndum, ss, dummy);
	out.mean = mean;
	out.std = sqrt(ss / (double)nf);
	if (converged && nested_last) {
		out.med = med;  // buffer == final set: the last median is the answer
	} else {
		auto eachf = [&](auto f) {
#pragma unroll
			for (int e = 0; e < TBK_VPT; ++e) {
				double d = (double)v[e];
				if ((valid >> e & 1u) && d >= lo_last && d <= hi_last) f(d);
			}
		};
		const double a = fmax(lo_last, mn), b = fmin(hi_last, mx);
		const double q1 = block_select(sm.sel, sm.red, eachf, (nf - 1) >> 1, a, b);
		double q2 = q1;
		if ((nf & 1) == 0) {
			int cle = 0; double nxt = INFINITY, dmy = -INFINITY;
			eachf([&](double d) { if (d <= q1) ++cle; else nxt = fmin(nxt, d); });
			block_sum_min_max(sm.red, cle, nxt, dmy);
			q2 = (cle > (nf >> 1)) ? q1 : nxt;
		}
		out.med = 0.5 * (q1 + q2);
	}
	return out;
}
