// tbk_tile_zone.cuh -- sigma-clipped mesh statistics without sorting the mesh ("zone" algorithm, one warp per mesh).
//
// Same result as tbk_tile_warp.cuh / tbk_tile.cuh (astropy 5.1 SigmaClip(3, maxiters=5, median / std) followed by the
// nan-aware median / mean / std of the survivors; photometry/backgrounds.py:105-106, 200-205) -- but only the elements
// that can matter individually are ever stored:
//   * a 64-element sample gives a rough centre m and width s of the sky distribution;
//   * every clip bound the iteration produces lies in the tails (beyond m -+ 2 s) unless the sample misjudged the
//     mesh, so the elements inside [A, B] = m -+ 2 s ("bulk") only enter through their count and their float64
//     moments about a pivot; the elements outside ("tails", a few per cent) are kept one by one;
//   * every median the iteration needs is the element of a known global rank R = below + (n - 1) / 2 with
//     (n - 1 - nB) / 2 <= R <= (n - 1 + nA) / 2 + 1 (nA, nB = tail counts).  Pass 1 counts the elements below m
//     exactly, which locates that rank interval in value to within a few dozen ranks; the elements of a "zone"
//     [ZL, ZH] around it (a few per cent again) are kept one by one, together with the exact number of elements
//     below ZL, so a rank inside the zone is resolved exactly from a small counting sort.
// Pass 1 (mask, minima, counts, bulk moments) streams the mesh from HBM, pass 2 (collect tails + zone) re-reads it from
// L2.  A clip bound that enters the bulk (the sample overestimated the width) re-plans the mesh around the exact statistics
// of the failing iteration: one more run for the raw-pixel kernel, the previous round's width for the residual kernel.
// Anything the two lists still cannot answer exactly -- a median rank outside the zone, a list overflow, a degenerate
// sample, a sparsely populated mesh -- sends the mesh to a queue that the bucketed kernels (tbk_tile_warp.cuh) work off
// afterwards.  The result of a mesh that does not fall back is exactly the reference's: the same membership decisions on
// the same float64 comparisons, exact medians, float64 moments.
#pragma once
#include "tbk_tile_warp.cuh"

#define ZN_TCAP 40        // tail elements per lane
#define ZN_ZCAP 32        // zone elements per lane
#define ZN_BINS 128       // counting-sort bins over the zone
#define ZN_CT 2.0         // bulk = sample centre -+ ZN_CT sample sigma
#define ZN_MIN_N 256      // sparser meshes go to the bucketed path

// key traits on top of TwF32 / TwF64: Raw = what pass 2 stores (float bits / double), K = ordered integer key
struct Zn32 : TwF32 {
	typedef uint32_t Raw;
	__device__ static __forceinline__ K key_of(Raw r) { return r; }
	__device__ static __forceinline__ double rval(Raw r) { return (double)__uint_as_float(r); }
	__device__ static __forceinline__ float offs(Raw r, Raw zl) { return (float)(r - zl); }
	static constexpr bool kStoreD = true;     // the clip sweeps keep value - pivot (float64) next to the 32-bit raw value
	__device__ static __forceinline__ Raw padraw() { return 0xFFFFFFFFu; }   // compares false against every range
	// inclusive value range [lo, hi] as raw thresholds: raw >= tlo && raw <= thi  <=>  lo <= value <= hi
	struct Range { Raw tlo, thi; };
	__device__ static __forceinline__ Range range(double lo, double hi)
	{
		Range r;
		r.tlo = key_ceil(lo);
		if (!key_floor(hi, r.thi)) { r.tlo = 1u; r.thi = 0u; }   // nothing is <= a negative bound
		return r;
	}
};
struct Zn64 : TwF64 {
	typedef double Raw;
	__device__ static __forceinline__ K key_of(Raw r) { return dkey(r); }
	__device__ static __forceinline__ double rval(Raw r) { return r; }
	__device__ static __forceinline__ float offs(Raw r, Raw zl) { return (float)(r - zl); }
	static constexpr bool kStoreD = false;
	__device__ static __forceinline__ Raw padraw() { return nan_d(); }
	struct Range { Raw tlo, thi; };
	__device__ static __forceinline__ Range range(double lo, double hi) { Range r; r.tlo = lo; r.thi = hi; return r; }
};

template <typename T, int TCAP = ZN_TCAP>
struct ZoneSmem {
	typename T::Raw tails[TCAP * 32];   // pass 2: per-lane lists [j * 32 + lane]; then dense
	typename T::K zone[ZN_ZCAP * 32];      // pass 2: per-lane lists of Raw (same size as K); then bin-sorted keys
	uint32_t cnt[ZN_BINS];                 // zone bin counts -> starts -> ends
};

// What the sample decides (warp-uniform).
struct ZonePlan {
	double mhat, shat, pivot;
	double A, B;          // bulk range in value
	bool ok;
};

// two sorted 32-element sample sets (invalid entries sort last as T::padkey) -> centre, width
template <typename T>
__device__ __forceinline__ ZonePlan zone_plan(typename T::K sa, typename T::K sb, int lane)
{
	ZonePlan zp;
	double med = 0.0, iqr = 0.0; int sets = 0;
#pragma unroll
	for (int t = 0; t < 2; ++t) {
		const typename T::K k = warp_bitonic32<typename T::K>(t ? sb : sa, lane);
		const int m = __popc(__ballot_sync(0xffffffffu, k != T::padkey()));
		if (m >= 8) {
			med += T::val(__shfl_sync(0xffffffffu, k, m >> 1));
			iqr += T::val(__shfl_sync(0xffffffffu, k, (3 * m) >> 2)) - T::val(__shfl_sync(0xffffffffu, k, m >> 2));
			++sets;
		}
	}
	zp.ok = sets > 0 && iqr > 0.0;
	zp.mhat = zp.ok ? med / (double)sets : 0.0;
	zp.shat = zp.ok ? (iqr / (double)sets) / 1.349 : 1.0;
	zp.pivot = T::pivot_of(zp.mhat);
	zp.A = zp.mhat - ZN_CT * zp.shat;
	zp.B = zp.mhat + ZN_CT * zp.shat;
	return zp;
}

// Zone [ZL, ZH] in value from the exact counts of pass 1: n valid, nA / nB in the tails, nM below the sample centre m and
// nC inside m -+ ZN_CW s.  nM says how many ranks the needed interval lies from m, nC gives the local density that turns
// ranks into a value offset; the margins cover the Poisson noise of that conversion and the curvature of the density.
#define ZN_CW 0.25
__device__ __forceinline__ void zone_range(const ZonePlan& zp, int n, int nA, int nB, int nM, int nC, double& ZL, double& ZH, bool nc_sampled)
{
	const int nbulk = n - nA - nB;
	const double r_lo = (double)((n - 1 - nB) >> 1), r_hi = (double)(((n - 1 + nA) >> 1) + 1);
	const bool local = nC >= 64;
	// elements per unit value near the centre: counted, or (sparse centre) a Gaussian of the sample width
	const double rho = local ? (double)nC / (2.0 * ZN_CW * zp.shat) : (double)max(nbulk, 1) * 0.418 / zp.shat;
	const double dlo = r_lo - (double)nM, dhi = r_hi - (double)nM;
	// relative margin: 0.10 covers the curvature of the density; an nC extrapolated from a quarter of the elements carries +-7 % (1 sigma) more
	const double m0 = local ? 16.0 : 24.0, m1 = local ? (nc_sampled ? 0.15 : 0.10) : 0.35;
	const double mlo = m0 + m1 * fabs(dlo), mhi = m0 + m1 * fabs(dhi);
	ZL = fmax(zp.mhat + (dlo - mlo) / rho, zp.A);
	ZH = fmin(zp.mhat + (dhi + mhi) / rho, zp.B);
}

// ---- pieces of the finish phase (one warp) ----------------------------------------------------------------------
// why a mesh went to the bucketed path (diagnostic counters, Workspace::fb_count[16 + 8 * kind + why])
#define ZN_WHY_SAMPLE 1    // unusable sample or fewer than ZN_MIN_N valid pixels
#define ZN_WHY_RANGE 2     // empty zone range
#define ZN_WHY_LIST 3      // tail / zone list overflow
#define ZN_WHY_SORT 4      // a lane's zone bins too full
#define ZN_WHY_RANK 5      // a median rank outside the zone
#define ZN_WHY_BOUND 6     // a clip bound entered the bulk
#define ZN_WHY_EMPTY 7     // empty bulk or zone
#define ZN_TREG 12        // tail elements per lane held in registers across the clip iterations (more: re-read from memory)
#define ZN_LANEMAX 64     // a lane never sorts more zone keys than this (4 bins); fuller bins send the mesh to the bucketed path

// Exclusive scan of the ZN_BINS bin counts in cnt[] (4 consecutive bins per lane): cnt becomes the bin starts, bend[]
// (registers) the ends of this lane's bins.  Returns the start of the lane's first bin.
__device__ __forceinline__ uint32_t zone_scan_bins(uint32_t* cnt, int lane, uint32_t (&bend)[ZN_BINS / 32])
{
	uint32_t c[ZN_BINS / 32], tot = 0;
#pragma unroll
	for (int j = 0; j < ZN_BINS / 32; ++j) { c[j] = cnt[lane * (ZN_BINS / 32) + j]; tot += c[j]; }
	uint32_t inc = tot;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	uint32_t run = inc - tot;
	const uint32_t start0 = run;
#pragma unroll
	for (int j = 0; j < ZN_BINS / 32; ++j) { cnt[lane * (ZN_BINS / 32) + j] = run; run += c[j]; bend[j] = run; }
	return start0;
}

// The keys are grouped by bin; every lane sorts the keys of its own 4 consecutive bins (a handful): the first 8 through a
// 19-comparator network in registers, any further ones by insertion into that sorted run.  Afterwards zone[] is sorted
// as a whole and the key of a rank is a single load.
template <typename K>
__device__ __forceinline__ bool zone_sort_bins(K* zone, uint32_t s, uint32_t e, K pad)
{
	const uint32_t m = e - s;
	const bool ok = m <= (uint32_t)ZN_LANEMAX;
	if (m > 1u && ok) {
		K a[8];
#pragma unroll
		for (int i = 0; i < 8; ++i) a[i] = (uint32_t)i < m ? zone[s + i] : pad;
#define ZN_CX(i, j) { const K lo_ = min(a[i], a[j]), hi_ = max(a[i], a[j]); a[i] = lo_; a[j] = hi_; }
		ZN_CX(0, 1) ZN_CX(2, 3) ZN_CX(4, 5) ZN_CX(6, 7)
		ZN_CX(0, 2) ZN_CX(1, 3) ZN_CX(4, 6) ZN_CX(5, 7)
		ZN_CX(1, 2) ZN_CX(5, 6)
		ZN_CX(0, 4) ZN_CX(1, 5) ZN_CX(2, 6) ZN_CX(3, 7)
		ZN_CX(2, 4) ZN_CX(3, 5)
		ZN_CX(1, 2) ZN_CX(3, 4) ZN_CX(5, 6)
#undef ZN_CX
#pragma unroll
		for (int i = 0; i < 8; ++i) if ((uint32_t)i < m) zone[s + i] = a[i];
		for (uint32_t i = s + 8u; i < e; ++i) {
			const K v = zone[i];
			uint32_t p = i;
			while (p > s && zone[p - 1] > v) { zone[p] = zone[p - 1]; --p; }
			zone[p] = v;
		}
	}
	return !__any_sync(0xffffffffu, !ok);
}

// The clip iterations and the final statistics.  ``zone``: the nZ zone keys, sorted (shared memory); ``tail_at(i)``: tail
// element i of the nA + nB (any order).  The first 32 * ZN_TREG tails live in registers for all sweeps.  vA / vB: smallest /
// largest value a bulk element can have.  Returns false when the lists cannot answer (the caller queues the mesh for the
// bucketed path).
template <typename T, typename TailAt>
__device__ bool zone_iterate(TailAt tail_at, const typename T::K* zone, int lane,
	int n, int nA, int nB, int nZL, int nZ, double s1b, double s2b, double pivot, double vA, double vB, TileStat& out, int& why)
{
	typedef typename T::K K;
	typedef typename T::Raw Raw;
	const int nbulk = n - nA - nB, nT = nA + nB;
	Raw tr[ZN_TREG];
	double td[T::kStoreD ? ZN_TREG : 1];
#pragma unroll
	for (int j = 0; j < ZN_TREG; ++j) {
		tr[j] = T::padraw();
		if (j * 32 < nT) {
			const int i = j * 32 + lane;
			if (i < nT) tr[j] = tail_at(i);
		}
		if (T::kStoreD) td[j] = T::rval(tr[j]) - pivot;
	}
	auto zone_median = [&](int q, bool two) -> double {
		const K k1 = zone[q];
		const K k2 = two ? zone[q + 1] : k1;
		return 0.5 * (T::val(k1) + T::val(k2));
	};
	// statistics of { bulk } + { tails inside [lo, hi] }; ``below`` = tails under lo
	auto tail_stats = [&](double lo, double hi, int& cnt, int& below, double& t1, double& t2) {
		int c = 0, bl = 0; double a1 = 0.0, a2 = 0.0;
		const typename T::Range rg = T::range(lo, hi);
#pragma unroll
		for (int j = 0; j < ZN_TREG; ++j) {
			if (j * 32 < nT) {   // warp-uniform; padded slots compare false on both tests
				const Raw r = tr[j];
				const double d = T::kStoreD ? td[T::kStoreD ? j : 0] : T::rval(r) - pivot;
				const bool lw = r < rg.tlo, in = !lw && r <= rg.thi;
				bl += lw ? 1 : 0; c += in ? 1 : 0;
				if (in) { a1 += d; a2 = fma(d, d, a2); }
			}
		}
		for (int i = ZN_TREG * 32 + lane; i < nT; i += 32) {
			const Raw r = tail_at(i);
			const double d = T::rval(r) - pivot;
			if (r < rg.tlo) ++bl;
			else if (r <= rg.thi) { ++c; a1 += d; a2 = fma(d, d, a2); }
		}
		cnt = __reduce_add_sync(0xffffffffu, c); below = __reduce_add_sync(0xffffffffu, bl);
		t1 = 0.0; t2 = 0.0;
		if (nT) { t1 = warp_sum_d(a1); t2 = warp_sum_d(a2); }
	};
	double lo_run = -INFINITY, hi_run = INFINITY, lo_last = 0.0, hi_last = 0.0;
	int n_prev = -1;
	bool nested_last = true;
	double mean_c = 0.0, sd_c = 0.0, med_c = 0.0;       // statistics of the buffer at the last bound computation
#pragma unroll 1
	for (int it = 0; it < 6; ++it) {
		// it < 5: a clip iteration on the running bounds.  it == 5 (reached only without the shortcut below): the final
		// statistics of the ORIGINAL valid values inside the last bounds (lo_last <= lo_run <= vA, hi_last >= vB).
		const bool fin = it == 5;
		int c, below; double t1, t2;
		tail_stats(fin ? lo_last : lo_run, fin ? hi_last : hi_run, c, below, t1, t2);
		const int ni = nbulk + c;
		if (!fin && ni == n_prev) {
			// the previous clip removed nothing: its bounds are the last ones.  When they also lie inside the running range the
			// final set IS the buffer whose count / mean / median / std were just computed with those bounds
			if (nested_last) { out.nfin = n_prev; out.mean = mean_c; out.std = sd_c; out.med = med_c; return true; }
			it = 4; continue;
		}
		const double m1 = (s1b + t1) / (double)ni;
		const double sd = sqrt(fmax((s2b + t2) / (double)ni - m1 * m1, 0.0));
		const int q = below + ((ni - 1) >> 1) - nZL;
		const bool two = (ni & 1) == 0;
		if (q < 0 || q + (two ? 1 : 0) >= nZ) { why = ZN_WHY_RANK; return false; }
		const double med = zone_median(q, two);
		if (fin) { out.nfin = ni; out.mean = pivot + m1; out.std = sd; out.med = med; return true; }
		lo_last = med - 3.0 * sd; hi_last = med + 3.0 * sd;
		nested_last = lo_last >= lo_run && hi_last <= hi_run;
		lo_run = fmax(lo_run, lo_last); hi_run = fmin(hi_run, hi_last);
		if (lo_run > vA || hi_run < vB) {   // a bound entered the bulk: the statistics of this iteration are still exact, the caller may re-plan around them
			why = ZN_WHY_BOUND; out.med = med; out.std = sd; out.nfin = ni; return false;
		}
		n_prev = ni; mean_c = pivot + m1; sd_c = sd; med_c = med;
	}
	why = ZN_WHY_EMPTY; return false;   // not reached
}

// Finish from per-lane lists in shared memory (the fused raw-pixel kernel).  tcnt / zcnt: this lane's list lengths;
// zl / zscale: zone bin map; the other arguments are warp-uniform.
template <typename T, int TCAP>
__device__ bool zone_finish(ZoneSmem<T, TCAP>& sm, int lane, int n, int nA, int nB, int nZL, int tcnt, int zcnt,
	double s1b, double s2b, double pivot, double vA, double vB, typename T::Raw zl, float zscale, TileStat& out, int& why)
{
	typedef typename T::Raw Raw;
	out.mean = out.med = out.std = nan_d();
	out.nfin = 0; out.pad = 0;
	if (n - nA - nB <= 0) { why = ZN_WHY_EMPTY; return false; }
	if (__any_sync(0xffffffffu, tcnt > TCAP || zcnt > ZN_ZCAP)) { why = ZN_WHY_LIST; return false; }
	const int nT = nA + nB;
	const int nZ = __reduce_add_sync(0xffffffffu, zcnt);
	if (nZ == 0) { why = ZN_WHY_EMPTY; return false; }
	__syncwarp();
	// ---- tails: per-lane lists -> dense, in place (row j is read completely before anything is written, and the
	// dense positions of row j never lie beyond row j)
	{
		const int maxc = __reduce_max_sync(0xffffffffu, tcnt);
		int base = 0;
#pragma unroll 2
		for (int j = 0; j < maxc; ++j) {
			const bool p = j < tcnt;
			Raw v = Raw();
			if (p) v = sm.tails[j * 32 + lane];
			const unsigned m = __ballot_sync(0xffffffffu, p);
			__syncwarp();
			if (p) sm.tails[base + __popc(m & ((1u << lane) - 1u))] = v;
			base += __popc(m);
			__syncwarp();
		}
	}
	// ---- zone: counting sort by bin through registers, in place (groups of 8 list slots, skipped when no lane uses them)
	uint32_t bend[ZN_BINS / 32];
	{
		Raw zr[ZN_ZCAP];
		const Raw* zraw = reinterpret_cast<const Raw*>(sm.zone);
		const int maxz = __reduce_max_sync(0xffffffffu, zcnt);
#pragma unroll
		for (int g8 = 0; g8 < ZN_ZCAP / 8; ++g8) {
			if (8 * g8 < maxz) {
#pragma unroll
				for (int j = 8 * g8; j < 8 * g8 + 8; ++j) zr[j] = (j < zcnt) ? zraw[j * 32 + lane] : zl;
			}
		}
#pragma unroll
		for (int j = 0; j < ZN_BINS / 32; ++j) sm.cnt[lane + 32 * j] = 0u;
		__syncwarp();
#pragma unroll
		for (int g8 = 0; g8 < ZN_ZCAP / 8; ++g8) {
			if (8 * g8 < maxz) {
#pragma unroll
				for (int j = 8 * g8; j < 8 * g8 + 8; ++j) {
					const int b = min(ZN_BINS - 1, (int)(T::offs(zr[j], zl) * zscale));
					if (j < zcnt) atomicAdd(&sm.cnt[b], 1u);
				}
			}
		}
		__syncwarp();
		const uint32_t start0 = zone_scan_bins(sm.cnt, lane, bend);
		__syncwarp();
#pragma unroll
		for (int g8 = 0; g8 < ZN_ZCAP / 8; ++g8) {
			if (8 * g8 < maxz) {
#pragma unroll
				for (int j = 8 * g8; j < 8 * g8 + 8; ++j) {
					const int b = min(ZN_BINS - 1, (int)(T::offs(zr[j], zl) * zscale));
					if (j < zcnt) sm.zone[atomicAdd(&sm.cnt[b], 1u)] = T::key_of(zr[j]);
				}
			}
		}
		__syncwarp();
		if (!zone_sort_bins(sm.zone, start0, bend[ZN_BINS / 32 - 1], T::padkey())) { why = ZN_WHY_SORT; return false; }
		__syncwarp();
	}
	const Raw* tl = sm.tails;
	return zone_iterate<T>([&](int i) { return tl[i]; }, sm.zone, lane, n, nA, nB, nZL, nZ, s1b, s2b, pivot, vA, vB, out, why);
}

// Finish from lists that are not per-lane: tails in TSEG segments of capacity tcap with tq[s] entries, zone elements in
// ZSEG segments of capacity zcap with zq[s] entries (Raw; global or shared memory).  ``zone`` (>= sum zq keys) and ``cnt``
// (ZN_BINS words) are this warp's shared memory.
template <typename T, int TSEG, int ZSEG>
__device__ bool zone_finish_seg(const typename T::Raw* gt, int tcap, const int (&tq)[TSEG],
	const typename T::Raw* gz, int zcap, const int (&zq)[ZSEG], typename T::K* zone, uint32_t* cnt, int lane,
	int n, int nA, int nB, int nZL, double s1b, double s2b, double pivot, double vA, double vB, typename T::Raw zl, float zscale,
	TileStat& out, int& why)
{
	typedef typename T::Raw Raw;
	out.mean = out.med = out.std = nan_d();
	out.nfin = 0; out.pad = 0;
	if (n - nA - nB <= 0) { why = ZN_WHY_EMPTY; return false; }
	int nZ = 0;
#pragma unroll
	for (int s = 0; s < ZSEG; ++s) nZ += zq[s];
	if (nZ == 0) { why = ZN_WHY_EMPTY; return false; }
	uint32_t bend[ZN_BINS / 32];
#pragma unroll
	for (int j = 0; j < ZN_BINS / 32; ++j) cnt[lane + 32 * j] = 0u;
	__syncwarp();
	auto zone_each = [&](auto f) {
#pragma unroll
		for (int s = 0; s < ZSEG; ++s) for (int i = lane; i < zq[s]; i += 32) f(gz[s * zcap + i]);
	};
	zone_each([&](Raw r) { atomicAdd(&cnt[min(ZN_BINS - 1, (int)(T::offs(r, zl) * zscale))], 1u); });
	__syncwarp();
	const uint32_t start0 = zone_scan_bins(cnt, lane, bend);
	__syncwarp();
	zone_each([&](Raw r) { zone[atomicAdd(&cnt[min(ZN_BINS - 1, (int)(T::offs(r, zl) * zscale))], 1u)] = T::key_of(r); });
	__syncwarp();
	if (!zone_sort_bins(zone, start0, bend[ZN_BINS / 32 - 1], T::padkey())) { why = ZN_WHY_SORT; return false; }
	__syncwarp();
	// tail element i of the dense order "segment 0, segment 1, ..."
	auto tail_at = [&](int i) -> Raw {
		int s = 0;
#pragma unroll
		for (int t = 0; t < TSEG - 1; ++t) { if (s == t && i >= tq[t]) { i -= tq[t]; s = t + 1; } }
		return gt[s * tcap + i];
	};
	return zone_iterate<T>(tail_at, zone, lane, n, nA, nB, nZL, nZ, s1b, s2b, pivot, vA, vB, out, why);
}
