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
// L2.  Anything the two lists cannot answer exactly -- a bound that enters the bulk, a median rank outside the zone, a
// list overflow, a degenerate sample, a sparsely populated mesh -- sends the mesh to a queue that the bucketed kernels
// (tbk_tile_warp.cuh) work off afterwards.  The result of a mesh that does not fall back is exactly the reference's:
// the same membership decisions on the same float64 comparisons, exact medians, float64 moments.
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
	struct Range { Raw tlo, thi; };
	__device__ static __forceinline__ Range range(double lo, double hi) { Range r; r.tlo = lo; r.thi = hi; return r; }
};

template <typename T>
struct ZoneSmem {
	typename T::Raw tails[ZN_TCAP * 32];   // pass 2: per-lane lists [j * 32 + lane]; then dense
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
__device__ __forceinline__ void zone_range(const ZonePlan& zp, int n, int nA, int nB, int nM, int nC, double& ZL, double& ZH)
{
	const int nbulk = n - nA - nB;
	const double r_lo = (double)((n - 1 - nB) >> 1), r_hi = (double)(((n - 1 + nA) >> 1) + 1);
	const bool local = nC >= 64;
	// elements per unit value near the centre: counted, or (sparse centre) a Gaussian of the sample width
	const double rho = local ? (double)nC / (2.0 * ZN_CW * zp.shat) : (double)max(nbulk, 1) * 0.418 / zp.shat;
	const double dlo = r_lo - (double)nM, dhi = r_hi - (double)nM;
	const double m0 = local ? 16.0 : 24.0, m1 = local ? 0.10 : 0.35;
	const double mlo = m0 + m1 * fabs(dlo), mhi = m0 + m1 * fabs(dhi);
	ZL = fmax(zp.mhat + (dlo - mlo) / rho, zp.A);
	ZH = fmin(zp.mhat + (dhi + mhi) / rho, zp.B);
}

// ---- pieces of the finish phase (one warp) ----------------------------------------------------------------------

// Exclusive scan of the ZN_BINS bin counts in cnt[] (4 consecutive bins per lane): cnt becomes the bin starts, bend[]
// (registers) the ends of this lane's bins.
__device__ __forceinline__ void zone_scan_bins(uint32_t* cnt, int lane, uint32_t (&bend)[ZN_BINS / 32])
{
	uint32_t c[ZN_BINS / 32], tot = 0;
#pragma unroll
	for (int j = 0; j < ZN_BINS / 32; ++j) { c[j] = cnt[lane * (ZN_BINS / 32) + j]; tot += c[j]; }
	uint32_t inc = tot;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	uint32_t run = inc - tot;
#pragma unroll
	for (int j = 0; j < ZN_BINS / 32; ++j) { cnt[lane * (ZN_BINS / 32) + j] = run; run += c[j]; bend[j] = run; }
}

// The clip iterations and the final statistics.  ``zone``: the nZ zone keys sorted by bin (shared memory), ``bend``: the
// bin ends (registers, 4 bins per lane); ``tails_each(f)`` calls f(raw) for every tail element this lane is responsible for
// (all lanes together cover each of the nA + nB tail elements once).  vA / vB: smallest / largest value a bulk element
// can have.  Returns false when the lists cannot answer (the caller queues the mesh for the bucketed path).
template <typename T, typename TailsEach>
__device__ bool zone_iterate(TailsEach tails_each, const typename T::K* zone, const uint32_t (&bend)[ZN_BINS / 32], int lane,
	int n, int nA, int nB, int nZL, int nZ, double s1b, double s2b, double pivot, double vA, double vB, TileStat& out)
{
	typedef typename T::K K;
	typedef typename T::Raw Raw;
	const int nbulk = n - nA - nB, nT = nA + nB;
	// key of zone rank r: bin lookup over the ends, then the ranks inside the bin by counting (a bin holds a few keys)
	auto zone_key = [&](uint32_t r) -> K {
		const unsigned m4 = __ballot_sync(0xffffffffu, bend[ZN_BINS / 32 - 1] > r);
		const int g = __ffs(m4) - 1;
		uint32_t bs = __shfl_sync(0xffffffffu, bend[ZN_BINS / 32 - 1], max(g - 1, 0));
		if (g == 0) bs = 0u;
		uint32_t be = 0u;
		bool found = false;
#pragma unroll
		for (int j = 0; j < ZN_BINS / 32; ++j) {
			const uint32_t ej = __shfl_sync(0xffffffffu, bend[j], g);
			if (!found) { if (ej > r) { be = ej; found = true; } else bs = ej; }
		}
		const uint32_t m = be - bs;
		if (m > 32u) {   // many equal / near-equal keys: exact radix selection
			K k1, k2;
			tw_select_in_span<T>(zone, bs, be, r - bs, false, lane, k1, k2);
			return k1;
		}
		const K mine = ((uint32_t)lane < m) ? zone[bs + lane] : T::padkey();
		uint32_t rank = 0u;
		for (uint32_t j = 0; j < m; ++j) {
			const K kj = __shfl_sync(0xffffffffu, mine, (int)j);
			rank += (kj < mine || (kj == mine && j < (uint32_t)lane)) ? 1u : 0u;
		}
		const unsigned hit = __ballot_sync(0xffffffffu, (uint32_t)lane < m && rank == r - bs);
		return __shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
	};
	auto zone_median = [&](int q, bool two) -> double {
		const K k1 = zone_key((uint32_t)q);
		const K k2 = two ? zone_key((uint32_t)q + 1u) : k1;
		return 0.5 * (T::val(k1) + T::val(k2));
	};
	// statistics of { bulk } + { tails inside [lo, hi] }; ``below`` = tails under lo
	auto tail_stats = [&](double lo, double hi, int& cnt, int& below, double& t1, double& t2) {
		int c = 0, bl = 0; double a1 = 0.0, a2 = 0.0;
		const typename T::Range rg = T::range(lo, hi);
		tails_each([&](Raw r) {
			const double d = T::rval(r) - pivot;
			if (r < rg.tlo) ++bl;
			else if (r <= rg.thi) { ++c; a1 += d; a2 = fma(d, d, a2); }
		});
		cnt = __reduce_add_sync(0xffffffffu, c); below = __reduce_add_sync(0xffffffffu, bl);
		t1 = 0.0; t2 = 0.0;
		if (nT) { t1 = warp_sum_d(a1); t2 = warp_sum_d(a2); }
	};
	double lo_run = -INFINITY, hi_run = INFINITY, lo_last = 0.0, hi_last = 0.0;
	int n_prev = -1;
	bool converged = false, nested_last = true;
	double mean_c = 0.0, sd_c = 0.0, med_c = 0.0;       // statistics of the buffer at the last bound computation
	for (int it = 0; it < 5; ++it) {
		int c, below; double t1, t2;
		tail_stats(lo_run, hi_run, c, below, t1, t2);
		const int ni = nbulk + c;
		if (ni == n_prev) { converged = true; break; }   // the previous clip removed nothing: its bounds are the last ones
		const double m1 = (s1b + t1) / (double)ni;
		const double sd = sqrt(fmax((s2b + t2) / (double)ni - m1 * m1, 0.0));
		const int q = below + ((ni - 1) >> 1) - nZL;
		const bool two = (ni & 1) == 0;
		if (q < 0 || q + (two ? 1 : 0) >= nZ) return false;
		const double med = zone_median(q, two);
		lo_last = med - 3.0 * sd; hi_last = med + 3.0 * sd;
		nested_last = lo_last >= lo_run && hi_last <= hi_run;
		lo_run = fmax(lo_run, lo_last); hi_run = fmin(hi_run, hi_last);
		if (lo_run > vA || hi_run < vB) return false;   // a bound entered the bulk
		n_prev = ni; mean_c = pivot + m1; sd_c = sd; med_c = med;
	}
	// ---- final statistics: ORIGINAL valid values inside the last bounds (lo_last <= lo_run <= vA, hi_last >= vB)
	if (converged && nested_last) {
		// the last clip removed nothing and its bounds lie inside the running range: the final set IS the buffer whose
		// count / mean / median / std were computed with those bounds
		out.nfin = n_prev; out.mean = mean_c; out.std = sd_c; out.med = med_c;
		return true;
	}
	{
		int c, below; double t1, t2;
		tail_stats(lo_last, hi_last, c, below, t1, t2);
		const int nf = nbulk + c;
		const double m1 = (s1b + t1) / (double)nf;
		const int q = below + ((nf - 1) >> 1) - nZL;
		const bool two = (nf & 1) == 0;
		if (q < 0 || q + (two ? 1 : 0) >= nZ) return false;
		out.nfin = nf;
		out.mean = pivot + m1;
		out.std = sqrt(fmax((s2b + t2) / (double)nf - m1 * m1, 0.0));
		out.med = zone_median(q, two);
	}
	return true;
}

// Finish from per-lane lists in shared memory (the fused raw-pixel kernel).  tcnt / zcnt: this lane's list lengths;
// zl / zscale: zone bin map; the other arguments are warp-uniform.
template <typename T>
__device__ bool zone_finish(ZoneSmem<T>& sm, int lane, int n, int nA, int nB, int nZL, int tcnt, int zcnt,
	double s1b, double s2b, double pivot, double vA, double vB, typename T::Raw zl, float zscale, TileStat& out)
{
	typedef typename T::Raw Raw;
	out.mean = out.med = out.std = nan_d();
	out.nfin = 0; out.pad = 0;
	if (n - nA - nB <= 0) return false;
	if (__any_sync(0xffffffffu, tcnt > ZN_TCAP || zcnt > ZN_ZCAP)) return false;
	const int nT = nA + nB;
	const int nZ = __reduce_add_sync(0xffffffffu, zcnt);
	if (nZ == 0) return false;
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
		zone_scan_bins(sm.cnt, lane, bend);
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
	}
	const Raw* tl = sm.tails;
	auto tails_each = [&](auto f) { for (int i = lane; i < nT; i += 32) f(tl[i]); };
	return zone_iterate<T>(tails_each, sm.zone, bend, lane, n, nA, nB, nZL, nZ, s1b, s2b, pivot, vA, vB, out);
}

// Finish from lists that are not per-lane: tails in TSEG segments of capacity tcap with tq[s] entries, zone elements in
// ZSEG segments of capacity zcap with zq[s] entries (Raw; global or shared memory).  ``zone`` (>= sum zq keys) and ``cnt``
// (ZN_BINS words) are this warp's shared memory.
template <typename T, int TSEG, int ZSEG>
__device__ bool zone_finish_seg(const typename T::Raw* gt, int tcap, const int (&tq)[TSEG],
	const typename T::Raw* gz, int zcap, const int (&zq)[ZSEG], typename T::K* zone, uint32_t* cnt, int lane,
	int n, int nA, int nB, int nZL, double s1b, double s2b, double pivot, double vA, double vB, typename T::Raw zl, float zscale,
	TileStat& out)
{
	typedef typename T::Raw Raw;
	out.mean = out.med = out.std = nan_d();
	out.nfin = 0; out.pad = 0;
	if (n - nA - nB <= 0) return false;
	int nZ = 0;
#pragma unroll
	for (int s = 0; s < ZSEG; ++s) nZ += zq[s];
	if (nZ == 0) return false;
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
	zone_scan_bins(cnt, lane, bend);
	__syncwarp();
	zone_each([&](Raw r) { zone[atomicAdd(&cnt[min(ZN_BINS - 1, (int)(T::offs(r, zl) * zscale))], 1u)] = T::key_of(r); });
	__syncwarp();
	auto tails_each = [&](auto f) {
#pragma unroll
		for (int s = 0; s < TSEG; ++s) for (int i = lane; i < tq[s]; i += 32) f(gt[s * tcap + i]);
	};
	return zone_iterate<T>(tails_each, zone, bend, lane, n, nA, nB, nZL, nZ, s1b, s2b, pivot, vA, vB, out);
}
