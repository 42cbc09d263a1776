// tbk_tile_warp.cuh -- warp-per-mesh sigma-clipped statistics on float32 pixels (the hot kernel).
//
// Same arithmetic as tbk_tile.cuh (astropy 5.1 SigmaClip(3, maxiters=5, median / std) + nan-aware
// median / mean / std of the survivors; photometry/backgrounds.py:105-106, 200-205) but organised for
// throughput:
//   * one warp owns one 64x64 mesh; the 4096 pixels live in registers (128 per lane), so there is no
//     block barrier anywhere;
//   * the valid pixels are counting-sorted ONCE into 2048 fine bins laid over a robust window
//     (sample median +- 10 sample sigma; everything outside goes to the first / last bin), keyed by
//     their float32 bit patterns (non-negative floats order like unsigned integers);
//   * after that every clip iteration only touches the few elements that leave the buffer (they are
//     contiguous in bin order) and the one bin that holds the median rank, so the five iterations and
//     the exact medians cost a few hundred instructions per mesh instead of several passes.
// Moments are accumulated in float64 about a float32 pivot (x - pivot is exact in float64).  Elements
// of the two overflow bins are never subtracted from running sums -- their in-range part is summed
// directly each iteration -- so removing bright outliers leaves no rounding residue in the core sums.
#pragma once
#include "tbk_common.cuh"

#define TW_NB 2048
#define TW_WORDS (TW_NB / 2)
#define TW_INVALID 0x7f800000u   // +inf as key: masked pixel

// counter word w lives at w + (w >> 5): one pad word per 32 keeps the per-lane chunked scan (lane l owns
// words 32l .. 32l+31) free of shared-memory bank conflicts
#define TW_CIDX(w) ((w) + ((w) >> 5))
struct TileWarpSmem {
	uint32_t keys[TBK_NPIX_TILE];         // bucketed float bit patterns
	uint32_t cnt[TW_WORDS + TW_WORDS / 32]; // packed uint16 pairs: counts -> starts -> ends
};

struct TwBinMap {
	float scale, off;
};

// Monotone non-decreasing map value -> bin: one FFMA onto the 2^23 "magic" range, then an integer clamp.
// off >= 2^22 and x >= 0, scale > 0 guarantee t > 0, so the float bit pattern is monotone in t.
__device__ __forceinline__ int tw_bin(const TwBinMap& m, float x)
{
	const float t = fmaf(x, m.scale, m.off);
	const int b = __float_as_int(t) - 0x4B000000;
	return max(0, min(TW_NB - 1, b));
}

__device__ __forceinline__ uint32_t tw_cend(const TileWarpSmem& sm, int b)
{
	return (sm.cnt[TW_CIDX(b >> 1)] >> ((b & 1) << 4)) & 0xFFFFu;
}
__device__ __forceinline__ uint32_t tw_cstart(const TileWarpSmem& sm, int b)
{
	return b ? tw_cend(sm, b - 1) : 0u;
}

// smallest bin b with cend(b) > P  (P < total)
__device__ __forceinline__ int tw_find_bin(const TileWarpSmem& sm, uint32_t P, int lane)
{
	unsigned m = __ballot_sync(0xffffffffu, tw_cend(sm, 64 * lane + 63) > P);
	const int g = __ffs(m) - 1;
	m = __ballot_sync(0xffffffffu, tw_cend(sm, 64 * g + 2 * lane + 1) > P);
	const int h = __ffs(m) - 1;
	const int b = 64 * g + 2 * h;
	return (tw_cend(sm, b) > P) ? b : b + 1;
}

__device__ __forceinline__ uint32_t warp_bitonic32(uint32_t v, int lane)
{
#pragma unroll
	for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
			const bool up = (lane & k) == 0, lower = (lane & j) == 0;
			v = (lower == up) ? min(v, o) : max(v, o);
		}
	}
	return v;
}

// keys at sorted ranks q and q+1 (when want2) among the span [s, e) of one bin (unsorted inside).
__device__ __noinline__ void tw_select_in_span(const TileWarpSmem& sm, uint32_t s, uint32_t e, uint32_t q, bool want2,
	int lane, uint32_t& k1, uint32_t& k2)
{
	const uint32_t m = e - s;
	if (m <= 32u) {
		uint32_t v = (s + lane < e) ? sm.keys[s + lane] : 0xFFFFFFFFu;
		v = warp_bitonic32(v, lane);
		k1 = __shfl_sync(0xffffffffu, v, q);
		k2 = want2 ? __shfl_sync(0xffffffffu, v, min(q + 1u, 31u)) : k1;
		return;
	}
	// large bin: exact radix selection on the key bits (MSB first), one sweep of the span per bit
	for (int r = 0; r < (want2 ? 2 : 1); ++r) {
		uint32_t target = q + r, prefix = 0;
		for (int bit = 31; bit >= 0; --bit) {
			const uint32_t mask = ~((1u << bit) - 1u);      // bits above and including ``bit``
			int c0 = 0;
			for (uint32_t p = s + lane; p < e; p += 32) {
				const uint32_t k = sm.keys[p];
				c0 += ((k & mask) == prefix) ? 1 : 0;          // same upper bits, this bit = 0
			}
			c0 = __reduce_add_sync(0xffffffffu, c0);
			if (target >= (uint32_t)c0) { target -= c0; prefix |= (1u << bit); }
		}
		if (r == 0) k1 = prefix; else k2 = prefix;
	}
	if (!want2) k2 = k1;
}

// median (mean of the keys at sorted ranks P and P+1 when ``even``) by bin lookup
__device__ __forceinline__ double tw_median_at(const TileWarpSmem& sm, uint32_t P, bool even, int lane)
{
	const int b = tw_find_bin(sm, P, lane);
	const uint32_t bs = tw_cstart(sm, b), be = tw_cend(sm, b);
	uint32_t k1, k2;
	const bool second_here = even && (P + 1u < be);
	tw_select_in_span(sm, bs, be, P - bs, second_here, lane, k1, k2);
	if (even && !second_here) {
		const int b2 = tw_find_bin(sm, P + 1u, lane);
		uint32_t dummy;
		tw_select_in_span(sm, tw_cstart(sm, b2), tw_cend(sm, b2), 0u, false, lane, k2, dummy);
	}
	return 0.5 * ((double)__uint_as_float(k1) + (double)__uint_as_float(k2));
}

// smallest float32 >= d and largest float32 <= d, as keys clamped to the non-negative finite range
__device__ __forceinline__ uint32_t tw_key_ceil(double d)
{
	if (!(d > 0.0)) return 0u;
	float f = (float)d;                       // round to nearest
	if ((double)f < d) f = __uint_as_float(__float_as_uint(f) + 1u);
	return min(__float_as_uint(f), 0x7f7fffffu + 1u);  // may become +inf key: nothing is >= it
}
__device__ __forceinline__ bool tw_key_floor(double d, uint32_t& key)
{
	if (d < 0.0) return false;                // nothing is <= a negative bound
	if (d == 0.0) { key = 0u; return true; }  // +-0: only zeros qualify
	float f = (float)d;
	if ((double)f > d) f = __uint_as_float(__float_as_uint(f) - 1u);  // f > d >= 0 so f > 0
	key = min(__float_as_uint(f), 0x7f7fffffu);
	return true;
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// The per-warp statistics.  ``v`` holds the lane's 128 pixels as float bit patterns (TW_INVALID = masked),
// nvalid / kmin are warp-uniform.  Returns the result in every lane.
__device__ TileStat tile_warp_stats(const uint32_t (&v)[128], int nvalid, uint32_t kmin,
	TileWarpSmem& sm, int lane)
{
	TileStat out;
	out.mean = out.med = out.std = nan_d();
	out.nfin = 0; out.pad = 0;
	if (nvalid == 0) return out;
	const float vmin = __uint_as_float(kmin);

	// ---- robust window from two 32-element samples spread over the mesh
	float w0, w1, pivot_f;
	{
		const int sel = lane & 3;
		const uint32_t sa = sel == 0 ? v[0] : sel == 1 ? v[37] : sel == 2 ? v[74] : v[111];
		const uint32_t sb = sel == 0 ? v[58] : sel == 1 ? v[95] : sel == 2 ? v[4] : v[41];
		float med = 0.f, iqr = 0.f; int sets = 0;
#pragma unroll
		for (int t = 0; t < 2; ++t) {
			const uint32_t k = warp_bitonic32(t ? sb : sa, lane);
			const int m = __popc(__ballot_sync(0xffffffffu, k != TW_INVALID));
			if (m >= 8) {
				med += __uint_as_float(__shfl_sync(0xffffffffu, k, m >> 1));
				iqr += __uint_as_float(__shfl_sync(0xffffffffu, k, (3 * m) >> 2)) - __uint_as_float(__shfl_sync(0xffffffffu, k, m >> 2));
				++sets;
			}
		}
		if (sets && iqr > 0.f) {
			med /= (float)sets;
			const float half = 10.0f * (iqr / (float)sets) / 1.349f;
			w0 = med - half; w1 = med + half;
			pivot_f = fmaxf(med, 0.f);
		} else {
			// degenerate sample: use the full data range (constant meshes end here)
			uint32_t kmax = 0u;
#pragma unroll
			for (int e = 0; e < 128; ++e) kmax = max(kmax, v[e] == TW_INVALID ? 0u : v[e]);
			kmax = __reduce_max_sync(0xffffffffu, kmax);
			if (kmax == kmin) {  // constant mesh: sigma = 0, nothing is clipped
				out.mean = out.med = (double)vmin; out.std = 0.0; out.nfin = nvalid;
				return out;
			}
			w0 = vmin; w1 = __uint_as_float(kmax); pivot_f = vmin;
		}
	}
	TwBinMap bm;
	bm.scale = (float)(TW_NB - 2) / (w1 - w0);
	if (!(bm.scale < 1e30f)) bm.scale = 1e30f;
	if (w0 > 0.f && w0 * bm.scale > 4194304.0f) bm.scale = 4194304.0f / w0;  // keeps off >= 2^22
	bm.off = fmaf(-w0, bm.scale, 8388609.0f);
	const double pivot = (double)pivot_f;

	// ---- pass 1: bin counts
	for (int i = lane; i < TW_WORDS + TW_WORDS / 32; i += 32) sm.cnt[i] = 0u;
	__syncwarp();
#pragma unroll
	for (int e = 0; e < 128; ++e) {
		const uint32_t k = v[e];
		const int b = tw_bin(bm, __uint_as_float(k));
		if (k != TW_INVALID) atomicAdd(&sm.cnt[TW_CIDX(b >> 1)], 1u << ((b & 1) << 4));
	}
	__syncwarp();
	// ---- exclusive scan of the packed counters (32 words = 64 bins per lane)
	{
		uint32_t tot = 0;
#pragma unroll 8
		for (int j = 0; j < 32; ++j) { const uint32_t w = sm.cnt[lane * 33 + j]; tot += (w & 0xFFFFu) + (w >> 16); }
		uint32_t inc = tot;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
		uint32_t run = inc - tot;
#pragma unroll 8
		for (int j = 0; j < 32; ++j) {
			const uint32_t w = sm.cnt[lane * 33 + j];
			const uint32_t c0 = w & 0xFFFFu, c1 = w >> 16;
			sm.cnt[lane * 33 + j] = run | ((run + c0) << 16);
			run += c0 + c1;
		}
	}
	__syncwarp();
	// ---- pass 2: scatter in groups of 8 (the counters advance from bin starts to bin ends)
#pragma unroll
	for (int g = 0; g < 16; ++g) {
		uint32_t old[8]; int sh[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const uint32_t k = v[8 * g + j];
			const int b = tw_bin(bm, __uint_as_float(k));
			sh[j] = (b & 1) << 4;
			old[j] = 0u;
			if (k != TW_INVALID) old[j] = atomicAdd(&sm.cnt[TW_CIDX(b >> 1)], 1u << sh[j]);
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const uint32_t k = v[8 * g + j];
			if (k != TW_INVALID) sm.keys[(old[j] >> sh[j]) & 0xFFFFu] = k;
		}
	}
	__syncwarp();

	// ---- moments about the pivot: core bins [t0e, t1s) and the two overflow bins
	const uint32_t t0e = tw_cend(sm, 0);                // [0, t0e)       = low overflow bin
	const uint32_t t1s = tw_cstart(sm, TW_NB - 1);      // [t1s, nvalid)  = high overflow bin
	const uint32_t ntail = t0e + ((uint32_t)nvalid - t1s);
	double s1c = 0.0, s2c = 0.0;
	{
		double a1 = 0.0, a2 = 0.0, b1 = 0.0, b2 = 0.0;
		uint32_t p = t0e + lane;
		for (; p + 32 < t1s; p += 64) {
			const double d0 = (double)__uint_as_float(sm.keys[p]) - pivot;
			const double d1 = (double)__uint_as_float(sm.keys[p + 32]) - pivot;
			a1 += d0; a2 = fma(d0, d0, a2); b1 += d1; b2 = fma(d1, d1, b2);
		}
		if (p < t1s) { const double d0 = (double)__uint_as_float(sm.keys[p]) - pivot; a1 += d0; a2 = fma(d0, d0, a2); }
		s1c = warp_sum_d(a1 + b1); s2c = warp_sum_d(a2 + b2);
	}
	int nc = (int)(t1s - t0e);

	uint32_t lo_key = 0u, hi_key = 0x7f7fffffu;         // running intersection (inclusive)
	uint32_t lo_last = 0u, hi_last = 0x7f7fffffu;       // last bounds (inclusive)
	bool hi_last_ok = true, nested_last = true, converged = false, exhausted = false, empty_run = false;
	uint32_t below = 0u;                                // valid elements with key < lo_key
	int n_prev = -1, n = 0;
	double med = 0.0, mean = 0.0, sd = 0.0;
	// in-range part of the overflow bins (direct sums, recomputed whenever the bounds move)
	int tn = 0; double t1 = 0.0, t2 = 0.0;
	for (uint32_t j = lane; j < ntail; j += 32) {
		const uint32_t p = j < t0e ? j : t1s + (j - t0e);
		const double d = (double)__uint_as_float(sm.keys[p]) - pivot;
		++tn; t1 += d; t2 = fma(d, d, t2);
	}
	if (ntail) { tn = __reduce_add_sync(0xffffffffu, tn); t1 = warp_sum_d(t1); t2 = warp_sum_d(t2); }

	for (int it = 0; it < 6; ++it) {
		const int ncur = nc + tn;
		if (it > 0 && ncur == n_prev) { converged = true; break; }   // stats of the previous pass stand
		n = ncur;
		if (n == 0) { empty_run = true; break; }
		const double m1 = (s1c + t1) / (double)n;
		mean = pivot + m1;
		sd = sqrt(fmax((s2c + t2) / (double)n - m1 * m1, 0.0));
		med = tw_median_at(sm, below + (uint32_t)((n - 1) >> 1), (n & 1) == 0, lane);
		if (it == 5) { exhausted = true; break; }   // five bound computations done: buffer after the last clip
		lo_last = tw_key_ceil(med - 3.0 * sd);
		hi_last_ok = tw_key_floor(med + 3.0 * sd, hi_last);
		nested_last = (lo_last >= lo_key) && hi_last_ok && (hi_last <= hi_key);
		const uint32_t new_lo = max(lo_key, lo_last);
		const uint32_t new_hi = hi_last_ok ? min(hi_key, hi_last) : 0u;
		n_prev = n;
		if (!hi_last_ok || new_lo > new_hi) { empty_run = true; break; }  // cannot happen for finite data
		if (new_lo == lo_key && new_hi == hi_key) continue;               // nothing can leave: next pass converges
		// one sweep: (a) core-bin elements leaving the buffer, (b) everything removed below (for ``below``),
		// (c) the overflow-bin elements still inside the new bounds
		uint32_t sL = 0, eL = 0, sH = 0, eH = 0;
		if (new_lo > lo_key) { sL = tw_cstart(sm, tw_bin(bm, __uint_as_float(lo_key))); eL = tw_cend(sm, tw_bin(bm, __uint_as_float(new_lo))); }
		if (new_hi < hi_key) { sH = tw_cstart(sm, tw_bin(bm, __uint_as_float(new_hi))); eH = tw_cend(sm, tw_bin(bm, __uint_as_float(hi_key))); }
		const uint32_t nL = eL - sL, nH = eH - sH;
		int rn = 0, rc = 0; double r1 = 0.0, r2 = 0.0;
		for (uint32_t j = lane; j < nL + nH; j += 32) {
			const uint32_t p = j < nL ? sL + j : sH + (j - nL);
			const uint32_t k = sm.keys[p];
			const bool gone_lo = (j < nL) && k >= lo_key && k < new_lo;
			const bool gone_hi = (j >= nL) && k > new_hi && k <= hi_key;
			if (gone_lo) ++rn;
			if ((gone_lo || gone_hi) && p >= t0e && p < t1s) {
				const double d = (double)__uint_as_float(k) - pivot; ++rc; r1 += d; r2 = fma(d, d, r2);
			}
		}
		tn = 0; t1 = 0.0; t2 = 0.0;
		for (uint32_t j = lane; j < ntail; j += 32) {
			const uint32_t p = j < t0e ? j : t1s + (j - t0e);
			const uint32_t k = sm.keys[p];
			if (k >= new_lo && k <= new_hi) { const double d = (double)__uint_as_float(k) - pivot; ++tn; t1 += d; t2 = fma(d, d, t2); }
		}
		rn = __reduce_add_sync(0xffffffffu, rn); rc = __reduce_add_sync(0xffffffffu, rc);
		if (rc) { r1 = warp_sum_d(r1); r2 = warp_sum_d(r2); nc -= rc; s1c -= r1; s2c -= r2; }
		if (ntail) { tn = __reduce_add_sync(0xffffffffu, tn); t1 = warp_sum_d(t1); t2 = warp_sum_d(t2); }
		below += (uint32_t)rn;
		lo_key = new_lo; hi_key = new_hi;
	}

	// ---- final statistics: ORIGINAL valid values inside the last bounds
	if (empty_run || !hi_last_ok || lo_last > hi_last) return out;
	if ((converged || exhausted) && nested_last) {
		// last bounds lie inside the previous buffer range, so the final set IS the current buffer, whose
		// count / mean / median / std(ddof=0 about the mean) were just computed
		out.nfin = n; out.mean = mean; out.med = med; out.std = sd;
		return out;
	}
	// general case (bounds not nested): direct evaluation over the bin range of the last bounds
	{
		const int b0 = tw_bin(bm, __uint_as_float(lo_last)), b1 = tw_bin(bm, __uint_as_float(hi_last));
		const uint32_t s = tw_cstart(sm, b0), e = tw_cend(sm, b1);
		int fn = 0, nb = 0; double f1 = 0.0, f2 = 0.0;
		for (uint32_t p = s + lane; p < e; p += 32) {
			const uint32_t k = sm.keys[p];
			if (k < lo_last) ++nb;
			else if (k <= hi_last) { const double d = (double)__uint_as_float(k) - pivot; ++fn; f1 += d; f2 = fma(d, d, f2); }
		}
		fn = __reduce_add_sync(0xffffffffu, fn); nb = __reduce_add_sync(0xffffffffu, nb);
		f1 = warp_sum_d(f1); f2 = warp_sum_d(f2);
		out.nfin = fn;
		if (fn == 0) return out;
		const double m1 = f1 / (double)fn;
		out.mean = pivot + m1;
		out.std = sqrt(fmax(f2 / (double)fn - m1 * m1, 0.0));
		out.med = tw_median_at(sm, s + (uint32_t)nb + (uint32_t)((fn - 1) >> 1), (fn & 1) == 0, lane);
	}
	return out;
}
