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
#define TW_INVALID 0x7f800000u   // +inf as key: masked pixel (float32 keys)

// counter word w lives at w + (w >> 5): one pad word per 32 keeps the per-lane chunked scan (lane l owns
// words 32l .. 32l+31) free of shared-memory bank conflicts
#define TW_CIDX(w) ((w) + ((w) >> 5))
#define TW_CNT_WORDS (TW_WORDS + TW_WORDS / 32)

template <typename K>
struct TwSmem {
	K keys[TBK_NPIX_TILE];          // bucketed keys
	uint32_t cnt[TW_CNT_WORDS];     // packed uint16 pairs: counts -> starts -> ends
};
typedef TwSmem<uint32_t> TileWarpSmem;

struct TwBinMap {
	float scale, off;
};
struct TwBinMap64 {
	double w0;              // window start; bins are laid over [w0, w0 + span]
	float scale, lo, hi;    // bins per unit, clamp range of (value - w0)
};

// ---- key traits: float32 pixels (non-negative) and float64 residuals (any sign) ---------------
struct TwF32 {
	typedef uint32_t K;
	typedef TwBinMap Map;
	static constexpr int NBITS = 32;
	__device__ static __forceinline__ K invalid() { return TW_INVALID; }
	__device__ static __forceinline__ K minkey() { return 0u; }
	__device__ static __forceinline__ K maxkey() { return 0x7f7fffffu; }
	__device__ static __forceinline__ K padkey() { return 0xFFFFFFFFu; }
	__device__ static __forceinline__ double val(K k) { return (double)__uint_as_float(k); }
	// Monotone non-decreasing map value -> bin: one FFMA onto the 2^23 "magic" range, then an integer clamp.
	// off >= 2^22 and x >= 0, scale > 0 guarantee t > 0, so the float bit pattern is monotone in t.
	__device__ static __forceinline__ int bin(const Map& m, K k)
	{
		const float t = fmaf(__uint_as_float(k), m.scale, m.off);
		const int b = __float_as_int(t) - 0x4B000000;
		return max(0, min(TW_NB - 1, b));
	}
	// smallest float32 >= d / largest float32 <= d as keys (exact translation of a float64 bound)
	__device__ static __forceinline__ Map make_map(double w0d, double w1d)
	{
		const float w0 = (float)w0d, w1 = (float)w1d;
		Map bm;
		bm.scale = (float)(TW_NB - 2) / (w1 - w0);
		if (!(bm.scale < 1e30f)) bm.scale = 1e30f;
		if (w0 > 0.f && w0 * bm.scale > 4194304.0f) bm.scale = 4194304.0f / w0;  // keeps off >= 2^22
		bm.off = fmaf(-w0, bm.scale, 8388609.0f);
		return bm;
	}
	__device__ static __forceinline__ double pivot_of(double med) { return (double)fmaxf((float)med, 0.f); }
	__device__ static __forceinline__ K key_ceil(double d)
	{
		if (!(d > 0.0)) return 0u;
		float f = (float)d;
		if ((double)f < d) f = __uint_as_float(__float_as_uint(f) + 1u);
		return min(__float_as_uint(f), 0x7f7fffffu + 1u);
	}
	__device__ static __forceinline__ bool key_floor(double d, K& key)
	{
		if (d < 0.0) return false;                // nothing is <= a negative bound
		if (d == 0.0) { key = 0u; return true; }  // +-0: only zeros qualify
		float f = (float)d;
		if ((double)f > d) f = __uint_as_float(__float_as_uint(f) - 1u);
		key = min(__float_as_uint(f), 0x7f7fffffu);
		return true;
	}
};

struct TwF64 {
	typedef unsigned long long K;
	typedef TwBinMap64 Map;
	static constexpr int NBITS = 64;
	__device__ static __forceinline__ K invalid() { return ~0ULL; }
	__device__ static __forceinline__ K minkey() { return 0ULL; }
	__device__ static __forceinline__ K maxkey() { return 0xFFEFFFFFFFFFFFFFULL; }  // dkey(DBL_MAX)
	__device__ static __forceinline__ K padkey() { return ~0ULL; }
	__device__ static __forceinline__ double val(K k) { return dkey_inv(k); }
	// Monotone non-decreasing map value -> bin.  Only the partition has to be monotone (membership tests inside
	// a bin compare the float64 keys), so the offset from the window start is rounded to float32 and binned with
	// one FFMA onto the 2^23 "magic" range like the float32 keys.
	__device__ static __forceinline__ int bin(const Map& m, K k)
	{
		const float e = (float)(dkey_inv(k) - m.w0);
		const float t = fmaf(fminf(fmaxf(e, m.lo), m.hi), m.scale, 8388609.0f);   // 2^23 + 1 + [-1.5, NB - 0.5]
		const int b = __float_as_int(t) - 0x4B000000;
		return max(0, min(TW_NB - 1, b));
	}
	__device__ static __forceinline__ Map make_map(double w0, double w1)
	{
		Map bm;
		float span = (float)(w1 - w0);
		if (!(span < 1e30f)) span = 1e30f;
		bm.w0 = w0;
		bm.scale = (float)(TW_NB - 2) / span;
		if (!(bm.scale < 1e30f)) bm.scale = 1e30f;
		bm.lo = -1.5f / bm.scale; bm.hi = span + 1.5f / bm.scale;
		return bm;
	}
	__device__ static __forceinline__ double pivot_of(double med) { return med; }
	__device__ static __forceinline__ K key_ceil(double d) { return dkey(d); }
	__device__ static __forceinline__ bool key_floor(double d, K& key) { key = dkey(d); return true; }
};

__device__ __forceinline__ uint32_t tw_cend(const uint32_t* cnt, int b)
{
	return (cnt[TW_CIDX(b >> 1)] >> ((b & 1) << 4)) & 0xFFFFu;
}
__device__ __forceinline__ uint32_t tw_cstart(const uint32_t* cnt, int b)
{
	return b ? tw_cend(cnt, b - 1) : 0u;
}

// smallest bin b with cend(b) > P  (P < total)
__device__ __forceinline__ int tw_find_bin(const uint32_t* cnt, uint32_t P, int lane)
{
	unsigned m = __ballot_sync(0xffffffffu, tw_cend(cnt, 64 * lane + 63) > P);
	const int g = __ffs(m) - 1;
	m = __ballot_sync(0xffffffffu, tw_cend(cnt, 64 * g + 2 * lane + 1) > P);
	const int h = __ffs(m) - 1;
	const int b = 64 * g + 2 * h;
	return (tw_cend(cnt, b) > P) ? b : b + 1;
}

template <typename K>
__device__ __forceinline__ K warp_bitonic32(K v, int lane)
{
#pragma unroll
	for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			const K o = __shfl_xor_sync(0xffffffffu, v, j);
			const bool up = (lane & k) == 0, lower = (lane & j) == 0;
			v = (lower == up) ? min(v, o) : max(v, o);
		}
	}
	return v;
}

// keys at sorted ranks q and q+1 (when want2) among the span [s, e) of one bin (unsorted inside).
template <typename T, typename Keys>
__device__ __noinline__ void tw_select_in_span(const Keys& keys, uint32_t s, uint32_t e, uint32_t q, bool want2,
	int lane, typename T::K& k1, typename T::K& k2)
{
	typedef typename T::K K;
	const uint32_t m = e - s;
	if (m <= 32u) {
		K v = (s + lane < e) ? keys[s + lane] : T::padkey();
		v = warp_bitonic32<K>(v, lane);
		k1 = __shfl_sync(0xffffffffu, v, q);
		k2 = want2 ? __shfl_sync(0xffffffffu, v, min(q + 1u, 31u)) : k1;
		return;
	}
	// large bin: exact radix selection on the key bits (MSB first), one sweep of the span per bit
	for (int r = 0; r < (want2 ? 2 : 1); ++r) {
		uint32_t target = q + r;
		K prefix = 0;
		for (int bit = T::NBITS - 1; bit >= 0; --bit) {
			const K mask = ~((((K)1) << bit) - (K)1);     // bits above and including ``bit``
			int c0 = 0;
			for (uint32_t p = s + lane; p < e; p += 32) c0 += ((keys[p] & mask) == prefix) ? 1 : 0;
			c0 = __reduce_add_sync(0xffffffffu, c0);
			if (target >= (uint32_t)c0) { target -= c0; prefix |= (((K)1) << bit); }
		}
		if (r == 0) k1 = prefix; else k2 = prefix;
	}
	if (!want2) k2 = k1;
}

// median (mean of the values at sorted ranks P and P+1 when ``even``) by bin lookup
template <typename T, typename Keys>
__device__ __forceinline__ double tw_median_at(const Keys& keys, const uint32_t* cnt, uint32_t P, bool even, int lane)
{
	typedef typename T::K K;
	const int b = tw_find_bin(cnt, P, lane);
	const uint32_t bs = tw_cstart(cnt, b), be = tw_cend(cnt, b);
	K k1, k2;
	const bool second_here = even && (P + 1u < be);
	tw_select_in_span<T, Keys>(keys, bs, be, P - bs, second_here, lane, k1, k2);
	if (even && !second_here) {
		const int b2 = tw_find_bin(cnt, P + 1u, lane);
		K dummy;
		tw_select_in_span<T, Keys>(keys, tw_cstart(cnt, b2), tw_cend(cnt, b2), 0u, false, lane, k2, dummy);
	}
	return 0.5 * (T::val(k1) + T::val(k2));
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// Exclusive scan of the packed counters by one warp (32 words = 64 bins per lane): counts -> bin starts.
__device__ __forceinline__ void tw_scan_counts(uint32_t* cnt, int lane)
{
	uint32_t tot = 0;
#pragma unroll 8
	for (int j = 0; j < 32; ++j) { const uint32_t w = cnt[lane * 33 + j]; tot += (w & 0xFFFFu) + (w >> 16); }
	uint32_t inc = tot;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
	uint32_t run = inc - tot;
#pragma unroll 8
	for (int j = 0; j < 32; ++j) {
		const uint32_t w = cnt[lane * 33 + j];
		const uint32_t c0 = w & 0xFFFFu, c1 = w >> 16;
		cnt[lane * 33 + j] = run | ((run + c0) << 16);
		run += c0 + c1;
	}
}

// The clip iterations + final statistics on bucketed keys (one warp).  Inputs: the moments about
// ``pivot`` of the core bins [t0e, t1s) (s1c, s2c) and of the two overflow bins (tn, t1, t2).
template <typename T, typename Keys>
__device__ TileStat tw_iterate(const Keys& keys, const uint32_t* cnt, const typename T::Map& bm,
	int nvalid, double pivot, double s1c, double s2c, int tn, double t1, double t2, int lane)
{
	typedef typename T::K K;
	TileStat out;
	out.mean = out.med = out.std = nan_d();
	out.nfin = 0; out.pad = 0;
	const uint32_t t0e = tw_cend(cnt, 0);                // [0, t0e)       = low overflow bin
	const uint32_t t1s = tw_cstart(cnt, TW_NB - 1);      // [t1s, nvalid)  = high overflow bin
	const uint32_t ntail = t0e + ((uint32_t)nvalid - t1s);
	int nc = (int)(t1s - t0e);

	K lo_key = T::minkey(), hi_key = T::maxkey();        // running intersection (inclusive)
	K lo_last = T::minkey(), hi_last = T::maxkey();      // last bounds (inclusive)
	bool hi_last_ok = true, nested_last = true, converged = false, exhausted = false, empty_run = false;
	uint32_t below = 0u;                                 // valid elements with key < lo_key
	int n_prev = -1, n = 0;
	double med = 0.0, mean = 0.0, sd = 0.0;

	for (int it = 0; it < 6; ++it) {
		const int ncur = nc + tn;
		if (it > 0 && ncur == n_prev) { converged = true; break; }   // stats of the previous pass stand
		n = ncur;
		if (n == 0) { empty_run = true; break; }
		const double m1 = (s1c + t1) / (double)n;
		mean = pivot + m1;
		sd = sqrt(fmax((s2c + t2) / (double)n - m1 * m1, 0.0));
		med = tw_median_at<T, Keys>(keys, cnt, below + (uint32_t)((n - 1) >> 1), (n & 1) == 0, lane);
		if (it == 5) { exhausted = true; break; }   // five bound computations done: buffer after the last clip
		lo_last = T::key_ceil(med - 3.0 * sd);
		hi_last_ok = T::key_floor(med + 3.0 * sd, hi_last);
		nested_last = (lo_last >= lo_key) && hi_last_ok && (hi_last <= hi_key);
		const K new_lo = max(lo_key, lo_last);
		const K new_hi = hi_last_ok ? min(hi_key, hi_last) : T::minkey();
		n_prev = n;
		if (!hi_last_ok || new_lo > new_hi) { empty_run = true; break; }  // cannot happen for finite data
		if (new_lo == lo_key && new_hi == hi_key) continue;               // nothing can leave: next pass converges
		// one sweep: (a) core-bin elements leaving the buffer, (b) everything removed below (for ``below``),
		// (c) the overflow-bin elements still inside the new bounds
		uint32_t sL = 0, eL = 0, sH = 0, eH = 0;
		if (new_lo > lo_key) { sL = tw_cstart(cnt, T::bin(bm, lo_key)); eL = tw_cend(cnt, T::bin(bm, new_lo)); }
		if (new_hi < hi_key) { sH = tw_cstart(cnt, T::bin(bm, new_hi)); eH = tw_cend(cnt, T::bin(bm, hi_key)); }
		const uint32_t nL = eL - sL, nH = eH - sH;
		int rn = 0, rc = 0; double r1 = 0.0, r2 = 0.0;
		for (uint32_t j = lane; j < nL + nH; j += 32) {
			const uint32_t p = j < nL ? sL + j : sH + (j - nL);
			const K k = keys[p];
			const bool gone_lo = (j < nL) && k >= lo_key && k < new_lo;
			const bool gone_hi = (j >= nL) && k > new_hi && k <= hi_key;
			if (gone_lo) ++rn;
			if ((gone_lo || gone_hi) && p >= t0e && p < t1s) {
				const double d = T::val(k) - pivot; ++rc; r1 += d; r2 = fma(d, d, r2);
			}
		}
		tn = 0; t1 = 0.0; t2 = 0.0;
		for (uint32_t j = lane; j < ntail; j += 32) {
			const uint32_t p = j < t0e ? j : t1s + (j - t0e);
			const K k = keys[p];
			if (k >= new_lo && k <= new_hi) { const double d = T::val(k) - pivot; ++tn; t1 += d; t2 = fma(d, d, t2); }
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
		const int b0 = T::bin(bm, lo_last), b1 = T::bin(bm, hi_last);
		const uint32_t s = tw_cstart(cnt, b0), e = tw_cend(cnt, b1);
		int fn = 0, nb = 0; double f1 = 0.0, f2 = 0.0;
		for (uint32_t p = s + lane; p < e; p += 32) {
			const K k = keys[p];
			if (k < lo_last) ++nb;
			else if (k <= hi_last) { const double d = T::val(k) - pivot; ++fn; f1 += d; f2 = fma(d, d, f2); }
		}
		fn = __reduce_add_sync(0xffffffffu, fn); nb = __reduce_add_sync(0xffffffffu, nb);
		f1 = warp_sum_d(f1); f2 = warp_sum_d(f2);
		out.nfin = fn;
		if (fn == 0) return out;
		const double m1 = f1 / (double)fn;
		out.mean = pivot + m1;
		out.std = sqrt(fmax(f2 / (double)fn - m1 * m1, 0.0));
		out.med = tw_median_at<T, Keys>(keys, cnt, s + (uint32_t)nb + (uint32_t)((fn - 1) >> 1), (fn & 1) == 0, lane);
	}
	return out;
}

// ---------------------------------------------------------------------------------------------
// Block version: NW warps share one mesh.  The warps cooperate on the two bucketing passes and the moment sweep;
// warp 0 alone runs the iterations.  Used with NW = 4 for float64 residuals (x - radial) and NW = 2 for float32 pixels.
template <typename T, int NW>
struct TwBlockSmem {
	TwSmem<typename T::K> tw;
	typename T::Map bm;
	double pivot;
	double red[2][NW][2];
	typename T::K kmin, kmax;
	int nvalid, ntl[NW], constant;
};

// ---------------------------------------------------------------------------------------------
// The (validated) keys of the mesh sit in sm.tw.keys in any order when this is called (INVALID entries allowed),
// nvalid is known.  The per-element passes are rolled loops over shared memory (a fully unrolled register version is
// ~115 KB of SASS and stalls on instruction fetch); only the in-place scatter holds the keys in registers.
template <typename T, int NW>
__device__ void tile_block_stats_staged(TwBlockSmem<T, NW>& sm, int nvalid, TileStat& out, bool& writer)
{
	typedef typename T::K K;
	constexpr int NT = 32 * NW, VPT = TBK_NPIX_TILE / NT;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const K INV = T::invalid();
	writer = (tid == 0);
	out.mean = out.med = out.std = nan_d();
	out.nfin = 0; out.pad = 0;
	if (nvalid == 0) return;
	for (int i = tid; i < TW_CNT_WORDS; i += NT) sm.tw.cnt[i] = 0u;
	if (tid == 0) sm.constant = 0;
	// ---- robust window from 2 x 32 samples spread over the mesh (warp 0)
	if (w == 0) {
		double med = 0.0, iqr = 0.0; int sets = 0;
#pragma unroll
		for (int t = 0; t < 2; ++t) {
			const K k = warp_bitonic32<K>(sm.tw.keys[(lane * 131 + 17 + t * 2053) & (TBK_NPIX_TILE - 1)], lane);
			const int m = __popc(__ballot_sync(0xffffffffu, k != INV));
			if (m >= 8) {
				med += T::val(__shfl_sync(0xffffffffu, k, m >> 1));
				iqr += T::val(__shfl_sync(0xffffffffu, k, (3 * m) >> 2)) - T::val(__shfl_sync(0xffffffffu, k, m >> 2));
				++sets;
			}
		}
		double w0, w1, pv;
		if (sets && iqr > 0.0) {
			med /= (double)sets;
			const double half = 10.0 * (iqr / (double)sets) / 1.349;
			w0 = med - half; w1 = med + half; pv = T::pivot_of(med);
		} else {
			// degenerate sample: full data range (constant meshes end here)
			K kmin = T::padkey(), kmax = 0;
			for (int i = lane; i < TBK_NPIX_TILE; i += 32) { const K k = sm.tw.keys[i]; if (k != INV) { kmin = min(kmin, k); kmax = max(kmax, k); } }
			for (int o = 16; o > 0; o >>= 1) { kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o)); kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o)); }
			w0 = T::val(kmin); w1 = T::val(kmax); pv = w0;
			if (lane == 0) { sm.kmin = kmin; if (!(w1 > w0)) sm.constant = 1; }
		}
		if (lane == 0) { sm.bm = T::make_map(w0, w1); sm.pivot = pv; }
	}
	__syncthreads();
	if (sm.constant) {  // all values equal: sigma = 0, nothing is clipped
		out.mean = out.med = T::val(sm.kmin); out.std = 0.0; out.nfin = nvalid;
		return;
	}
	const typename T::Map bm = sm.bm;
	const double pivot = sm.pivot;
	// ---- pass 1: counts (rolled)
#pragma unroll 4
	for (int j = 0; j < VPT; ++j) {
		const K k = sm.tw.keys[tid + NT * j];
		const int b = T::bin(bm, k);
		if (k != INV) atomicAdd(&sm.tw.cnt[TW_CIDX(b >> 1)], 1u << ((b & 1) << 4));
	}
	// ---- the keys move to registers for the in-place scatter
	K key[VPT];
#pragma unroll
	for (int j = 0; j < VPT; ++j) key[j] = sm.tw.keys[tid + NT * j];
	__syncthreads();
	if (w == 0) tw_scan_counts(sm.tw.cnt, lane);
	__syncthreads();
	// ---- pass 2: scatter in groups of 8
#pragma unroll
	for (int g = 0; g < VPT / 8; ++g) {
		uint32_t old[8]; int sh[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const K k = key[8 * g + j];
			const int b = T::bin(bm, k);
			sh[j] = (b & 1) << 4;
			old[j] = 0u;
			if (k != INV) old[j] = atomicAdd(&sm.tw.cnt[TW_CIDX(b >> 1)], 1u << sh[j]);
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const K k = key[8 * g + j];
			if (k != INV) sm.tw.keys[(old[j] >> sh[j]) & 0xFFFFu] = k;
		}
	}
	__syncthreads();
	// ---- moments: all warps sweep the bucketed keys
	const uint32_t t0e = tw_cend(sm.tw.cnt, 0), t1s = tw_cstart(sm.tw.cnt, TW_NB - 1);
	double c1 = 0.0, c2 = 0.0, q1 = 0.0, q2 = 0.0; int tn = 0;
	for (uint32_t p = tid; p < (uint32_t)nvalid; p += NT) {
		const double d = T::val(sm.tw.keys[p]) - pivot;
		if (p >= t0e && p < t1s) { c1 += d; c2 = fma(d, d, c2); }
		else { ++tn; q1 += d; q2 = fma(d, d, q2); }
	}
	c1 = warp_sum_d(c1); c2 = warp_sum_d(c2);
	tn = __reduce_add_sync(0xffffffffu, tn);
	if (__any_sync(0xffffffffu, tn != 0)) { q1 = warp_sum_d(q1); q2 = warp_sum_d(q2); }
	if (lane == 0) { sm.red[0][w][0] = c1; sm.red[0][w][1] = c2; sm.red[1][w][0] = q1; sm.red[1][w][1] = q2; sm.ntl[w] = tn; }
	__syncthreads();
	if (w != 0) { writer = false; return; }
	double s1c = 0.0, s2c = 0.0, t1 = 0.0, t2 = 0.0; tn = 0;
#pragma unroll
	for (int q = 0; q < NW; ++q) { s1c += sm.red[0][q][0]; s2c += sm.red[0][q][1]; t1 += sm.red[1][q][0]; t2 += sm.red[1][q][1]; tn += sm.ntl[q]; }
	out = tw_iterate<T>(sm.tw.keys, sm.tw.cnt, bm, nvalid, pivot, s1c, s2c, tn, t1, t2, lane);
}
