// tbk_shenanigans.cu -- background-shenanigans detection (photometry/pixel_flags.py:61-79 and the driver in
// photometry/prepare.py:514-622):
//   1. indicator[k] = float32(median_filter(images[k] - SumImage, size=15))      (pixel_flags.py:74-77, prepare.py:537-549)
//   2. mean_shenanigans = mean over shuffled blocks of 25 images of nanmedian(block), NaN -> 0   (prepare.py:556-576)
//   3. flags[k] |= BackgroundShenanigans where abs(indicator[k] - mean) > threshold            (prepare.py:581-612)
//
// Step 1 is the heavy one (a 15 x 15 median of float64 differences for every pixel of every cadence).  Two exact
// reductions make it cheap:
//   * rounding to float32 is monotone, so the float32 cast of the float64 median (what the reference stores) equals
//     the median of the float32-rounded differences: the filter runs on 32-bit ordered keys;
//   * the windows of vertically adjacent pixels share 14 of their 15 rows, so the rank of the previous median in the
//     new window is at most 15 away from the new median.  Each CTA walks down a strip of columns, keeps the 15
//     window values of every column sorted in shared memory (one removal + one insertion per step), and a thread
//     finds its median by locating the previous one in the 15 sorted columns (15 binary searches) and stepping
//     to the wanted rank over the merged heads / tails (on average two or three steps).
// Windows that contain NaN (unspecified in the reference, see oracle/shenanigans_oracle.py) give the median of the
// non-NaN values (mean of the two middle values for an even count), all-NaN -> NaN.
#include "tbk_common.cuh"
#include "tbk_internal.h"
#include "tbk_zoom.cuh"   // reflect_fold (scipy mode='reflect' is the same half-sample symmetric fold)

#define SHE_R 7                        // window radius (size 15)
#define SHE_K (2 * SHE_R + 1)          // 15
#define SHE_W 64                       // output columns per CTA
#define SHE_NC (SHE_W + 2 * SHE_R)     // 78 window columns per CTA
#define SHE_CS 81                      // shared-memory row stride in words (odd: spreads data-dependent probes)
#define SHE_SEG 128                    // output rows per CTA
#define SHE_NT 128                     // threads per CTA
#define SHE_NANKEY 0xFFFFFFFFu

// float32 -> order-preserving uint32 (NaN -> largest key) and back
__device__ __forceinline__ uint32_t she_key(float f)
{
	if (f != f) return SHE_NANKEY;
	const uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float she_val(uint32_t k)
{
	return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct SheSmem {
	uint32_t col[SHE_K][SHE_CS];    // col[i][c]: i-th smallest key of window column c (NaN keys last)
	uint32_t ring[SHE_K][SHE_CS];   // ring[row mod 15][c]: key of (row, c) for the 15 rows of the window
	uint32_t kold[SHE_CS];          // key that left column c in the last slide ...
	uint32_t knew[SHE_CS];          // ... and the key that replaced it (equal: column unchanged)
	int nv[SHE_CS];                 // non-NaN keys of column c (first row only)
	uint32_t bd[SHE_K][SHE_W];      // per output column: the 15 keys next to its cut (selection scratch)
	int nanflag[2];                 // a NaN key entered or left the window in the slide to an odd / even row
};

__device__ __forceinline__ uint32_t she_load_key(const float* __restrict__ img, const double* __restrict__ sum,
	int H, int W, int y, int gx)
{
	const size_t off = (size_t)reflect_fold(y, H) * W + gx;
	const float v = __ldg(img + off);
	if (!sum) return she_key(v);
	return she_key((float)((double)v - __ldg(sum + off)));
}

// #keys <= pivot in the sorted column that starts at cj (stride SHE_CS)
__device__ __forceinline__ int she_count_le(const uint32_t* cj, uint32_t pivot)
{
	int p = (cj[7 * SHE_CS] <= pivot) ? 8 : 0;
	p += (cj[(p + 3) * SHE_CS] <= pivot) ? 4 : 0;
	p += (cj[(p + 1) * SHE_CS] <= pivot) ? 2 : 0;
	p += (cj[p * SHE_CS] <= pivot) ? 1 : 0;
	return p;
}

// max / min of 15 keys as a balanced tree (depth 4: the walk is a chain of dependent steps, so the latency of each
// step matters), and the index of a key known to be present
__device__ __forceinline__ uint32_t she_max15(const uint32_t (&v)[SHE_K])
{
	const uint32_t a0 = max(v[0], v[1]), a1 = max(v[2], v[3]), a2 = max(v[4], v[5]), a3 = max(v[6], v[7]);
	const uint32_t a4 = max(v[8], v[9]), a5 = max(v[10], v[11]), a6 = max(v[12], v[13]);
	return max(max(max(a0, a1), max(a2, a3)), max(max(a4, a5), max(a6, v[14])));
}
__device__ __forceinline__ uint32_t she_min15(const uint32_t (&v)[SHE_K])
{
	const uint32_t a0 = min(v[0], v[1]), a1 = min(v[2], v[3]), a2 = min(v[4], v[5]), a3 = min(v[6], v[7]);
	const uint32_t a4 = min(v[8], v[9]), a5 = min(v[10], v[11]), a6 = min(v[12], v[13]);
	return min(min(min(a0, a1), min(a2, a3)), min(min(a4, a5), min(a6, v[14])));
}
__device__ __forceinline__ int she_find15(const uint32_t (&v)[SHE_K], uint32_t key)
{
	unsigned mask = 0u;
#pragma unroll
	for (int j = 0; j < SHE_K; ++j) mask |= (unsigned)(v[j] == key) << j;
	return __ffs(mask) - 1;
}

// The selection state of an output column lives in registers from row to row: ``pp`` holds, 4 bits per window
// column, how many keys of that column lie below the cut; the cut is valid when every key below it is <= every key
// above it, ``r`` = number of keys below, ``a`` = the previous median (largest key below the cut).
__global__ void __launch_bounds__(SHE_NT) k_bkgshe_median(const float* __restrict__ images,
	const double* __restrict__ sum, int H, int W, float* __restrict__ out)
{
	__shared__ SheSmem sm;
	const int tid = threadIdx.x;
	const int x0 = blockIdx.x * SHE_W, y0 = blockIdx.y * SHE_SEG;
	const size_t img_off = (size_t)blockIdx.z * H * W;
	const float* img = images + img_off;
	const int y1 = min(y0 + SHE_SEG, H);

	if (tid < 2) sm.nanflag[tid] = 0;
	// ---- window of the first row: rows y0-7 .. y0+7 of every column, sorted
	int gx = 0;
	if (tid < SHE_NC) {
		gx = reflect_fold(x0 - SHE_R + tid, W);
		uint32_t a[SHE_K];
#pragma unroll
		for (int i = 0; i < SHE_K; ++i) {
			const int row = y0 - SHE_R + i;
			a[i] = she_load_key(img, sum, H, W, row, gx);
			sm.ring[(row + 2 * SHE_K) % SHE_K][tid] = a[i];
		}
		// insertion sort in registers (static indices)
#pragma unroll
		for (int i = 1; i < SHE_K; ++i) {
#pragma unroll
			for (int j = i; j > 0; --j) {
				const uint32_t lo = min(a[j - 1], a[j]), hi = max(a[j - 1], a[j]);
				a[j - 1] = lo; a[j] = hi;
			}
		}
		int nv = 0;
#pragma unroll
		for (int i = 0; i < SHE_K; ++i) { sm.col[i][tid] = a[i]; nv += a[i] != SHE_NANKEY; }
		sm.nv[tid] = nv;
	}
	__syncthreads();

	const bool sel = tid < SHE_W && x0 + tid < W;
	uint32_t plo = 0u, phi = 0u;   // cut position per window column, 4 bits each: columns 0..7 / 8..14
	int r = 0, m = 0;
	uint32_t a = 0u;
	if (sel) {
		// first row: cut at the median of the centre column
		a = sm.col[SHE_R][tid + SHE_R];
#pragma unroll
		for (int j = 0; j < SHE_K; ++j) {
			m += sm.nv[tid + j];
			const int p = she_count_le(&sm.col[0][tid + j], a);
			if (j < 8) plo |= (uint32_t)p << (4 * j); else phi |= (uint32_t)p << (4 * (j - 8));
			r += p;
		}
	}

	for (int y = y0; y < y1; ++y) {
		if (sel) {
			float res = nan_f();
			if (m == 0) { plo = phi = 0u; r = 0; }   // nothing but NaN: every key is above the cut
			else {
				// Walk the cut to t + 1 keys below it.  Down: the largest key below the cut moves above it; up: the
				// smallest key above moves below.  Keys of lanes that walk up are complemented, so both directions take
				// a maximum and the lanes of a warp share one loop.  bd[j][tid] = key of column j next to the cut on the
				// side the walk eats from (0 = none).
				const int t = (m - 1) >> 1;
				const bool dn = r >= t + 1;
				const uint32_t flip = dn ? 0u : 0xFFFFFFFFu;
				const int off = dn ? -1 : 0;
				int steps = dn ? r - (t + 1) : (t + 1) - r;
				const bool moved = steps > 0;
#pragma unroll
				for (int j = 0; j < SHE_K; ++j) {
					const int pj = (int)((j < 8 ? plo >> (4 * j) : phi >> (4 * (j - 8))) & 15u);
					const int idx = pj + off;
					sm.bd[j][tid] = (idx >= 0 && idx < SHE_K) ? (sm.col[idx][tid + j] ^ flip) : 0u;
				}
				uint32_t last = 0u;
				while (steps > 0) {
					uint32_t v[SHE_K];
#pragma unroll
					for (int j = 0; j < SHE_K; ++j) v[j] = sm.bd[j][tid];
					const uint32_t best = she_max15(v);
					int arg = SHE_K - 1;
#pragma unroll
					for (int j = SHE_K - 2; j >= 0; --j) arg = (v[j] == best) ? j : arg;
					const int sh = (arg & 7) << 2;
					const uint32_t inc = dn ? (0u - (1u << sh)) : (1u << sh);   // -1 / +1 in the nibble (it stays within 0..15)
					if (arg < 8) plo += inc; else phi += inc;
					last = best;
					const int pj = (int)(((arg < 8 ? plo : phi) >> sh) & 15u);
					const int idx = pj + off;
					sm.bd[arg][tid] = (idx >= 0 && idx < SHE_K) ? (sm.col[idx][tid + arg] ^ flip) : 0u;
					--steps;
				}
				r = t + 1;
				uint32_t v[SHE_K];
#pragma unroll
				for (int j = 0; j < SHE_K; ++j) v[j] = sm.bd[j][tid];
				const uint32_t top = she_max15(v);
				a = dn ? top : ~last;                 // largest key below the cut
				if (m & 1) res = she_val(a);
				else {
					uint32_t b;                       // smallest key above the cut
					if (!dn) b = ~top;
					else if (moved) b = last;
					else {
						b = SHE_NANKEY;
#pragma unroll
						for (int j = 0; j < SHE_K; ++j) {
							const int pj = (int)((j < 8 ? plo >> (4 * j) : phi >> (4 * (j - 8))) & 15u);
							if (pj < SHE_K) b = min(b, sm.col[pj][tid + j]);
						}
					}
					res = (float)(0.5 * ((double)she_val(a) + (double)she_val(b)));
				}
			}
			out[img_off + (size_t)y * W + x0 + tid] = res;
		}
		if (y + 1 >= y1) break;
		__syncthreads();
		if (tid == SHE_NT - 1) sm.nanflag[y & 1] = 0;   // last read in the previous iteration's carry
		if (tid < SHE_NC) {
			// slide the window down: row y+8 replaces row y-7 in every column
			const int rnew = y + 1 + SHE_R, slot = (rnew + 2 * SHE_K) % SHE_K;
			const uint32_t kold = sm.ring[slot][tid];
			const uint32_t knew = she_load_key(img, sum, H, W, rnew, gx);
			sm.ring[slot][tid] = knew;
			sm.kold[tid] = kold; sm.knew[tid] = knew;
			if ((kold == SHE_NANKEY) != (knew == SHE_NANKEY)) sm.nanflag[(y + 1) & 1] = 1;
			if (knew != kold) {
				uint32_t c[SHE_K];
				int pold = 0;
#pragma unroll
				for (int i = 0; i < SHE_K; ++i) { c[i] = sm.col[i][tid]; pold += c[i] < kold; }
				int q = 0;   // position of the new key in the list without the old one
#pragma unroll
				for (int i = 0; i < SHE_K; ++i) q += (i != pold) && (c[i] < knew);
#pragma unroll
				for (int i = 0; i < SHE_K; ++i) {
					// element i of the new list = the list without c[pold], with knew inserted at position q:
					// below q it is reduced[i], above q reduced[i - 1], where reduced[s] = s < pold ? c[s] : c[s + 1]
					const uint32_t up = c[i < SHE_K - 1 ? i + 1 : i], dn = c[i > 0 ? i - 1 : 0];
					const uint32_t lo = (i < pold) ? c[i] : up;          // reduced[i]
					const uint32_t hi = (i - 1 < pold) ? dn : c[i];      // reduced[i - 1]
					sm.col[i][tid] = (i == q) ? knew : (i < q ? lo : hi);
				}
			}
		}
		__syncthreads();
		if (sel) {
			// carry the cut over to the new window: a key that left from below the cut takes one off, a key that
			// arrived below it adds one; a key equal to the previous median may sit on either side, so that column is
			// cut afresh (keys <= a below)
			const bool nanrow = sm.nanflag[(y + 1) & 1] != 0;
#pragma unroll
			for (int j = 0; j < SHE_K; ++j) {
				const uint32_t ko = sm.kold[tid + j], kn = sm.knew[tid + j];
				if (nanrow) m += (int)(kn != SHE_NANKEY) - (int)(ko != SHE_NANKEY);
				int d = (int)(kn < a) - (int)(ko < a);
				if (ko == a || kn == a) {
					const int pj = (int)((j < 8 ? plo >> (4 * j) : phi >> (4 * (j - 8))) & 15u);
					d = she_count_le(&sm.col[0][tid + j], a) - pj;
				}
				if (j < 8) plo += (uint32_t)d << (4 * j); else phi += (uint32_t)d << (4 * (j - 8));
				r += d;
			}
		}
	}
}

int tbk_launch_bkgshe_indicator(const float* images, const double* sum, int B, int H, int W, float* out, cudaStream_t st)
{
	dim3 grid((W + SHE_W - 1) / SHE_W, (H + SHE_SEG - 1) / SHE_SEG, B);
	k_bkgshe_median<<<grid, SHE_NT, 0, st>>>(images, sum, H, W, out);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_bkgshe_median: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
// Step 2.  One thread per pixel walks the shuffled order in blocks of 25; the 25 slots persist between blocks, so the
// last, partial block keeps the trailing images of the block before it (the reference's buffer is allocated once,
// prepare.py:564); slots never written (stacks shorter than a block) hold 0.  nanmedian by ranking (exact, no
// dynamic register indexing), NaN -> 0, float64 accumulation in block order.
#define SHE_BLOCK 25
__global__ void __launch_bounds__(128) k_bkgshe_mean(const float* __restrict__ ind, size_t stride, size_t npix,
	int n, const int* __restrict__ order, double* __restrict__ mean)
{
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npix) return;
	float v[SHE_BLOCK];
#pragma unroll
	for (int j = 0; j < SHE_BLOCK; ++j) v[j] = 0.f;
	double acc = 0.0;
	int nblocks = 0;
	for (int k = 0; k < n; k += SHE_BLOCK, ++nblocks) {
#pragma unroll
		for (int j = 0; j < SHE_BLOCK; ++j)
			if (k + j < n) v[j] = __ldg(ind + (size_t)__ldg(order + k + j) * stride + p);
		uint32_t key[SHE_BLOCK];
		int nv = 0;
#pragma unroll
		for (int j = 0; j < SHE_BLOCK; ++j) { key[j] = she_key(v[j]); nv += key[j] != SHE_NANKEY; }
		if (nv == 0) continue;   // all NaN -> NaN -> 0
		const int t0 = (nv - 1) >> 1, t1 = nv >> 1;
		float a = 0.f, b = 0.f;
#pragma unroll
		for (int i = 0; i < SHE_BLOCK; ++i) {
			int rk = 0;
#pragma unroll
			for (int j = 0; j < SHE_BLOCK; ++j) rk += (key[j] < key[i]) || (key[j] == key[i] && j < i);
			if (rk == t0) a = v[i];
			if (rk == t1) b = v[i];
		}
		acc += (t0 == t1) ? (double)a : 0.5 * ((double)a + (double)b);
	}
	mean[p] = acc / (double)nblocks;
}

int tbk_launch_bkgshe_mean(const float* ind, size_t stride, size_t npix, int n, const int* order, double* mean, cudaStream_t st)
{
	k_bkgshe_mean<<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(ind, stride, npix, n, order, mean);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_bkgshe_mean: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
// Step 3.  flags[k] = (flags[k] & ~bit) | (abs(indicator[k] - mean) > threshold ? bit : 0); NaN compares false.
__global__ void __launch_bounds__(256) k_bkgshe_flag(const float* __restrict__ ind, const double* __restrict__ mean,
	size_t npix, double threshold, uint8_t bit, uint8_t* __restrict__ flags)
{
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npix) return;
	const size_t off = (size_t)blockIdx.y * npix + p;
	const double d = fabs((double)__ldg(ind + off) - __ldg(mean + p));
	uint8_t f = flags[off] & (uint8_t)~bit;
	if (d > threshold) f |= bit;
	flags[off] = f;
}

// four pixels per thread (npix % 4 == 0, 16-byte aligned stacks)
__global__ void __launch_bounds__(256) k_bkgshe_flag4(const float4* __restrict__ ind, const double2* __restrict__ mean,
	size_t npix4, double threshold, uint8_t bit, uchar4* __restrict__ flags)
{
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npix4) return;
	const size_t off = (size_t)blockIdx.y * npix4 + p;
	const float4 v = __ldg(ind + off);
	const double2 m0 = __ldg(mean + 2 * p), m1 = __ldg(mean + 2 * p + 1);
	uchar4 f = flags[off];
	const uint8_t clr = (uint8_t)~bit;
	f.x = (f.x & clr) | (fabs((double)v.x - m0.x) > threshold ? bit : 0);
	f.y = (f.y & clr) | (fabs((double)v.y - m0.y) > threshold ? bit : 0);
	f.z = (f.z & clr) | (fabs((double)v.z - m1.x) > threshold ? bit : 0);
	f.w = (f.w & clr) | (fabs((double)v.w - m1.y) > threshold ? bit : 0);
	flags[off] = f;
}

int tbk_launch_bkgshe_flag(const float* ind, const double* mean, int B, size_t npix, double threshold, int bit,
	uint8_t* flags, cudaStream_t st)
{
	if (npix % 4 == 0 && (((uintptr_t)ind | (uintptr_t)mean) & 15) == 0 && ((uintptr_t)flags & 3) == 0) {
		dim3 grid((unsigned)((npix / 4 + 255) / 256), B);
		k_bkgshe_flag4<<<grid, 256, 0, st>>>((const float4*)ind, (const double2*)mean, npix / 4, threshold, (uint8_t)bit, (uchar4*)flags);
	} else {
		dim3 grid((unsigned)((npix + 255) / 256), B);
		k_bkgshe_flag<<<grid, 256, 0, st>>>(ind, mean, npix, threshold, (uint8_t)bit, flags);
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_bkgshe_flag: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}
