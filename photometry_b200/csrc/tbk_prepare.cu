// tbk_prepare.cu -- the prepare-stage loops around fit_background: background time smoothing
// (photometry/prepare.py:317-335) and the fused final per-image loop / sumimage accumulation
// (prepare.py:408-470).  Pure streaming kernels: HBM-bound, float4 / uchar4 accesses.
#include "tbk_common.cuh"
#include "tbk_internal.h"

// ---------------------------------------------------------------------------------------------
// out[k] = bottleneck.nanmean(float32 block[k-w .. k+w], axis=2): float32 accumulator, frames added
// in index order, divided by the non-NaN count (all-NaN -> NaN).  Grid: x = cadence (fastest, so the
// 2w+1 CTAs that share an input line run together and hit L2), y = pixel chunk.
__device__ __forceinline__ void nanacc(float v, float& s, int& c) { if (v == v) { s += v; ++c; } }

__global__ void __launch_bounds__(256) k_time_smooth(size_t npix4, const float4* __restrict__ bkg, int n, int w,
	const float4* __restrict__ halo_lo, int n_lo, const float4* __restrict__ halo_hi, int n_hi,
	float4* __restrict__ out)
{
	const int k = blockIdx.x;
	const size_t p = (size_t)blockIdx.y * blockDim.x + threadIdx.x;
	if (p >= npix4) return;
	float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
	int cx = 0, cy = 0, cz = 0, cw = 0;
	for (int j = k - w; j <= k + w; ++j) {
		const float4* src;
		if (j < 0) { if (n_lo + j < 0) continue; src = halo_lo + (size_t)(n_lo + j) * npix4; }
		else if (j >= n) { if (j - n >= n_hi) continue; src = halo_hi + (size_t)(j - n) * npix4; }
		else src = bkg + (size_t)j * npix4;
		const float4 v = __ldg(src + p);
		nanacc(v.x, s.x, cx); nanacc(v.y, s.y, cy); nanacc(v.z, s.z, cz); nanacc(v.w, s.w, cw);
	}
	float4 o;
	o.x = cx ? s.x / (float)cx : nan_f();
	o.y = cy ? s.y / (float)cy : nan_f();
	o.z = cz ? s.z / (float)cz : nan_f();
	o.w = cw ? s.w / (float)cw : nan_f();
	out[(size_t)k * npix4 + p] = o;
}

// The same result with every frame read once: a thread owns a few horizontally adjacent pixels and walks a segment of the
// cadence axis with the 2w+1 frames of the window in registers (a ring buffer with compile-time slots: the walk is unrolled
// by the window length).  Every output is still summed from scratch over its window in index order, in float32, so the
// result is bit-identical to k_time_smooth; what changes is the traffic -- 1 read + 1 write per frame instead of 2w+1 reads
// (the L2 only partly absorbs those: 14 us per FFI at w = 4 and 33 us at w = 13 against 8 us at w = 1).
template <typename V> struct SmoothVec;
template <> struct SmoothVec<float4> {
	static constexpr int N = 4;
	__device__ static __forceinline__ float get(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
	__device__ static __forceinline__ float4 make(const float* a) { return make_float4(a[0], a[1], a[2], a[3]); }
	__device__ static __forceinline__ float4 nanv() { const float q = nan_f(); return make_float4(q, q, q, q); }
};
template <> struct SmoothVec<float2> {
	static constexpr int N = 2;
	__device__ static __forceinline__ float get(const float2& v, int i) { return i == 0 ? v.x : v.y; }
	__device__ static __forceinline__ float2 make(const float* a) { return make_float2(a[0], a[1]); }
	__device__ static __forceinline__ float2 nanv() { const float q = nan_f(); return make_float2(q, q); }
};

template <> struct SmoothVec<float> {
	static constexpr int N = 1;
	__device__ static __forceinline__ float get(const float& v, int) { return v; }
	__device__ static __forceinline__ float make(const float* a) { return a[0]; }
	__device__ static __forceinline__ float nanv() { return nan_f(); }
};

template <int WH, typename V, int SEG>
__global__ void __launch_bounds__(256) k_time_smooth_slide(size_t npixv, const V* __restrict__ bkg, int n,
	const V* __restrict__ halo_lo, int n_lo, const V* __restrict__ halo_hi, int n_hi, V* __restrict__ out)
{
	constexpr int L = 2 * WH + 1;
	typedef SmoothVec<V> SV;
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npixv) return;
	const int k_lo = blockIdx.y * SEG, k_hi = min(n, k_lo + SEG);
	// frame j of the (halo-extended) stack; frames that do not exist read as NaN, which the nan-mean skips
	auto frame = [&](int j) -> V {
		if (j < 0) return (n_lo + j >= 0) ? __ldg(halo_lo + (size_t)(n_lo + j) * npixv + p) : SV::nanv();
		if (j >= n) return (j - n < n_hi) ? __ldg(halo_hi + (size_t)(j - n) * npixv + p) : SV::nanv();
		return __ldg(bkg + (size_t)j * npixv + p);
	};
	// ring of L + D slots: the frame a step adds to its window was requested D steps earlier
	constexpr int D = 2, R = L + D;
	V buf[R];
#pragma unroll
	for (int u = 0; u < 2 * WH + D; ++u) buf[u] = frame(k_lo - WH + u);
	for (int k = k_lo; k < k_hi; k += R) {
#pragma unroll
		for (int u = 0; u < R; ++u) {
			const int kk = k + u;
			if (kk < k_hi) {
				buf[(2 * WH + D + u) % R] = frame(kk + WH + D);
				float o[SV::N];
#pragma unroll
				for (int c = 0; c < SV::N; ++c) {
					// frames kk - w .. kk + w in index order.  Without a NaN in the window the nan-mean is the plain sequential
					// sum (the same additions in the same order) over L; a NaN result sends the pixel through the NaN-aware sum
					float sum = 0.f; int cnt = L;
#pragma unroll
					for (int t = 0; t < L; ++t) sum += SV::get(buf[(u + t) % R], c);
					if (!(sum == sum)) {
						sum = 0.f; cnt = 0;
#pragma unroll
						for (int t = 0; t < L; ++t) nanacc(SV::get(buf[(u + t) % R], c), sum, cnt);
					}
					o[c] = cnt ? sum / (float)cnt : nan_f();
				}
				out[(size_t)kk * npixv + p] = SV::make(o);
			}
		}
	}
}

template <int WH, typename V, int SEG>
static void launch_smooth_slide(size_t npix, const float* bkg, int n, const float* halo_lo, int n_lo, const float* halo_hi, int n_hi,
	float* out, cudaStream_t st)
{
	const size_t npixv = npix / SmoothVec<V>::N;
	dim3 grid((unsigned)((npixv + 255) / 256), (unsigned)((n + SEG - 1) / SEG));
	k_time_smooth_slide<WH, V, SEG><<<grid, 256, 0, st>>>(npixv, (const V*)bkg, n, (const V*)halo_lo, n_lo, (const V*)halo_hi, n_hi, (V*)out);
}

int tbk_launch_time_smooth(int H, int W, const float* bkg, int n, int w,
	const float* halo_lo, int n_lo, const float* halo_hi, int n_hi, float* out, cudaStream_t st)
{
	// the windows of the TESS cadences (prepare.py:258 and the 200-s extension) walk the cadence axis; any other w, or
	// TBK_SMOOTH_KERNEL=0, takes the frame-parallel kernel
	static const bool slide = !(getenv("TBK_SMOOTH_KERNEL") && atoi(getenv("TBK_SMOOTH_KERNEL")) == 0);
	if (slide && n > 0 && (w == 1 || w == 4 || w == 13)) {
		const size_t npix = (size_t)H * W;
		if (w == 1) launch_smooth_slide<1, float4, 128>(npix, bkg, n, halo_lo, n_lo, halo_hi, n_hi, out, st);
		else if (w == 4) launch_smooth_slide<4, float4, 128>(npix, bkg, n, halo_lo, n_lo, halo_hi, n_hi, out, st);
		else launch_smooth_slide<13, float, 256>(npix, bkg, n, halo_lo, n_lo, halo_hi, n_hi, out, st);   // one pixel per thread: the unrolled walk of 29 steps stays within the instruction cache
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) { tbk_set_error("k_time_smooth_slide: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
		return TBK_OK;
	}
	const size_t npix4 = (size_t)H * W / 4;
	dim3 grid(n, (unsigned)((npix4 + 255) / 256));
	k_time_smooth<<<grid, 256, 0, st>>>(npix4, (const float4*)bkg, n, w, (const float4*)halo_lo, n_lo,
		(const float4*)halo_hi, n_hi, (float4*)out);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_time_smooth: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
// "Whole image is zero" manual exclude (pixel_flags.py:54-56): zero_flags[k] = 1 unless some pixel != 0.
__global__ void k_zero_init(int* zero_flags, int n)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n) zero_flags[k] = 1;
}
__global__ void __launch_bounds__(256) k_zero_detect(size_t npix4, const float4* __restrict__ cube, int* zero_flags)
{
	const int k = blockIdx.y;
	bool nz = false;
	// a frame with data answers on the first load of any warp; only an all-zero frame is read to the end
	for (size_t p0 = (size_t)blockIdx.x * blockDim.x; p0 < npix4; p0 += (size_t)gridDim.x * blockDim.x) {
		const size_t p = p0 + threadIdx.x;
		if (p < npix4) {
			const float4 v = __ldg(cube + (size_t)k * npix4 + p);
			nz |= !(v.x == 0.f) || !(v.y == 0.f) || !(v.z == 0.f) || !(v.w == 0.f);
		}
		if (__any_sync(0xffffffffu, nz) || *((volatile int*)&zero_flags[k]) == 0) break;
	}
	if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) zero_flags[k] = 0;
}

// One thread owns four horizontally adjacent pixels and walks the cadence axis; the accumulators
// stay in registers and are folded into the caller's running sums at the end.
__global__ void __launch_bounds__(256) k_sum_accumulate(PlanDev P, const float4* __restrict__ cube,
	const float4* __restrict__ bkg, uchar4* flags, const tbk_ffi_meta* __restrict__ meta,
	const int* __restrict__ zero_flags, int n, float4* flux_out,
	double* sum, int32_t* nimg, int32_t* used)
{
	const size_t npix4 = (size_t)P.H * P.W / 4;
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npix4) return;
	const int gx = (int)((p * 4) % P.W);
	double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
	int n0 = 0, n1 = 0, n2 = 0, n3 = 0, u0 = 0, u1 = 0, u2 = 0, u3 = 0;
	auto step = [&](int k, uchar4 f, float4 x, const float4 bk) {
		const tbk_ffi_meta m = meta[k];
		bool excl_all = false, excl_cols = false;
		if (P.is_tess) {
			const double time = 0.5 * (m.tstart + m.tstop);
			const int cad = m.cadenceno;
			if (P.camera == 1 && P.ccd == 4 && (cad <= 4724 || m.tstart <= 1325.881282301840)) excl_cols = gx >= 1536;
			else if (P.camera == 1 && ((cad >= 11354 && cad <= 11366) || (time >= 1464.0158778 && time <= 1464.265871))) excl_all = true;
			if (zero_flags[k]) excl_all = true;
		}
		const size_t o = (size_t)k * npix4 + p;
		if (excl_all || excl_cols) {
			f.x |= 2; f.y |= 2; f.z |= 2; f.w |= 2;   // PixelQualityFlags.ManualExclude
			flags[o] = f;
		}
		if (!m.backapp) { x.x -= bk.x; x.y -= bk.y; x.z -= bk.z; x.w -= bk.w; }
		if (f.x & 2) x.x = nan_f();
		if (f.y & 2) x.y = nan_f();
		if (f.z & 2) x.z = nan_f();
		if (f.w & 2) x.w = nan_f();
		if (flux_out) flux_out[o] = x;
		if ((m.dquality & 4335) == 0) {   // TESSQualityFlags.DEFAULT_BITMASK (quality.py:123-124)
			n0 += isfinite(x.x) ? 1 : 0; n1 += isfinite(x.y) ? 1 : 0; n2 += isfinite(x.z) ? 1 : 0; n3 += isfinite(x.w) ? 1 : 0;
			s0 += (x.x == x.x) ? (double)x.x : 0.0;
			s1 += (x.y == x.y) ? (double)x.y : 0.0;
			s2 += (x.z == x.z) ? (double)x.z : 0.0;
			s3 += (x.w == x.w) ? (double)x.w : 0.0;
		}
		u0 += (f.x & 1) == 0; u1 += (f.y & 1) == 0; u2 += (f.z & 1) == 0; u3 += (f.w & 1) == 0;
	};
	// four cadences per trip: their twelve loads are in flight together (the walk is bound by memory latency otherwise); the
	// cadences are still folded into the float64 sums one after the other, in index order
	int k = 0;
	for (; k + 4 <= n; k += 4) {
		uchar4 f[4]; float4 x[4], bk[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const size_t o = (size_t)(k + u) * npix4 + p;
			f[u] = flags[o]; x[u] = __ldg(cube + o); bk[u] = __ldg(bkg + o);
		}
#pragma unroll
		for (int u = 0; u < 4; ++u) step(k + u, f[u], x[u], bk[u]);
	}
	for (; k < n; ++k) {
		const size_t o = (size_t)k * npix4 + p;
		step(k, flags[o], __ldg(cube + o), __ldg(bkg + o));
	}
	double* sp = sum + p * 4; int32_t* np_ = nimg + p * 4; int32_t* up = used + p * 4;
	sp[0] += s0; sp[1] += s1; sp[2] += s2; sp[3] += s3;
	np_[0] += n0; np_[1] += n1; np_[2] += n2; np_[3] += n3;
	up[0] += u0; up[1] += u1; up[2] += u2; up[3] += u3;
}

int tbk_launch_sum_accumulate(const PlanDev& P, const float* cube, const float* bkg_smooth,
	uint8_t* flags, const tbk_ffi_meta* meta, int n, float* flux_out,
	double* sum, int32_t* nimg, int32_t* used, int* zero_flags, cudaStream_t st)
{
	const size_t npix4 = (size_t)P.H * P.W / 4;
	k_zero_init<<<(n + 255) / 256, 256, 0, st>>>(zero_flags, n);
	if (P.is_tess) {
		dim3 g(148 * 2, n);
		k_zero_detect<<<g, 256, 0, st>>>(npix4, (const float4*)cube, zero_flags);
	} else {
		cudaMemsetAsync(zero_flags, 0, sizeof(int) * n, st);
	}
	k_sum_accumulate<<<(unsigned)((npix4 + 255) / 256), 256, 0, st>>>(P, (const float4*)cube,
		(const float4*)bkg_smooth, (uchar4*)flags, meta, zero_flags, n, (float4*)flux_out, sum, nimg, used);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_sum_accumulate: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sum_finalize(size_t npix, const double* __restrict__ sum,
	const int32_t* __restrict__ nimg, const int32_t* __restrict__ used, int numfiles, double threshold,
	double* __restrict__ sumimage, uint8_t* __restrict__ pixels_used)
{
	const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npix) return;
	sumimage[p] = sum[p] / (double)nimg[p];                       // SumImage /= Nimg (0/0 -> NaN)
	pixels_used[p] = ((double)used[p] / (double)numfiles > threshold) ? 1 : 0;
}

int tbk_launch_sum_finalize(int H, int W, const double* sum, const int32_t* nimg, const int32_t* used,
	int numfiles, double threshold, double* sumimage, uint8_t* pixels_used, cudaStream_t st)
{
	const size_t npix = (size_t)H * W;
	k_sum_finalize<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(npix, sum, nimg, used, numfiles, threshold, sumimage, pixels_used);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_sum_finalize: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
// FITS image HDU (big-endian float32) -> native float32 science crop (io.py:46-48).  One thread per output
// float4; the source row is only 4-byte aligned (col0 = 44 pixels = 176 bytes is, the row pitch 8544 bytes is
// 16-byte aligned, so 128-bit loads are fine whenever col0 % 4 == 0; otherwise scalar loads).
__global__ void __launch_bounds__(256) k_decode_ffi_be(const uint32_t* __restrict__ raw, int naxis1, size_t hdu_words,
	int row0, int col0, int H, int W, float* __restrict__ out)
{
	const int b = blockIdx.z, y = blockIdx.y;
	const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if (x >= W) return;
	const uint32_t* src = raw + (size_t)b * hdu_words + (size_t)(row0 + y) * naxis1 + col0 + x;
	uint32_t w[4];
	if ((((uintptr_t)src) & 15) == 0) {
		const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
		w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
	} else {
		for (int q = 0; q < 4; ++q) w[q] = __ldg(src + q);
	}
	float4 o;
	o.x = __uint_as_float(__byte_perm(w[0], 0, 0x0123)); o.y = __uint_as_float(__byte_perm(w[1], 0, 0x0123));
	o.z = __uint_as_float(__byte_perm(w[2], 0, 0x0123)); o.w = __uint_as_float(__byte_perm(w[3], 0, 0x0123));
	*reinterpret_cast<float4*>(out + ((size_t)b * H + y) * W + x) = o;
}

int tbk_launch_decode(const uint8_t* raw, int B, int naxis1, int naxis2, int row0, int col0, int H, int W, float* out, cudaStream_t st)
{
	dim3 grid((W / 4 + 255) / 256, H, B);
	k_decode_ffi_be<<<grid, 256, 0, st>>>((const uint32_t*)raw, naxis1, (size_t)naxis1 * naxis2, row0, col0, H, W, out);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_decode_ffi_be: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
// Mask bytes -> bits for the device-to-host copy of the end-to-end path (the link is the bottleneck there and the mask is
// a fifth of the result bytes): bit (7 - j) of byte i = mask[8 i + j] != 0, the order of numpy.packbits.  One thread per
// four output bytes (32 mask bytes, two 16-byte loads).
__global__ void __launch_bounds__(256) k_pack_mask(const uint4* __restrict__ mask, size_t nwords, uint32_t* __restrict__ bits)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nwords) return;
	const uint4 a = __ldg(mask + 2 * i), b = __ldg(mask + 2 * i + 1);
	const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
	uint32_t out = 0u;
#pragma unroll
	for (int q = 0; q < 4; ++q) {       // output byte q <- mask bytes 8q .. 8q+7 = words 2q, 2q+1
		uint32_t byte = 0u;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const uint32_t m = (w[2 * q + (j >> 2)] >> (8 * (j & 3))) & 0xFFu;
			byte |= (m ? 1u : 0u) << (7 - j);
		}
		out |= byte << (8 * q);
	}
	bits[i] = out;
}

int tbk_launch_pack_mask(const uint8_t* mask, size_t nbytes, uint8_t* bits, cudaStream_t st)
{
	const size_t nwords = nbytes / 32;
	if (nwords) k_pack_mask<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>((const uint4*)mask, nwords, (uint32_t*)bits);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_pack_mask: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
// Stamp gather (consumer side, photometry/BasePhotometry.py:720-751 _load_cube): for a target's stamp
// (rows r0..r1, columns c0..c1 of the CCD) build cube[r][c][k] = stack[k][r0 + r][c0 + c] -- the (rows, cols, times)
// array every photometry method works on.  Reads run along the columns of a frame, writes along time, so 32 x 32
// (time x column) tiles are transposed through shared memory.  Grid: x = tiles of a stamp (grid-stride), y = stamp.
template <typename T>
__global__ void __launch_bounds__(256) k_gather_stamps(const T* __restrict__ stack, int N, int H, int W,
	const int4* __restrict__ stamps, const long long* __restrict__ offs, T* __restrict__ out)
{
	__shared__ T tile[32][33];
	const int4 st = stamps[blockIdx.y];   // r0, r1, c0, c1
	const int h = st.y - st.x, w = st.w - st.z;
	const int ntk = (N + 31) / 32, ntc = (w + 31) / 32;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	T* dst = out + offs[blockIdx.y];
	for (int t = blockIdx.x; t < h * ntc * ntk; t += gridDim.x) {
		const int r = t / (ntc * ntk), rem = t - r * (ntc * ntk);
		const int cb = (rem / ntk) * 32, kb = (rem % ntk) * 32;
#pragma unroll
		for (int kk = ty; kk < 32; kk += 8) {
			const int k = kb + kk, c = cb + tx;
			if (k < N && c < w) tile[kk][tx] = __ldg(stack + ((size_t)k * H + st.x + r) * W + st.z + c);
		}
		__syncthreads();
#pragma unroll
		for (int cc = ty; cc < 32; cc += 8) {
			const int c = cb + cc, k = kb + tx;
			if (k < N && c < w) dst[((size_t)r * w + c) * N + k] = tile[tx][cc];
		}
		__syncthreads();
	}
}

int tbk_launch_gather_stamps(const void* stack, int elem_bytes, int N, int H, int W, const int* stamps,
	const long long* offs, int S, int tiles_x, void* out, cudaStream_t st)
{
	dim3 grid(tiles_x, S);
	if (elem_bytes == 4)
		k_gather_stamps<float><<<grid, 256, 0, st>>>((const float*)stack, N, H, W, (const int4*)stamps, offs, (float*)out);
	else
		k_gather_stamps<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)stack, N, H, W, (const int4*)stamps, offs, (uint8_t*)out);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_gather_stamps: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
// Catalog-driven star mask (extension; the reference's ``catalog`` argument is a TODO, photometry/backgrounds.py:64-65, 90):
// pixel (row y, column x) is masked when (x - sx)^2 + (y - sy)^2 <= r^2 for some star (sx, sy, r).  One CTA per star walks the
// star's bounding box; the mask is OR-ed into (several stars may cover a pixel: every writer stores the same 1).
__global__ void __launch_bounds__(128) k_star_mask(const double* __restrict__ stars, int S, int H, int W, uint8_t* __restrict__ mask)
{
	const int s = blockIdx.x;
	if (s >= S) return;
	const double sx = stars[3 * s], sy = stars[3 * s + 1], r = stars[3 * s + 2];
	if (!(r >= 0.0) || !(sx == sx) || !(sy == sy)) return;
	const int x0 = max(0, (int)ceil(sx - r)), x1 = min(W - 1, (int)floor(sx + r));
	const int y0 = max(0, (int)ceil(sy - r)), y1 = min(H - 1, (int)floor(sy + r));
	if (x1 < x0 || y1 < y0) return;
	const int bw = x1 - x0 + 1, n = bw * (y1 - y0 + 1);
	const double r2 = r * r;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		const int y = y0 + i / bw, x = x0 + i % bw;
		const double dx = (double)x - sx, dy = (double)y - sy;
		if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= r2) mask[(size_t)y * W + x] = 1;
	}
}

int tbk_launch_star_mask(const double* stars, int S, int H, int W, uint8_t* mask, cudaStream_t st)
{
	k_star_mask<<<S, 128, 0, st>>>(stars, S, H, W, mask);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_star_mask: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}
