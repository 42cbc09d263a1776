// tbk_motion.cu -- image movement kernels (photometry/image_motion.py:74-111, 182-258; photometry/prepare.py:678-698):
// for every background-subtracted frame, the translation (dx, dy) against a reference frame by ECC maximisation.
//
//   prepare (``_prepare_flux``):  log10(flux - nanmin + 1)  ->  scaled to [-1, 1]  ->  Scharr gradient magnitude
//                                 (skimage 0.19: sqrt((h^2 + v^2) / 2), 3 x 3, mode='reflect')  ->  NaN -> 0, float32
//   ECC (``cv2.findTransformECC(ref, img, eye(2, 3), MOTION_TRANSLATION, (10000, 1e-6), mask, 5)``):
//       Gaussian 5-tap pre-filter of both images (fixed table 1/16 [1 4 6 4 1], BORDER_REFLECT_101), central differences of the
//       image, then Gauss-Newton iterations: warp image + gradients by the current translation (cv::warpAffine: source
//       coordinates in fixed point, rounded to 1/32 px, float32 bilinear table, zero outside), zero-mean correlation over the
//       pixels the warp keeps inside the frame, 2 x 2 normal equations, update; stop when the correlation coefficient moves
//       by less than eps.
// The CPU restatement oracle/image_motion_oracle.py reproduces the real cv2 to its printed digits; these kernels follow it,
// with the per-pixel float32 roundings of OpenCV's zero-mean images replaced by float64 sums (one pass per iteration).
// One iteration = one launch of k_ecc_sums (grid: ECC_CTAS x B, 17 float64 sums per CTA, written -- not atomically added --
// to a partials array) + one launch of k_ecc_update (one warp per frame adds the partials in a fixed order and updates the
// translation): deterministic, and every frame of the batch iterates in lockstep until its own criterion is met.
#include "tbk_common.cuh"
#include "tbk_internal.h"

#define ECC_CTAS 148
#define ECC_NS 17

__device__ __forceinline__ int reflect101(int i, int n) { if (n == 1) return 0; i = i < 0 ? -i : i; return i >= n ? 2 * n - 2 - i : i; }
// scipy.ndimage mode='reflect' (half-sample symmetric): (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect_hs(int i, int n) { i = i < 0 ? -i - 1 : i; return i >= n ? 2 * n - 1 - i : i; }

// ---- prepare ---------------------------------------------------------------------------------------------------
// minmax[b] = {ordered key of nanmin(flux), ...} via atomics on ordered uint keys (float32)
__device__ __forceinline__ unsigned fkey(float v) { const unsigned b = __float_as_uint(v); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float fkey_inv(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void k_pf_init(unsigned* mm, int B)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < B) { mm[4 * i] = 0xFFFFFFFFu; mm[4 * i + 1] = 0u; mm[4 * i + 2] = 0xFFFFFFFFu; mm[4 * i + 3] = 0u; }
}

// pass 1: nanmin of the flux; pass 2 (LOG): nanmin / nanmax of log10(flux - min + 1)
template <bool LOG>
__global__ void __launch_bounds__(256) k_pf_minmax(const float* __restrict__ flux, size_t npix, unsigned* mm)
{
	const int b = blockIdx.y;
	const float* f = flux + (size_t)b * npix;
	const float fmin0 = LOG ? fkey_inv(mm[4 * b]) : 0.f;
	unsigned lo = 0xFFFFFFFFu, hi = 0u;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
		float v = f[i];
		if (LOG) v = log10f(v - fmin0 + 1.0f);
		if (v == v) { const unsigned k = fkey(v); lo = min(lo, k); hi = max(hi, k); }
	}
	lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
	if ((threadIdx.x & 31) == 0) {
		if (lo != 0xFFFFFFFFu) { atomicMin(&mm[4 * b + (LOG ? 2 : 0)], lo); atomicMax(&mm[4 * b + (LOG ? 3 : 1)], hi); }
	}
}

// scaled log image at (y, x) with reflect indexing
__device__ __forceinline__ float pf_scaled(const float* f, int H, int W, int y, int x, float fmin0, float lmin, float ran)
{
	const float v = f[(size_t)reflect_hs(y, H) * W + reflect_hs(x, W)];
	const float l = log10f(v - fmin0 + 1.0f);
	return -1.0f + 2.0f * ((l - lmin) / ran);
}

__global__ void __launch_bounds__(256) k_pf_scharr(const float* __restrict__ flux, int H, int W, const unsigned* mm, float* __restrict__ out)
{
	const int b = blockIdx.z;
	const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
	if (x >= W || y >= H) return;
	const float* f = flux + (size_t)b * H * W;
	const float fmin0 = fkey_inv(mm[4 * b]), lmin = fkey_inv(mm[4 * b + 2]), lmax = fkey_inv(mm[4 * b + 3]);
	const float ran = fabsf(lmax - lmin);
	float s[3][3];
#pragma unroll
	for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
		for (int dx = -1; dx <= 1; ++dx) s[dy + 1][dx + 1] = pf_scaled(f, H, W, y + dy, x + dx, fmin0, lmin, ran);
	// ndimage.convolve flips the kernel: edge weights [1, 0, -1] along an axis -> value(-1) - value(+1)
	const double sm0 = 3.0 / 16, sm1 = 10.0 / 16;
	const float v = (float)(sm0 * ((double)s[0][0] - (double)s[2][0]) + sm1 * ((double)s[0][1] - (double)s[2][1]) + sm0 * ((double)s[0][2] - (double)s[2][2]));
	const float h = (float)(sm0 * ((double)s[0][0] - (double)s[0][2]) + sm1 * ((double)s[1][0] - (double)s[1][2]) + sm0 * ((double)s[2][0] - (double)s[2][2]));
	float r = sqrtf(v * v + h * h) / sqrtf(2.0f);
	if (!(r == r)) r = 0.f;
	out[(size_t)b * H * W + (size_t)y * W + x] = r;
}

// ---- ECC: pre-filter + gradients -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ecc_gauss_rows(const float* __restrict__ in, int H, int W, float* __restrict__ out)
{
	const int b = blockIdx.z;
	const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
	if (x >= W || y >= H) return;
	const float* p = in + (size_t)b * H * W + (size_t)y * W;
	const float g[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
	float a = 0.f;
#pragma unroll
	for (int k = 0; k < 5; ++k) a += g[k] * p[reflect101(x + k - 2, W)];
	out[(size_t)b * H * W + (size_t)y * W + x] = a;
}
__global__ void __launch_bounds__(256) k_ecc_gauss_cols(const float* __restrict__ in, int H, int W, float* __restrict__ out)
{
	const int b = blockIdx.z;
	const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
	if (x >= W || y >= H) return;
	const float* p = in + (size_t)b * H * W;
	const float g[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
	float a = 0.f;
#pragma unroll
	for (int k = 0; k < 5; ++k) a += g[k] * p[(size_t)reflect101(y + k - 2, H) * W + x];
	out[(size_t)b * H * W + (size_t)y * W + x] = a;
}
__global__ void __launch_bounds__(256) k_ecc_grad(const float* __restrict__ img, int H, int W, float* __restrict__ gx, float* __restrict__ gy)
{
	const int b = blockIdx.z;
	const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
	if (x >= W || y >= H) return;
	const float* p = img + (size_t)b * H * W;
	const size_t o = (size_t)b * H * W + (size_t)y * W + x;
	gx[o] = 0.5f * p[(size_t)y * W + reflect101(x + 1, W)] - 0.5f * p[(size_t)y * W + reflect101(x - 1, W)];
	gy[o] = 0.5f * p[(size_t)reflect101(y + 1, H) * W + x] - 0.5f * p[(size_t)reflect101(y - 1, H) * W + x];
}

// ---- ECC: iteration ------------------------------------------------------------------------------------------------
struct EccState {      // per frame
	float tx, ty;      // current translation (the warp matrix of OpenCV is float32)
	int done, iters;   // done: 1 converged / iteration limit, 2 failed (OpenCV raises: correlation not increasing, NaN)
	double rho, last_rho;
};

__device__ __forceinline__ int cv_round(double v) { return __double2int_rn(v); }

// bilinear sample of cv::warpAffine (float32 table weights, zero outside)
__device__ __forceinline__ float ecc_bilin(const float* __restrict__ p, int H, int W, int sx, int sy, float w0, float w1, float w2, float w3)
{
	const bool x0 = sx >= 0 && sx < W, x1 = sx + 1 >= 0 && sx + 1 < W, y0 = sy >= 0 && sy < H, y1 = sy + 1 >= 0 && sy + 1 < H;
	const float a = (x0 && y0) ? p[(size_t)sy * W + sx] : 0.f, bq = (x1 && y0) ? p[(size_t)sy * W + sx + 1] : 0.f;
	const float c = (x0 && y1) ? p[(size_t)(sy + 1) * W + sx] : 0.f, d = (x1 && y1) ? p[(size_t)(sy + 1) * W + sx + 1] : 0.f;
	return a * w0 + bq * w1 + c * w2 + d * w3;
}

__global__ void __launch_bounds__(256) k_ecc_sums(const float* __restrict__ tmpl, const float* __restrict__ img, const float* __restrict__ gx,
	const float* __restrict__ gy, int H, int W, const EccState* __restrict__ st, double* __restrict__ partial)
{
	__shared__ double red[8][ECC_NS];
	const int b = blockIdx.y;
	const EccState s = st[b];
	if (s.done) return;
	const size_t npix = (size_t)H * W;
	const float* I = img + (size_t)b * npix; const float* GX = gx + (size_t)b * npix; const float* GY = gy + (size_t)b * npix;
	// cv::warpAffine, M = [[1, 0, tx], [0, 1, ty]] in double, AB_BITS = 10
	const int X0l = cv_round((double)s.tx * 1024.0) + 16, Y0l_ = 16, X0n = cv_round((double)s.tx * 1024.0) + 512, Y0n_ = 512;
	double a[ECC_NS];
#pragma unroll
	for (int k = 0; k < ECC_NS; ++k) a[k] = 0.0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
		const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
		const int Yb = cv_round(((double)y + (double)s.ty) * 1024.0);
		const int X = (X0l + x * 1024) >> 5, Y = (Yb + Y0l_) >> 5;
		const int sx = X >> 5, sy = Y >> 5;
		const float fx = (float)(X & 31) / 32.0f, fy = (float)(Y & 31) / 32.0f;
		const float w0 = (1.f - fy) * (1.f - fx), w1 = (1.f - fy) * fx, w2 = fy * (1.f - fx), w3 = fy * fx;
		const double Iw = (double)ecc_bilin(I, H, W, sx, sy, w0, w1, w2, w3);
		const double gxw = (double)ecc_bilin(GX, H, W, sx, sy, w0, w1, w2, w3), gyw = (double)ecc_bilin(GY, H, W, sx, sy, w0, w1, w2, w3);
		const int xn = (X0n + x * 1024) >> 10, yn = (Yb + Y0n_) >> 10;
		const bool m = xn >= 0 && xn < W && yn >= 0 && yn < H;      // nearest-neighbour warp of the (all ones) mask
		const double T = (double)tmpl[i];
		a[6] += gxw * gxw; a[7] += gxw * gyw; a[8] += gyw * gyw;
		if (m) {
			a[0] += 1.0; a[1] += Iw; a[2] += Iw * Iw; a[3] += T; a[4] += T * T; a[5] += T * Iw;
			a[9] += gxw * Iw; a[10] += gxw; a[12] += gyw * Iw; a[13] += gyw; a[15] += gxw * T; a[16] += gyw * T;
		} else { a[11] += gxw * Iw; a[14] += gyw * Iw; }
	}
#pragma unroll
	for (int k = 0; k < ECC_NS; ++k) a[k] = warp_sum(a[k]);
	if ((threadIdx.x & 31) == 0) for (int k = 0; k < ECC_NS; ++k) red[threadIdx.x >> 5][k] = a[k];
	__syncthreads();
	if (threadIdx.x < ECC_NS) {
		double t = 0.0;
		for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
		partial[((size_t)b * gridDim.x + blockIdx.x) * ECC_NS + threadIdx.x] = t;
	}
}

__global__ void __launch_bounds__(32) k_ecc_update(EccState* st, const double* __restrict__ partial, int nctas, int max_iter, double eps, int* all_done)
{
	const int b = blockIdx.x, lane = threadIdx.x;
	EccState& s = st[b];
	if (s.done) return;
	double v = 0.0;
	if (lane < ECC_NS) for (int c = 0; c < nctas; ++c) v += partial[((size_t)b * nctas + c) * ECC_NS + lane];
	double a[ECC_NS];
#pragma unroll
	for (int k = 0; k < ECC_NS; ++k) a[k] = __shfl_sync(0xffffffffu, v, k);
	if (lane != 0) return;
	const double n = a[0], mI = a[1] / n, mT = a[3] / n;
	const double img2 = a[2] - n * mI * mI, tmp2 = a[4] - n * mT * mT, corr = a[5] - n * mI * mT;
	const double hxx = (double)(float)a[6], hxy = (double)(float)a[7], hyy = (double)(float)a[8];     // the Hessian is float32 in OpenCV
	const double det = hxx * hyy - hxy * hxy;
	const double i00 = hyy / det, i01 = -hxy / det, i11 = hxx / det;
	const double ipx = (a[9] - mI * a[10]) + a[11], ipy = (a[12] - mI * a[13]) + a[14];
	const double tpx = a[15] - mT * a[10], tpy = a[16] - mT * a[13];
	s.last_rho = s.rho;
	s.rho = corr / (sqrt(fmax(img2, 0.0)) * sqrt(fmax(tmp2, 0.0)));
	s.iters += 1;
	bool fail = !(s.rho == s.rho) || !(n > 0.0);
	const double iphx = i00 * ipx + i01 * ipy, iphy = i01 * ipx + i11 * ipy;
	const double lam_n = img2 - (ipx * iphx + ipy * iphy), lam_d = corr - (tpx * iphx + tpy * iphy);
	if (!(lam_d > 0.0)) fail = true;
	if (fail) { s.done = 2; atomicAdd(all_done, 1); return; }
	const double lam = lam_n / lam_d;
	const double epx = lam * tpx - ipx, epy = lam * tpy - ipy;
	s.tx = (float)((double)s.tx + (float)(i00 * epx + i01 * epy));
	s.ty = (float)((double)s.ty + (float)(i01 * epx + i11 * epy));
	// OpenCV tests the criterion at the top of the next iteration: stop when |rho - last_rho| < eps or the count is reached
	if (fabs(s.rho - s.last_rho) < eps || s.iters >= max_iter) { s.done = 1; atomicAdd(all_done, 1); }
}

__global__ void k_ecc_init(EccState* st, int B, double eps, int* all_done)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b == 0) *all_done = 0;
	if (b < B) { st[b].tx = 0.f; st[b].ty = 0.f; st[b].done = 0; st[b].iters = 0; st[b].rho = -1.0; st[b].last_rho = -eps; }
}

__global__ void k_ecc_result(const EccState* st, int B, double* out)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	const bool ok = st[b].done == 1;
	out[4 * b] = ok ? (double)st[b].tx : nan_d();
	out[4 * b + 1] = ok ? (double)st[b].ty : nan_d();
	out[4 * b + 2] = st[b].rho;
	out[4 * b + 3] = (double)st[b].iters;
}

// ---- host side -----------------------------------------------------------------------------------------------------
static bool ok_launch(const char* what)
{
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("%s: %s", what, cudaGetErrorString(e)); return false; }
	return true;
}

int tbk_launch_motion_prepare(const float* flux, int B, int H, int W, float* out, unsigned* scratch, cudaStream_t st)
{
	const size_t npix = (size_t)H * W;
	k_pf_init<<<(B + 127) / 128, 128, 0, st>>>(scratch, B);
	const dim3 gr(128, B);
	k_pf_minmax<false><<<gr, 256, 0, st>>>(flux, npix, scratch);
	k_pf_minmax<true><<<gr, 256, 0, st>>>(flux, npix, scratch);
	const dim3 g2((W + 31) / 32, (H + 7) / 8, B);
	k_pf_scharr<<<g2, 256, 0, st>>>(flux, H, W, scratch, out);
	return ok_launch("motion_prepare") ? TBK_OK : TBK_ERR_CUDA;
}

size_t tbk_motion_workspace(int B, int H, int W)
{
	const size_t npix = (size_t)H * W;
	return 256 + sizeof(float) * npix * (2 + 4 * (size_t)B) + sizeof(EccState) * (size_t)B + 256 + sizeof(double) * ECC_NS * ECC_CTAS * (size_t)B + 256;
}

int tbk_launch_motion_ecc(const float* ref_prepared, const float* prepared, int B, int H, int W, int max_iter, double eps,
	void* workspace, double* out, cudaStream_t st)
{
	const size_t npix = (size_t)H * W;
	char* w = (char*)workspace;
	int* all_done = (int*)w; w += 256;
	float* tmpl = (float*)w; w += sizeof(float) * npix;
	float* tmp1 = (float*)w; w += sizeof(float) * npix * (size_t)(B > 1 ? B : 1) ;
	float* img = (float*)w; w += sizeof(float) * npix * B;
	float* gx = (float*)w; w += sizeof(float) * npix * B;
	float* gy = (float*)w; w += sizeof(float) * npix * B;
	w += sizeof(float) * npix;   // slack of the layout above (tmp1 holds B frames)
	EccState* state = (EccState*)(((uintptr_t)w + 255) & ~(uintptr_t)255); w = (char*)(state + B);
	double* partial = (double*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
	const dim3 g1((W + 31) / 32, (H + 7) / 8, 1), gB((W + 31) / 32, (H + 7) / 8, B);
	k_ecc_gauss_rows<<<g1, 256, 0, st>>>(ref_prepared, H, W, tmp1);
	k_ecc_gauss_cols<<<g1, 256, 0, st>>>(tmp1, H, W, tmpl);
	k_ecc_gauss_rows<<<gB, 256, 0, st>>>(prepared, H, W, tmp1);
	k_ecc_gauss_cols<<<gB, 256, 0, st>>>(tmp1, H, W, img);
	k_ecc_grad<<<gB, 256, 0, st>>>(img, H, W, gx, gy);
	k_ecc_init<<<(B + 127) / 128, 128, 0, st>>>(state, B, eps, all_done);
	if (!ok_launch("motion_ecc setup")) return TBK_ERR_CUDA;
	int it = 0;
	while (it < max_iter) {
		const int burst = it < 32 ? 8 : 32;
		for (int k = 0; k < burst && it < max_iter; ++k, ++it) {
			k_ecc_sums<<<dim3(ECC_CTAS, B), 256, 0, st>>>(tmpl, img, gx, gy, H, W, state, partial);
			k_ecc_update<<<B, 32, 0, st>>>(state, partial, ECC_CTAS, max_iter, eps, all_done);
		}
		int done = 0;
		if (cudaMemcpyAsync(&done, all_done, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
			tbk_set_error("motion_ecc: %s", cudaGetErrorString(cudaGetLastError())); return TBK_ERR_CUDA;
		}
		if (done >= B) break;
	}
	k_ecc_result<<<(B + 127) / 128, 128, 0, st>>>(state, B, out);
	return ok_launch("motion_ecc") ? TBK_OK : TBK_ERR_CUDA;
}
