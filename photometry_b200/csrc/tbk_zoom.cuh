// tbk_zoom.cuh -- mesh -> pixel interpolation.
//
// Restates photutils 1.3.0 BkgZoomInterpolator(order=3, mode='reflect', grid_mode=True, clip=True):
// scipy.ndimage.zoom of the (already prefiltered) 3x3-median-filtered mesh by the box size, then a
// clip to [mesh.min(), mesh.max()]; ptp(mesh) == 0 short-circuits to the constant.  Output pixel o
// samples input coordinate u = (o + 0.5)/64 - 0.5 with 4 cubic B-spline taps floor(u)-1 .. floor(u)+2,
// out-of-range taps folded by half-sample reflection (verified against scipy).
#pragma once
#include "tbk_common.cuh"

__device__ __forceinline__ int reflect_fold(int i, int n)
{
	while (i < 0 || i >= n) {
		if (i < 0) i = -i - 1;
		if (i >= n) i = 2 * n - 1 - i;
	}
	return i;
}

// 5x5 neighbourhood of prefiltered coefficients around tile (ty, tx), rows/cols ty-2 .. ty+2.
struct ZoomTile {
	double c[5][5];
	double w[64 * 4];          // zoom weights staged from global (one LDS instead of an LDG per tap)
	double mesh_min, mesh_max, c_flat;
	int mesh_const, radial_ok;
};

__device__ __forceinline__ void zoom_tile_load(ZoomTile& z, const double* __restrict__ coef,
	int ty, int tx, int ny, int nx)
{
	// call with all threads of the CTA; followed by __syncthreads() at the caller
	for (int i = threadIdx.x; i < 25; i += blockDim.x) {
		int a = i / 5, b = i % 5;
		z.c[a][b] = coef[reflect_fold(ty - 2 + a, ny) * nx + reflect_fold(tx - 2 + b, nx)];
	}
}

__device__ __forceinline__ void zoom_tile_stage(ZoomTile& z, const FfiCtl& c, const double* __restrict__ zw)
{
	for (int i = threadIdx.x; i < 256; i += blockDim.x) z.w[i] = __ldg(zw + i);
	if (threadIdx.x == 0) {
		z.mesh_min = c.mesh_min; z.mesh_max = c.mesh_max; z.mesh_const = c.mesh_const;
		z.radial_ok = c.radial_ok; z.c_flat = c.c_flat;
	}
}

__device__ __forceinline__ double zoom_clip_s(const ZoomTile& z, double v)
{
	if (z.mesh_const) return z.mesh_min;
	return fmin(fmax(v, z.mesh_min), z.mesh_max);
}

// Interpolated (unclipped) mesh value at tile-local pixel (lrow, lcol).
__device__ __forceinline__ double zoom_eval(const ZoomTile& z, const double* __restrict__ zw,
	int lrow, int lcol)
{
	const int oy = lrow >> 5, ox = lcol >> 5;  // taps start one mesh later in the second half
	const double* wy = zw + 4 * lrow;
	const double* wx = zw + 4 * lcol;
	double acc = 0.0;
#pragma unroll
	for (int a = 0; a < 4; ++a) {
		double ra = 0.0;
#pragma unroll
		for (int b = 0; b < 4; ++b) ra += wx[b] * z.c[oy + a][ox + b];
		acc += wy[a] * ra;
	}
	return acc;
}

// Same, four horizontally adjacent pixels lcol .. lcol+3 (lcol % 4 == 0, so one half of the tile).
__device__ __forceinline__ void zoom_eval4(const ZoomTile& z, const double* __restrict__ zw,
	int lrow, int lcol, double (&out)[4])
{
	const int oy = lrow >> 5, ox = lcol >> 5;
	const double* wy = zw + 4 * lrow;
	double r[4];
#pragma unroll
	for (int b = 0; b < 4; ++b) {
		double t = 0.0;
#pragma unroll
		for (int a = 0; a < 4; ++a) t += wy[a] * z.c[oy + a][ox + b];
		r[b] = t;
	}
#pragma unroll
	for (int q = 0; q < 4; ++q) {
		const double* wx = zw + 4 * (lcol + q);
		out[q] = wx[0] * r[0] + wx[1] * r[1] + wx[2] * r[2] + wx[3] * r[3];
	}
}

// Single pixel straight from global coefficients (ring gather: scattered pixels).
__device__ __forceinline__ double zoom_eval_global(const double* __restrict__ coef,
	const double* __restrict__ zw, int y, int x, int ny, int nx)
{
	const int ty = y >> 6, tx = x >> 6, lrow = y & 63, lcol = x & 63;
	const int sy = ty - 2 + (lrow >> 5), sx = tx - 2 + (lcol >> 5);
	const double* wy = zw + 4 * lrow;
	const double* wx = zw + 4 * lcol;
	const bool interior = sy >= 0 && sy + 3 < ny && sx >= 0 && sx + 3 < nx;
	int cx[4];
#pragma unroll
	for (int b = 0; b < 4; ++b) cx[b] = interior ? sx + b : reflect_fold(sx + b, nx);
	double acc = 0.0;
#pragma unroll
	for (int a = 0; a < 4; ++a) {
		const double* row = coef + (interior ? sy + a : reflect_fold(sy + a, ny)) * nx;
		double ra = 0.0;
#pragma unroll
		for (int b = 0; b < 4; ++b) ra += wx[b] * row[cx[b]];
		acc += wy[a] * ra;
	}
	return acc;
}

__device__ __forceinline__ double zoom_clip(const FfiCtl& c, double v)
{
	if (c.mesh_const) return c.mesh_min;
	return fmin(fmax(v, c.mesh_min), c.mesh_max);
}

// Register-resident evaluator for a thread that owns columns lcol .. lcol+3 in several rows of one mesh:
// the column weights and the 5x4 coefficient block are loaded once; a row then costs 4 weight loads.
struct ZoomCols {
	double c[5][4];   // coefficient rows 0..4 of the neighbourhood, the 4 columns this half needs
	double wx[4][4];  // column weights of the 4 pixels
};

__device__ __forceinline__ void zoom_cols_load(ZoomCols& zc, const ZoomTile& z, int lcol)
{
	const int ox = lcol >> 5;
#pragma unroll
	for (int a = 0; a < 5; ++a)
#pragma unroll
		for (int b = 0; b < 4; ++b) zc.c[a][b] = z.c[a][ox + b];
#pragma unroll
	for (int q = 0; q < 4; ++q)
#pragma unroll
		for (int b = 0; b < 4; ++b) zc.wx[q][b] = z.w[4 * (lcol + q) + b];
}

__device__ __forceinline__ void zoom_cols_eval4(const ZoomCols& zc, const ZoomTile& z, int lrow, double (&out)[4])
{
	const double* wy = z.w + 4 * lrow;
	const double w0 = wy[0], w1 = wy[1], w2 = wy[2], w3 = wy[3];
	double r[4];
	if (lrow < 32) {
#pragma unroll
		for (int b = 0; b < 4; ++b) r[b] = w0 * zc.c[0][b] + w1 * zc.c[1][b] + w2 * zc.c[2][b] + w3 * zc.c[3][b];
	} else {
#pragma unroll
		for (int b = 0; b < 4; ++b) r[b] = w0 * zc.c[1][b] + w1 * zc.c[2][b] + w2 * zc.c[3][b] + w3 * zc.c[4][b];
	}
#pragma unroll
	for (int q = 0; q < 4; ++q) out[q] = zc.wx[q][0] * r[0] + zc.wx[q][1] * r[1] + zc.wx[q][2] * r[2] + zc.wx[q][3] * r[3];
}
