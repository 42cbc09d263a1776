// tbk_common.cuh -- shared device structures and block-level primitives (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/tbk.h"

#define TBK_NT 256            // threads per tile CTA
#define TBK_VPT 16            // pixels per thread in a tile CTA (64*64 / 256)
#define TBK_NBINS 1024        // histogram bins for exact selection
#define TBK_CAND 256          // max candidates resolved by rank counting
#define TBK_KDE_M 2048        // KDE grid (2**ceil(log2(2000)), statsmodels kdensityfft)
#define TBK_NPIX_TILE (TBK_TILE * TBK_TILE)
#define TBK_IDW_CACHE 16      // good-mesh patterns whose IDW neighbour tables a plan remembers

// ---------------------------------------------------------------------------------------------
// Static per-plan description, passed by value to kernels.
struct PlanDev {
	int H, W, ny, nx, ntiles;
	int is_tess, use_radial, camera, ccd;
	int bkgiters, radial_smooth;
	float flux_cutoff;          // compared in float32 like ``img0.data > flux_cutoff``
	double xc, yc;              // camera centre (backgrounds.py:121-138)
	double radial_cutoff, step; // ring edges: cutoff + i*step
	int nrings;                 // len(bins) - 1
	int nringpix;               // pixels with a ring id
	int n_nonflat;              // tiles with max(r) > first ring centre
	// device tables
	const int* ring_ptr;        // [nrings + 1] CSR offsets into ring_pix
	const int* ring_order;      // [nrings] ring ids by decreasing pixel count (launch order of the KDE CTAs)
	const int* ring_pix;        // [nringpix] (y << 16) | x, row-major within a ring
	const int* nonflat_tiles;   // [n_nonflat] tile ids
	const int* tile_slot;       // [ntiles] index into nonflat list or -1 (flat tile)
	const double* nonflat_r;    // [n_nonflat][64*64] pixel radius of the non-flat meshes (static, bit-equal to pixel_radius)
	const double2* nonflat_rr;  // [n_nonflat] (min, max) of that radius over the mesh
	const double* nonflat_uj;   // [n_nonflat][64*64] offset from the Taylor-piece centre | local piece index (see radial_tab_eval_uj)
	const int* nonflat_jlo;     // [n_nonflat] first Taylor piece the mesh can see (-1: more than TBK_RTAB_ROWS pieces)
	int n_ringtiles;            // meshes that contain at least one ring pixel
	const int* ringtile_id;     // [n_ringtiles] mesh id
	const int* ringtile_ptr;    // [n_ringtiles + 1] CSR offsets into ringtile_ent
	const unsigned* ringtile_ent; // [nringpix] (index into the ring-ordered sample array << 12) | (row << 6 | col) within the mesh
	const int* ringtile_idx;    // [n_ringtiles][64*64] the same as a dense map: index into the ring-ordered sample array, -1 = no ring pixel
	uint32_t* idw_cache;        // [1 + TBK_IDW_CACHE entries]: entries used, then {state, good-mesh bitmap, neighbour table} each
	const double* zoom_w;       // [64][4] cubic B-spline weights per sub-tile phase
	const double2* twiddle;     // [TBK_KDE_M/2] exp(-2 pi i k / M)
};

// Sigma-clipped statistics of one mesh (values are about the tile's own data, before any shift).
struct TileStat {
	double mean, med, std;
	int nfin;                   // pixels surviving mask + clip; nbad = 4096 - nfin
	int pad;
};

// What k_tile_round_z leaves for k_tile_round_fin per non-flat mesh: this header, then the tails (4 quarters of ZR_TQ
// float64) and the zone elements (4 quarters of ZR_ZQ float64).
#define ZR_TQ 320
#define ZR_ZQ 256
struct ZoneRec {
	int n, nA, nB, nZL;
	int tq[4], zq[4];
	int state, pad[3];          // 1 = lists complete, to be finished
	double s1, s2, pivot, A, B, ZL, ZH, pad2;
};
#define ZR_REC_BYTES (sizeof(ZoneRec) + sizeof(double) * 4 * (ZR_TQ + ZR_ZQ))

// Per-FFI dynamic state.
struct FfiCtl {
	unsigned int min_bits;      // min over valid pixels of x (float bits, x >= 0)
	int any_nonzero;            // some pixel != 0 (NaN counts), pixel_flags.py:54
	int n_valid;
	int mars, earth;            // manual excludes decided from the header scalars
	int all_masked;
	int no_good_mesh;
	int radial_ok;              // current round: spline available
	int npts;                   // spline knots
	int mesh_const;             // ptp(mesh) == 0
	int kde_fallbacks;          // diagnostics: quartile ranks resolved by the generic iterated selection
	int pad0;
	unsigned long long min_key; // rounds >= 2: ordered key of min(x - sq)
	unsigned long long min_ub;  // rounds >= 2: ordered key of an upper bound of that minimum (pruning)
	double zp;                  // zeropoint of the current round
	double c_flat;              // radial value for r <= x0 (10**y0 - zp), 0 when !radial_ok
	double x0, xlast;           // spline abscissa range (ext=3 clamp)
	double mesh_min, mesh_max;
	double kx[TBK_MAX_RINGS];       // knot abscissae
	double pp[TBK_MAX_RINGS][4];    // piecewise cubic: y = pp0 + s*(pp1 + s*(pp2 + s*pp3)), s = t - kx[i]
	double seg[TBK_MAX_RINGS][6];   // dense table per ring-centre interval: {knot, pp0..pp3, 10**pp0} of the covering piece
	short seg_of_ring[TBK_MAX_RINGS]; // ring-centre interval -> spline piece
};

// Workspace carve-up (all pointers into the caller's scratch buffer).
struct Workspace {
	FfiCtl* ctl;            // [B]
	TileStat* tile_base;    // [B][ntiles]
	TileStat* tile_nf;      // [B][n_nonflat]
	double* coef;           // [B][ntiles] prefiltered cubic-spline coefficients of the current round
	double* mesh_hist;      // [B][rounds][ntiles] filtered mesh per round (diagnostics)
	double* s2_raw;         // [B][nrings] ring modes of the current round
	double* s2_hist;        // [B][rounds][nrings] smoothed ring values per round (diagnostics)
	double* ring_v;         // [B][nringpix] ring samples (NaN = masked)
	float* sbmin;           // [B][ntiles][64] minimum valid pixel of every 8x8 sub-block (+inf = none)
	float* sblow;           // [B][ntiles][64] lower bound of min(x - sq) per sub-block (zeropoint pruning)
	unsigned char* zrec;    // [B][n_nonflat][ZR_REC_BYTES] lists of the residual statistics (ZoneRec + tails + zone)
	double* rtab;           // [B][TBK_RSUB * max(nrings - 1, 1)][8] radial profile as Taylor pieces (see RadialTab)
	int* fb_list2;          // [B * n_nonflat] queue of the residual statistics (entries b * n_nonflat + slot)
	int* rt_list;           // [B * ntiles] retry queue of the raw-pixel statistics (entries b * ntiles + tile, count in fb_count[8])
	uint32_t* idw_bits;     // [B][1 + ceil(ntiles / 32)] valid flag + good-mesh bitmap the neighbour table was built for
	uint16_t* idw_tab;      // [B][ntiles][10] IDW neighbours (mesh ids, 0xFFFF = none) of the excluded meshes
	int* fb_count;          // [1] meshes queued for the full-buffer statistics
	int* fb_list;           // [B * ntiles] queue entries b * ntiles + tile
};

// ---------------------------------------------------------------------------------------------
// numeric helpers

__device__ __forceinline__ double nan_d() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ float nan_f() { return __int_as_float(0x7fc00000); }

// r exactly as numpy computes sqrt((xx - xc)**2 + (yy - yc)**2): no FMA contraction.
__device__ __forceinline__ double pixel_radius(const PlanDev& P, int y, int x)
{
	double dx = __dsub_rn((double)(x + 44), P.xc);
	double dy = __dsub_rn((double)y, P.yc);
	return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// Centre of ring i: (bins[1:] - step/2)[i] with bins = arange(cutoff, ..., step)  (backgrounds.py:152-154).
__device__ __forceinline__ double ring_center(const PlanDev& P, int i)
{
	return __dsub_rn(__dadd_rn(P.radial_cutoff, __dmul_rn((double)(i + 1), P.step)), P.step / 2);
}

// Order-preserving map double -> uint64 (for atomicMin on mixed-sign doubles).
__device__ __forceinline__ unsigned long long dkey(double v)
{
	unsigned long long b = (unsigned long long)__double_as_longlong(v);
	return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k)
{
	unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
	return __longlong_as_double((long long)b);
}

// Validity of a pixel: backgrounds.py:91-97 (finite, <= cutoff, >= 0, manual excludes, extra mask).
__device__ __forceinline__ bool pixel_valid(float x, float cutoff)
{
	// NaN fails both comparisons; +inf fails x <= cutoff.
	return (x >= 0.0f) && (x <= cutoff);
}

// Radial component at radius r for the current round (10**spline(clamp(r)) - zp), backgrounds.py:191.
__device__ __forceinline__ double radial_value(const FfiCtl& c, const PlanDev& P, double r)
{
	if (!c.radial_ok) return 0.0;
	if (r <= c.x0) return c.c_flat;
	double t = fmin(r, c.xlast);
	// ring centres are cutoff + step/2 + i*step; seg_of_ring maps the centre interval to the piece
	int i = (int)floor((t - ring_center(P, 0)) / P.step);
	i = max(0, min(i, P.nrings - 1));
	int s = c.seg_of_ring[i];
	// guard against rounding at interval ends
	while (s + 1 < c.npts - 1 && t >= c.kx[s + 1]) ++s;
	while (s > 0 && t < c.kx[s]) --s;
	double u = t - c.kx[s];
	double y = c.pp[s][0] + u * (c.pp[s][1] + u * (c.pp[s][2] + u * c.pp[s][3]));
	return exp10(y) - c.zp;
}

// Shared-memory copy of the per-FFI radial profile for kernels that evaluate it per pixel: the dense
// ring-centre-interval table makes the piece lookup one multiply + one conversion (no search).
struct RadialSmem2 {
	double seg[TBK_MAX_RINGS][6];
	double x0, xlast, c_flat, zp, center0, inv_step;
	int radial_ok, nseg;
};

__device__ __forceinline__ void radial_stage(RadialSmem2& rs, const FfiCtl& c, const PlanDev& P)
{
	const int nseg = max(P.nrings - 1, 1);
	for (int i = threadIdx.x; i < nseg * 6; i += blockDim.x) rs.seg[i / 6][i % 6] = c.seg[i / 6][i % 6];
	if (threadIdx.x == 0) {
		rs.x0 = c.x0; rs.xlast = c.xlast; rs.c_flat = c.c_flat; rs.zp = c.zp;
		rs.center0 = ring_center(P, 0); rs.inv_step = 1.0 / P.step;
		rs.radial_ok = c.radial_ok; rs.nseg = nseg;
	}
}

// clamp without the NaN handling of fmin/fmax (v is finite)
__device__ __forceinline__ double clamp_d(double v, double lo, double hi)
{
	return v < lo ? lo : (v > hi ? hi : v);
}

// Radial profile table for the per-pixel evaluation in the residual statistics.  Every interval between two adjacent
// ring centres is cut into TBK_RSUB pieces of width h = step / TBK_RSUB; a piece lies inside one spline piece, so
// 10**spline(t) = exp(g(u)) with g a cubic in u = t - u0 (u0 = piece centre, |u| <= h / 2 < 1).  The row holds the Taylor
// coefficients of exp(g) about u0 up to degree 6 (from the power-series recurrence n b_n = sum k a_k b_(n-k)), with the
// zeropoint already subtracted from b_0, and u0:  radial(t) = b0 + u (b1 + u (b2 + ... + u b6)).  The truncated term is
// below 1e-15 of the value for any profile the ring statistic can produce (|g'| ~ 1e-3 / px).
#define TBK_RSUB 8
#define TBK_RTAB_ROWS 56   // pieces one mesh can see (64 px diagonal = 91 px of radius < 56 pieces of step / 8 >= 1.75 px)
struct RadialTab {
	const double* rows;     // [nsub][8]: b0 - zp, b1 .. b6, u0
	double x0, xlast, center0, inv_h;
	int nsub, radial_ok;
};
__device__ __forceinline__ RadialTab radial_tab(const double* rtab_b, const FfiCtl& c, const PlanDev& P)
{
	RadialTab t;
	t.rows = rtab_b; t.x0 = c.x0; t.xlast = c.xlast; t.center0 = ring_center(P, 0); t.inv_h = (double)TBK_RSUB / P.step;
	t.nsub = TBK_RSUB * max(P.nrings - 1, 1); t.radial_ok = c.radial_ok;
	return t;
}
// the same evaluation from rows [jlo, jlo + nrows) staged in shared memory (r must map into that range).  The staged rows
// are TBK_RROW = 10 doubles apart: with the natural stride of 8 (64 B) the lanes of a quarter warp, which read the same
// coefficient of DIFFERENT rows, would fall on two 16-byte bank groups only (4-way conflicts); 80 B spreads them over all eight.
#define TBK_RROW 10
__device__ __forceinline__ double radial_tab_eval_s(const RadialTab& t, const double* srows, int jlo, double r)
{
	const double tc = clamp_d(r, t.x0, t.xlast);
	const int j = max(0, min(t.nsub - 1, (int)((tc - t.center0) * t.inv_h))) - jlo;
	const double2* row = reinterpret_cast<const double2*>(srows + TBK_RROW * j);
	const double2 c01 = row[0], c23 = row[1], c45 = row[2], c6u = row[3];
	const double u = tc - c6u.y;
	return fma(u, fma(u, fma(u, fma(u, fma(u, fma(u, c6u.x, c45.y), c45.x), c23.y), c23.x), c01.y), c01.x);
}
// The evaluation from the static per-pixel word of PlanDev::nonflat_uj (offset from the piece centre with the local piece
// index in the low 6 mantissa bits; 62 / 63 = clamped below / above): no clamp, no index arithmetic, no radius.
//   j0 / j1: pieces [j0, j1) lie inside [x0, xlast] of this FFI's spline; cflat / clast: the profile at x0 / xlast.
__device__ __forceinline__ int radial_tab_j0(const RadialTab& t) { return (int)((t.x0 - t.center0) * t.inv_h + 0.5); }
__device__ __forceinline__ int radial_tab_j1(const RadialTab& t) { return (int)((t.xlast - t.center0) * t.inv_h + 0.5); }
__device__ __forceinline__ double radial_tab_eval_uj(const double* srows, int jlo, int j0, int j1, double cflat, double clast, double uj)
{
	const int jl = (int)((unsigned long long)__double_as_longlong(uj) & 63ull);
	const int ja = jlo + jl;
	const double2* row = reinterpret_cast<const double2*>(srows + TBK_RROW * min(jl, TBK_RTAB_ROWS - 1));
	const double2 c01 = row[0], c23 = row[1], c45 = row[2], c6u = row[3];
	const double v = fma(uj, fma(uj, fma(uj, fma(uj, fma(uj, fma(uj, c6u.x, c45.y), c45.x), c23.y), c23.x), c01.y), c01.x);
	return (jl == 62 || ja < j0) ? cflat : ((jl == 63 || ja >= j1) ? clast : v);
}
__device__ __forceinline__ double radial_tab_eval(const RadialTab& t, double r)
{
	const double tc = clamp_d(r, t.x0, t.xlast);
	const int j = max(0, min(t.nsub - 1, (int)((tc - t.center0) * t.inv_h)));
	const double2* row = reinterpret_cast<const double2*>(t.rows + 8 * (size_t)j);
	const double2 c01 = __ldg(row), c23 = __ldg(row + 1), c45 = __ldg(row + 2), c6u = __ldg(row + 3);
	const double u = tc - c6u.y;
	return fma(u, fma(u, fma(u, fma(u, fma(u, fma(u, c6u.x, c45.y), c45.x), c23.y), c23.x), c01.y), c01.x);
}

// ---- TMA bulk copy (cp.async.bulk, global -> shared, completion on an mbarrier) ------------------------------------
// One thread arms the barrier with the byte count and issues the copy; the copy engine moves the bytes without any
// register staging and flips the barrier phase when they have landed; consumers wait on the phase parity.
__device__ __forceinline__ void tma_bar_init(unsigned long long* bar, int arrivals)
{
	const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(arrivals) : "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the init must be visible to the async proxy
}
__device__ __forceinline__ void tma_bar_expect(unsigned long long* bar, unsigned bytes)
{
	const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
// bytes: multiple of 16; src / dst 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), a = (unsigned)__cvta_generic_to_shared(bar);
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(d), "l"(src_gmem), "r"(bytes), "r"(a) : "memory");
}
__device__ __forceinline__ void tma_bar_wait(unsigned long long* bar, unsigned parity)
{
	const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
	asm volatile("{\n\t.reg .pred p;\n\tTMA_WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@!p bra TMA_WAIT_%=;\n\t}" :: "r"(a), "r"(parity) : "memory");
}

#include "tbk_log10_table.cuh"

// Table-driven float64 log10 for positive normal arguments (anything else takes the library path).
// s = 2^k z with z in [0.6875, 1.375); the 7 leading mantissa bits of z pick 1/c and log10(c);
// r = z/c - 1 (|r| < 2^-7, one fma) and log10(s) = k log10(2) + log10(c) + log1p(r)/ln(10).  The heads of
// log10(2) and log10(c) carry 32 / 40 bits so their combination is exact; the degree-8 polynomial truncates below
// 1e-20.  Error < 1 ulp (tests/test_gpu_parity.py::test_device_log10), at about a fifth of the instructions of log10().
__device__ __forceinline__ double tbk_log10(double s, const double (*tab)[4])
{
	const int hi = __double2hiint(s);
	if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log10(s);
	const int t = hi - 0x3fe60000;
	const int i = (t >> 13) & 127;
	const double kd = (double)(t >> 20);
	const double z = __hiloint2double(hi - (t & 0xfff00000), __double2loint(s));
	const double2 e = *reinterpret_cast<const double2*>(tab[i]);   // 1/c, head
	const double tail = tab[i][2];
	const double r = fma(z, e.x, -1.0);
	const double r2 = r * r;
	const double a = fma(r, c_log10_poly[7], c_log10_poly[6]), b = fma(r, c_log10_poly[5], c_log10_poly[4]);
	const double c = fma(r, c_log10_poly[3], c_log10_poly[2]), d = fma(r, c_log10_poly[1], c_log10_poly[0]);
	const double p = fma(r2 * r2, fma(r2, a, b), fma(r2, c, d));
	const double head = fma(kd, TBK_LOG10_2_HI, e.y);
	return head + fma(r, p, fma(kd, TBK_LOG10_2_LO, tail));
}


static __constant__ double c_exp_taylor[7] = {1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};

// 10**spline(clamp(r)) - zp.  Within a piece y = y_i + dy with |dy| small, so 10**y = 10**y_i * exp(ln10 dy):
// the table holds 10**y_i and exp() is a degree-9 Taylor polynomial (relative error < 1e-16 for |ln10 dy| < 0.1;
// steeper profiles take the exp10 path).
__device__ __forceinline__ double radial_value_s(const RadialSmem2& rs, double r)
{
	if (!rs.radial_ok) return 0.0;
	if (r <= rs.x0) return rs.c_flat;
	const double t = fmin(r, rs.xlast);
	const int i = max(0, min(rs.nseg - 1, (int)((t - rs.center0) * rs.inv_step)));
	const double* e = rs.seg[i];
	const double u = t - e[0];
	const double dy = u * (e[2] + u * (e[3] + u * e[4]));
	const double x = 2.302585092994045684 * dy;
	if (fabs(x) > 0.1) return exp10(e[1] + dy) - rs.zp;
	// coefficients come from the constant bank (a DFMA operand) -- as literals each costs two moves per use
	double p = c_exp_taylor[0];
	p = fma(p, x, c_exp_taylor[1]); p = fma(p, x, c_exp_taylor[2]); p = fma(p, x, c_exp_taylor[3]); p = fma(p, x, c_exp_taylor[4]);
	p = fma(p, x, c_exp_taylor[5]); p = fma(p, x, c_exp_taylor[6]); p = fma(p, x, 0.5); p = fma(p, x, 1.0); p = fma(p, x, 1.0);
	return fma(e[5], p, -rs.zp);
}

// ---------------------------------------------------------------------------------------------
// Block reductions for TBK_NT (or any multiple of 32 up to 1024) threads.  Every thread returns the
// same value (partials are combined in a fixed order, so results are deterministic).
struct RedSmem {
	double d[3][32];
	int i[2][32];
};

__device__ __forceinline__ double warp_sum(double v)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ int warp_sum(int v)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ double warp_min(double v)
{
	for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ double warp_max(double v)
{
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// sum of (int n, double a, double b)
__device__ __forceinline__ void block_sum3(RedSmem& s, int& n, double& a, double& b)
{
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	n = warp_sum(n); a = warp_sum(a); b = warp_sum(b);
	__syncthreads();
	if (lane == 0) { s.i[0][w] = n; s.d[0][w] = a; s.d[1][w] = b; }
	__syncthreads();
	n = 0; a = 0.0; b = 0.0;
	for (int k = 0; k < nw; ++k) { n += s.i[0][k]; a += s.d[0][k]; b += s.d[1][k]; }
}
// (int sum, double min, double max)
__device__ __forceinline__ void block_sum_min_max(RedSmem& s, int& n, double& mn, double& mx)
{
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	n = warp_sum(n); mn = warp_min(mn); mx = warp_max(mx);
	__syncthreads();
	if (lane == 0) { s.i[0][w] = n; s.d[0][w] = mn; s.d[1][w] = mx; }
	__syncthreads();
	n = 0; mn = INFINITY; mx = -INFINITY;
	for (int k = 0; k < nw; ++k) { n += s.i[0][k]; mn = fmin(mn, s.d[0][k]); mx = fmax(mx, s.d[1][k]); }
}

// ---------------------------------------------------------------------------------------------
// Exact k-th smallest (0-based) of a distributed multiset by iterated histogram refinement.
//
// The caller supplies a functor ``each(f)`` that calls ``f(value_as_double)`` for every element this
// thread owns (register arrays or a strided sweep over global memory).  ``a <= elements <= b`` must
// hold.  Bins are a monotone non-decreasing function of the value, so after locating the bin that
// holds rank k the search recurses into the exact [min, max] of that bin; a bin with at most
// TBK_CAND members is resolved by rank counting.  Works for float32 data too (float -> double is exact).
struct SelectSmem {
	unsigned int hist[TBK_NBINS];
	double cand[TBK_CAND];
	unsigned int wsum[32];
	int ncand;
	int sel_bin, sel_excl, sel_cnt;
	double result;
};

template <typename Each>
__device__ double block_select(SelectSmem& sm, RedSmem& rs, Each each, int k, double a, double b)
{
	const int tid = threadIdx.x, nt = blockDim.x;
	const int lane = tid & 31, w = tid >> 5, nw = (nt + 31) >> 5;
	for (int level = 0; level < 64; ++level) {
		if (!(a < b)) return a;
		const double scale = (double)TBK_NBINS / (b - a);
		for (int i = tid; i < TBK_NBINS; i += nt) sm.hist[i] = 0u;
		if (tid == 0) sm.ncand = 0;
		__syncthreads();
		each([&](double v) {
			if (v >= a && v <= b) {
				int bin = min(TBK_NBINS - 1, (int)((v - a) * scale));
				atomicAdd(&sm.hist[bin], 1u);
			}
		});
		__syncthreads();
		// exclusive scan over bins: thread t owns bins [t*per, (t+1)*per)
		const int per = (TBK_NBINS + nt - 1) / nt;
		unsigned int loc[8];
		unsigned int tsum = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (j < per) {
				int bi = tid * per + j;
				loc[j] = (bi < TBK_NBINS) ? sm.hist[bi] : 0u;
				tsum += loc[j];
			}
		}
		unsigned int inc = tsum;
		for (int o = 1; o < 32; o <<= 1) {
			unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc += t;
		}
		if (lane == 31) sm.wsum[w] = inc;
		__syncthreads();
		unsigned int base = 0;
		for (int q = 0; q < w; ++q) base += sm.wsum[q];
		unsigned int excl = base + inc - tsum;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			if (j < per) {
				if ((unsigned)k >= excl && (unsigned)k < excl + loc[j]) {
					sm.sel_bin = tid * per + j; sm.sel_excl = (int)excl; sm.sel_cnt = (int)loc[j];
				}
				excl += loc[j];
			}
		}
		__syncthreads();
		const int sbin = sm.sel_bin, sexcl = sm.sel_excl, scnt = sm.sel_cnt;
		if (scnt <= TBK_CAND) {
			each([&](double v) {
				if (v >= a && v <= b) {
					int bin = min(TBK_NBINS - 1, (int)((v - a) * scale));
					if (bin == sbin) { int p = atomicAdd(&sm.ncand, 1); sm.cand[p] = v; }
				}
			});
			__syncthreads();
			const int kk = k - sexcl;
			for (int j = tid; j < scnt; j += nt) {
				const double cj = sm.cand[j];
				int r = 0;
				for (int i = 0; i < scnt; ++i) {
					const double ci = sm.cand[i];
					r += (ci < cj) || (ci == cj && i < j);
				}
				if (r == kk) sm.result = cj;
			}
			__syncthreads();
			return sm.result;
		}
		// recurse into the tight range of the selected bin
		int cnt = 0; double mn = INFINITY, mx = -INFINITY;
		each([&](double v) {
			if (v >= a && v <= b) {
				int bin = min(TBK_NBINS - 1, (int)((v - a) * scale));
				if (bin == sbin) { ++cnt; mn = fmin(mn, v); mx = fmax(mx, v); }
			}
		});
		block_sum_min_max(rs, cnt, mn, mx);
		k -= sexcl; a = mn; b = mx;
		__syncthreads();
	}
	return a;
}
