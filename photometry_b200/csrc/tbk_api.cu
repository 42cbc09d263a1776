// tbk_api.cu -- C ABI (include/tbk.h): plan construction (static geometry tables) and entry points.
#include <cstdarg>
#include <climits>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <map>
#include <mutex>
#include "tbk_common.cuh"
#include "tbk_internal.h"
#include "tbk_kdtree.cuh"

static thread_local char g_err[512] = "";

void tbk_set_error(const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char* tbk_last_error(void) { return g_err; }
extern "C" int tbk_version(void) { return TBK_VERSION; }

// Makes the plan's device current for the duration of an entry point and restores the caller's afterwards.
struct DeviceGuard {
	int prev;
	bool switched;
	explicit DeviceGuard(int dev) : prev(-1), switched(false)
	{
		if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
	}
	~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
	tbk_set_error("%s: %s", #x, cudaGetErrorString(e_)); return TBK_ERR_CUDA; } } while (0)

struct tbk_plan {
	PlanDev dev;
	int device;
	std::vector<void*> allocs;
	int tile_kernel;   // TBK_TILE_KERNEL (development / cross-check switch): 0 = generic CTA-per-mesh kernels, 3 = bucketed kernels, 6 = zone kernels with the bucketed ones as fallback (default), 7 = 6 with the raw mesh staged by TMA bulk copies (measured alternative)
	std::map<cudaStream_t, TbkSide> sides;   // per caller stream (see TbkSide)
	std::mutex mtx;
};

// Byte offsets of the workspace sections for a batch of B (a value, not plan state: a plan may serve several streams
// and host threads with different batch sizes at once).
struct WsLayout {
	size_t off_ctl, off_base, off_nf, off_coef, off_mesh, off_s2raw, off_s2hist, off_ringv, off_sbmin, off_sblow, off_fb, off_idwbits, off_idwtab, off_rtab, off_fb2, off_zrec, off_rt;
	size_t total;
};

// photometry/backgrounds.py:121-138
static const double XYCEN[4][4][2] = {
	{{2158.222313, 2099.523364}, {-5.653058, 2098.018608}, {2141.511437, 2099.868226}, {-22.406442, 2100.116443}},
	{{2148.588316, 2094.033024}, {-16.806140, 2095.810070}, {2151.351646, 2105.747100}, {-13.118570, 2105.982211}},
	{{2152.175481, 2092.337442}, {-10.494413, 2093.108135}, {2145.029218, 2107.883573}, {-17.374782, 2105.296746}},
	{{2149.259760, 2091.433315}, {-12.906931, 2093.350054}, {2148.906766, 2110.730620}, {-14.629676, 2111.341670}},
};

// volatile stores keep the host compiler from contracting a*b+c into an FMA: the ring membership
// must match numpy's sqrt((xx - xc)**2 + (yy - yc)**2) bit for bit.
static double host_radius(double xc, double yc, int y, int x)
{
	volatile double dx = (double)(x + 44) - xc;
	volatile double dy = (double)y - yc;
	volatile double a = dx * dx;
	volatile double b = dy * dy;
	volatile double s = a + b;
	return std::sqrt(s);
}
static double host_edge(double cutoff, double step, int i)
{
	volatile double t = (double)i * step;
	volatile double e = cutoff + t;
	return e;
}

template <typename T>
static int upload(tbk_plan* p, const std::vector<T>& v, const T** out)
{
	void* d = nullptr;
	size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
	CUDA_TRY(cudaMalloc(&d, bytes));
	p->allocs.push_back(d);
	if (!v.empty()) CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	*out = (const T*)d;
	return TBK_OK;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static WsLayout layout(const tbk_plan* plan, int B)
{
	const PlanDev& P = plan->dev;
	WsLayout L;
	WsLayout* p = &L;
	size_t o = 0;
	p->off_ctl = o;    o = align_up(o + sizeof(FfiCtl) * (size_t)B);
	p->off_base = o;   o = align_up(o + sizeof(TileStat) * (size_t)B * P.ntiles);
	p->off_nf = o;     o = align_up(o + sizeof(TileStat) * (size_t)B * std::max(P.n_nonflat, 1));
	p->off_coef = o;   o = align_up(o + sizeof(double) * (size_t)B * P.ntiles);
	p->off_mesh = o;   o = align_up(o + sizeof(double) * (size_t)B * P.bkgiters * P.ntiles);
	p->off_s2raw = o;  o = align_up(o + sizeof(double) * (size_t)B * std::max(P.nrings, 1));
	p->off_s2hist = o; o = align_up(o + sizeof(double) * (size_t)B * P.bkgiters * std::max(P.nrings, 1));
	p->off_ringv = o;  o = align_up(o + sizeof(double) * (size_t)B * std::max(P.nringpix, 1));
	p->off_sbmin = o;  o = align_up(o + sizeof(float) * (size_t)B * P.ntiles * 64);
	p->off_sblow = o;  o = align_up(o + sizeof(float) * (size_t)B * P.ntiles * 64);
	p->off_fb = o;     o = align_up(o + 256 + sizeof(int) * (size_t)B * P.ntiles);
	p->off_idwbits = o; o = align_up(o + sizeof(uint32_t) * (size_t)B * ((P.ntiles + 31) / 32 + 1));
	p->off_idwtab = o; o = align_up(o + sizeof(uint16_t) * (size_t)B * P.ntiles * 10);
	p->off_rtab = o;   o = align_up(o + sizeof(double) * 8 * (size_t)B * TBK_RSUB * std::max(P.nrings - 1, 1));
	p->off_fb2 = o;    o = align_up(o + sizeof(int) * (size_t)B * std::max(P.n_nonflat, 1));
	p->off_zrec = o;   o = align_up(o + ZR_REC_BYTES * (size_t)B * std::max(P.n_nonflat, 1));
	p->off_rt = o;     o = align_up(o + sizeof(int) * (size_t)B * P.ntiles);
	L.total = o;
	return L;
}

static Workspace carve(const tbk_plan* plan, void* base, int B)
{
	const WsLayout L = layout(plan, B);
	const WsLayout* p = &L;
	char* b = (char*)base;
	Workspace ws;
	ws.ctl = (FfiCtl*)(b + p->off_ctl);
	ws.tile_base = (TileStat*)(b + p->off_base);
	ws.tile_nf = (TileStat*)(b + p->off_nf);
	ws.coef = (double*)(b + p->off_coef);
	ws.mesh_hist = (double*)(b + p->off_mesh);
	ws.s2_raw = (double*)(b + p->off_s2raw);
	ws.s2_hist = (double*)(b + p->off_s2hist);
	ws.ring_v = (double*)(b + p->off_ringv);
	ws.sbmin = (float*)(b + p->off_sbmin);
	ws.sblow = (float*)(b + p->off_sblow);
	ws.fb_count = (int*)(b + p->off_fb);
	ws.fb_list = (int*)(b + p->off_fb + 256);
	ws.idw_bits = (uint32_t*)(b + p->off_idwbits);
	ws.idw_tab = (uint16_t*)(b + p->off_idwtab);
	ws.rtab = (double*)(b + p->off_rtab);
	ws.fb_list2 = (int*)(b + p->off_fb2);
	ws.zrec = (unsigned char*)(b + p->off_zrec);
	ws.rt_list = (int*)(b + p->off_rt);
	return ws;
}

extern "C" int tbk_plan_create(tbk_plan** out, int H, int W, int is_tess, int camera, int ccd,
	double flux_cutoff, int bkgiters, double radial_cutoff, double radial_pixel_step,
	int radial_smooth, const double* xycen_override, int device)
{
	if (!out) { tbk_set_error("plan pointer is NULL"); return TBK_ERR_INVALID; }
	*out = nullptr;
	if (H <= 0 || W <= 0 || H % TBK_TILE || W % TBK_TILE) {
		tbk_set_error("image shape (%d, %d) must be positive multiples of %d", H, W, TBK_TILE);
		return TBK_ERR_INVALID;
	}
	if ((H / TBK_TILE) * (W / TBK_TILE) > 4096) { tbk_set_error("too many meshes (max 4096)"); return TBK_ERR_INVALID; }
	if (bkgiters < 1 || bkgiters > TBK_MAX_ROUNDS) { tbk_set_error("bkgiters must be in 1..%d", TBK_MAX_ROUNDS); return TBK_ERR_INVALID; }
	if (radial_smooth < 0 || radial_smooth > 63) { tbk_set_error("radial_smooth must be in 0..63"); return TBK_ERR_INVALID; }
	double xc = 0, yc = 0;
	if (is_tess) {
		if (xycen_override) { xc = xycen_override[0]; yc = xycen_override[1]; }
		else if (camera >= 1 && camera <= 4 && ccd >= 1 && ccd <= 4) { xc = XYCEN[camera - 1][ccd - 1][0]; yc = XYCEN[camera - 1][ccd - 1][1]; }
		else { tbk_set_error("Invalid CAMERA or CCD in header: CAMERA=%d, CCD=%d", camera, ccd); return TBK_ERR_INVALID; }
		if (!(radial_pixel_step > 0)) { tbk_set_error("radial_pixel_step must be positive"); return TBK_ERR_INVALID; }
	}
	{ int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { tbk_set_error("no CUDA device %d", device); return TBK_ERR_CUDA; } }
	DeviceGuard guard(device);
	if (tbk_fit_configure() != TBK_OK) return TBK_ERR_CUDA;

	tbk_plan* p = new tbk_plan();
	p->device = device;
	{ const char* tk = getenv("TBK_TILE_KERNEL"); p->tile_kernel = tk ? (tk[0] - '0') : 6; if (p->tile_kernel != 0 && p->tile_kernel != 3 && p->tile_kernel != 7) p->tile_kernel = 6; }
	PlanDev& P = p->dev;
	memset(&P, 0, sizeof(P));
	P.H = H; P.W = W; P.ny = H / TBK_TILE; P.nx = W / TBK_TILE; P.ntiles = P.ny * P.nx;
	P.is_tess = is_tess ? 1 : 0; P.use_radial = P.is_tess; P.camera = camera; P.ccd = ccd;
	P.bkgiters = is_tess ? bkgiters : 1;  // backgrounds.py:155-157
	P.radial_smooth = radial_smooth;
	P.flux_cutoff = (float)flux_cutoff;
	P.xc = xc; P.yc = yc; P.radial_cutoff = radial_cutoff; P.step = radial_pixel_step;

	std::vector<int> ring_ptr(1, 0), ring_pix, nonflat, tile_slot(P.ntiles, -1), ringtile_id, ringtile_ptr(1, 0);
	std::vector<unsigned> ringtile_ent;
	std::vector<double> nonflat_r;
	std::vector<double2> nonflat_rr;
	std::vector<double> nonflat_uj;
	std::vector<int> ringtile_idx;
	std::vector<int> nonflat_jlo;
	if (P.use_radial) {
		// backgrounds.py:145-154
		std::vector<double> r((size_t)H * W);
		double rmax = 0;
		for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
			double v = host_radius(xc, yc, y, x);
			r[(size_t)y * W + x] = v;
			rmax = std::max(rmax, v);
		}
		const double radial_max = rmax + radial_pixel_step;
		const long nedges = (long)std::ceil((radial_max - radial_cutoff) / radial_pixel_step);  // len(np.arange)
		const int nrings = (int)nedges - 1;
		if (nrings < 1) {
			delete p;
			tbk_set_error("radial_cutoff=%g leaves no radial bins inside the image (max r = %g)", radial_cutoff, rmax);
			return TBK_ERR_INVALID;
		}
		if (nrings > TBK_MAX_RINGS) { delete p; tbk_set_error("too many radial rings (%d > %d)", nrings, TBK_MAX_RINGS); return TBK_ERR_INVALID; }
		if (radial_smooth / 2 + 1 > nrings) { delete p; tbk_set_error("radial_smooth too wide for %d rings", nrings); return TBK_ERR_INVALID; }
		P.nrings = nrings;
		// ring id = searchsorted(bins, r, 'right') - 1, r == bins[-1] joins the last ring
		std::vector<int> rid((size_t)H * W, -1);
		std::vector<int> count(nrings, 0);
		const double last_edge = host_edge(radial_cutoff, radial_pixel_step, nrings);
		for (size_t i = 0; i < r.size(); ++i) {
			const double v = r[i];
			if (v < radial_cutoff) continue;
			long k = (long)std::floor((v - radial_cutoff) / radial_pixel_step);
			while (k + 1 <= nrings && host_edge(radial_cutoff, radial_pixel_step, (int)k + 1) <= v) ++k;
			while (k > 0 && host_edge(radial_cutoff, radial_pixel_step, (int)k) > v) --k;
			if (k >= nrings) { if (v == last_edge) k = nrings - 1; else continue; }
			rid[i] = (int)k; ++count[k];
		}
		ring_ptr.assign(nrings + 1, 0);
		for (int k = 0; k < nrings; ++k) ring_ptr[k + 1] = ring_ptr[k] + count[k];
		ring_pix.resize(ring_ptr[nrings]);
		std::vector<int> fill(ring_ptr.begin(), ring_ptr.end() - 1);
		// ring pixels are stored as (y << 16) | x, row-major within a ring
		for (size_t i = 0; i < rid.size(); ++i) if (rid[i] >= 0) ring_pix[fill[rid[i]]++] = (int)(((i / W) << 16) | (i % W));
		P.nringpix = (int)ring_pix.size();
		if (ring_pix.size() >= (1u << 20)) { delete p; tbk_set_error("too many ring pixels (%zu)", ring_pix.size()); return TBK_ERR_INVALID; }
		// the same pixels grouped by mesh (for the gather kernel, which stages one mesh's spline coefficients)
		{
			std::vector<int> tcount(P.ntiles, 0);
			for (size_t j = 0; j < ring_pix.size(); ++j) {
				const int y = ring_pix[j] >> 16, x = ring_pix[j] & 0xFFFF;
				++tcount[(y / TBK_TILE) * P.nx + x / TBK_TILE];
			}
			std::vector<int> slot_of(P.ntiles, -1);
			for (int t = 0; t < P.ntiles; ++t) if (tcount[t]) { slot_of[t] = (int)ringtile_id.size(); ringtile_id.push_back(t); }
			ringtile_ptr.assign(ringtile_id.size() + 1, 0);
			for (size_t k = 0; k < ringtile_id.size(); ++k) ringtile_ptr[k + 1] = ringtile_ptr[k] + tcount[ringtile_id[k]];
			ringtile_ent.resize(ring_pix.size());
			std::vector<int> fill2(ringtile_ptr.begin(), ringtile_ptr.end() - 1);
			for (size_t j = 0; j < ring_pix.size(); ++j) {
				const int y = ring_pix[j] >> 16, x = ring_pix[j] & 0xFFFF;
				const int t = (y / TBK_TILE) * P.nx + x / TBK_TILE;
				ringtile_ent[fill2[slot_of[t]]++] = ((unsigned)j << 12) | (unsigned)(((y % TBK_TILE) << 6) | (x % TBK_TILE));
			}
			P.n_ringtiles = (int)ringtile_id.size();
			// dense form of the same map for the gather kernel (87 % of the pixels of these meshes are ring pixels)
			ringtile_idx.assign(ringtile_id.size() * (size_t)TBK_NPIX_TILE, -1);
			for (size_t k = 0; k < ringtile_id.size(); ++k)
				for (int e = ringtile_ptr[k]; e < ringtile_ptr[k + 1]; ++e)
					ringtile_idx[k * TBK_NPIX_TILE + (ringtile_ent[e] & 4095u)] = (int)(ringtile_ent[e] >> 12);
		}
		// meshes that reach beyond the first ring centre see a non-constant radial component
		const double c0 = host_edge(radial_cutoff, radial_pixel_step, 1) - radial_pixel_step / 2;
		for (int t = 0; t < P.ntiles; ++t) {
			const int ty = t / P.nx, tx = t % P.nx;
			double m = 0;
			const int ys[2] = {ty * TBK_TILE, ty * TBK_TILE + TBK_TILE - 1}, xs[2] = {tx * TBK_TILE, tx * TBK_TILE + TBK_TILE - 1};
			for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) m = std::max(m, r[(size_t)ys[a] * W + xs[b]]);
			if (m > c0) { tile_slot[t] = (int)nonflat.size(); nonflat.push_back(t); }
		}
		P.n_nonflat = (int)nonflat.size();
		nonflat_r.resize((size_t)nonflat.size() * TBK_NPIX_TILE);
		nonflat_rr.resize(nonflat.size());
		for (size_t k = 0; k < nonflat.size(); ++k) {
			const int ty = nonflat[k] / P.nx, tx = nonflat[k] % P.nx;
			double lo = 1e300, hi = 0;
			for (int a = 0; a < TBK_TILE; ++a) for (int bb = 0; bb < TBK_TILE; ++bb) {
				const double v = r[(size_t)(ty * TBK_TILE + a) * W + tx * TBK_TILE + bb];
				nonflat_r[k * TBK_NPIX_TILE + a * TBK_TILE + bb] = v;
				lo = std::min(lo, v); hi = std::max(hi, v);
			}
			nonflat_rr[k] = make_double2(lo, hi);
		}
		// Static part of the Taylor-piece evaluation of the radial profile (RadialTab, tbk_common.cuh): the piece a pixel
		// falls into and its offset from the piece centre do not depend on the FFI.  Per pixel one float64 = the offset u
		// with the low 6 mantissa bits replaced by the piece index relative to the first piece of the mesh (|du| < 2^-46 |u|);
		// 62 / 63 mark pixels below the first / at or beyond the last ring centre (the spline is clamped there, ext=3).
		nonflat_uj.assign((size_t)nonflat.size() * TBK_NPIX_TILE, 0.0);
		nonflat_jlo.assign(nonflat.size(), 0);
		{
			const int nsub = TBK_RSUB * std::max(nrings - 1, 1);
			const double hsub = radial_pixel_step / (double)TBK_RSUB, inv_h = (double)TBK_RSUB / radial_pixel_step;
			const double clast = c0 + (double)(nrings - 1) * radial_pixel_step;
			auto piece = [&](double v) { return std::max(0, std::min(nsub - 1, (int)((v - c0) * inv_h))); };
			for (size_t k = 0; k < nonflat.size(); ++k) {
				const double* rk = &nonflat_r[k * TBK_NPIX_TILE];
				int jmin = INT_MAX, jmax = -1;
				for (int e = 0; e < TBK_NPIX_TILE; ++e)
					if (rk[e] >= c0 && rk[e] < clast) { const int j = piece(rk[e]); jmin = std::min(jmin, j); jmax = std::max(jmax, j); }
				if (jmax < 0) jmin = jmax = 0;
				nonflat_jlo[k] = (jmax - jmin < TBK_RTAB_ROWS) ? jmin : -1;   // -1: the mesh sees too many pieces, bucketed path
				for (int e = 0; e < TBK_NPIX_TILE; ++e) {
					unsigned long long bits;
					if (rk[e] < c0) bits = 62ull;
					else if (!(rk[e] < clast)) bits = 63ull;
					else {
						const int j = piece(rk[e]);
						const double u = rk[e] - (c0 + ((double)j + 0.5) * hsub);
						std::memcpy(&bits, &u, 8);
						bits = (bits & ~63ull) | (unsigned long long)std::min(j - jmin, TBK_RTAB_ROWS - 1);
					}
					std::memcpy(&nonflat_uj[k * TBK_NPIX_TILE + e], &bits, 8);
				}
			}
		}
	}
	// cubic B-spline weights per sub-tile phase (scipy ni_interpolation.c, order 3):
	// output o samples u = (o + 0.5)/64 - 0.5; x = u - floor(u)
	std::vector<double> zw(64 * 4);
	for (int o = 0; o < 64; ++o) {
		const double u = ((double)o + 0.5) / 64.0 - 0.5;
		const double x = u - std::floor(u);
		const double y = x, z = 1.0 - x;
		double w0 = z * z * z / 6.0;
		double w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
		double w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
		double w3 = 1.0 - w0 - w1 - w2;
		zw[o * 4 + 0] = w0; zw[o * 4 + 1] = w1; zw[o * 4 + 2] = w2; zw[o * 4 + 3] = w3;
	}
	std::vector<double2> tw(TBK_KDE_M / 2);
	for (int k = 0; k < TBK_KDE_M / 2; ++k) {
		const double ang = -2.0 * M_PI * (double)k / (double)TBK_KDE_M;
		tw[k] = make_double2(std::cos(ang), std::sin(ang));
	}
	{
		// IDW pattern cache: zeroed device words, written only by k_mesh_finalize
		const size_t nwords = (P.ntiles + 31) / 32;
		const size_t words = 1 + TBK_IDW_CACHE * (1 + nwords + ((size_t)P.ntiles * 10 + 1) / 2);
		void* d = nullptr;
		CUDA_TRY(cudaMalloc(&d, words * sizeof(uint32_t)));
		p->allocs.push_back(d);
		CUDA_TRY(cudaMemset(d, 0, words * sizeof(uint32_t)));
		P.idw_cache = (uint32_t*)d;
	}
	// rings by decreasing sample count: the KDE grid starts its longest CTAs first, so the tail of the launch is short work
	std::vector<int> ring_order(std::max(P.nrings, 1), 0);
	for (int k = 0; k < P.nrings; ++k) ring_order[k] = k;
	if (P.nrings > 0) std::stable_sort(ring_order.begin(), ring_order.end(), [&](int a, int b) { return ring_ptr[a + 1] - ring_ptr[a] > ring_ptr[b + 1] - ring_ptr[b]; });
	int rc;
	if ((rc = upload(p, ring_order, &P.ring_order)) || (rc = upload(p, ring_ptr, &P.ring_ptr)) || (rc = upload(p, ring_pix, &P.ring_pix)) ||
		(rc = upload(p, nonflat, &P.nonflat_tiles)) || (rc = upload(p, tile_slot, &P.tile_slot)) ||
		(rc = upload(p, zw, &P.zoom_w)) || (rc = upload(p, tw, &P.twiddle)) ||
		(rc = upload(p, nonflat_r, &P.nonflat_r)) || (rc = upload(p, nonflat_rr, &P.nonflat_rr)) || (rc = upload(p, nonflat_uj, &P.nonflat_uj)) || (rc = upload(p, nonflat_jlo, &P.nonflat_jlo)) || (rc = upload(p, ringtile_id, &P.ringtile_id)) || (rc = upload(p, ringtile_ptr, &P.ringtile_ptr)) || (rc = upload(p, ringtile_ent, &P.ringtile_ent)) || (rc = upload(p, ringtile_idx, &P.ringtile_idx))) {
		tbk_plan_destroy(p);
		return rc;
	}
	*out = p;
	return TBK_OK;
}

// side stream + events for the caller's stream (created on first use)
static const TbkSide* side_for(tbk_plan* p, cudaStream_t st)
{
	std::lock_guard<std::mutex> lock(p->mtx);
	auto it = p->sides.find(st);
	if (it != p->sides.end()) return &it->second;
	TbkSide s;
	if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
	if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
	return &(p->sides[st] = s);
}

extern "C" int tbk_plan_destroy(tbk_plan* p)
{
	if (!p) return TBK_OK;
	DeviceGuard guard(p->device);
	for (auto& kv : p->sides) { cudaStreamDestroy(kv.second.stream); cudaEventDestroy(kv.second.fork); cudaEventDestroy(kv.second.join); }
	for (void* d : p->allocs) cudaFree(d);
	delete p;
	return TBK_OK;
}

extern "C" int tbk_plan_num_rings(const tbk_plan* p) { return p ? p->dev.nrings : 0; }

extern "C" size_t tbk_workspace_bytes(const tbk_plan* p, int B)
{
	if (!p || B <= 0) return 0;
	return layout(p, B).total;
}

extern "C" int tbk_fit_batch(tbk_plan* p, const float* cube, int B, const tbk_ffi_meta* meta,
	const uint8_t* extra_mask, float* bkg_out, uint8_t* mask_out, tbk_ffi_status* status,
	void* workspace, void* stream)
{
	if (!p || !cube || !bkg_out || !mask_out || !workspace || B <= 0) { tbk_set_error("tbk_fit_batch: NULL argument or B <= 0"); return TBK_ERR_INVALID; }
	DeviceGuard guard(p->device);
	if (p->dev.is_tess && !meta) { tbk_set_error("tbk_fit_batch: meta is required for TESS plans"); return TBK_ERR_INVALID; }
	if (((uintptr_t)cube | (uintptr_t)bkg_out) & 15 || ((uintptr_t)mask_out & 3) || ((uintptr_t)extra_mask & 3) || ((uintptr_t)workspace & 255)) {
		tbk_set_error("tbk_fit_batch: misaligned pointer (cube/bkg 16 B, masks 4 B, workspace 256 B)");
		return TBK_ERR_INVALID;
	}
	Workspace ws = carve(p, workspace, B);
	return tbk_launch_fit(p->dev, ws, cube, B, meta, extra_mask, bkg_out, mask_out, status, (cudaStream_t)stream, nullptr, p->tile_kernel,
		side_for(p, (cudaStream_t)stream));
}

extern "C" int tbk_fit_batch_profiled(tbk_plan* p, const float* cube, int B, const tbk_ffi_meta* meta,
	const uint8_t* extra_mask, float* bkg_out, uint8_t* mask_out, tbk_ffi_status* status,
	void* workspace, void* stream, float* ms)
{
	if (!p || !cube || !bkg_out || !mask_out || !workspace || !ms || B <= 0) { tbk_set_error("tbk_fit_batch_profiled: NULL argument or B <= 0"); return TBK_ERR_INVALID; }
	DeviceGuard guard(p->device);
	Workspace ws = carve(p, workspace, B);
	return tbk_launch_fit(p->dev, ws, cube, B, meta, extra_mask, bkg_out, mask_out, status, (cudaStream_t)stream, ms, p->tile_kernel, nullptr);
}

extern "C" unsigned long long tbk_launch_count(void) { return tbk_launch_counter(); }

extern "C" int tbk_time_smooth(tbk_plan* p, const float* bkg, int n, int w,
	const float* halo_lo, int n_lo, const float* halo_hi, int n_hi, float* out, void* stream)
{
	if (!p || !bkg || !out || n <= 0 || w < 0) { tbk_set_error("tbk_time_smooth: bad argument"); return TBK_ERR_INVALID; }
	DeviceGuard guard(p->device);
	if ((n_lo > 0 && !halo_lo) || (n_hi > 0 && !halo_hi) || n_lo < 0 || n_hi < 0) { tbk_set_error("tbk_time_smooth: bad halo"); return TBK_ERR_INVALID; }
	return tbk_launch_time_smooth(p->dev.H, p->dev.W, bkg, n, w, halo_lo, n_lo, halo_hi, n_hi, out, (cudaStream_t)stream);
}

extern "C" int tbk_sum_accumulate(tbk_plan* p, const float* cube, const float* bkg_smooth,
	uint8_t* flags, const tbk_ffi_meta* meta, int n, float* flux_out,
	double* sum, int32_t* nimg, int32_t* used, void* stream)
{
	if (!p || !cube || !bkg_smooth || !flags || !meta || !sum || !nimg || !used || n <= 0) { tbk_set_error("tbk_sum_accumulate: bad argument"); return TBK_ERR_INVALID; }
	DeviceGuard guard(p->device);
	// per-call scratch (one flag per cadence) from the stream-ordered allocator: nothing shared between concurrent calls
	int* zero_flags = nullptr;
	CUDA_TRY(cudaMallocAsync((void**)&zero_flags, sizeof(int) * (size_t)n, (cudaStream_t)stream));
	const int rc = tbk_launch_sum_accumulate(p->dev, cube, bkg_smooth, flags, meta, n, flux_out, sum, nimg, used, zero_flags, (cudaStream_t)stream);
	cudaFreeAsync(zero_flags, (cudaStream_t)stream);
	return rc;
}

extern "C" int tbk_pack_mask(const uint8_t* mask, size_t nbytes, uint8_t* bits, void* stream)
{
	if (!mask || !bits || (nbytes % 32) != 0 || (((uintptr_t)mask) & 15) || (((uintptr_t)bits) & 3)) { tbk_set_error("tbk_pack_mask: bad argument (nbytes must be a multiple of 32, mask 16-byte aligned)"); return TBK_ERR_INVALID; }
	return tbk_launch_pack_mask(mask, nbytes, bits, (cudaStream_t)stream);
}

// host side of the same format: out[8 i + j] = bit (7 - j) of bits[i] (0 / 1 bytes); runs on the calling CPU thread
extern "C" int tbk_unpack_mask_host(const uint8_t* bits, size_t nbits_bytes, uint8_t* out)
{
	if (!bits || !out) { tbk_set_error("tbk_unpack_mask_host: bad argument"); return TBK_ERR_INVALID; }
	for (size_t i = 0; i < nbits_bytes; ++i) {
		// byte j of (b * 0x8040201008040201 & 0x8080808080808080) >> 7 is bit (7 - j) of b
		const unsigned long long x = ((((unsigned long long)bits[i] * 0x8040201008040201ULL) & 0x8080808080808080ULL) >> 7);
		std::memcpy(out + 8 * i, &x, 8);
	}
	return TBK_OK;
}

extern "C" int tbk_sum_finalize(tbk_plan* p, const double* sum, const int32_t* nimg, const int32_t* used,
	int numfiles, double threshold, double* sumimage, uint8_t* pixels_used, void* stream)
{
	if (!p || !sum || !nimg || !used || !sumimage || !pixels_used || numfiles <= 0) { tbk_set_error("tbk_sum_finalize: bad argument"); return TBK_ERR_INVALID; }
	DeviceGuard guard(p->device);
	return tbk_launch_sum_finalize(p->dev.H, p->dev.W, sum, nimg, used, numfiles, threshold, sumimage, pixels_used, (cudaStream_t)stream);
}

extern "C" int tbk_debug_fetch(tbk_plan* p, const void* workspace, int B, int b, int round, double* s2, double* mesh)
{
	if (!p || !workspace || b < 0 || b >= B || round < 0 || round >= p->dev.bkgiters) { tbk_set_error("tbk_debug_fetch: bad argument"); return TBK_ERR_INVALID; }
	DeviceGuard guard(p->device);
	Workspace ws = carve(p, const_cast<void*>(workspace), B);
	const PlanDev& P = p->dev;
	if (s2 && P.nrings > 0)
		CUDA_TRY(cudaMemcpy(s2, ws.s2_hist + ((size_t)b * P.bkgiters + round) * P.nrings, sizeof(double) * P.nrings, cudaMemcpyDeviceToHost));
	if (mesh)
		CUDA_TRY(cudaMemcpy(mesh, ws.mesh_hist + ((size_t)b * P.bkgiters + round) * P.ntiles, sizeof(double) * P.ntiles, cudaMemcpyDeviceToHost));
	return TBK_OK;
}

extern "C" int tbk_workspace_layout(const tbk_plan* p, int B, size_t* offsets, size_t* sizes)
{
	if (!p || B <= 0 || !offsets || !sizes) { tbk_set_error("tbk_workspace_layout: bad argument"); return TBK_ERR_INVALID; }
	const WsLayout L = layout(p, B);
	const WsLayout* q = &L;
	offsets[0] = q->off_ctl; offsets[1] = q->off_base; offsets[2] = q->off_nf; offsets[3] = q->off_coef;
	offsets[4] = q->off_mesh; offsets[5] = q->off_s2raw; offsets[6] = q->off_s2hist; offsets[7] = q->off_ringv; offsets[8] = q->off_fb;
	sizes[0] = sizeof(FfiCtl); sizes[1] = sizeof(TileStat); sizes[2] = (size_t)p->dev.n_nonflat;
	return TBK_OK;
}

extern "C" int tbk_bkgshe_indicator(const float* images, const double* sumimage, int B, int H, int W, float* ind_out, void* stream)
{
	if (!images || !ind_out || B <= 0 || H <= 0 || W <= 0 || B > 65535) { tbk_set_error("tbk_bkgshe_indicator: bad argument"); return TBK_ERR_INVALID; }
	return tbk_launch_bkgshe_indicator(images, sumimage, B, H, W, ind_out, (cudaStream_t)stream);
}

extern "C" int tbk_bkgshe_mean(const float* ind, size_t stride, size_t npix, int n, const int32_t* order, double* mean_out, void* stream)
{
	if (!ind || !order || !mean_out || n <= 0 || npix == 0 || stride < npix) { tbk_set_error("tbk_bkgshe_mean: bad argument"); return TBK_ERR_INVALID; }
	return tbk_launch_bkgshe_mean(ind, stride, npix, n, order, mean_out, (cudaStream_t)stream);
}

extern "C" int tbk_bkgshe_flag(const float* ind, const double* mean, int B, size_t npix, double threshold, int bit,
	uint8_t* flags, void* stream)
{
	if (!ind || !mean || !flags || B <= 0 || B > 65535 || npix == 0 || bit <= 0 || bit > 255) { tbk_set_error("tbk_bkgshe_flag: bad argument"); return TBK_ERR_INVALID; }
	return tbk_launch_bkgshe_flag(ind, mean, B, npix, threshold, bit, flags, (cudaStream_t)stream);
}

extern "C" int tbk_gather_stamps(const void* stack, int elem_bytes, int N, int H, int W, const int32_t* stamps,
	const int64_t* out_offsets, int S, void* out, void* stream)
{
	if (!stack || !stamps || !out_offsets || !out || (elem_bytes != 1 && elem_bytes != 4) || N <= 0 || H <= 0 || W <= 0 || S <= 0 || S > 65535
		|| ((uintptr_t)stamps & 15)) {
		tbk_set_error("tbk_gather_stamps: bad argument"); return TBK_ERR_INVALID;
	}
	// enough CTAs per stamp to fill the device for a handful of stamps, few enough that thousands of stamps stay cheap
	const int tiles_x = S >= 1184 ? 1 : (1184 + S - 1) / S;
	return tbk_launch_gather_stamps(stack, elem_bytes, N, H, W, (const int*)stamps, (const long long*)out_offsets, S, tiles_x, out, (cudaStream_t)stream);
}

extern "C" int tbk_star_mask(const double* stars, int S, int H, int W, uint8_t* mask, void* stream)
{
	if (!stars || !mask || S <= 0 || H <= 0 || W <= 0) { tbk_set_error("tbk_star_mask: bad argument"); return TBK_ERR_INVALID; }
	return tbk_launch_star_mask(stars, S, H, W, mask, (cudaStream_t)stream);
}

extern "C" int tbk_motion_prepare(const float* images, int B, int H, int W, float* prepared, void* scratch, void* stream)
{
	if (!images || !prepared || !scratch || B <= 0 || H < 3 || W < 3 || B > 65535) { tbk_set_error("tbk_motion_prepare: bad argument"); return TBK_ERR_INVALID; }
	return tbk_launch_motion_prepare(images, B, H, W, prepared, (unsigned*)scratch, (cudaStream_t)stream);
}

extern "C" size_t tbk_motion_workspace_bytes(int B, int H, int W)
{
	if (B <= 0 || H <= 0 || W <= 0) return 0;
	return tbk_motion_workspace(B, H, W);
}

extern "C" int tbk_motion_ecc(const float* ref_prepared, const float* prepared, int B, int H, int W, int max_iter, double eps,
	void* workspace, double* out, void* stream)
{
	if (!ref_prepared || !prepared || !workspace || !out || B <= 0 || B > 65535 || H < 5 || W < 5 || max_iter < 1 || ((uintptr_t)workspace & 255)) {
		tbk_set_error("tbk_motion_ecc: bad argument"); return TBK_ERR_INVALID;
	}
	return tbk_launch_motion_ecc(ref_prepared, prepared, B, H, W, max_iter, eps, workspace, out, (cudaStream_t)stream);
}

extern "C" int tbk_debug_log10(const double* in, double* out, int n, void* stream)
{
	if (!in || !out || n <= 0) { tbk_set_error("tbk_debug_log10: bad argument"); return TBK_ERR_INVALID; }
	return tbk_launch_log10(in, out, n, (cudaStream_t)stream);
}

extern "C" int tbk_decode_ffi_be(const uint8_t* raw, int B, int naxis1, int naxis2, int row0, int col0, int H, int W,
	float* cube_out, void* stream)
{
	if (!raw || !cube_out || B <= 0 || H <= 0 || W <= 0 || W % 4 || row0 < 0 || col0 < 0 || row0 + H > naxis2 || col0 + W > naxis1
		|| ((uintptr_t)raw & 3) || ((uintptr_t)cube_out & 15) || H > 65535 || B > 65535) {
		tbk_set_error("tbk_decode_ffi_be: bad argument"); return TBK_ERR_INVALID;
	}
	return tbk_launch_decode(raw, B, naxis1, naxis2, row0, col0, H, W, cube_out, (cudaStream_t)stream);
}

// Host execution of the kd-tree restatement (tbk_kdtree.cuh) for the CPU tests against scipy.spatial.cKDTree: the same
// functions k_mesh_finalize runs on the device.  good: uint8 [ny*nx] (non-zero = good mesh).  Outputs (host pointers, any
// may be NULL): nbr_id / nbr_d2 int32 [ny*nx][10] (-1 / -1 beyond the number of good meshes), idx_out int32 [ngood] the
// tree's index permutation as good-point indices, nodes_out int32 [nnodes][4] = (dim, split, a, b) in creation order.
extern "C" int tbk_debug_idw_neighbors(const uint8_t* good, int ny, int nx, int32_t* nbr_id, int32_t* nbr_d2,
	int32_t* idx_out, int32_t* nodes_out, int32_t* nnodes_out)
{
	if (!good || ny <= 0 || nx <= 0 || ny * nx > 4096) { tbk_set_error("tbk_debug_idw_neighbors: bad argument"); return TBK_ERR_INVALID; }
	const int nt = ny * nx;
	std::vector<uint32_t> idx;
	std::vector<int> rank_of(nt, -1);
	for (int g = 0; g < nt; ++g) if (good[g]) { rank_of[g] = (int)idx.size(); idx.push_back(kdt_pack(g, nx)); }
	std::vector<KdtNode> nodes(2 * idx.size() + 2);
	KdtTree t;
	t.idx = idx.data(); t.nodes = nodes.data(); t.npts = (int)idx.size(); t.nx = nx;
	int stack[3 * 64];
	kdt_build(t, stack, (int)nodes.size());
	if (t.overflow) { tbk_set_error("tbk_debug_idw_neighbors: tree overflow"); return TBK_ERR_INVALID; }
	if (idx_out) for (int i = 0; i < t.npts; ++i) idx_out[i] = rank_of[kdt_id(idx[i])];
	if (nnodes_out) *nnodes_out = t.nnodes;
	if (nodes_out) for (int i = 0; i < t.nnodes; ++i) {
		nodes_out[4 * i] = nodes[i].dim; nodes_out[4 * i + 1] = nodes[i].split; nodes_out[4 * i + 2] = nodes[i].a; nodes_out[4 * i + 3] = nodes[i].b;
	}
	if (nbr_id || nbr_d2) {
		for (int g = 0; g < nt; ++g) {
			int id[KDT_K], d2[KDT_K], ovf = 0;
			const int m = kdt_query(t, g / nx, g % nx, KDT_K, id, d2, &ovf);
			if (ovf) { tbk_set_error("tbk_debug_idw_neighbors: query queue overflow"); return TBK_ERR_INVALID; }
			for (int j = 0; j < KDT_K; ++j) {
				if (nbr_id) nbr_id[g * KDT_K + j] = j < m ? id[j] : -1;
				if (nbr_d2) nbr_d2[g * KDT_K + j] = j < m ? d2[j] : -1;
			}
		}
	}
	return TBK_OK;
}
