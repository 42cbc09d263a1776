// tbk_fit.cu -- kernels of the batched fit_background path (photometry/backgrounds.py:86-211).
#include "tbk_common.cuh"
#include "tbk_tile.cuh"
#include "tbk_tile_warp.cuh"
#include "tbk_tile_zone.cuh"
#include "tbk_zoom.cuh"
#include "tbk_internal.h"
#include "tbk_kdtree.cuh"

// ---------------------------------------------------------------------------------------------
// K_init: per-FFI control block; manual excludes from header scalars (pixel_flags.py:34-50).
__global__ void k_init_ctl(PlanDev P, Workspace ws, const tbk_ffi_meta* __restrict__ meta, int B)
{
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b < 64) ws.fb_count[b] = 0;
	if (b >= B) return;
	FfiCtl& c = ws.ctl[b];
	c.min_bits = 0x7f800000u;
	c.any_nonzero = 0;
	c.n_valid = 0;
	c.all_masked = 0;
	c.no_good_mesh = 0;
	c.radial_ok = 0;
	c.npts = 0;
	c.mesh_const = 0;
	c.kde_fallbacks = 0;
	c.min_key = ~0ULL;
	c.min_ub = ~0ULL;
	ws.idw_bits[(size_t)b * ((P.ntiles + 31) / 32 + 1)] = 0u;   // no IDW neighbour table yet for this FFI
	c.zp = 0.0; c.c_flat = 0.0; c.x0 = 0.0; c.xlast = 0.0;
	c.mesh_min = 0.0; c.mesh_max = 0.0;
	int mars = 0, earth = 0;
	if (P.is_tess) {
		const tbk_ffi_meta m = meta[b];
		const double time = 0.5 * (m.tstart + m.tstop);
		const int cad = m.cadenceno;
		if (P.camera == 1 && P.ccd == 4 && (cad <= 4724 || m.tstart <= 1325.881282301840)) mars = 1;
		else if (P.camera == 1 && ((cad >= 11354 && cad <= 11366) || (time >= 1464.0158778 && time <= 1464.265871))) earth = 1;
	}
	c.mars = mars; c.earth = earth;
}

// ---------------------------------------------------------------------------------------------
// K_tile_base: mask build (backgrounds.py:91-97), per-FFI min / flags, and sigma-clipped statistics
// of every mesh on the raw pixels.  Meshes that can see r > first ring centre ("non-flat") are
// re-evaluated each round by k_tile_round instead.
__global__ void __launch_bounds__(TBK_NT) k_tile_base(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ extra, uint8_t* __restrict__ mask_out)
{
	__shared__ TileSmem sm;
	__shared__ int s_flag;
	const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
	const int ty = tile / P.nx, tx = tile % P.nx;
	FfiCtl& c = ws.ctl[b];
	const size_t img = (size_t)b * P.H * P.W;
	const int lcol = tile_lcol(tid);
	const int gx = tx * TBK_TILE + lcol;
	const bool mars_cols = c.mars && gx >= 1536;
	const bool earth = c.earth;

	float v[TBK_VPT];
	unsigned valid = 0;
	bool nonzero = false;
	float4 raw[4];
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int gy = ty * TBK_TILE + tile_lrow(tid, j);
		raw[j] = __ldg(reinterpret_cast<const float4*>(cube + img + (size_t)gy * P.W + gx));
	}
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int gy = ty * TBK_TILE + tile_lrow(tid, j);
		const size_t off = img + (size_t)gy * P.W + gx;
		const float x4[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
		uchar4 ex = make_uchar4(0, 0, 0, 0);
		if (extra) ex = __ldg(reinterpret_cast<const uchar4*>(extra + off));
		const unsigned char e4[4] = {ex.x, ex.y, ex.z, ex.w};
		unsigned char m4[4];
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const float x = x4[q];
			nonzero |= !(x == 0.0f);
			const bool ok = pixel_valid(x, P.flux_cutoff) && !mars_cols && !earth && !e4[q];
			v[4 * j + q] = x + 0.0f;  // -0.0 -> +0.0
			valid |= (ok ? 1u : 0u) << (4 * j + q);
			m4[q] = ok ? 0 : 1;
		}
		*reinterpret_cast<uchar4*>(mask_out + off) = make_uchar4(m4[0], m4[1], m4[2], m4[3]);
	}

	// per-FFI reductions: any pixel != 0, number of valid pixels, min of valid pixels
	if (tid == 0) s_flag = 0;
	__syncthreads();
	if (nonzero) s_flag = 1;
	int cnt = __popc(valid);
	float mnf = INFINITY;
#pragma unroll
	for (int e = 0; e < TBK_VPT; ++e) if (valid >> e & 1u) mnf = fminf(mnf, v[e]);
	double mn = mnf, mx = 0.0;
	block_sum_min_max(sm.red, cnt, mn, mx);
	if (tid == 0) {
		if (s_flag && !c.any_nonzero) atomicOr(&c.any_nonzero, 1);
		if (cnt > 0) {
			atomicAdd(&c.n_valid, cnt);
			atomicMin(&c.min_bits, __float_as_uint((float)mn));
		}
	}

	const bool nonflat = P.use_radial && P.tile_slot[tile] >= 0;
	if (nonflat) {   // re-evaluated every round by k_tile_round; the slot gets a defined "no statistics" entry
		if (tid == 0) { TileStat z; z.mean = z.med = z.std = nan_d(); z.nfin = 0; z.pad = 0; ws.tile_base[(size_t)b * P.ntiles + tile] = z; }
		return;
	}
	TileStat st = tile_sigma_clip<float>(v, valid, sm);
	if (tid == 0) ws.tile_base[(size_t)b * P.ntiles + tile] = st;
}

// K_tile_base_w3: the same outputs as k_tile_base with the bucketed algorithm of tbk_tile_warp.cuh: two warps per mesh,
// the validated keys staged in shared memory, per-pixel loops rolled (small code: a fully unrolled register version is
// bound by instruction fetch).  Also records the minimum valid pixel of every 8x8 sub-block (ws.sbmin) for the pruned
// zeropoint pass.  Thread t, load i (0..15) owns row 4i + 2(t/32) + (t%32)/16, columns 4(t%16) .. +3.
template <bool HAS_EXTRA, int NW>
__global__ void __launch_bounds__(32 * NW, 20 / NW) k_tile_base_w3(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ extra, uint8_t* __restrict__ mask_out)
{
	__shared__ TwBlockSmem<TwF32, NW> sm;
	constexpr int LPA = 4 / NW;   // loads per 8-row band: a load step covers 2 NW rows
	__shared__ uint32_t s_sb[64];
	__shared__ int s_nbad, s_nz;
	const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int ty = tile / P.nx, tx = tile % P.nx;
	FfiCtl& c = ws.ctl[b];
	const int lrow0 = 2 * w + (lane >> 4), lcol = (lane & 15) << 2;
	const size_t base = (size_t)b * P.H * P.W + (size_t)(ty * TBK_TILE + lrow0) * P.W + tx * TBK_TILE + lcol;
	const size_t step = (size_t)(2 * NW) * P.W;
	const bool excl = (c.mars && tx * TBK_TILE >= 1536) || c.earth;
	const uint32_t cut = excl ? 0u : __float_as_uint(P.flux_cutoff);
	if (tid < 64) s_sb[tid] = TW_INVALID;
	if (tid == 0) { s_nbad = 0; s_nz = 0; }
	__syncthreads();
	uint32_t nz = 0u;
	int nbad = 0;
	// all of this thread's pixels are fetched at once with cp.async (16 B each) into the key buffer, so the
	// HBM latency is paid once per mesh; every thread later reads back only what it copied itself
	{
#pragma unroll
		for (int i = 0; i < 8 * LPA; ++i) {
			const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.tw.keys[(lrow0 + 2 * NW * i) * TBK_TILE + lcol]);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(cube + base + (size_t)i * step));
		}
		asm volatile("cp.async.commit_group;");
		asm volatile("cp.async.wait_group 0;" ::: "memory");
	}
#pragma unroll 1
	for (int a = 0; a < 8; ++a) {   // rows 8a .. 8a+7 of the mesh: loads 2a and 2a+1 of both warps
		float4 r[LPA];
#pragma unroll
		for (int h = 0; h < LPA; ++h) r[h] = *reinterpret_cast<const float4*>(&sm.tw.keys[(lrow0 + 2 * NW * (LPA * a + h)) * TBK_TILE + lcol]);
		uint32_t smin = TW_INVALID;
#pragma unroll
		for (int h = 0; h < LPA; ++h) {
			const size_t off = base + (size_t)(LPA * a + h) * step;
			uint32_t ex = 0u;
			if (HAS_EXTRA) ex = __ldg(reinterpret_cast<const unsigned int*>(extra + off));
			const float x4[4] = {r[h].x, r[h].y, r[h].z, r[h].w};
			uint32_t kk[4], m = 0u;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t k = __float_as_uint(x4[q] + 0.0f);   // -0.0 -> +0.0
				nz |= k;
				bool ok = (k <= cut) && !excl;
				if (HAS_EXTRA) ok = ok && !((ex >> (8 * q)) & 0xFFu);
				m |= ok ? 0u : (1u << (8 * q));
				kk[q] = ok ? k : TW_INVALID;
				smin = min(smin, kk[q]);
			}
			nbad += __popc(m);
			*reinterpret_cast<unsigned int*>(mask_out + off) = m;
			*reinterpret_cast<uint4*>(&sm.tw.keys[(lrow0 + 2 * NW * (LPA * a + h)) * TBK_TILE + lcol]) = make_uint4(kk[0], kk[1], kk[2], kk[3]);
		}
		smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, 1));
		smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, 16));
		if ((lane & 17) == 0) atomicMin(&s_sb[a * 8 + (lane >> 1)], smin);
	}
	nbad = __reduce_add_sync(0xffffffffu, nbad);
	nz = __reduce_or_sync(0xffffffffu, nz);
	if (lane == 0) { atomicAdd(&s_nbad, nbad); if (nz) atomicOr(&s_nz, 1); }
	__syncthreads();
	const uint32_t mysb = s_sb[tid & 63];
	if (tid < 64) ws.sbmin[((size_t)b * P.ntiles + tile) * 64 + tid] = __uint_as_float(mysb);
	const int n = 4096 - s_nbad;
	if (tid == 0) {
		if (s_nz && !c.any_nonzero) atomicOr(&c.any_nonzero, 1);
		if (n > 0) atomicAdd(&c.n_valid, n);
	}
	if (w == 0) {
		uint32_t kmin = min(s_sb[lane], s_sb[lane + 32]);
		kmin = __reduce_min_sync(0xffffffffu, kmin);
		if (lane == 0 && kmin != TW_INVALID) atomicMin(&c.min_bits, kmin);
	}
	if (P.use_radial && P.tile_slot[tile] >= 0) {   // re-evaluated every round by k_tile_round; defined "no statistics" entry
		if (tid == 0) { TileStat z; z.mean = z.med = z.std = nan_d(); z.nfin = 0; z.pad = 0; ws.tile_base[(size_t)b * P.ntiles + tile] = z; }
		return;
	}
	TileStat st; bool writer;
	tile_block_stats_staged<TwF32, NW>(sm, n, st, writer);
	if (writer) ws.tile_base[(size_t)b * P.ntiles + tile] = st;
}

// K_tile_base_z: the same outputs as k_tile_base_w3 with the zone algorithm of tbk_tile_zone.cuh: one warp per mesh
// (four meshes per CTA, no block barrier), pass 1 from HBM straight into registers, pass 2 from L2.  Lane l, load i
// (0..31) owns row 2i + l/16, columns 4(l%16) .. +3, so every warp load covers two full 256 B row segments.
// Meshes the lists cannot answer are queued in ws.fb_list for k_tile_base_fb.
//
// The per-pixel bodies are written as predicated PTX: the compiler's own if-conversion of the C++ form spends ~1.7x
// the instructions (select-based counters, recomputed predicates, per-append address arithmetic).
#define ZB_WARPS 4
#define ZB_RETRY_TCAP 136   // tail elements per lane in the retry launch (a lane sees 128 pixels of its mesh)

// Pass 1, one pixel.  k = float bits of x (x >= +0 when valid), ex = extra-mask byte (0 = usable).
//   valid  = k <= cut && !ex;  kv = valid ? k : +inf key;  mask byte Q set when !valid;  smin = min(smin, kv)
//   bulk   = kA <= kv <= kA + spanAB:  s1 += x - pivot, s2 += (x - pivot)^2  (float64; a non-bulk pixel enters as the
//            pivot itself -- the pivot is a float32 value -- i.e. as an exact zero, which keeps the float64 adds unpredicated)
//   tail   = valid && !bulk: appended to this lane's list (pointer tptr, stride 128 B, not stored beyond tend)
//   nA    += kv < kA;  nM += kv < kM;  every fourth pixel: nC += 4 * (kC <= kv <= kC + spanC)
template <int Q, bool HAS_EXTRA>
__device__ __forceinline__ void zb_p1(float x, uint32_t ex, uint32_t cut, uint32_t kA, uint32_t spanAB, uint32_t kM, double pivot,
	uint32_t& m, uint32_t& smin, int& nA, int& nM, double& s1, double& s2, uint32_t& tptr, uint32_t tend, float pivot_f,
	uint32_t kC, uint32_t spanC, int& nC)
{
#define ZB_P1_HEAD \
		".reg .pred pok, pbulk, plow, pm, pt, pst, pc;\n\t" \
		".reg .u32 k, kv, t;\n\t" \
		".reg .f64 d;\n\t" \
		".reg .f32 xm;\n\t" \
		"mov.b32 k, %9;\n\t" \
		"setp.le.u32 pok, k, %11;\n\t"
#define ZB_P1_TAIL \
		"selp.u32 kv, k, 0x7f800000, pok;\n\t" \
		"min.u32 %1, %1, kv;\n\t" \
		"sub.u32 t, kv, %12;\n\t" \
		"setp.le.u32 pbulk, t, %13;\n\t" \
		"setp.lt.u32 plow, kv, %12;\n\t" \
		"setp.lt.u32 pm, kv, %14;\n\t" \
		"@plow add.s32 %2, %2, 1;\n\t" \
		"@pm add.s32 %3, %3, 1;\n\t" \
		"selp.f32 xm, %9, %17, pbulk;\n\t" \
		"cvt.f64.f32 d, xm;\n\t" \
		"sub.f64 d, d, %15;\n\t" \
		"add.f64 %4, %4, d;\n\t" \
		"fma.rn.f64 %5, d, d, %5;\n\t" \
		"and.pred pt, pok, !pbulk;\n\t" \
		"setp.lt.and.u32 pst, %6, %16, pt;\n\t" \
		"@pst st.shared.u32 [%6], kv;\n\t" \
		"@pt add.u32 %6, %6, 128;\n\t" \
		"@!pok or.b32 %0, %0, %8;\n\t"
	// the centre count nC only sizes the zone (local density), so every fourth pixel is enough for it
#define ZB_P1_NC \
		"sub.u32 t, kv, %18;\n\t" \
		"setp.le.u32 pc, t, %19;\n\t" \
		"@pc add.s32 %7, %7, 4;\n\t"
	if (HAS_EXTRA && Q == 0)
		asm volatile("{\n\t" ZB_P1_HEAD "setp.eq.and.u32 pok, %10, 0, pok;\n\t" ZB_P1_TAIL ZB_P1_NC "}"
			: "+r"(m), "+r"(smin), "+r"(nA), "+r"(nM), "+d"(s1), "+d"(s2), "+r"(tptr), "+r"(nC)
			: "n"(1u << (8 * Q)), "f"(x), "r"(ex), "r"(cut), "r"(kA), "r"(spanAB), "r"(kM), "d"(pivot), "r"(tend), "f"(pivot_f), "r"(kC), "r"(spanC) : "memory");
	else if (Q == 0)
		asm volatile("{\n\t" ZB_P1_HEAD ZB_P1_TAIL ZB_P1_NC "}"
			: "+r"(m), "+r"(smin), "+r"(nA), "+r"(nM), "+d"(s1), "+d"(s2), "+r"(tptr), "+r"(nC)
			: "n"(1u << (8 * Q)), "f"(x), "r"(ex), "r"(cut), "r"(kA), "r"(spanAB), "r"(kM), "d"(pivot), "r"(tend), "f"(pivot_f), "r"(kC), "r"(spanC) : "memory");
	else if (HAS_EXTRA)
		asm volatile("{\n\t" ZB_P1_HEAD "setp.eq.and.u32 pok, %10, 0, pok;\n\t" ZB_P1_TAIL "}"
			: "+r"(m), "+r"(smin), "+r"(nA), "+r"(nM), "+d"(s1), "+d"(s2), "+r"(tptr), "+r"(nC)
			: "n"(1u << (8 * Q)), "f"(x), "r"(ex), "r"(cut), "r"(kA), "r"(spanAB), "r"(kM), "d"(pivot), "r"(tend), "f"(pivot_f), "r"(kC), "r"(spanC) : "memory");
	else
		asm volatile("{\n\t" ZB_P1_HEAD ZB_P1_TAIL "}"
			: "+r"(m), "+r"(smin), "+r"(nA), "+r"(nM), "+d"(s1), "+d"(s2), "+r"(tptr), "+r"(nC)
			: "n"(1u << (8 * Q)), "f"(x), "r"(ex), "r"(cut), "r"(kA), "r"(spanAB), "r"(kM), "d"(pivot), "r"(tend), "f"(pivot_f), "r"(kC), "r"(spanC) : "memory");
}

// Pass 2, one pixel: zone = kZL <= kv <= kZL + zspan appended to this lane's zone list; nZL += kv < kZL.
template <bool HAS_EXTRA>
__device__ __forceinline__ void zb_p2(float x, uint32_t ex, uint32_t cut, uint32_t kZL, uint32_t zspan, int& nZL, uint32_t& zptr, uint32_t zend)
{
	if (HAS_EXTRA) {
		asm volatile("{\n\t"
			".reg .pred pok, pz, pl, pst;\n\t"
			".reg .u32 k, kv, t;\n\t"
			"mov.b32 k, %2;\n\t"
			"setp.le.u32 pok, k, %4;\n\t"
			"setp.eq.and.u32 pok, %3, 0, pok;\n\t"
			"selp.u32 kv, k, 0x7f800000, pok;\n\t"
			"sub.u32 t, kv, %5;\n\t"
			"setp.le.u32 pz, t, %6;\n\t"
			"setp.lt.u32 pl, kv, %5;\n\t"
			"@pl add.s32 %0, %0, 1;\n\t"
			"setp.lt.and.u32 pst, %1, %7, pz;\n\t"
			"@pst st.shared.u32 [%1], kv;\n\t"
			"@pz add.u32 %1, %1, 128;\n\t"
			"}"
			: "+r"(nZL), "+r"(zptr) : "f"(x), "r"(ex), "r"(cut), "r"(kZL), "r"(zspan), "r"(zend) : "memory");
	} else {
		// without an extra mask every key inside the zone, and every key below it, is a valid pixel (the zone lies
		// inside [0, cutoff]; negative / NaN / inf bit patterns compare above the cutoff)
		asm volatile("{\n\t"
			".reg .pred pz, pl, pst;\n\t"
			".reg .u32 k, t;\n\t"
			"mov.b32 k, %2;\n\t"
			"sub.u32 t, k, %3;\n\t"
			"setp.le.u32 pz, t, %4;\n\t"
			"setp.lt.u32 pl, k, %3;\n\t"
			"@pl add.s32 %0, %0, 1;\n\t"
			"setp.lt.and.u32 pst, %1, %5, pz;\n\t"
			"@pst st.shared.u32 [%1], k;\n\t"
			"@pz add.u32 %1, %1, 128;\n\t"
			"}"
			: "+r"(nZL), "+r"(zptr) : "f"(x), "r"(kZL), "r"(zspan), "r"(zend) : "memory");
	}
}

template <bool STAGED>
__device__ __forceinline__ float4 zb_ld(const float* p)
{
	return STAGED ? *reinterpret_cast<const float4*>(p) : __ldg(reinterpret_cast<const float4*>(p));
}
// STAGED = true is the measured alternative (TBK_TILE_KERNEL=7): the 16 KB mesh is brought into shared memory by 64
// TMA bulk row copies (cp.async.bulk + mbarrier, one barrier per warp) and both passes read it from there instead of
// from global memory / L2.  The tile costs 16 KB per warp on top of the 9.75 KB of lists: 8 resident warps per SM
// instead of 20 (DESIGN.md "Decisions recorded with numbers").
struct ZbStage {
	float px[ZB_WARPS][TBK_NPIX_TILE];
	unsigned long long bar[ZB_WARPS];
};
// RETRY = true: the launch works off the retry queue (ws.rt_list, count in fb_count[8]); a
// warp takes queue entry blockIdx.x * ZB_WARPS + w, the plan comes from the statistics parked in the mesh's TileStat.
template <bool HAS_EXTRA, bool STAGED, bool RETRY>
__global__ void __launch_bounds__(32 * ZB_WARPS, STAGED ? 2 : 5) k_tile_base_z(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ extra, uint8_t* __restrict__ mask_out, int rt_cap)
{
	// the retry launch runs one warp per CTA with tail lists long enough for any mesh (128 pixels per lane)
	constexpr int TCAP = RETRY ? ZB_RETRY_TCAP : ZN_TCAP, NW = RETRY ? 1 : ZB_WARPS;
	__shared__ ZoneSmem<Zn32, TCAP> smw[NW];
	extern __shared__ __align__(128) unsigned char zb_dyn[];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	int tile = blockIdx.x * ZB_WARPS + w, b = blockIdx.y;
	if (RETRY) {
		const int e = blockIdx.x;
		if (e >= min(ws.fb_count[8], rt_cap)) return;
		const int ent = ws.rt_list[e];
		b = ent / P.ntiles; tile = ent % P.ntiles;
	}
	if (tile >= P.ntiles) return;
	ZoneSmem<Zn32, TCAP>& sm = smw[w];
	const int ty = tile / P.nx, tx = tile % P.nx;
	FfiCtl& c = ws.ctl[b];
	const size_t tile0 = (size_t)b * P.H * P.W + (size_t)(ty * TBK_TILE) * P.W + tx * TBK_TILE;
	const size_t base = tile0 + (size_t)(lane >> 4) * P.W + ((lane & 15) << 2);
	const size_t step = (size_t)2 * P.W;
	// where the two passes read the pixels: global memory, or the staged tile (row stride 64)
	const float* px0 = cube + base;
	size_t pstep = step;
	if (STAGED) {
		ZbStage& stg = *reinterpret_cast<ZbStage*>(zb_dyn);
		if (lane == 0) { tma_bar_init(&stg.bar[w], 1); tma_bar_expect(&stg.bar[w], (unsigned)(TBK_NPIX_TILE * sizeof(float))); }
		__syncwarp();
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			const int row = lane + 32 * h;
			tma_load_1d(&stg.px[w][row * TBK_TILE], cube + tile0 + (size_t)row * P.W, (unsigned)(TBK_TILE * sizeof(float)), &stg.bar[w]);
		}
		px0 = &stg.px[w][(lane >> 4) * TBK_TILE + ((lane & 15) << 2)];
		pstep = 2 * TBK_TILE;
		tma_bar_wait(&stg.bar[w], 0u);
	}
	const uint32_t cut = __float_as_uint(P.flux_cutoff);
	float* sbdst = ws.sbmin + ((size_t)b * P.ntiles + tile) * 64;
	// manual excludes are mesh-uniform (the Mars boundary, column 1536, is a multiple of the mesh size): nothing is valid
	if ((c.mars && tx * TBK_TILE >= 1536) || c.earth) {
		uint32_t nz = 0u;
		for (int i = 0; i < 32; ++i) {
			const float4 r = zb_ld<STAGED>(px0 + (size_t)i * pstep);
			nz |= __float_as_uint(r.x + 0.0f) | __float_as_uint(r.y + 0.0f) | __float_as_uint(r.z + 0.0f) | __float_as_uint(r.w + 0.0f);
			*reinterpret_cast<unsigned int*>(mask_out + base + (size_t)i * step) = 0x01010101u;
		}
		sbdst[lane] = __uint_as_float(TW_INVALID); sbdst[lane + 32] = __uint_as_float(TW_INVALID);
		nz = __reduce_or_sync(0xffffffffu, nz);
		if (lane == 0) {
			if (nz && !c.any_nonzero) atomicOr(&c.any_nonzero, 1);
			TileStat st; st.mean = st.med = st.std = nan_d(); st.nfin = 0; st.pad = 0;
			ws.tile_base[(size_t)b * P.ntiles + tile] = st;
		}
		return;
	}
	const bool do_stats = !(P.use_radial && P.tile_slot[tile] >= 0);   // those are re-evaluated every round

	// ---- sample: 2 x 32 pixels spread over the mesh
	ZonePlan zp;
	zp.ok = false; zp.mhat = 0.0; zp.shat = 1.0; zp.pivot = 0.0; zp.A = 0.0; zp.B = 0.0;
	if (RETRY) {
		const TileStat pv = ws.tile_base[(size_t)b * P.ntiles + tile];   // the plan of the retry, parked by the first run
		zp.ok = pv.std > 0.0; zp.mhat = pv.med; zp.shat = pv.std; zp.pivot = Zn32::pivot_of(zp.mhat);
		zp.A = zp.mhat - ZN_CT * zp.shat; zp.B = zp.mhat + ZN_CT * zp.shat;
	} else if (do_stats) {
		uint32_t sk[2];
#pragma unroll
		for (int t = 0; t < 2; ++t) {
			const int idx = (lane * 131 + 17 + t * 2053) & (TBK_NPIX_TILE - 1);
			const size_t off = tile0 + (size_t)(idx >> 6) * P.W + (idx & 63);
			const float xs = STAGED ? reinterpret_cast<const ZbStage*>(zb_dyn)->px[w][idx] : __ldg(cube + off);
			const uint32_t k = __float_as_uint(xs + 0.0f);
			bool ok = k <= cut;
			if (HAS_EXTRA) ok = ok && !__ldg(extra + off);
			sk[t] = ok ? k : Zn32::padkey();
		}
		zp = zone_plan<Zn32>(sk[0], sk[1], lane);
	}
	TileStat st;
	TileStat* dst = ws.tile_base + (size_t)b * P.ntiles + tile;
	bool good = false;
	int why = ZN_WHY_SAMPLE;
	// bulk = [kA, kB] as keys; kM: keys below the sample centre.  Without a usable sample: empty bulk, every valid
	// pixel is a "tail" (the lists overflow harmlessly and the mesh goes to the bucketed path).
	uint32_t kA = 1u, kB = 0u, kM = 0u;
	if (zp.ok) {
		kA = Zn32::key_ceil(zp.A);
		if (!Zn32::key_floor(zp.B, kB)) zp.ok = false;
		kB = min(kB, cut);
		kM = Zn32::key_ceil(zp.mhat);
		if (kA > kB) zp.ok = false;
	}
	if (!zp.ok) { kA = 0xFFFFFFFFu; kB = 0xFFFFFFFFu; kM = 0u; }   // spanAB = 0 and kv - kA != 0 for every key: no bulk
	const uint32_t spanAB = kB - kA;
	// keys inside the sample centre -+ ZN_CW sigma (local density for the zone placement)
	uint32_t kC = 0xFFFFFFFFu, kC1 = 0u;
	if (zp.ok) { kC = Zn32::key_ceil(zp.mhat - ZN_CW * zp.shat); if (!Zn32::key_floor(zp.mhat + ZN_CW * zp.shat, kC1) || kC1 < kC) kC = 0xFFFFFFFFu; }
	const uint32_t spanC = kC == 0xFFFFFFFFu ? 0u : kC1 - kC;
	const double pivot = zp.pivot;
	const float pivot_f = (float)zp.pivot;   // exact: Zn32::pivot_of returns a float32 value
	const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(&sm.tails[lane]);
	const uint32_t zbase = (uint32_t)__cvta_generic_to_shared(&sm.zone[lane]);
	uint32_t tptr = tbase;
	const uint32_t tend = do_stats ? tbase + 128u * TCAP : tbase;

	// ---- pass 1: mask, flags, sub-block minima, counts, bulk moments, tails
	uint32_t nz = 0u, kmin = TW_INVALID;
	int nbad = 0, nA = 0, nM = 0, nC = 0;
	double s1 = 0.0, s2 = 0.0;
	float4 nx4[4]; uint32_t nex[4] = {0u, 0u, 0u, 0u};
	const float* pc = px0;                          // this lane's pixels of the current band (bumped by 8 rows per band)
	const uint8_t* pe = HAS_EXTRA ? extra + base : nullptr;
	uint8_t* pm = mask_out + base;
	const size_t band = 4 * step, pband = 4 * pstep;
#pragma unroll
	for (int h = 0; h < 4; ++h) {
		nx4[h] = zb_ld<STAGED>(pc + h * pstep);
		if (HAS_EXTRA) nex[h] = __ldg(reinterpret_cast<const unsigned int*>(pe + h * step));
	}
#pragma unroll 1
	for (int a = 0; a < 8; ++a) {   // rows 8a .. 8a+7 of the mesh = loads 4a .. 4a+3; the next band's loads are issued first
		float4 r[4]; uint32_t exr[4];
#pragma unroll
		for (int h = 0; h < 4; ++h) { r[h] = nx4[h]; exr[h] = nex[h]; }
		pc += pband; if (HAS_EXTRA) pe += band;
		if (a < 7) {
#pragma unroll
			for (int h = 0; h < 4; ++h) {
				nx4[h] = zb_ld<STAGED>(pc + h * pstep);
				if (HAS_EXTRA) nex[h] = __ldg(reinterpret_cast<const unsigned int*>(pe + h * step));
			}
		}
		uint32_t smin = TW_INVALID;
#pragma unroll
		for (int h = 0; h < 4; ++h) {
			// x + 0.0f maps -0.0 to +0.0: non-negative floats then order like unsigned integers, and negative / NaN / inf
			// bit patterns all compare above the cutoff
			const float x0 = r[h].x + 0.0f, x1 = r[h].y + 0.0f, x2 = r[h].z + 0.0f, x3 = r[h].w + 0.0f;
			nz |= __float_as_uint(x0) | __float_as_uint(x1) | __float_as_uint(x2) | __float_as_uint(x3);
			uint32_t m = 0u;
			zb_p1<0, HAS_EXTRA>(x0, exr[h] & 0xFFu, cut, kA, spanAB, kM, pivot, m, smin, nA, nM, s1, s2, tptr, tend, pivot_f, kC, spanC, nC);
			zb_p1<1, HAS_EXTRA>(x1, exr[h] & 0xFF00u, cut, kA, spanAB, kM, pivot, m, smin, nA, nM, s1, s2, tptr, tend, pivot_f, kC, spanC, nC);
			zb_p1<2, HAS_EXTRA>(x2, exr[h] & 0xFF0000u, cut, kA, spanAB, kM, pivot, m, smin, nA, nM, s1, s2, tptr, tend, pivot_f, kC, spanC, nC);
			zb_p1<3, HAS_EXTRA>(x3, exr[h] & 0xFF000000u, cut, kA, spanAB, kM, pivot, m, smin, nA, nM, s1, s2, tptr, tend, pivot_f, kC, spanC, nC);
			nbad += __popc(m);
			*reinterpret_cast<unsigned int*>(pm + h * step) = m;
		}
		pm += band;
		// sub-block minima: columns 8c .. 8c+7 are lanes {2c, 2c+1} + {0, 16}
		smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, 1));
		smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, 16));
		if ((lane & 17) == 0) sbdst[a * 8 + (lane >> 1)] = __uint_as_float(smin);
		kmin = min(kmin, smin);
	}
	kmin = __reduce_min_sync(0xffffffffu, kmin);
	const int n = 4096 - __reduce_add_sync(0xffffffffu, nbad);
	nz = __reduce_or_sync(0xffffffffu, nz);
	if (lane == 0 && !RETRY) {
		if (nz && !c.any_nonzero) atomicOr(&c.any_nonzero, 1);
		if (n > 0) { atomicAdd(&c.n_valid, n); atomicMin(&c.min_bits, kmin); }
	}
	st.mean = st.med = st.std = nan_d(); st.nfin = 0; st.pad = 0;
	if (!do_stats || n == 0) { if (lane == 0) *dst = st; return; }
	good = zp.ok && n >= ZN_MIN_N;
	why = ZN_WHY_SAMPLE;
	if (good) {
		const int tcnt = (int)((tptr - tbase) >> 7);
		const int nT = __reduce_add_sync(0xffffffffu, tcnt);
		nA = __reduce_add_sync(0xffffffffu, nA); nM = __reduce_add_sync(0xffffffffu, nM); nC = __reduce_add_sync(0xffffffffu, nC);
		const int nB = nT - nA;
		s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
		double ZL, ZH;
		zone_range(zp, n, nA, nB, nM, nC, ZL, ZH, true);
		uint32_t kZL = max(Zn32::key_ceil(ZL), kA), kZH = 0u;
		good = Zn32::key_floor(ZH, kZH);
		kZH = min(kZH, kB);
		good = good && kZL <= kZH;
		why = ZN_WHY_RANGE;
		if (good) {
			// ---- pass 2 (L2): zone elements into the per-lane lists; elements below the zone are counted
			const uint32_t zspan = kZH - kZL;
			uint32_t zptr = zbase;
			const uint32_t zend = zbase + 128u * ZN_ZCAP;
			int nZL = 0;
			pc = px0; if (HAS_EXTRA) pe = extra + base;
#pragma unroll
			for (int h = 0; h < 4; ++h) {
				nx4[h] = zb_ld<STAGED>(pc + h * pstep);
				if (HAS_EXTRA) nex[h] = __ldg(reinterpret_cast<const unsigned int*>(pe + h * step));
			}
#pragma unroll 1
			for (int a = 0; a < 8; ++a) {
				float4 r[4]; uint32_t exr[4];
#pragma unroll
				for (int h = 0; h < 4; ++h) { r[h] = nx4[h]; exr[h] = nex[h]; }
				pc += pband; if (HAS_EXTRA) pe += band;
				if (a < 7) {
#pragma unroll
					for (int h = 0; h < 4; ++h) {
						nx4[h] = zb_ld<STAGED>(pc + h * pstep);
						if (HAS_EXTRA) nex[h] = __ldg(reinterpret_cast<const unsigned int*>(pe + h * step));
					}
				}
#pragma unroll
				for (int h = 0; h < 4; ++h) {
					zb_p2<HAS_EXTRA>(r[h].x + 0.0f, exr[h] & 0xFFu, cut, kZL, zspan, nZL, zptr, zend);
					zb_p2<HAS_EXTRA>(r[h].y + 0.0f, exr[h] & 0xFF00u, cut, kZL, zspan, nZL, zptr, zend);
					zb_p2<HAS_EXTRA>(r[h].z + 0.0f, exr[h] & 0xFF0000u, cut, kZL, zspan, nZL, zptr, zend);
					zb_p2<HAS_EXTRA>(r[h].w + 0.0f, exr[h] & 0xFF000000u, cut, kZL, zspan, nZL, zptr, zend);
				}
			}
			nZL = __reduce_add_sync(0xffffffffu, nZL);
			const int zcnt = (int)((zptr - zbase) >> 7);
			const float zscale = (float)ZN_BINS / ((float)zspan + 1.0f);
			__syncwarp();
			good = zone_finish<Zn32>(sm, lane, n, nA, nB, nZL, tcnt, zcnt, s1, s2, pivot,
				(double)__uint_as_float(kA), (double)__uint_as_float(kB), kZL, zscale, st, why);
		}
	}
	if (lane == 0) {
		if (good) *dst = st;
		else {
			// Two failures get ONE more run of this kernel (RETRY) before the bucketed path; the plan of that run travels in *dst.
			//  * a clip bound entered the bulk (the 64-pixel sample overestimated the width: star wings, gradients): the
			//    statistics of the failing iteration are exact, the bulk is placed around them (median -+ 1.5 sigma of that
			//    iteration, half its clip range);
			//  * a tail list overflowed (crowded meshes: a quarter of the pixels above the bulk): the same plan again -- the retry
			//    launch has room for every pixel of the mesh in its tail lists.
			bool retry = false;
			if (!RETRY && zp.ok && ((why == ZN_WHY_BOUND && st.std > 0.0) || why == ZN_WHY_LIST)) {
				const int pos = atomicAdd(ws.fb_count + 8, 1);
				if (pos < rt_cap) {
					if (why == ZN_WHY_BOUND) st.std = 0.75 * st.std;
					else { st.med = zp.mhat; st.std = zp.shat; }
					*dst = st; ws.rt_list[pos] = b * P.ntiles + tile; retry = true;
				}
			}
			if (!retry) { ws.fb_list[atomicAdd(ws.fb_count, 1)] = b * P.ntiles + tile; atomicAdd(ws.fb_count + 16 + why, 1); }
		}
	}
}

// K_tile_base_fb: bucketed statistics (tile_block_stats_staged) for the meshes k_tile_base_z queued.
template <bool HAS_EXTRA>
__global__ void __launch_bounds__(64) k_tile_base_fb(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ extra)
{
	__shared__ TwBlockSmem<TwF32, 2> sm;
	__shared__ int s_n;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int count = *ws.fb_count;
	for (int e = blockIdx.x; e < count; e += gridDim.x) {
		const int id = ws.fb_list[e], b = id / P.ntiles, tile = id % P.ntiles;
		const int ty = tile / P.nx, tx = tile % P.nx;
		const FfiCtl& c = ws.ctl[b];
		const bool excl = (c.mars && tx * TBK_TILE >= 1536) || c.earth;
		const uint32_t cut = excl ? 0u : __float_as_uint(P.flux_cutoff);
		const int lrow0 = 2 * w + (lane >> 4), lcol = (lane & 15) << 2;
		const size_t base = (size_t)b * P.H * P.W + (size_t)(ty * TBK_TILE + lrow0) * P.W + tx * TBK_TILE + lcol;
		if (tid == 0) s_n = 0;
		__syncthreads();
		int n = 0;
		for (int i = 0; i < 16; ++i) {
			const size_t off = base + (size_t)i * 4 * P.W;
			const float4 r = __ldg(reinterpret_cast<const float4*>(cube + off));
			uint32_t ex = 0u;
			if (HAS_EXTRA) ex = __ldg(reinterpret_cast<const unsigned int*>(extra + off));
			const float x4[4] = {r.x, r.y, r.z, r.w};
			uint32_t kk[4];
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t k = __float_as_uint(x4[q] + 0.0f);
				bool ok = k <= cut;
				if (HAS_EXTRA) ok = ok && !((ex >> (8 * q)) & 0xFFu);
				kk[q] = ok ? k : TW_INVALID;
				n += ok;
			}
			*reinterpret_cast<uint4*>(&sm.tw.keys[(lrow0 + 4 * i) * TBK_TILE + lcol]) = make_uint4(kk[0], kk[1], kk[2], kk[3]);
		}
		n = __reduce_add_sync(0xffffffffu, n);
		if (lane == 0) atomicAdd(&s_n, n);
		__syncthreads();
		TileStat st; bool writer;
		tile_block_stats_staged<TwF32, 2>(sm, s_n, st, writer);
		if (writer) ws.tile_base[(size_t)b * P.ntiles + tile] = st;
		__syncthreads();
	}
}

// K_post_base: all-zero rule (pixel_flags.py:54-56), all-masked early-out (backgrounds.py:101-102),
// zeropoint of round 1 (backgrounds.py:171).
__global__ void k_post_base(PlanDev P, Workspace ws, tbk_ffi_status* status, int B)
{
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	FfiCtl& c = ws.ctl[b];
	if (P.is_tess && !c.any_nonzero) { c.all_masked = 1; c.n_valid = 0; }
	if (c.n_valid == 0) c.all_masked = 1;
	c.zp = 1.0 - (double)__uint_as_float(c.min_bits);
	if (status) {
		tbk_ffi_status& s = status[b];
		s.all_masked = c.all_masked;
		s.no_good_mesh = 0;
		s.n_valid = c.n_valid;
		s.rounds = 0;
		for (int i = 0; i < TBK_MAX_ROUNDS; ++i) {
			s.n_excluded[i] = 0; s.n_ring_valid[i] = 0; s.radial_ok[i] = 0; s.zeropoint[i] = 0.0;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K_zp_min (rounds >= 2): min over valid pixels of x - img_bkg_square (backgrounds.py:165-171).
__global__ void __launch_bounds__(TBK_NT) k_zp_min(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ mask)
{
	__shared__ ZoomTile z;
	__shared__ RedSmem red;
	const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
	FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	const int ty = tile / P.nx, tx = tile % P.nx;
	zoom_tile_load(z, ws.coef + (size_t)b * P.ntiles, ty, tx, P.ny, P.nx);
	zoom_tile_stage(z, c, P.zoom_w);
	__syncthreads();
	const size_t img = (size_t)b * P.H * P.W;
	const int lcol = tile_lcol(tid);
	double mn = INFINITY, mx = 0.0; int cnt = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int lrow = tile_lrow(tid, j);
		const size_t off = img + (size_t)(ty * TBK_TILE + lrow) * P.W + tx * TBK_TILE + lcol;
		const float4 x = __ldg(reinterpret_cast<const float4*>(cube + off));
		const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(mask + off));
		double sq[4];
		zoom_eval4(z, z.w, lrow, lcol, sq);
		if (!m.x) mn = fmin(mn, (double)x.x - zoom_clip_s(z, sq[0]));
		if (!m.y) mn = fmin(mn, (double)x.y - zoom_clip_s(z, sq[1]));
		if (!m.z) mn = fmin(mn, (double)x.z - zoom_clip_s(z, sq[2]));
		if (!m.w) mn = fmin(mn, (double)x.w - zoom_clip_s(z, sq[3]));
	}
	block_sum_min_max(red, cnt, mn, mx);
	if (tid == 0 && mn < INFINITY) atomicMin(&c.min_key, dkey(mn));
}

// K_zp_bound / K_zp_exact: the same minimum without touching every pixel.  k_tile_base_warp left the
// minimum valid pixel of every 8x8 sub-block in ws.sbmin.  Inside a mesh the interpolated background is
// Lipschitz: |sq(p) - sq(p0)| <= (Gx |dx| + Gy |dy|) / 64 with Gx, Gy the largest differences of adjacent
// spline coefficients in the 5x5 neighbourhood (the derivative of a cubic B-spline surface is a convex
// combination of coefficient differences; the clip to [mesh_min, mesh_max] is 1-Lipschitz).  So with
// sq0 = sq(sub-block pixel (4,4)) and slack = 4 (Gx + Gy) / 64:
//     xmin - (sq0 + slack)  <=  min over the sub-block of (x - sq)  <=  xmin - (sq0 - slack).
// Pass 1 takes the smallest upper bound over the FFI; pass 2 evaluates exactly only the sub-blocks whose
// lower bound does not exceed it -- the true minimiser is always among them.
// One warp per mesh (8 meshes per CTA); lane l owns sub-blocks l and l + 32.
#define ZP_WARPS 8
struct ZpWarpSmem {
	double c[5][5];
	double wc[8][4];   // zoom weights of the 8 sub-block sample phases (pixel 8a + 4)
};

// sq at the sample pixel of sub-block sbk and the Lipschitz slack of the mesh (warp-uniform)
__device__ __forceinline__ void zp_mesh_setup(ZpWarpSmem& sm, const PlanDev& P, const double* __restrict__ coef,
	const FfiCtl& c, int tile, int lane, double& slack)
{
	const int ty = tile / P.nx, tx = tile % P.nx;
	if (lane < 25) sm.c[lane / 5][lane % 5] = coef[reflect_fold(ty - 2 + lane / 5, P.ny) * P.nx + reflect_fold(tx - 2 + lane % 5, P.nx)];
	sm.wc[lane >> 2][lane & 3] = __ldg(P.zoom_w + 4 * (8 * (lane >> 2) + 4) + (lane & 3));
	__syncwarp();
	double g = 0.0;
	if (lane < 20) g = fabs(sm.c[lane / 4][lane % 4 + 1] - sm.c[lane / 4][lane % 4]);          // x differences
	double gx = g;
	for (int o = 16; o > 0; o >>= 1) gx = fmax(gx, __shfl_xor_sync(0xffffffffu, gx, o));
	g = 0.0;
	if (lane < 20) g = fabs(sm.c[lane % 4 + 1][lane / 4] - sm.c[lane % 4][lane / 4]);          // y differences
	double gy = g;
	for (int o = 16; o > 0; o >>= 1) gy = fmax(gy, __shfl_xor_sync(0xffffffffu, gy, o));
	slack = c.mesh_const ? 0.0 : (4.0 * (gx + gy) / 64.0) * (1.0 + 1e-9) + 1e-9;
}

__device__ __forceinline__ double zp_sample_sq(const ZpWarpSmem& sm, const FfiCtl& c, int sbk)
{
	const int a = sbk >> 3, bcol = sbk & 7;   // sample pixel (8a + 4, 8 bcol + 4): first half of the mesh iff a < 4
	const int oy = a >> 2, ox = bcol >> 2;
	double acc = 0.0;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		double ra = 0.0;
#pragma unroll
		for (int j = 0; j < 4; ++j) ra += sm.wc[bcol][j] * sm.c[oy + i][ox + j];
		acc += sm.wc[a][i] * ra;
	}
	if (c.mesh_const) return c.mesh_min;
	return fmin(fmax(acc, c.mesh_min), c.mesh_max);
}

// pass 1: per sub-block bounds; the lower bounds are kept (rounded down to float32) for pass 2
__global__ void __launch_bounds__(32 * ZP_WARPS) k_zp_bound(PlanDev P, Workspace ws)
{
	__shared__ ZpWarpSmem smw[ZP_WARPS];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int tile = blockIdx.x * ZP_WARPS + w, b = blockIdx.y;
	FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh || tile >= P.ntiles) return;
	double slack;
	zp_mesh_setup(smw[w], P, ws.coef + (size_t)b * P.ntiles, c, tile, lane, slack);
	float* sb = ws.sbmin + ((size_t)b * P.ntiles + tile) * 64;
	float* lb = ws.sblow + ((size_t)b * P.ntiles + tile) * 64;
	double ub = INFINITY;
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const int sbk = lane + 32 * h;
		const float xmin = sb[sbk];
		float low = INFINITY;
		if (xmin < INFINITY) {
			const double sq0 = zp_sample_sq(smw[w], c, sbk);
			ub = fmin(ub, (double)xmin - (sq0 - slack));
			low = __double2float_rd((double)xmin - (sq0 + slack));
		}
		lb[sbk] = low;
	}
	for (int o = 16; o > 0; o >>= 1) ub = fmin(ub, __shfl_xor_sync(0xffffffffu, ub, o));
	if (lane == 0 && ub < INFINITY) atomicMin(&c.min_ub, dkey(ub));
}

// pass 2: exact evaluation of the candidate sub-blocks (lower bound <= smallest upper bound)
__global__ void __launch_bounds__(32 * ZP_WARPS) k_zp_exact(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ mask)
{
	__shared__ ZpWarpSmem smw[ZP_WARPS];
	__shared__ double wT[4][64];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int tile = blockIdx.x * ZP_WARPS + w, b = blockIdx.y;
	FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	const double U = dkey_inv(c.min_ub);
	unsigned cand[2] = {0u, 0u};
	if (tile < P.ntiles) {
		const float* lb = ws.sblow + ((size_t)b * P.ntiles + tile) * 64;
		cand[0] = __ballot_sync(0xffffffffu, (double)lb[lane] <= U);
		cand[1] = __ballot_sync(0xffffffffu, (double)lb[lane + 32] <= U);
	}
	if (!__syncthreads_or(cand[0] | cand[1])) return;
	wT[threadIdx.x & 3][threadIdx.x >> 2] = __ldg(P.zoom_w + threadIdx.x);
	__syncthreads();
	if (!(cand[0] | cand[1])) return;
	double slack;
	zp_mesh_setup(smw[w], P, ws.coef + (size_t)b * P.ntiles, c, tile, lane, slack);
	const int ty = tile / P.nx, tx = tile % P.nx;
	const size_t img = (size_t)b * P.H * P.W;
	double mn = INFINITY;
	for (int h = 0; h < 2; ++h) {
		unsigned m = cand[h];
		while (m) {
			const int sbk = 32 * h + __ffs(m) - 1;
			m &= m - 1;
			const int r0 = 8 * (sbk >> 3), c0 = 8 * (sbk & 7);
			// 64 pixels of the sub-block, two per lane
#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const int lr = r0 + ((lane + 32 * e) >> 3), lc = c0 + ((lane + 32 * e) & 7);
				const size_t off = img + (size_t)(ty * TBK_TILE + lr) * P.W + tx * TBK_TILE + lc;
				if (!__ldg(mask + off)) {
					const int oy = lr >> 5, ox = lc >> 5;
					double acc = 0.0;
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						double ra = 0.0;
#pragma unroll
						for (int j = 0; j < 4; ++j) ra += wT[j][lc] * smw[w].c[oy + i][ox + j];
						acc += wT[i][lr] * ra;
					}
					const double sq = c.mesh_const ? c.mesh_min : fmin(fmax(acc, c.mesh_min), c.mesh_max);
					mn = fmin(mn, (double)__ldg(cube + off) - sq);
				}
			}
		}
	}
	for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
	if (lane == 0 && mn < INFINITY) atomicMin(&c.min_key, dkey(mn));
}

__global__ void k_set_zp(Workspace ws, int B)
{
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	c.zp = 1.0 - dkey_inv(c.min_key);  // zeropoint = -min(pix) + 1.0
	c.min_key = ~0ULL;
	c.min_ub = ~0ULL;
}

// ---------------------------------------------------------------------------------------------
// K_ring_gather: log-flux samples of every ring pixel (backgrounds.py:165-172).
// round 0: float32 arithmetic with a correctly rounded float32 log10; later rounds float64.
__global__ void k_ring_gather(PlanDev P, Workspace ws, const float* __restrict__ cube,
	const uint8_t* __restrict__ mask, int round)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int b = blockIdx.y;
	if (i >= P.nringpix) return;
	const FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	const unsigned yx = (unsigned)__ldg(P.ring_pix + i);
	const int py = (int)(yx >> 16), px = (int)(yx & 0xFFFFu);
	const size_t off = (size_t)b * P.H * P.W + (size_t)py * P.W + px;
	double val = nan_d();
	if (!__ldg(mask + off)) {
		const float x = __ldg(cube + off);
		if (round == 0) {
			const float s = (x + 0.0f) + (float)c.zp;
			val = (double)(float)log10((double)s);
		} else {
			const double sq = zoom_clip(c, zoom_eval_global(ws.coef + (size_t)b * P.ntiles, P.zoom_w, py, px, P.ny, P.nx));
			val = log10(((double)x - sq) + c.zp);
		}
	}
	ws.ring_v[(size_t)b * P.nringpix + i] = val;
}

// K_ring_gather_t: the same samples, pixels grouped by mesh: one CTA stages the mesh's 5x5 spline
// coefficients, reduces them with the row weights once (R[row][B], see k_final) and evaluates its ring pixels
// from shared memory with 4 DFMA each; log10 is the table-driven tbk_log10.
__global__ void __launch_bounds__(256) k_ring_gather_t(PlanDev P, Workspace ws, const float* __restrict__ cube,
	const uint8_t* __restrict__ mask, int round)
{
	__shared__ double sc[5][6];
	__shared__ double wT[4][64];
	__shared__ __align__(16) double R[2][64][4];
	__shared__ __align__(16) double ltab[128][4];
	const int slot = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
	const FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	const int tile = P.ringtile_id[slot];
	const int ty = tile / P.nx, tx = tile % P.nx;
	reinterpret_cast<double2*>(&ltab[0][0])[tid] = reinterpret_cast<const double2*>(&tbk_log10_tab[0][0])[tid];
	if (round > 0) {
		const double* coef = ws.coef + (size_t)b * P.ntiles;
		if (tid < 25) sc[tid / 5][tid % 5] = coef[reflect_fold(ty - 2 + tid / 5, P.ny) * P.nx + reflect_fold(tx - 2 + tid % 5, P.nx)];
		wT[tid & 3][tid >> 2] = __ldg(P.zoom_w + tid);
		__syncthreads();
		for (int e = tid; e < 64 * 5; e += 256) {
			const int row = e & 63, B = e >> 6, oy = row >> 5;
			double r = 0.0;
#pragma unroll
			for (int a = 0; a < 4; ++a) r = fma(wT[a][row], sc[oy + a][B], r);
			if (B < 4) R[0][row][B] = r;
			if (B > 0) R[1][row][B - 1] = r;
		}
	}
	__syncthreads();
	const double zp = c.zp, mmin = c.mesh_min, mmax = c.mesh_max;
	const float zp32 = (float)c.zp;
	const bool mconst = c.mesh_const != 0;
	const size_t img = (size_t)b * P.H * P.W + (size_t)(ty * TBK_TILE) * P.W + tx * TBK_TILE;
	double* __restrict__ dst = ws.ring_v + (size_t)b * P.nringpix;
	const int lo = P.ringtile_ptr[slot], hi = P.ringtile_ptr[slot + 1];
	// four entries per thread and step: the entry, mask and pixel loads of a step are issued back to back (the
	// kernel is bound by the latency of these dependent, scattered loads, not by arithmetic)
	for (int i0 = lo + tid; i0 < hi; i0 += 4 * 256) {
		unsigned ent[4]; uint8_t m[4]; float x[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) ent[u] = (i0 + 256 * u < hi) ? __ldg(P.ringtile_ent + i0 + 256 * u) : 0u;
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const size_t off = img + (size_t)((ent[u] >> 6) & 63) * P.W + (ent[u] & 63);
			m[u] = __ldg(mask + off); x[u] = __ldg(cube + off);
		}
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			if (i0 + 256 * u >= hi) break;
			const int lr = (ent[u] >> 6) & 63, lc = ent[u] & 63;
			double val = nan_d();
			if (!m[u]) {
				if (round == 0) {
					const float s = (x[u] + 0.0f) + zp32;
					val = (double)(float)tbk_log10((double)s, ltab);
				} else {
					const double2 ra = *reinterpret_cast<const double2*>(&R[lc >> 5][lr][0]);
					const double2 rb = *reinterpret_cast<const double2*>(&R[lc >> 5][lr][2]);
					const double acc = wT[0][lc] * ra.x + wT[1][lc] * ra.y + wT[2][lc] * rb.x + wT[3][lc] * rb.y;
					const double sq = mconst ? mmin : clamp_d(acc, mmin, mmax);
					val = tbk_log10(((double)x[u] - sq) + zp, ltab);
				}
			}
			dst[ent[u] >> 12] = val;
		}
	}
}

// K_ring_gather_d: the same samples from the dense per-mesh map (PlanDev::ringtile_idx).  Almost every pixel of a mesh
// that touches the rings is a ring pixel, so the mesh is walked densely -- a thread owns one column and 16 rows, a warp
// reads 32 consecutive indices / mask bytes / pixels per row, the four column weights stay in registers and the row
// part of the zoom is one shared-memory broadcast per warp and row -- instead of through the entry list, whose loads
// scatter over the mesh.
__global__ void __launch_bounds__(256, 4) k_ring_gather_d(PlanDev P, Workspace ws, const float* __restrict__ cube,
	const uint8_t* __restrict__ mask, int round)
{
	__shared__ double sc[5][6];
	__shared__ double wT[4][64];
	__shared__ __align__(16) double R[2][64][4];
	__shared__ __align__(16) double ltab[128][4];
	const int slot = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
	const FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	const int tile = P.ringtile_id[slot];
	const int ty = tile / P.nx, tx = tile % P.nx;
	const int lc = tid & 63, row0 = (tid >> 6) * 16, ox = lc >> 5;
	const size_t img = (size_t)b * P.H * P.W + (size_t)(ty * TBK_TILE) * P.W + tx * TBK_TILE + lc;
	const int* __restrict__ idxmap = P.ringtile_idx + (size_t)slot * TBK_NPIX_TILE + lc;
	// the loads of the first four rows are issued before the staging of the zoom rows
	int id[4]; uint8_t mk[4]; float px[4];
	auto load_rows = [&](int r0) {
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int lrow = r0 + j;
			id[j] = __ldg(idxmap + lrow * TBK_TILE);
			const size_t off = img + (size_t)lrow * P.W;
			mk[j] = __ldg(mask + off);
			px[j] = __ldg(cube + off);
		}
	};
	load_rows(row0);
	reinterpret_cast<double2*>(&ltab[0][0])[tid] = reinterpret_cast<const double2*>(&tbk_log10_tab[0][0])[tid];
	if (round > 0) {
		const double* coef = ws.coef + (size_t)b * P.ntiles;
		if (tid < 25) sc[tid / 5][tid % 5] = coef[reflect_fold(ty - 2 + tid / 5, P.ny) * P.nx + reflect_fold(tx - 2 + tid % 5, P.nx)];
		wT[tid & 3][tid >> 2] = __ldg(P.zoom_w + tid);
		__syncthreads();
		for (int e = tid; e < 64 * 5; e += 256) {
			const int row = e & 63, B = e >> 6, oy = row >> 5;
			double r = 0.0;
#pragma unroll
			for (int a = 0; a < 4; ++a) r = fma(wT[a][row], sc[oy + a][B], r);
			if (B < 4) R[0][row][B] = r;
			if (B > 0) R[1][row][B - 1] = r;
		}
	}
	__syncthreads();
	const double zp = c.zp, mmin = c.mesh_min, mmax = c.mesh_max;
	const float zp32 = (float)c.zp;
	const bool mconst = c.mesh_const != 0;
	double* __restrict__ dst = ws.ring_v + (size_t)b * P.nringpix;
	double wx[4] = {0.0, 0.0, 0.0, 0.0};
	if (round > 0) {
#pragma unroll
		for (int a = 0; a < 4; ++a) wx[a] = wT[a][lc];
	}
#pragma unroll 1
	for (int g = 0; g < 4; ++g) {
		int cid[4]; uint8_t cmk[4]; float cpx[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) { cid[j] = id[j]; cmk[j] = mk[j]; cpx[j] = px[j]; }
		if (g < 3) load_rows(row0 + 4 * (g + 1));
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			if (cid[j] < 0) continue;
			double val = nan_d();
			if (!cmk[j]) {
				if (round == 0) {
					const float s = (cpx[j] + 0.0f) + zp32;
					val = (double)(float)tbk_log10((double)s, ltab);
				} else {
					const int lrow = row0 + 4 * g + j;
					const double2 ra = *reinterpret_cast<const double2*>(&R[ox][lrow][0]);
					const double2 rb = *reinterpret_cast<const double2*>(&R[ox][lrow][2]);
					const double acc = wx[0] * ra.x + wx[1] * ra.y + wx[2] * rb.x + wx[3] * rb.y;
					const double sq = mconst ? mmin : clamp_d(acc, mmin, mmax);
					val = tbk_log10(((double)cpx[j] - sq) + zp, ltab);
				}
			}
			dst[cid[j]] = val;
		}
	}
}

__global__ void k_debug_log10(const double* __restrict__ in, double* __restrict__ out, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = tbk_log10(in[i], tbk_log10_tab);
}

int tbk_launch_log10(const double* in, double* out, int n, cudaStream_t st)
{
	k_debug_log10<<<(n + 255) / 256, 256, 0, st>>>(in, out, n);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("k_debug_log10: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}

// ---------------------------------------------------------------------------------------------
// K_ring_kde: mode of a Gaussian FFT-KDE per ring (backgrounds.py:21-33 + statsmodels 0.13.2
// KDEUnivariate.fit(gridsize=2000); see oracle/backgrounds_oracle.py:kde_density).
#ifndef TBK_KDE_NT
#define TBK_KDE_NT 256   // 4 CTAs per SM (64 registers, 49 KB shared memory each)
#endif
#define TBK_KDE_CAND 256
struct KdeSmem {
	double2 x[TBK_KDE_M];
	SelectSmem sel;
	RedSmem red;
	double bestv[32];
	int besti[32];
	double qcand[4][TBK_KDE_CAND];   // members of the bins that hold the quartile ranks
	double qres[4];
	int qbin[4], qexcl[4], qcnt[4], qn[4];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// exp(-2 pi i q / 2048) (conjugated for the inverse transform) from the 1024-entry table
template <bool INV>
__device__ __forceinline__ double2 kde_tw(const double2* __restrict__ tw, int q)
{
	q &= TBK_KDE_M - 1;
	double2 w = __ldg(tw + (q & (TBK_KDE_M / 2 - 1)));
	if (q >= TBK_KDE_M / 2) { w.x = -w.x; w.y = -w.y; }
	if (INV) w.y = -w.y;
	return w;
}

// 1024-point complex FFT, Stockham autosort radix-4: 5 stages, 256 butterflies each, natural order in and out.
// Input in a[], result in b[] (a is used as scratch).  Unnormalised; call with all threads of the CTA.
template <bool INV>
__device__ __forceinline__ void kde_fft1024(double2* a, double2* b, const double2* __restrict__ tw)
{
	double2* x = a; double2* y = b;
#pragma unroll
	for (int p = 1; p < 1024; p <<= 2) {
		for (int i = threadIdx.x; i < 256; i += blockDim.x) {
			const int k = i & (p - 1), j = ((i - k) << 2) + k, q = k * (512 / p);
			const double2 u0 = x[i];
			const double2 u1 = cmul(x[i + 256], kde_tw<INV>(tw, q));
			const double2 u2 = cmul(x[i + 512], kde_tw<INV>(tw, 2 * q));
			const double2 u3 = cmul(x[i + 768], kde_tw<INV>(tw, 3 * q));
			const double2 v0 = make_double2(u0.x + u2.x, u0.y + u2.y), v1 = make_double2(u0.x - u2.x, u0.y - u2.y);
			const double2 v2 = make_double2(u1.x + u3.x, u1.y + u3.y), d = make_double2(u1.x - u3.x, u1.y - u3.y);
			const double2 v3 = INV ? make_double2(-d.y, d.x) : make_double2(d.y, -d.x);   // (+i) d or (-i) d
			y[j] = make_double2(v0.x + v2.x, v0.y + v2.y);
			y[j + p] = make_double2(v1.x + v3.x, v1.y + v3.y);
			y[j + 2 * p] = make_double2(v0.x - v2.x, v0.y - v2.y);
			y[j + 3 * p] = make_double2(v1.x - v3.x, v1.y - v3.y);
		}
		__syncthreads();
		double2* t = x; x = y; y = t;
	}
}

__global__ void __launch_bounds__(TBK_KDE_NT, 1024 / TBK_KDE_NT) k_ring_kde(PlanDev P, Workspace ws)
{
	extern __shared__ __align__(16) unsigned char smraw[];
	KdeSmem& sm = *reinterpret_cast<KdeSmem*>(smraw);
	// grid (B, nrings): the block index runs over the FFIs first, so the launch works through the rings from the longest to
	// the shortest (ring_order) and ends on short CTAs
	const int ring = P.ring_order[blockIdx.y], b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
	const FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	const int lo = P.ring_ptr[ring], hi = P.ring_ptr[ring + 1];
	const double* __restrict__ v = ws.ring_v + (size_t)b * P.nringpix;
	double* out = ws.s2_raw + (size_t)b * P.nrings + ring;

	// sweep over the ring's samples: four independent loads in flight per thread (the array lives in L2)
	auto each = [&](auto f) {
		int i = lo + tid;
		for (; i + 3 * nt < hi; i += 4 * nt) {
			const double d0 = v[i], d1 = v[i + nt], d2 = v[i + 2 * nt], d3 = v[i + 3 * nt];
			if (d0 == d0) f(d0);
			if (d1 == d1) f(d1);
			if (d2 == d2) f(d2);
			if (d3 == d3) f(d3);
		}
		for (; i < hi; i += nt) { const double d = v[i]; if (d == d) f(d); }
	};

	// n, min, max, mean and std(ddof=1) in one sweep: moments about a pivot sample (the first finite one among the
	// leading entries), so the one-pass variance loses no more than a few ulps to cancellation
	double pv = 0.0;
	for (int t = 0; t < 4 && lo + 32 * t < hi; ++t) {
		const int i = lo + 32 * t + (tid & 31);
		const double d = i < hi ? v[i] : nan_d();
		const unsigned fin = __ballot_sync(0xffffffffu, d == d);
		if (fin) { pv = __shfl_sync(0xffffffffu, d, __ffs(fin) - 1); break; }
	}
	// Window of the quartile histogram from a 32-element sample spread over the ring (median -+ 2.5 sample sigma: the
	// quartiles sit at -+ 0.67 sigma).  The window only decides which bins the order statistics are looked up in -- the
	// statistics themselves are exact -- so a poor sample costs a fallback, never accuracy.
	const int len = hi - lo;
	double w0 = 0.0, w1 = 0.0;
	{
		const int i = lo + (int)(((long long)len * (2 * (tid & 31) + 1)) >> 6);
		const double d = len > 0 ? v[i] : nan_d();
		const unsigned long long k = warp_bitonic32<unsigned long long>(d == d ? dkey(d) : ~0ULL, tid & 31);
		const int m = __popc(__ballot_sync(0xffffffffu, k != ~0ULL));
		if (m >= 16) {
			const double med = dkey_inv(__shfl_sync(0xffffffffu, k, m >> 1));
			const double sg = (dkey_inv(__shfl_sync(0xffffffffu, k, (3 * m) >> 2)) - dkey_inv(__shfl_sync(0xffffffffu, k, m >> 2))) / 1.349;
			w0 = med - 2.5 * sg; w1 = med + 2.5 * sg;
		}
	}
	const double hscale = (double)(TBK_NBINS - 2) / (w1 - w0);
	const bool fast = (w1 > w0) && (hscale < 1e300);
	const bool staged = len <= 8 * TBK_KDE_M;   // 16-bit bin per sample, overlaid on the (still unused) FFT buffer
	uint16_t* sbin = reinterpret_cast<uint16_t*>(sm.x);
	auto binof = [&](double d) {
		const double t = (d - w0) * hscale;
		return d < w0 ? 0 : (d > w1 ? TBK_NBINS - 1 : 1 + min(TBK_NBINS - 3, (int)t));
	};
	for (int i = tid; i < TBK_NBINS; i += nt) sm.sel.hist[i] = 0u;
	if (tid < 4) sm.qn[tid] = 0;
	__syncthreads();
	// ONE sweep: n, min, max, moments about the pivot, and the quartile histogram (bins kept per sample when they fit)
	int n = 0; double mn = INFINITY, mx = -INFINITY, s1 = 0.0, ss = 0.0;
	{
		auto put = [&](int i, double d) {
			int bin = 0xFFFF;
			if (d == d) {
				++n; mn = fmin(mn, d); mx = fmax(mx, d); const double e = d - pv; s1 += e; ss = fma(e, e, ss);
				if (fast) { bin = binof(d); atomicAdd(&sm.sel.hist[bin], 1u); }
			}
			if (staged) sbin[i] = (uint16_t)bin;
		};
		int i = tid;
		for (; i + 3 * nt < len; i += 4 * nt) {
			const double d0 = v[lo + i], d1 = v[lo + i + nt], d2 = v[lo + i + 2 * nt], d3 = v[lo + i + 3 * nt];
			put(i, d0); put(i + nt, d1); put(i + 2 * nt, d2); put(i + 3 * nt, d3);
		}
		for (; i < len; i += nt) put(i, v[lo + i]);
	}
	int nd = 0;
	block_sum_min_max(sm.red, n, mn, mx);
	block_sum3(sm.red, nd, s1, ss);
	if (n <= 1) { if (tid == 0) *out = nan_d(); return; }  // reduce_mode([]) = NaN; one sample -> NaN
	const double mean = pv + s1 / (double)n;
	const double sd = sqrt(fmax(ss - s1 * s1 / (double)n, 0.0) / (double)(n - 1));
	(void)mean;

	// scipy.stats.scoreatpercentile(x, 25 / 75): linear interpolation at (n-1)*p.  The (up to) four order
	// statistics are resolved together from the histogram of the sweep above and one collecting sweep; a rank whose
	// bin is an overflow bin or too full falls back to the generic iterated selection.
	double q[2];
	{
		int rk[4]; double fr[2];
		for (int t = 0; t < 2; ++t) {
			const double idx = (t == 0 ? 0.25 : 0.75) * (double)(n - 1);
			const int i0 = (int)idx;
			fr[t] = idx - (double)i0;
			rk[2 * t] = i0; rk[2 * t + 1] = (fr[t] > 0.0) ? i0 + 1 : i0;
		}
		double ord[4];
		bool done[4] = {false, false, false, false};
		if (fast) {
			// exclusive scan (thread t owns bins 2t, 2t+1 for 512 threads)
			const int per = TBK_NBINS / TBK_KDE_NT;
			unsigned loc[TBK_NBINS / TBK_KDE_NT], tsum = 0;
			for (int j = 0; j < per; ++j) { loc[j] = sm.sel.hist[tid * per + j]; tsum += loc[j]; }
			unsigned inc = tsum;
			for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += u; }
			if ((tid & 31) == 31) sm.sel.wsum[tid >> 5] = inc;
			__syncthreads();
			unsigned base = 0;
			for (int w = 0; w < (tid >> 5); ++w) base += sm.sel.wsum[w];
			unsigned excl = base + inc - tsum;
			for (int j = 0; j < per; ++j) {
				for (int r = 0; r < 4; ++r)
					if ((unsigned)rk[r] >= excl && (unsigned)rk[r] < excl + loc[j]) { sm.qbin[r] = tid * per + j; sm.qexcl[r] = (int)excl; sm.qcnt[r] = (int)loc[j]; }
				excl += loc[j];
			}
			__syncthreads();
			int qb[4], qe[4], qc[4];
			for (int r = 0; r < 4; ++r) { qb[r] = sm.qbin[r]; qe[r] = sm.qexcl[r]; qc[r] = sm.qcnt[r]; }
			// collect the members of the target bins (ranks in the same bin share list r of the first of them)
			int lst[4];
			for (int r = 0; r < 4; ++r) { lst[r] = r; for (int p = 0; p < r; ++p) if (qb[p] == qb[r]) { lst[r] = lst[p]; break; } }
			int tb[4];   // bins to collect (-1: nothing to do for this rank)
			for (int r = 0; r < 4; ++r)
				tb[r] = (lst[r] == r && qc[r] <= TBK_KDE_CAND && qb[r] != 0 && qb[r] != TBK_NBINS - 1) ? qb[r] : -1;
			auto collect = [&](int bin, double d) {
				for (int r = 0; r < 4; ++r)
					if (bin == tb[r]) { const int p = atomicAdd(&sm.qn[r], 1); sm.qcand[r][p] = d; }
			};
			if (staged) {
				for (int i = tid; i < len; i += nt) {
					const int bin = sbin[i];
					if (bin == tb[0] || bin == tb[1] || bin == tb[2] || bin == tb[3]) collect(bin, v[lo + i]);
				}
			} else {
				each([&](double d) { const int bin = binof(d); if (bin == tb[0] || bin == tb[1] || bin == tb[2] || bin == tb[3]) collect(bin, d); });
			}
			__syncthreads();
			for (int r = 0; r < 4; ++r) {
				const int L = lst[r];
				if (qc[L] > TBK_KDE_CAND || qb[r] == 0 || qb[r] == TBK_NBINS - 1) continue;
				const int cntL = qc[L], kk = rk[r] - qe[r];
				for (int j = tid; j < cntL; j += nt) {
					const double cj = sm.qcand[L][j];
					int rr = 0;
					for (int i = 0; i < cntL; ++i) { const double ci = sm.qcand[L][i]; rr += (ci < cj) || (ci == cj && i < j); }
					if (rr == kk) sm.qres[r] = cj;
				}
				done[r] = true;
			}
			__syncthreads();
			for (int r = 0; r < 4; ++r) if (done[r]) ord[r] = sm.qres[r];
		}
		for (int r = 0; r < 4; ++r) {
			if (done[r]) continue;
			if (r > 0 && rk[r] == rk[r - 1]) { ord[r] = ord[r - 1]; continue; }
			if (tid == 0) atomicAdd(const_cast<int*>(&c.kde_fallbacks), 1);
			ord[r] = block_select(sm.sel, sm.red, each, rk[r], mn, mx);
		}
		for (int t = 0; t < 2; ++t) {
			if (fr[t] > 0.0) {
				const double w0q = (double)(rk[2 * t] + 1) - ((t == 0 ? 0.25 : 0.75) * (double)(n - 1)), w1q = fr[t];
				q[t] = __dadd_rn(__dmul_rn(ord[2 * t], w0q), __dmul_rn(ord[2 * t + 1], w1q)) / (w0q + w1q);
			} else q[t] = ord[2 * t];
		}
	}
	const double iqr = (q[1] - q[0]) / 1.349;
	const double sigma = (iqr > 0.0) ? fmin(sd, iqr) : sd;
	const double bw = 1.0592238410488122 * sigma * pow((double)n, -0.2);
	if (bw == 0.0) {
		// "Selected KDE bandwidth is 0" -> np.median(x)  (backgrounds.py:28-32)
		const double m1 = block_select(sm.sel, sm.red, each, (n - 1) >> 1, mn, mx);
		double m2 = m1;
		if ((n & 1) == 0) m2 = block_select(sm.sel, sm.red, each, n >> 1, mn, mx);
		if (tid == 0) *out = 0.5 * (m1 + m2);
		return;
	}

	const double a = mn - 3.0 * bw;
	const double bb = mx + 3.0 * bw;
	const double delta = (bb - a) / (double)(TBK_KDE_M - 1);
	const double range = bb - a;

	// linear binning (statsmodels linbin.fast_linbin): g[li] += 1 - rem, g[li+1] += rem, i.e. g[c] = count[c] - S[c] + S[c-1]
	// with S[c] = sum of rem over the samples of cell c.  Shared memory has no native 64-bit add, so S is kept in fixed point,
	// which also makes the result independent of the arrival order:
	//   packed (default): two 32-bit atomics per sample -- word A = count (8 bit) | sum of the top 16 bits of rem (24 bit),
	//     word B = sum of the next 24 bits of rem; exact to 2^-40 per sample as long as no cell holds more than 255
	//     samples, which is verified afterwards (the count fields must add up to the number of binned samples);
	//   long rings (n > 8192): three atomics -- count, two 20-bit halves of the fraction -- good up to 4095 samples per cell;
	//   fixed: four atomics per sample (count + 48-bit fraction in three 16-bit chunks), exact for n < 65536;
	//   otherwise float64 CAS adds.
	uint32_t* lb = reinterpret_cast<uint32_t*>(sm.x);   // [4][TBK_KDE_M] overlay on the FFT buffer
	// (d - a) / delta as reciprocal multiply + one fma correction (= the correctly rounded quotient); a quotient
	// that lands within 1e-9 of an integer is recomputed with the true division so the cell index cannot differ
	const double rdelta = 1.0 / delta;
	auto cell_of = [&](double d, int& li, double& rem) {
		const double xa = d - a;
		double lxi = xa * rdelta;
		lxi = fma(fma(-lxi, delta, xa), rdelta, lxi);
		li = (int)lxi;
		rem = lxi - (double)li;
		if (rem < 1e-9 || rem > 1.0 - 1e-9) { lxi = xa / delta; li = (int)lxi; rem = lxi - (double)li; }
		return li > 1 && li < TBK_KDE_M - 1;
	};
	bool packed = false;
	if (n <= 8192) {
		for (int i = tid; i < 2 * TBK_KDE_M; i += nt) lb[i] = 0u;
		__syncthreads();
		int nb = 0;
		each([&](double d) {
			int li; double rem;
			if (cell_of(d, li, rem)) {
				++nb;
				const double t = rem * 65536.0;
				const unsigned h16 = (unsigned)t;
				const unsigned l24 = (unsigned)((t - (double)h16) * 16777216.0);
				atomicAdd(&lb[li], (1u << 24) | h16);
				atomicAdd(&lb[TBK_KDE_M + li], l24);
			}
		});
		__syncthreads();
		int csum = 0;
		for (int j = tid; j < TBK_KDE_M; j += nt) csum += (int)(lb[j] >> 24);
		double z0 = 0.0, z1 = 0.0;
		block_sum3(sm.red, csum, z0, z1);
		int nbt = nb;
		block_sum3(sm.red, nbt, z0, z1);
		packed = csum == nbt;
		if (packed) {
			double g[TBK_KDE_M / TBK_KDE_NT];
			for (int j = 0; j < TBK_KDE_M / TBK_KDE_NT; ++j) {
				const int cidx = tid + j * nt;
				const uint32_t A = lb[cidx];
				const double S = (double)(A & 0xFFFFFFu) * 1.52587890625e-05 + (double)lb[TBK_KDE_M + cidx] * 9.094947017729282e-13;
				double Sm = 0.0;
				if (cidx > 0) Sm = (double)(lb[cidx - 1] & 0xFFFFFFu) * 1.52587890625e-05 + (double)lb[TBK_KDE_M + cidx - 1] * 9.094947017729282e-13;
				g[j] = ((double)(A >> 24) - S) + Sm;
			}
			__syncthreads();
			for (int j = 0; j < TBK_KDE_M / TBK_KDE_NT; ++j) sm.x[tid + j * nt] = make_double2(g[j], 0.0);
			__syncthreads();
		}
	} else {
		// long rings: three atomics per sample -- count, and the fraction as two 20-bit halves (exact to 2^-40 per sample
		// as long as no cell holds more than 4095 samples, verified afterwards)
		for (int i = tid; i < 3 * TBK_KDE_M; i += nt) lb[i] = 0u;
		__syncthreads();
		each([&](double d) {
			int li; double rem;
			if (cell_of(d, li, rem)) {
				const double t = rem * 1048576.0;
				const unsigned h20 = (unsigned)t;
				const unsigned l20 = (unsigned)((t - (double)h20) * 1048576.0);
				atomicAdd(&lb[li], 1u);
				atomicAdd(&lb[TBK_KDE_M + li], h20);
				atomicAdd(&lb[2 * TBK_KDE_M + li], l20);
			}
		});
		__syncthreads();
		int cmax = 0;
		for (int j = tid; j < TBK_KDE_M; j += nt) cmax = max(cmax, (int)lb[j]);
		double zmn = 0.0, zmx = (double)cmax;
		int zi = 0;
		block_sum_min_max(sm.red, zi, zmn, zmx);
		packed = zmx <= 4095.0;
		if (packed) {
			double g[TBK_KDE_M / TBK_KDE_NT];
			for (int j = 0; j < TBK_KDE_M / TBK_KDE_NT; ++j) {
				const int cidx = tid + j * nt;
				const double S = (double)lb[TBK_KDE_M + cidx] * 9.5367431640625e-07 + (double)lb[2 * TBK_KDE_M + cidx] * 9.094947017729282e-13;
				double Sm = 0.0;
				if (cidx > 0) Sm = (double)lb[TBK_KDE_M + cidx - 1] * 9.5367431640625e-07 + (double)lb[2 * TBK_KDE_M + cidx - 1] * 9.094947017729282e-13;
				g[j] = ((double)lb[cidx] - S) + Sm;
			}
			__syncthreads();
			for (int j = 0; j < TBK_KDE_M / TBK_KDE_NT; ++j) sm.x[tid + j * nt] = make_double2(g[j], 0.0);
			__syncthreads();
		}
	}
	if (!packed) {
		const bool fixed = n < 65536;
		__syncthreads();
		if (fixed) { for (int i = tid; i < 4 * TBK_KDE_M; i += nt) lb[i] = 0u; }
		else { for (int i = tid; i < TBK_KDE_M; i += nt) sm.x[i] = make_double2(0.0, 0.0); }
		__syncthreads();
		each([&](double d) {
			int li; double rem;
			if (cell_of(d, li, rem)) {
				if (fixed) {
					const unsigned long long fp = (unsigned long long)(rem * 281474976710656.0);  // rem * 2^48, rem in [0, 1)
					atomicAdd(&lb[li], 1u);
					atomicAdd(&lb[TBK_KDE_M + li], (uint32_t)(fp & 0xFFFFu));
					atomicAdd(&lb[2 * TBK_KDE_M + li], (uint32_t)((fp >> 16) & 0xFFFFu));
					atomicAdd(&lb[3 * TBK_KDE_M + li], (uint32_t)(fp >> 32));
				} else {
					atomicAdd(&sm.x[li].x, 1.0 - rem);
					atomicAdd(&sm.x[li + 1].x, rem);
				}
			}
		});
		__syncthreads();
		if (fixed) {
			double g[TBK_KDE_M / TBK_KDE_NT];
			for (int j = 0; j < TBK_KDE_M / TBK_KDE_NT; ++j) {
				const int cidx = tid + j * nt;
				const double S = (double)lb[TBK_KDE_M + cidx] * 3.552713678800501e-15 + (double)lb[2 * TBK_KDE_M + cidx] * 2.3283064365386963e-10
					+ (double)lb[3 * TBK_KDE_M + cidx] * 1.52587890625e-05;
				double Sm = 0.0;
				if (cidx > 0) Sm = (double)lb[TBK_KDE_M + cidx - 1] * 3.552713678800501e-15 + (double)lb[2 * TBK_KDE_M + cidx - 1] * 2.3283064365386963e-10
					+ (double)lb[3 * TBK_KDE_M + cidx - 1] * 1.52587890625e-05;
				g[j] = ((double)lb[cidx] - S) + Sm;
			}
			__syncthreads();
			for (int j = 0; j < TBK_KDE_M / TBK_KDE_NT; ++j) sm.x[tid + j * nt] = make_double2(g[j], 0.0);
			__syncthreads();
		}
	}
	// density = irfft(rfft(binned) * FAC): real-input transform through one 1024-point complex FFT each way
	// (z[j] = g[2j] + i g[2j+1]).  Only the argmax is needed, so positive scale factors are dropped
	// (binned = g / (delta nobs), the 1/M of the transforms).
	double2* fa = sm.x;                    // [1024]
	double2* fb = sm.x + TBK_KDE_M / 2;    // [1024]
	{
		constexpr int RP = TBK_KDE_M / 2 / TBK_KDE_NT;
		double ge[RP], go[RP];
		for (int j = 0; j < RP; ++j) { ge[j] = sm.x[2 * (tid + j * nt)].x; go[j] = sm.x[2 * (tid + j * nt) + 1].x; }
		__syncthreads();
		for (int j = 0; j < RP; ++j) fa[tid + j * nt] = make_double2(ge[j], go[j]);
		__syncthreads();
	}
	kde_fft1024<false>(fa, fb, P.twiddle);   // Z in fb
	// X[k] = E[k] + w^k O[k];  Y = X * FAC;  Z'[k] = E'[k] + i O'[k]  (pairs k, N-k handled together, in place)
	{
		const double fac1 = 2.0 * (M_PI * bw / range) * (M_PI * bw / range);
		const int N = TBK_KDE_M / 2;
		auto fac_of = [&](int J) {
			const double t = (double)J / (double)TBK_KDE_M * M_PI;
			return exp(-((double)J * (double)J * fac1)) / (1.0 - 1.0 / 3.0 * (t * t));
		};
		for (int k = tid; k <= N / 2; k += nt) {
			const int kn = (N - k) & (N - 1);
			const double2 zk = fb[k], zn = fb[kn];
			// forward post-processing for k and N-k
			const double2 w = kde_tw<false>(P.twiddle, k), wn = kde_tw<false>(P.twiddle, N - k);
			const double2 ek = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));      // (Z[k] + conj Z[N-k]) / 2
			const double2 ok = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));     // (Z[k] - conj Z[N-k]) / (2i)
			const double2 en = make_double2(ek.x, -ek.y), on = make_double2(ok.x, -ok.y);   // E[N-k] = conj E[k], O likewise
			double2 xk = cmul(w, ok); xk.x += ek.x; xk.y += ek.y;                           // X[k]
			double2 xn = cmul(wn, on); xn.x += en.x; xn.y += en.y;                          // X[N-k]
			const double fk = fac_of(k), fn = fac_of(N - k);
			const double2 yk = make_double2(xk.x * fk, xk.y * fk), yn = make_double2(xn.x * fn, xn.y * fn);
			if (k == 0) {
				// X[0] = Re Z0 + Im Z0, X[N] = Re Z0 - Im Z0 (both real); Z'[0] = (Y0 + YN)/2 + i (Y0 - YN)/2
				const double y0 = (zk.x + zk.y) * fac_of(0), yN = (zk.x - zk.y) * fac_of(N);
				fb[0] = make_double2(0.5 * (y0 + yN), 0.5 * (y0 - yN));
			} else {
				// inverse pre-processing: E' = (Y[k] + conj Y[N-k]) / 2, O' = (Y[k] - conj Y[N-k]) / 2 * conj(w^k)
				const double2 e2 = make_double2(0.5 * (yk.x + yn.x), 0.5 * (yk.y - yn.y));
				const double2 d2 = make_double2(0.5 * (yk.x - yn.x), 0.5 * (yk.y + yn.y));
				const double2 o2 = cmul(d2, make_double2(w.x, -w.y));
				fb[k] = make_double2(e2.x - o2.y, e2.y + o2.x);                               // E' + i O'
				if (kn != k) {
					const double2 e3 = make_double2(e2.x, -e2.y);                               // E'[N-k] = conj E'[k]
					const double2 d3 = make_double2(0.5 * (yn.x - yk.x), 0.5 * (yn.y + yk.y));
					const double2 o3 = cmul(d3, make_double2(wn.x, -wn.y));
					fb[kn] = make_double2(e3.x - o3.y, e3.y + o3.x);
				}
			}
		}
	}
	__syncthreads();
	kde_fft1024<true>(fb, fa, P.twiddle);    // z' in fa: density[2j] = Re, density[2j+1] = Im
	// argmax (first maximum) of the density
	double bv = -INFINITY; int bi = 0x7fffffff;
	for (int j = tid; j < TBK_KDE_M / 2; j += nt) {
		const double2 d = fa[j];
		if (d.x > bv) { bv = d.x; bi = 2 * j; }
		if (d.y > bv) { bv = d.y; bi = 2 * j + 1; }
	}
	for (int o = 16; o > 0; o >>= 1) {
		const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
		const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
		if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
	}
	if ((tid & 31) == 0) { sm.bestv[tid >> 5] = bv; sm.besti[tid >> 5] = bi; }
	__syncthreads();
	if (tid == 0) {
		for (int w = 1; w < nt / 32; ++w) {
			if (sm.bestv[w] > bv || (sm.bestv[w] == bv && sm.besti[w] < bi)) { bv = sm.bestv[w]; bi = sm.besti[w]; }
		}
		// support = linspace(a, b, M): arange * step + start, last point = stop
		const double g = (bi == TBK_KDE_M - 1) ? bb : __dadd_rn(__dmul_rn((double)bi, delta), a);
		*out = g;
	}
}

// ---------------------------------------------------------------------------------------------
// K_radial_fit: move_median_central (utilities.py:52-62), drop NaNs, not-a-knot cubic spline
// (InterpolatedUnivariateSpline k=3, backgrounds.py:186-197) in piecewise-polynomial form.
__device__ double nanmedian_small(const double* x, int lo, int hi)
{
	double t[64];
	int m = 0;
	for (int i = lo; i < hi; ++i) {
		const double d = x[i];
		if (d == d) {
			int j = m++;
			while (j > 0 && t[j - 1] > d) { t[j] = t[j - 1]; --j; }
			t[j] = d;
		}
	}
	if (m == 0) return nan_d();
	return (m & 1) ? t[m >> 1] : 0.5 * (t[(m >> 1) - 1] + t[m >> 1]);
}

struct RadialSmem {
	double raw[TBK_MAX_RINGS], s2[TBK_MAX_RINGS], kx[TBK_MAX_RINGS], ky[TBK_MAX_RINGS], h[TBK_MAX_RINGS], dl[TBK_MAX_RINGS],
		dg[TBK_MAX_RINGS], du[TBK_MAX_RINGS], rhs[TBK_MAX_RINGS], M2[TBK_MAX_RINGS];
	int m;
};

// one warp per FFI: the window medians, right-hand sides, piece coefficients and lookup tables are spread over
// the lanes; only the knot compaction and the tridiagonal solve (<= TBK_MAX_RINGS unknowns) are serial.
__global__ void __launch_bounds__(32) k_radial_fit(PlanDev P, Workspace ws, tbk_ffi_status* status, int round, int B)
{
	__shared__ RadialSmem rsm;
	const int b = blockIdx.x, lane = threadIdx.x;
	if (b >= B) return;
	FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	const int n = P.nrings;
	const double* raw_g = ws.s2_raw + (size_t)b * n;
	double* s2g = ws.s2_hist + ((size_t)b * P.bkgiters + round) * n;
	for (int i = lane; i < n; i += 32) rsm.raw[i] = raw_g[i];
	__syncwarp();
	const double* raw = rsm.raw;
	const int w = P.radial_smooth;
	if (w > 0) {
		const int ne = w / 2 + 1;                 // points fixed up at each end
		const int k = -(((-w) >> 1) + 1);         // python: -( (-w)//2 + 1 ), the np.roll shift
		if (n >= 2 * ne) {
			// bottleneck.move_median(min_count=1) rolled by -w//2+1, ends replaced by growing-window medians
			for (int i = lane; i < n; i += 32) {
				double v;
				if (i < ne) v = nanmedian_small(raw, 0, min(n, i + 2));
				else if (i >= n - ne) v = nanmedian_small(raw, max(0, n - ((n - 1 - i) + 2)), n);
				else { const int src = i + k; v = nanmedian_small(raw, max(0, src - w + 1), src + 1); }
				rsm.s2[i] = v;
			}
		} else if (lane == 0) {
			// tiny profile: the end fix-ups overlap, reproduce the assignment order literally
			for (int i = 0; i < n; ++i) { const int src = (i + k) % n; rsm.s2[i] = nanmedian_small(raw, max(0, src - w + 1), src + 1); }
			for (int e = 0; e < ne; ++e) {
				rsm.s2[e] = nanmedian_small(raw, 0, min(n, e + 2));
				rsm.s2[n - 1 - e] = nanmedian_small(raw, max(0, n - (e + 2)), n);
			}
		}
	} else {
		for (int i = lane; i < n; i += 32) rsm.s2[i] = raw[i];
	}
	__syncwarp();
	for (int i = lane; i < n; i += 32) s2g[i] = rsm.s2[i];
	// knots = finite ring values at the ring centres
	if (lane == 0) {
		int m = 0;
		for (int i = 0; i < n; ++i) {
			const double v = rsm.s2[i];
			if (v == v) { rsm.kx[m] = ring_center(P, i); rsm.ky[m] = v; ++m; }
		}
		rsm.m = m;
	}
	__syncwarp();
	const int m = rsm.m;
	const double *kx = rsm.kx, *ky = rsm.ky;
	int ok = 0;
	if (m >= 4) {
		double *h = rsm.h, *dl = rsm.dl, *dg = rsm.dg, *du = rsm.du, *rhs = rsm.rhs, *M2 = rsm.M2;
		for (int i = lane; i < m - 1; i += 32) h[i] = kx[i + 1] - kx[i];
		__syncwarp();
		const int nu = m - 2;  // unknowns M_1 .. M_{m-2} (second derivatives of the not-a-knot cubic interpolant)
		for (int i = 1 + lane; i <= nu; i += 32) {
			dl[i] = h[i - 1]; dg[i] = 2.0 * (h[i - 1] + h[i]); du[i] = h[i];
			rhs[i] = 6.0 * ((ky[i + 1] - ky[i]) / h[i] - (ky[i] - ky[i - 1]) / h[i - 1]);
		}
		__syncwarp();
		if (lane == 0) {
			// not-a-knot: M_0 = (1 + h0/h1) M_1 - (h0/h1) M_2 ; M_{m-1} likewise
			const double r0 = h[0] / h[1];
			dg[1] += dl[1] * (1.0 + r0);
			du[1] -= dl[1] * r0;
			const double r1 = h[m - 2] / h[m - 3];
			dg[nu] += du[nu] * (1.0 + r1);
			dl[nu] -= du[nu] * r1;
			for (int i = 2; i <= nu; ++i) {   // Thomas algorithm
				const double f = dl[i] / dg[i - 1];
				dg[i] -= f * du[i - 1];
				rhs[i] -= f * rhs[i - 1];
			}
			M2[nu] = rhs[nu] / dg[nu];
			for (int i = nu - 1; i >= 1; --i) M2[i] = (rhs[i] - du[i] * M2[i + 1]) / dg[i];
			M2[0] = (1.0 + r0) * M2[1] - r0 * M2[2];
			M2[m - 1] = (1.0 + r1) * M2[m - 2] - r1 * M2[m - 3];
		}
		__syncwarp();
		for (int i = lane; i < m - 1; i += 32) {
			c.kx[i] = kx[i];
			c.pp[i][0] = ky[i];
			c.pp[i][1] = (ky[i + 1] - ky[i]) / h[i] - h[i] * (2.0 * M2[i] + M2[i + 1]) / 6.0;
			c.pp[i][2] = 0.5 * M2[i];
			c.pp[i][3] = (M2[i + 1] - M2[i]) / (6.0 * h[i]);
		}
		if (lane == 0) c.kx[m - 1] = kx[m - 1];
		// dense table: ring-centre interval -> covering spline piece
		for (int i = lane; i < n; i += 32) {
			const double cen = ring_center(P, i);
			int sidx = 0;
			while (sidx + 1 < m - 1 && kx[sidx + 1] <= cen) ++sidx;
			c.seg_of_ring[i] = (short)sidx;
			c.seg[i][0] = kx[sidx];
			c.seg[i][1] = ky[sidx];
			c.seg[i][2] = (ky[sidx + 1] - ky[sidx]) / h[sidx] - h[sidx] * (2.0 * M2[sidx] + M2[sidx + 1]) / 6.0;
			c.seg[i][3] = 0.5 * M2[sidx];
			c.seg[i][4] = (M2[sidx + 1] - M2[sidx]) / (6.0 * h[sidx]);
			c.seg[i][5] = exp10(ky[sidx]);
		}
		if (lane == 0) { c.x0 = kx[0]; c.xlast = kx[m - 1]; c.c_flat = exp10(ky[0]) - c.zp; }
		// Taylor pieces of 10**spline - zp for the per-pixel evaluation of the residual statistics (RadialTab)
		{
			const int nsub = TBK_RSUB * max(n - 1, 1);
			const double hsub = P.step / (double)TBK_RSUB, c0 = ring_center(P, 0), zp = c.zp;
			double* rows = ws.rtab + (size_t)b * nsub * 8;
			for (int j = lane; j < nsub; j += 32) {
				const double u0 = c0 + ((double)j + 0.5) * hsub;
				const double tcl = clamp_d(u0, kx[0], kx[m - 1]);
				int sidx = 0;
				while (sidx + 1 < m - 1 && kx[sidx + 1] <= tcl) ++sidx;
				const double p0 = ky[sidx], p1 = (ky[sidx + 1] - ky[sidx]) / h[sidx] - h[sidx] * (2.0 * M2[sidx] + M2[sidx + 1]) / 6.0;
				const double p2 = 0.5 * M2[sidx], p3 = (M2[sidx + 1] - M2[sidx]) / (6.0 * h[sidx]);
				const double s0 = tcl - kx[sidx];
				const double LN10 = 2.302585092994045684;
				const double yv = p0 + s0 * (p1 + s0 * (p2 + s0 * p3));
				const double a1 = LN10 * (p1 + s0 * (2.0 * p2 + 3.0 * s0 * p3)), a2 = LN10 * (p2 + 3.0 * s0 * p3), a3 = LN10 * p3;
				double bb[7];
				bb[0] = exp10(yv);
				for (int q = 1; q <= 6; ++q) {
					double acc = a1 * bb[q - 1];
					if (q >= 2) acc += 2.0 * a2 * bb[q - 2];
					if (q >= 3) acc += 3.0 * a3 * bb[q - 3];
					bb[q] = acc / (double)q;
				}
				rows[8 * j] = bb[0] - zp;
				for (int q = 1; q <= 6; ++q) rows[8 * j + q] = bb[q];
				rows[8 * j + 7] = tcl;
			}
		}
		ok = 1;
	}
	// m == 3: FITPACK raises "m must be > k" (caught, backgrounds.py:192-194); m < 3: not enough points.
	if (lane == 0) {
		c.radial_ok = ok;
		c.npts = m;
		if (!ok) c.c_flat = 0.0;
		if (status) {
			status[b].n_ring_valid[round] = m;
			status[b].radial_ok[round] = ok;
			status[b].zeropoint[round] = c.zp;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K_tile_round: sigma-clipped statistics of the non-flat meshes on x - radial(r) (backgrounds.py:200).
__global__ void __launch_bounds__(TBK_NT) k_tile_round(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ mask)
{
	__shared__ TileSmem sm;
	__shared__ RadialSmem2 rs;
	const int slot = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
	const FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	radial_stage(rs, c, P);
	__syncthreads();
	const int tile = P.nonflat_tiles[slot];
	const int ty = tile / P.nx, tx = tile % P.nx;
	const size_t img = (size_t)b * P.H * P.W;
	const int lcol = tile_lcol(tid);
	const int gx = tx * TBK_TILE + lcol;
	double v[TBK_VPT];
	unsigned valid = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int gy = ty * TBK_TILE + tile_lrow(tid, j);
		const size_t off = img + (size_t)gy * P.W + gx;
		const float4 x = __ldg(reinterpret_cast<const float4*>(cube + off));
		const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(mask + off));
		const float x4[4] = {x.x, x.y, x.z, x.w};
		const unsigned char m4[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			double d = 0.0;
			if (!m4[q]) {
				d = (double)x4[q] - radial_value_s(rs, pixel_radius(P, gy, gx + q));
				valid |= 1u << (4 * j + q);
			}
			v[4 * j + q] = d;
		}
	}
	TileStat st = tile_sigma_clip<double>(v, valid, sm);
	if (tid == 0) ws.tile_nf[(size_t)b * P.n_nonflat + slot] = st;
}

// K_tile_round_w: same result as k_tile_round with the bucketed algorithm of tbk_tile_warp.cuh on float64
// residuals.  128 threads per mesh; thread t, load i (0..7) owns row 8i + 2(t/32) + (t%32)/16, columns 4(t%16)..+3.
#ifndef TWR_MINB
#define TWR_MINB 5
#endif
template <bool QUEUE>
__global__ void __launch_bounds__(128, TWR_MINB) k_tile_round_w(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ mask, int round)
{
	__shared__ TwBlockSmem<TwF64, 4> sm;
	__shared__ RadialSmem2 rs;
	const int tid = threadIdx.x;
	// QUEUE: work off the meshes k_tile_round_z queued (grid-stride over ws.fb_list2); otherwise one CTA per (mesh, FFI)
	const int count = QUEUE ? ws.fb_count[1 + round] : 1;
	for (int e = QUEUE ? blockIdx.x : 0; e < count; e += QUEUE ? gridDim.x : 1) {
	const int slot = QUEUE ? ws.fb_list2[e] % P.n_nonflat : blockIdx.x, b = QUEUE ? ws.fb_list2[e] / P.n_nonflat : blockIdx.y;
	const FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	__syncthreads();
	radial_stage(rs, c, P);
	__syncthreads();
	const int tile = P.nonflat_tiles[slot];
	const int ty = tile / P.nx, tx = tile % P.nx;
	const int lane = tid & 31, w = tid >> 5;
	const int lcol = (lane & 15) << 2, gx = tx * TBK_TILE + lcol;
	__shared__ int s_n;
	if (tid == 0) s_n = 0;
	__syncthreads();
	int n = 0;
	// rolled (the radial evaluation is ~100 instructions per pixel) with the next step's loads issued ahead
	const size_t img = (size_t)b * P.H * P.W;
	const double* rbase = P.nonflat_r + (size_t)slot * TBK_NPIX_TILE;
	const int lrow_first = 2 * w + (lane >> 4);
	float4 x = __ldg(reinterpret_cast<const float4*>(cube + img + (size_t)(ty * TBK_TILE + lrow_first) * P.W + gx));
	unsigned int m = __ldg(reinterpret_cast<const unsigned int*>(mask + img + (size_t)(ty * TBK_TILE + lrow_first) * P.W + gx));
	double2 r01 = __ldg(reinterpret_cast<const double2*>(rbase + lrow_first * TBK_TILE + lcol));
	double2 r23 = __ldg(reinterpret_cast<const double2*>(rbase + lrow_first * TBK_TILE + lcol + 2));
#pragma unroll 1
	for (int i = 0; i < 8; ++i) {
		const int lrow = 8 * i + lrow_first;
		const float x4[4] = {x.x, x.y, x.z, x.w};
		const double rr[4] = {r01.x, r01.y, r23.x, r23.y};
		const unsigned int mcur = m;
		if (i < 7) {
			const int lnext = lrow + 8;
			const size_t off = img + (size_t)(ty * TBK_TILE + lnext) * P.W + gx;
			x = __ldg(reinterpret_cast<const float4*>(cube + off));
			m = __ldg(reinterpret_cast<const unsigned int*>(mask + off));
			r01 = __ldg(reinterpret_cast<const double2*>(rbase + lnext * TBK_TILE + lcol));
			r23 = __ldg(reinterpret_cast<const double2*>(rbase + lnext * TBK_TILE + lcol + 2));
		}
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			unsigned long long k = ~0ULL;
			if (!((mcur >> (8 * q)) & 0xFFu)) { k = dkey((double)x4[q] - radial_value_s(rs, rr[q])); ++n; }
			sm.tw.keys[lrow * TBK_TILE + lcol + q] = k;
		}
	}
	n = __reduce_add_sync(0xffffffffu, n);
	if (lane == 0) atomicAdd(&s_n, n);
	__syncthreads();
	TileStat st; bool writer;
	tile_block_stats_staged<TwF64, 4>(sm, s_n, st, writer);
	if (writer) ws.tile_nf[(size_t)b * P.n_nonflat + slot] = st;
	}
}

// K_tile_round_z / K_tile_round_fin: the residual statistics (x - radial(r), float64) of the non-flat meshes with the zone
// algorithm of tbk_tile_zone.cuh, in two kernels.
//   k_tile_round_z   one CTA of 128 threads per mesh: the residuals are evaluated once (Taylor-piece table of the round,
//                    RadialTab) and staged in shared memory; the two classification passes of the four warps run over the
//                    staged values; every warp appends its tails / zone elements to its own quarter of the mesh's lists
//                    in global memory by ballot prefix (no atomics: the lists do not depend on timing).
//   k_tile_round_fin one warp per mesh finishes from those lists (clip iterations on the tails, medians from the zone).
// The finish phase is serial work of one warp (a few thousand dependent instructions); as a kernel of its own it runs at
// ~24 warps per SM instead of idling three of four warps of the producer CTA.
// Meshes the lists cannot answer are queued in ws.fb_list2 for k_tile_round_w<true>.
#define ZR_ROWS TBK_RTAB_ROWS
struct ZoneRoundSmem {
	double d[TBK_NPIX_TILE];          // staged residuals, NaN = masked
	double rows[ZR_ROWS][TBK_RROW];   // the Taylor pieces this mesh can see (RadialTab rows jlo .. jlo + ZR_ROWS - 1), padded
	unsigned long long bar;           // mbarrier of the TMA copy of the mesh's radius tile
	double red[4][2];
	int redi[4][4];
	double A, B, M, pivot, shat, mhat;
	int ok, n;
};

__global__ void __launch_bounds__(128, 5) k_tile_round_z(PlanDev P, Workspace ws,
	const float* __restrict__ cube, const uint8_t* __restrict__ mask, int round)
{
	__shared__ __align__(16) ZoneRoundSmem sm;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int slot = blockIdx.x, b = blockIdx.y;
	const FfiCtl& c = ws.ctl[b];
	ZoneRec& rec = *reinterpret_cast<ZoneRec*>(ws.zrec + ((size_t)b * P.n_nonflat + slot) * ZR_REC_BYTES);
	if (c.all_masked || c.no_good_mesh) { if (tid == 0) rec.state = 0; return; }
	const RadialTab rt = radial_tab(ws.rtab + (size_t)b * TBK_RSUB * max(P.nrings - 1, 1) * 8, c, P);
	const int tile = P.nonflat_tiles[slot];
	const int ty = tile / P.nx, tx = tile % P.nx;
	// The static Taylor-piece words of the mesh (PlanDev::nonflat_uj, 32 KB, contiguous) come in through the TMA engine (one
	// cp.async.bulk into sm.d, the residuals later overwrite them in place); meanwhile every thread has all of its pixel /
	// mask loads in flight at once and the Taylor pieces the mesh can see are staged (a mesh spans < 91 px).
	if (tid == 0) {
		tma_bar_init(&sm.bar, 1);
		tma_bar_expect(&sm.bar, (unsigned)(TBK_NPIX_TILE * sizeof(double)));
		tma_load_1d(sm.d, P.nonflat_uj + (size_t)slot * TBK_NPIX_TILE, (unsigned)(TBK_NPIX_TILE * sizeof(double)), &sm.bar);
	}
	// thread t, step i (0..7) owns row 8i + 2(t/32) + (t%32)/16, columns 4(t%16) .. +3
	const int lrow0 = 2 * w + (lane >> 4), lcol = (lane & 15) << 2;
	float4 xs[8]; unsigned int mks[8];
	{
		const size_t off0 = (size_t)b * P.H * P.W + (size_t)(ty * TBK_TILE + lrow0) * P.W + tx * TBK_TILE + lcol;
		const size_t step = (size_t)8 * P.W;
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			xs[i] = __ldg(reinterpret_cast<const float4*>(cube + off0 + i * step));
			mks[i] = __ldg(reinterpret_cast<const unsigned int*>(mask + off0 + i * step));
		}
	}
	const int jlo = __ldg(P.nonflat_jlo + slot);
	if (rt.radial_ok) {
		if (jlo < 0) {   // cannot happen for meshes of 64 px and step / 8 >= 1.75 px; other parameters: bucketed path
			__syncthreads();
			if (tid == 0) { tma_bar_wait(&sm.bar, 0u); rec.state = 0; ws.fb_list2[atomicAdd(ws.fb_count + 1 + round, 1)] = b * P.n_nonflat + slot; atomicAdd(ws.fb_count + 24 + ZN_WHY_EMPTY, 1); }
			return;
		}
		const int nrow = min(TBK_RTAB_ROWS, rt.nsub - jlo);
		for (int e = tid; e < nrow * 4; e += 128)
			reinterpret_cast<double2*>(&sm.rows[e >> 2][0])[e & 3] = __ldg(reinterpret_cast<const double2*>(rt.rows + 8 * (size_t)jlo) + e);
	}
	// the profile at the two ends of the spline (ext=3 clamp) and the pieces between them
	double cflat = 0.0, clast = 0.0;
	int j0 = 0, j1 = 0;
	if (rt.radial_ok) { cflat = radial_tab_eval(rt, rt.x0); clast = radial_tab_eval(rt, rt.xlast); j0 = radial_tab_j0(rt); j1 = radial_tab_j1(rt); }
	__syncthreads();                 // rows staged, barrier initialised
	tma_bar_wait(&sm.bar, 0u);       // piece words landed
	{
		int n = 0;
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			double2* cell = reinterpret_cast<double2*>(&sm.d[(lrow0 + 8 * i) * TBK_TILE + lcol]);
			const double2 r01 = cell[0], r23 = cell[1];
			const float x4[4] = {xs[i].x, xs[i].y, xs[i].z, xs[i].w};
			const double uj[4] = {r01.x, r01.y, r23.x, r23.y};
			double dd[4];
			// where every pixel of the warp's two rows is clamped to the same end the evaluation is skipped
			bool inner = false;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const int jl = (int)((unsigned long long)__double_as_longlong(uj[q]) & 63ull);
				inner |= jl < 62 && jlo + jl >= j0 && jlo + jl < j1;
			}
			const bool beyond = rt.radial_ok && __any_sync(0xffffffffu, inner);
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const bool ok = !((mks[i] >> (8 * q)) & 0xFFu);
				double rad = 0.0;
				if (beyond) rad = radial_tab_eval_uj(&sm.rows[0][0], jlo, j0, j1, cflat, clast, uj[q]);
				else if (rt.radial_ok) {
					const int jl = (int)((unsigned long long)__double_as_longlong(uj[q]) & 63ull);
					rad = (jl == 62 || jlo + jl < j0) ? cflat : clast;
				}
				dd[q] = ok ? (double)x4[q] - rad : nan_d();
				n += ok;
			}
			cell[0] = make_double2(dd[0], dd[1]); cell[1] = make_double2(dd[2], dd[3]);
		}
		n = __reduce_add_sync(0xffffffffu, n);
		if (lane == 0) sm.redi[w][0] = n;
	}
	__syncthreads();
	// ---- sample (warp 0): 2 x 32 staged residuals spread over the mesh
	if (w == 0) {
		unsigned long long sk[2];
#pragma unroll
		for (int t = 0; t < 2; ++t) {
			const double v = sm.d[(lane * 131 + 17 + t * 2053) & (TBK_NPIX_TILE - 1)];
			sk[t] = (v == v) ? dkey(v) : Zn64::padkey();
		}
		ZonePlan zp = zone_plan<Zn64>(sk[0], sk[1], lane);
		if (round > 0 && zp.ok) {
			// the width of the previous round's clipped distribution of this mesh is a far better scale than the IQR of 64
			// samples (star wings inflate it, and a bulk wider than the final clip range costs a fallback); the centre stays
			// the fresh sample's, because the radial profile moves between rounds
			const TileStat prev = ws.tile_nf[(size_t)b * P.n_nonflat + slot];
			if (prev.nfin >= ZN_MIN_N && prev.std > 0.0 && prev.std < zp.shat) {
				zp.shat = prev.std; zp.A = zp.mhat - ZN_CT * zp.shat; zp.B = zp.mhat + ZN_CT * zp.shat;
			}
		}
		if (lane == 0) {
			sm.ok = zp.ok; sm.A = zp.A; sm.B = zp.B; sm.M = zp.mhat; sm.mhat = zp.mhat; sm.shat = zp.shat; sm.pivot = zp.pivot;
			sm.n = sm.redi[0][0] + sm.redi[1][0] + sm.redi[2][0] + sm.redi[3][0];
		}
	}
	__syncthreads();
	const int n = sm.n;
	if (n == 0) {
		if (tid == 0) {
			TileStat st; st.mean = st.med = st.std = nan_d(); st.nfin = 0; st.pad = 0;
			ws.tile_nf[(size_t)b * P.n_nonflat + slot] = st;
			rec.state = 0;
		}
		return;
	}
	if (!sm.ok || n < ZN_MIN_N) {
		if (tid == 0) { rec.state = 0; ws.fb_list2[atomicAdd(ws.fb_count + 1 + round, 1)] = b * P.n_nonflat + slot; atomicAdd(ws.fb_count + 24 + ZN_WHY_SAMPLE, 1); }
		return;
	}
	const double A = sm.A, Bv = sm.B, M = sm.M, pivot = sm.pivot;
	const unsigned lt = (1u << lane) - 1u;
	double* gt = reinterpret_cast<double*>(reinterpret_cast<char*>(&rec) + sizeof(ZoneRec));
	double* gz = gt + 4 * ZR_TQ;
	// ---- pass 1 over the staged values (element j * 128 + t): counts, bulk moments, tails in (j, lane) order
	{
		int nA = 0, nM = 0, nC = 0, tpos = 0;
		double s1 = 0.0, s2 = 0.0;
		double* gtw = gt + w * ZR_TQ;
		const double C0 = M - ZN_CW * sm.shat, C1 = M + ZN_CW * sm.shat;
#pragma unroll 4
		for (int j = 0; j < TBK_NPIX_TILE / 128; ++j) {
			const double d = sm.d[j * 128 + tid];
			const bool bulk = d >= A && d <= Bv;       // false for NaN
			const bool isT = d == d && !bulk;
			nA += (d < A); nM += (d < M); nC += (d >= C0 && d <= C1);
			const double e = (bulk ? d : pivot) - pivot;
			s1 += e; s2 = fma(e, e, s2);
			const unsigned mT = __ballot_sync(0xffffffffu, isT);
			const int pos = tpos + __popc(mT & lt);
			if (isT && pos < ZR_TQ) gtw[pos] = d;
			tpos += __popc(mT);
		}
		nA = __reduce_add_sync(0xffffffffu, nA); nM = __reduce_add_sync(0xffffffffu, nM); nC = __reduce_add_sync(0xffffffffu, nC);
		s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
		if (lane == 0) { sm.redi[w][0] = nC; sm.redi[w][1] = nA; sm.redi[w][2] = nM; sm.redi[w][3] = tpos; sm.red[w][0] = s1; sm.red[w][1] = s2; }
	}
	__syncthreads();
	const int nC = sm.redi[0][0] + sm.redi[1][0] + sm.redi[2][0] + sm.redi[3][0];
	const int nA = sm.redi[0][1] + sm.redi[1][1] + sm.redi[2][1] + sm.redi[3][1];
	const int nM = sm.redi[0][2] + sm.redi[1][2] + sm.redi[2][2] + sm.redi[3][2];
	const int tq[4] = {sm.redi[0][3], sm.redi[1][3], sm.redi[2][3], sm.redi[3][3]};
	const int nB = tq[0] + tq[1] + tq[2] + tq[3] - nA;
	const double s1 = (sm.red[0][0] + sm.red[1][0]) + (sm.red[2][0] + sm.red[3][0]);
	const double s2 = (sm.red[0][1] + sm.red[1][1]) + (sm.red[2][1] + sm.red[3][1]);
	ZonePlan zp;
	zp.ok = true; zp.mhat = sm.mhat; zp.shat = sm.shat; zp.pivot = pivot; zp.A = A; zp.B = Bv;
	double ZL, ZH;
	zone_range(zp, n, nA, nB, nM, nC, ZL, ZH, false);
	const bool tail_ovf = tq[0] > ZR_TQ || tq[1] > ZR_TQ || tq[2] > ZR_TQ || tq[3] > ZR_TQ;
	if (!(ZL < ZH) || tail_ovf || n - nA - nB <= 0) {
		if (tid == 0) { rec.state = 0; ws.fb_list2[atomicAdd(ws.fb_count + 1 + round, 1)] = b * P.n_nonflat + slot; atomicAdd(ws.fb_count + 24 + ZN_WHY_RANGE, 1); }
		return;
	}
	__syncthreads();   // redi is reused below
	// ---- pass 2: zone elements into the warp's quarter of the zone list; elements below the zone are counted
	{
		int nZL = 0, zpos = 0;
		double* gzw = gz + w * ZR_ZQ;
#pragma unroll 4
		for (int j = 0; j < TBK_NPIX_TILE / 128; ++j) {
			const double d = sm.d[j * 128 + tid];
			nZL += (d < ZL);
			const bool isZ = d >= ZL && d <= ZH;
			const unsigned mZ = __ballot_sync(0xffffffffu, isZ);
			const int pos = zpos + __popc(mZ & lt);
			if (isZ && pos < ZR_ZQ) gzw[pos] = d;
			zpos += __popc(mZ);
		}
		nZL = __reduce_add_sync(0xffffffffu, nZL);
		if (lane == 0) { sm.redi[w][0] = nZL; sm.redi[w][1] = zpos; }
	}
	__syncthreads();
	if (tid != 0) return;
	const int zq[4] = {sm.redi[0][1], sm.redi[1][1], sm.redi[2][1], sm.redi[3][1]};
	if (zq[0] > ZR_ZQ || zq[1] > ZR_ZQ || zq[2] > ZR_ZQ || zq[3] > ZR_ZQ || zq[0] + zq[1] + zq[2] + zq[3] == 0) {
		rec.state = 0; ws.fb_list2[atomicAdd(ws.fb_count + 1 + round, 1)] = b * P.n_nonflat + slot; atomicAdd(ws.fb_count + 24 + ZN_WHY_LIST, 1);
		return;
	}
	rec.n = n; rec.nA = nA; rec.nB = nB; rec.nZL = sm.redi[0][0] + sm.redi[1][0] + sm.redi[2][0] + sm.redi[3][0];
	for (int q = 0; q < 4; ++q) { rec.tq[q] = tq[q]; rec.zq[q] = zq[q]; }
	rec.s1 = s1; rec.s2 = s2; rec.pivot = pivot; rec.A = A; rec.B = Bv; rec.ZL = ZL; rec.ZH = ZH;
	rec.state = 1;
}

#define ZF_WARPS 4
#define ZF_ZCAP 512
struct ZoneFinSmem {
	unsigned long long zone[ZF_ZCAP];
	uint32_t cnt[ZN_BINS];
};
__global__ void __launch_bounds__(32 * ZF_WARPS, 6) k_tile_round_fin(PlanDev P, Workspace ws, int round)
{
	__shared__ ZoneFinSmem smw[ZF_WARPS];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int slot = blockIdx.x * ZF_WARPS + w, b = blockIdx.y;
	if (slot >= P.n_nonflat) return;
	const ZoneRec& rec = *reinterpret_cast<const ZoneRec*>(ws.zrec + ((size_t)b * P.n_nonflat + slot) * ZR_REC_BYTES);
	if (rec.state != 1) return;
	ZoneFinSmem& sm = smw[w];
	const double* gt = reinterpret_cast<const double*>(reinterpret_cast<const char*>(&rec) + sizeof(ZoneRec));
	const double* gz = gt + 4 * ZR_TQ;
	const int tq[4] = {rec.tq[0], rec.tq[1], rec.tq[2], rec.tq[3]}, zq[4] = {rec.zq[0], rec.zq[1], rec.zq[2], rec.zq[3]};
	const int nZ = zq[0] + zq[1] + zq[2] + zq[3];
	const double ZL = rec.ZL, ZH = rec.ZH;
	const float zscale = (float)ZN_BINS / (float)(ZH - ZL) * 0.99999f;
	TileStat st;
	bool good = nZ <= ZF_ZCAP;
	int why = ZN_WHY_LIST;
	// the tails go from the four quarters (global memory, L2) straight into the registers of the clip sweeps
	if (good) good = zone_finish_seg<Zn64, 4, 4>(gt, ZR_TQ, tq, gz, ZR_ZQ, zq, sm.zone, sm.cnt, lane,
		rec.n, rec.nA, rec.nB, rec.nZL, rec.s1, rec.s2, rec.pivot, rec.A, rec.B, ZL, zscale, st, why);
	if (lane == 0) {
		if (good) ws.tile_nf[(size_t)b * P.n_nonflat + slot] = st;
		else { ws.fb_list2[atomicAdd(ws.fb_count + 1 + round, 1)] = b * P.n_nonflat + slot; atomicAdd(ws.fb_count + 24 + why, 1); }
	}
}

// ---------------------------------------------------------------------------------------------
// K_mesh_finalize: one CTA per FFI.  SExtractor estimator per mesh, mesh exclusion, IDW fill,
// 3x3 nan-median, cubic-spline prefilter (photutils 1.3.0 Background2D).
__device__ __forceinline__ double median_small(double* t, int m)
{
	for (int i = 1; i < m; ++i) {
		const double d = t[i];
		int j = i;
		while (j > 0 && t[j - 1] > d) { t[j] = t[j - 1]; --j; }
		t[j] = d;
	}
	return (m & 1) ? t[m >> 1] : 0.5 * (t[(m >> 1) - 1] + t[m >> 1]);
}

// dynamic shared memory: three double arrays + the kd-tree of the IDW fill (index permutation, nodes, build frontier)
static size_t mesh_finalize_smem(int ntiles)
{
	return 3 * (size_t)ntiles * sizeof(double) + sizeof(uint32_t) * (size_t)ntiles + sizeof(KdtNode) * (size_t)(2 * ntiles + 2)
		+ sizeof(KdtFrontier) * 2 * (size_t)(2 * (ntiles / (KDT_LEAFSIZE + 1)) + 2);
}

__global__ void __launch_bounds__(1024) k_mesh_finalize(PlanDev P, Workspace ws,
	tbk_ffi_status* status, int round)
{
	extern __shared__ __align__(16) unsigned char smraw[];
	double* val = reinterpret_cast<double*>(smraw);          // [ntiles] mesh statistic (NaN = excluded)
	double* fil = val + P.ntiles;                            // [ntiles] filled / filtered mesh
	double* tmp = fil + P.ntiles;                            // [ntiles] scratch
	__shared__ RedSmem red;
	__shared__ int s_nexcl;
	const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
	FfiCtl& c = ws.ctl[b];
	if (c.all_masked || c.no_good_mesh) return;
	const int nt_tiles = P.ntiles, ny = P.ny, nx = P.nx;
	if (tid == 0) s_nexcl = 0;
	__syncthreads();

	// (1) mesh statistic
	int nexcl = 0;
	for (int t = tid; t < nt_tiles; t += nt) {
		TileStat st;
		double shift = 0.0;
		const int slot = P.use_radial ? P.tile_slot[t] : -1;
		if (slot >= 0) st = ws.tile_nf[(size_t)b * P.n_nonflat + slot];
		else { st = ws.tile_base[(size_t)b * nt_tiles + t]; shift = c.radial_ok ? c.c_flat : 0.0; }
		const int nbad = TBK_NPIX_TILE - st.nfin;
		double r = nan_d();
		if (nbad <= TBK_NPIX_TILE / 2 && st.nfin > 0) {  // exclude_percentile = 50
			const double mean = st.mean - shift, med = st.med - shift, sd = st.std;
			r = 2.5 * med - 1.5 * mean;
			if (sd == 0.0) r = mean;
			else if (!(fabs(mean - med) / sd < 0.3)) r = med;
		} else ++nexcl;
		val[t] = r;
	}
	if (nexcl) atomicAdd(&s_nexcl, nexcl);
	__syncthreads();
	nexcl = s_nexcl;
	if (status && tid == 0) { status[b].n_excluded[round] = nexcl; status[b].rounds = round + 1; }
	if (nexcl == nt_tiles) {
		// photutils raises ValueError("All meshes contain > ... masked pixels"); report and give NaN.
		if (tid == 0) { c.no_good_mesh = 1; if (status) status[b].no_good_mesh = 1; }
		return;
	}

	// (2) IDW fill of excluded meshes (photutils ShepardIDWInterpolator over the 10 nearest good meshes, weights 1/d).
	// The neighbours are the ones scipy.spatial.cKDTree(good_yx, leafsize=10).query(k=10) returns, ties included: the tree
	// is rebuilt here exactly as scipy builds it (tbk_kdtree.cuh) -- one thread builds, every excluded mesh then searches.
	// The neighbour table of an FFI is kept with the good-mesh bitmap it belongs to; a later round whose bitmap is the
	// same (the usual case) reuses it.
	if (nexcl) {
		uint32_t* kidx = reinterpret_cast<uint32_t*>(tmp + nt_tiles);
		KdtNode* knodes = reinterpret_cast<KdtNode*>(kidx + nt_tiles);
		KdtFrontier* kfront = reinterpret_cast<KdtFrontier*>(knodes + 2 * nt_tiles + 2);
		const int kcap = 2 * (nt_tiles / (KDT_LEAFSIZE + 1)) + 2;
		__shared__ KdtTree tree;
		__shared__ int s_same, s_ovf, s_cnt[2], s_wcount[32];
		uint32_t* bits = ws.idw_bits + (size_t)b * ((P.ntiles + 31) / 32 + 1);   // [0] = valid flag, then the bitmap
		uint16_t* tab = ws.idw_tab + (size_t)b * P.ntiles * KDT_K;
		const int nwords = (nt_tiles + 31) / 32;
		if (tid == 0) { s_same = (round > 0 && bits[0] == 1u) ? 1 : 0; s_ovf = 0; }
		__syncthreads();
		for (int w = tid; w < nwords; w += nt) {
			uint32_t m = 0u;
			for (int j = 0; j < 32; ++j) { const int t = 32 * w + j; if (t < nt_tiles && val[t] == val[t]) m |= 1u << j; }
			if (bits[1 + w] != m) { s_same = 0; bits[1 + w] = m; }
		}
		__syncthreads();
		bool reuse = s_same != 0;
		// Not this FFI's previous pattern: look the bitmap up in the plan's pattern cache (P.idw_cache: a few entries of
		// {state, bitmap, neighbour table} shared by every FFI, batch and stream of the plan).  Whole stacks repeat one pattern
		// (the Mars frames of sector 1 exclude the same eight mesh columns in every cadence), so the tree is built once.
		const size_t cstride = 1 + (size_t)nwords + ((size_t)nt_tiles * KDT_K + 1) / 2;    // words per cache entry
		__shared__ int s_hit;
		if (!reuse) {
			if (tid == 0) s_hit = -1;
			__syncthreads();
			for (int e = 0; e < TBK_IDW_CACHE; ++e) {
				const uint32_t* ce = P.idw_cache + 1 + e * cstride;
				if (*(volatile const uint32_t*)ce != 2u) continue;     // uniform: every thread reads the same word
				__threadfence();
				int same = 1;
				for (int w = tid; w < nwords; w += nt) same &= (ce[1 + w] == bits[1 + w]);
				if (__syncthreads_and(same)) { if (tid == 0) s_hit = e; break; }
			}
			__syncthreads();
			if (s_hit >= 0) {
				const uint16_t* ctab = reinterpret_cast<const uint16_t*>(P.idw_cache + 1 + s_hit * cstride + 1 + nwords);
				for (int i = tid; i < nt_tiles * KDT_K; i += nt) tab[i] = ctab[i];
				if (tid == 0) bits[0] = 1u;
				__syncthreads();
				reuse = true;
			}
		}
		const bool built = !reuse;
		if (!reuse) {
			// good meshes in increasing mesh id (the reference's point order), packed with their coordinates
			int base = 0;
			for (int t0 = 0; t0 < nt_tiles; t0 += nt) {
				const int t = t0 + tid;
				const bool g = t < nt_tiles && val[t] == val[t];
				const unsigned bal = __ballot_sync(0xffffffffu, g);
				if ((tid & 31) == 0) s_wcount[tid >> 5] = __popc(bal);
				__syncthreads();
				int off = base;
				for (int w = 0; w < (tid >> 5); ++w) off += s_wcount[w];
				if (g) kidx[off + __popc(bal & ((1u << (tid & 31)) - 1u))] = kdt_pack(t, nx);
				for (int w = 0; w < (nt >> 5); ++w) base += s_wcount[w];
				__syncthreads();
			}
			if (tid == 0) { tree.idx = kidx; tree.nodes = knodes; tree.npts = base; tree.nx = nx; bits[0] = 1u; }
			__syncthreads();
			kdt_build_cta(tree, kfront, kcap, 2 * nt_tiles + 2, s_cnt);
			if (tid == 0 && tree.overflow) s_ovf = 1;
			__syncthreads();
		}
		for (int t = tid; t < nt_tiles; t += nt) {
			double r = val[t];
			if (!(r == r)) {
				int id[KDT_K], d2[KDT_K], m;
				if (reuse) {
					m = 0;
					for (int j = 0; j < KDT_K; ++j) {
						const int g = tab[(size_t)t * KDT_K + j];
						if (g == 0xFFFF) break;
						const int dy = g / nx - t / nx, dx = g % nx - t % nx;
						id[m] = g; d2[m] = dy * dy + dx * dx; ++m;
					}
				} else {
					int ovf = 0;
					m = kdt_query(tree, t / nx, t % nx, KDT_K, id, d2, &ovf);
					if (ovf) s_ovf = 1;
					for (int j = 0; j < KDT_K; ++j) tab[(size_t)t * KDT_K + j] = j < m ? (uint16_t)id[j] : (uint16_t)0xFFFF;
				}
				r = kdt_shepard(id, d2, m, [&](int g) { return val[g]; });
			}
			fil[t] = r;
		}
		__syncthreads();
		if (s_ovf && tid == 0 && status) status[b].n_excluded[round] = -1;   // cannot happen for <= 4096 meshes (checked by the tests)
		// publish a freshly built table (entries for good meshes are never read: fill them so the copy is complete)
		if (built && !s_ovf) {
			__shared__ int s_slot;
			if (tid == 0) { const uint32_t k = atomicAdd(P.idw_cache, 1u); s_slot = k < TBK_IDW_CACHE ? (int)k : -1; }
			__syncthreads();
			if (s_slot >= 0) {
				uint32_t* ce = P.idw_cache + 1 + s_slot * cstride;
				for (int w = tid; w < nwords; w += nt) ce[1 + w] = bits[1 + w];
				uint16_t* ctab = reinterpret_cast<uint16_t*>(ce + 1 + nwords);
				for (int t = tid; t < nt_tiles; t += nt) {
					const bool excluded = !(val[t] == val[t]);
					for (int j = 0; j < KDT_K; ++j) ctab[(size_t)t * KDT_K + j] = excluded ? tab[(size_t)t * KDT_K + j] : (uint16_t)0xFFFF;
				}
				__threadfence();
				__syncthreads();
				if (tid == 0) *(volatile uint32_t*)ce = 2u;
			}
		}
	} else {
		for (int t = tid; t < nt_tiles; t += nt) fil[t] = val[t];
	}
	__syncthreads();

	// (3) 3x3 nan-median, NaN padding (generic_filter(nanmedian, mode='constant', cval=nan))
	for (int t = tid; t < nt_tiles; t += nt) {
		const int iy = t / nx, ix = t % nx;
		double w9[9]; int m = 0;
		for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
			const int yy = iy + dy, xx = ix + dx;
			if (yy >= 0 && yy < ny && xx >= 0 && xx < nx) { const double d = fil[yy * nx + xx]; if (d == d) w9[m++] = d; }
		}
		tmp[t] = m ? median_small(w9, m) : nan_d();
	}
	__syncthreads();
	double mn = INFINITY, mx = -INFINITY; int dummy = 0;
	double* hist = ws.mesh_hist + ((size_t)b * P.bkgiters + round) * nt_tiles;
	for (int t = tid; t < nt_tiles; t += nt) { mn = fmin(mn, tmp[t]); mx = fmax(mx, tmp[t]); hist[t] = tmp[t]; }
	block_sum_min_max(red, dummy, mn, mx);
	if (tid == 0) { c.mesh_min = mn; c.mesh_max = mx; c.mesh_const = (mx - mn == 0.0) ? 1 : 0; }

	// (4) cubic B-spline prefilter, mode='reflect' (scipy ni_splines.c), axis 0 then axis 1
	const double z = sqrt(3.0) - 2.0;
	const double gain = (1.0 - z) * (1.0 - 1.0 / z);
	for (int axis = 0; axis < 2; ++axis) {
		const int nlines = axis == 0 ? nx : ny;
		const int len = axis == 0 ? ny : nx;
		const int stride = axis == 0 ? nx : 1;
		const int lstep = axis == 0 ? 1 : nx;
		for (int l = tid; l < nlines; l += nt) {
			double* p = tmp + (size_t)l * lstep;
			if (len > 1) {
				for (int i = 0; i < len; ++i) p[i * stride] *= gain;
				// _init_causal_reflect
				const double zn = pow(z, (double)len);
				const double c0 = p[0];
				double zi = z;
				double acc = p[0] + zn * p[(len - 1) * stride];
				for (int i = 1; i < len; ++i) {
					acc += zi * (p[i * stride] + zn * p[(len - 1 - i) * stride]);
					zi *= z;
				}
				acc *= z / (1.0 - zn * zn);
				acc += c0;
				p[0] = acc;
				for (int i = 1; i < len; ++i) p[i * stride] += z * p[(i - 1) * stride];
				// _init_anticausal_reflect
				p[(len - 1) * stride] *= z / (z - 1.0);
				for (int i = len - 2; i >= 0; --i) p[i * stride] = z * (p[(i + 1) * stride] - p[i * stride]);
			}
		}
		__syncthreads();
	}
	double* coef = ws.coef + (size_t)b * nt_tiles;
	for (int t = tid; t < nt_tiles; t += nt) coef[t] = tmp[t];
}

// ---------------------------------------------------------------------------------------------
// K_final: bkg = img_bkg_radial + img_bkg_square (backgrounds.py:209), float32.
// The zoom is separable: the row part  R[row][B] = sum_a wy[row][a] c[oy + a][B]  is the same for every pixel of a
// mesh row, so it is computed once per mesh (64 x 5 values) and staged; a pixel then costs 4 DFMA + add + convert.
// One CTA writes a strip of TBK_FINAL_NM horizontally adjacent meshes: the coefficient loads, the weight table and
// the two barriers of the prologue are paid once per strip (the kernel is bound by that latency and by the store
// stream, not by arithmetic).
// Mesh-uniform specialisations: the clip to [mesh_min, mesh_max] is skipped when the 5x5 coefficient
// neighbourhood already lies inside that range (a cubic B-spline value is a convex combination of its
// coefficients), and the radial term is a constant for every mesh that cannot see beyond the first ring centre.
#define TBK_FINAL_NM 4
struct FinalSmem {
	double blk[5][TBK_FINAL_NM + 4];      // coefficient rows ty-2..ty+2, columns tx0-2..tx0+NM+1
	double wT[4][64];                     // zoom weights transposed: wT[tap][phase] (conflict-free per-lane loads)
	double R[TBK_FINAL_NM][2][64][4];     // row part for the left (coefficient columns 0..3) / right (1..4) mesh half
	double rows[TBK_FINAL_NM][TBK_RTAB_ROWS][TBK_RROW];   // Taylor pieces of the radial profile seen by the non-flat meshes
	int need_clip[TBK_FINAL_NM];
};

// RADIAL: 0 = constant radial term (flat mesh), 1 = Taylor-piece table (static piece words, PlanDev::nonflat_uj),
// 2 = spline evaluation from the radius (meshes that see more pieces than the staged table holds: non-default parameters)
template <bool CLIP, int RADIAL>
__device__ __forceinline__ void final_rows(const FinalSmem& z, int m, const RadialSmem2& rs, const PlanDev& P,
	double lo, double hi, double cflat, double ctab0, double ctab1, int j0, int j1, float* __restrict__ bkg, size_t img, int ty, int tx, int tid)
{
	const int lcol = tile_lcol(tid), ox = lcol >> 5;
	const int gx = tx * TBK_TILE + lcol;
	double wx[4][4];
#pragma unroll
	for (int b = 0; b < 4; ++b) {
		const double2 u0 = *reinterpret_cast<const double2*>(&z.wT[b][lcol]);
		const double2 u1 = *reinterpret_cast<const double2*>(&z.wT[b][lcol + 2]);
		wx[0][b] = u0.x; wx[1][b] = u0.y; wx[2][b] = u1.x; wx[3][b] = u1.y;
	}
	const int slot = RADIAL ? P.tile_slot[ty * P.nx + tx] : 0;
	const int jlo = RADIAL == 1 ? __ldg(P.nonflat_jlo + slot) : 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		const int lrow = tile_lrow(tid, j);
		const int gy = ty * TBK_TILE + lrow;
		const double2 ra = *reinterpret_cast<const double2*>(&z.R[m][ox][lrow][0]);
		const double2 rb = *reinterpret_cast<const double2*>(&z.R[m][ox][lrow][2]);
		const double r[4] = {ra.x, ra.y, rb.x, rb.y};
		float o[4];
		double rr[4] = {0.0, 0.0, 0.0, 0.0};
		if (RADIAL) {
			const double* rp = (RADIAL == 1 ? P.nonflat_uj : P.nonflat_r) + (size_t)slot * TBK_NPIX_TILE + lrow * TBK_TILE + lcol;
			const double2 r01 = __ldg(reinterpret_cast<const double2*>(rp)), r23 = __ldg(reinterpret_cast<const double2*>(rp + 2));
			rr[0] = r01.x; rr[1] = r01.y; rr[2] = r23.x; rr[3] = r23.y;
		}
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			double sq = wx[q][0] * r[0] + wx[q][1] * r[1] + wx[q][2] * r[2] + wx[q][3] * r[3];
			if (CLIP) sq = clamp_d(sq, lo, hi);
			double rad = cflat;
			if (RADIAL == 1) rad = radial_tab_eval_uj(&z.rows[m][0][0], jlo, j0, j1, ctab0, ctab1, rr[q]);
			if (RADIAL == 2) rad = radial_value_s(rs, rr[q]);
			o[q] = (float)(rad + sq);
		}
		*reinterpret_cast<float4*>(bkg + img + (size_t)gy * P.W + gx) = make_float4(o[0], o[1], o[2], o[3]);
	}
}

template <int MINB>
__global__ void __launch_bounds__(TBK_NT, MINB) k_final(PlanDev P, Workspace ws,
	float* __restrict__ bkg, uint8_t* __restrict__ mask_out)
{
	__shared__ __align__(16) FinalSmem z;
	__shared__ RadialSmem2 rs;
	const int nstrip = (P.nx + TBK_FINAL_NM - 1) / TBK_FINAL_NM;
	const int ty = blockIdx.x / nstrip, tx0 = (blockIdx.x % nstrip) * TBK_FINAL_NM;
	const int nm = min(TBK_FINAL_NM, P.nx - tx0);
	const int b = blockIdx.y, tid = threadIdx.x;
	const FfiCtl& c = ws.ctl[b];
	const size_t img = (size_t)b * P.H * P.W;
	const int lcol = tile_lcol(tid);
	if (c.all_masked || c.no_good_mesh) {
		const float q = nan_f();
		for (int m = 0; m < nm; ++m) {
			const int gx = (tx0 + m) * TBK_TILE + lcol;
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const size_t off = img + (size_t)(ty * TBK_TILE + tile_lrow(tid, j)) * P.W + gx;
				*reinterpret_cast<float4*>(bkg + off) = make_float4(q, q, q, q);
				if (c.all_masked) *reinterpret_cast<uchar4*>(mask_out + off) = make_uchar4(1, 1, 1, 1);
			}
		}
		return;
	}
	const double* coef = ws.coef + (size_t)b * P.ntiles;
	if (tid < 5 * (TBK_FINAL_NM + 4)) {
		const int a = tid / (TBK_FINAL_NM + 4), bb = tid % (TBK_FINAL_NM + 4);
		z.blk[a][bb] = coef[reflect_fold(ty - 2 + a, P.ny) * P.nx + reflect_fold(tx0 - 2 + bb, P.nx)];
	}
	z.wT[tid & 3][tid >> 2] = __ldg(P.zoom_w + tid);
	const bool radial = P.use_radial && c.radial_ok;
	// radial term of the strip's meshes: 0 flat, 1 table, 2 spline
	int rmode[TBK_FINAL_NM];
	bool any_spline = false, any_tab = false;
#pragma unroll
	for (int m = 0; m < TBK_FINAL_NM; ++m) {
		rmode[m] = 0;
		if (radial && m < nm) {
			const int slot = P.tile_slot[ty * P.nx + tx0 + m];
			if (slot >= 0) rmode[m] = __ldg(P.nonflat_jlo + slot) >= 0 ? 1 : 2;
		}
		any_spline |= rmode[m] == 2; any_tab |= rmode[m] == 1;
	}
	if (any_spline) radial_stage(rs, c, P);
	double ctab0 = 0.0, ctab1 = 0.0;
	int j0 = 0, j1 = 0;
	if (any_tab) {
		const RadialTab rt = radial_tab(ws.rtab + (size_t)b * TBK_RSUB * max(P.nrings - 1, 1) * 8, c, P);
		ctab0 = radial_tab_eval(rt, rt.x0); ctab1 = radial_tab_eval(rt, rt.xlast);
		j0 = radial_tab_j0(rt); j1 = radial_tab_j1(rt);
#pragma unroll
		for (int m = 0; m < TBK_FINAL_NM; ++m) {
			if (rmode[m] != 1) continue;
			const int jlo = __ldg(P.nonflat_jlo + P.tile_slot[ty * P.nx + tx0 + m]);
			const int nrow = min(TBK_RTAB_ROWS, rt.nsub - jlo);
			for (int e = tid; e < nrow * 4; e += TBK_NT)
				reinterpret_cast<double2*>(&z.rows[m][e >> 2][0])[e & 3] = __ldg(reinterpret_cast<const double2*>(rt.rows + 8 * (size_t)jlo) + e);
		}
	}
	const double mesh_min = c.mesh_min, mesh_max = c.mesh_max;
	const int mesh_const = c.mesh_const;
	const double cflat = c.radial_ok ? c.c_flat : 0.0;
	__syncthreads();
	if (tid < 32 * TBK_FINAL_NM) {
		// warp m decides whether mesh m needs the clip
		const int m = tid >> 5, l = tid & 31;
		const double v = l < 25 ? z.blk[l / 5][m + l % 5] : z.blk[0][m];
		double cmin = v, cmax = v;
		for (int o = 16; o > 0; o >>= 1) { cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, o)); cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o)); }
		if (l == 0) {
			// small margin: the interpolated value can leave [cmin, cmax] only by rounding
			const double eps = 1e-12 * fmax(fabs(cmin), fabs(cmax));
			// ptp(mesh) == 0: the clip returns the constant (BkgZoomInterpolator short-circuit)
			z.need_clip[m] = mesh_const || !(cmin - eps >= mesh_min && cmax + eps <= mesh_max);
		}
	}
	for (int e = tid; e < nm * 64 * 5; e += TBK_NT) {
		const int row = e & 63, B = (e >> 6) % 5, m = e / 320, oy = row >> 5;
		double r = 0.0;
#pragma unroll
		for (int a = 0; a < 4; ++a) r = fma(z.wT[a][row], z.blk[oy + a][m + B], r);
		if (B < 4) z.R[m][0][row][B] = r;
		if (B > 0) z.R[m][1][row][B - 1] = r;
	}
	__syncthreads();
#pragma unroll
	for (int m = 0; m < TBK_FINAL_NM; ++m) {
		if (m >= nm) break;
		const int tx = tx0 + m;
#define FINAL_ROWS(CL, RM) final_rows<CL, RM>(z, m, rs, P, mesh_min, mesh_max, cflat, ctab0, ctab1, j0, j1, bkg, img, ty, tx, tid)
		if (z.need_clip[m]) {
			if (rmode[m] == 1) FINAL_ROWS(true, 1);
			else if (rmode[m] == 2) FINAL_ROWS(true, 2);
			else FINAL_ROWS(true, 0);
		} else {
			if (rmode[m] == 1) FINAL_ROWS(false, 1);
			else if (rmode[m] == 2) FINAL_ROWS(false, 2);
			else FINAL_ROWS(false, 0);
		}
#undef FINAL_ROWS
	}
}

// ---------------------------------------------------------------------------------------------
static inline bool launch_ok(const char* what)
{
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { tbk_set_error("%s: %s", what, cudaGetErrorString(e)); return false; }
	return true;
}

// Optional per-kernel-class timing: events are recorded between launches and read back afterwards.
struct FitProf {
	cudaEvent_t ev[128];
	int cls[128];
	int n;
};
static unsigned long long g_launches = 0;
unsigned long long tbk_launch_counter(void) { return g_launches; }

#define LAUNCH(cls_, ...) do { __VA_ARGS__; ++g_launches; if (prof) { cudaEventRecord(prof->ev[prof->n + 1], st); prof->cls[prof->n] = (cls_); ++prof->n; } } while (0)

int tbk_launch_fit(const PlanDev& P, const Workspace& ws, const float* cube, int B,
	const tbk_ffi_meta* meta, const uint8_t* extra, float* bkg, uint8_t* mask,
	tbk_ffi_status* status, cudaStream_t st, float* prof_ms, int tile_kernel, const TbkSide* side)
{
	FitProf* prof = nullptr;
	FitProf pstore;
	if (prof_ms) {
		prof = &pstore; prof->n = 0;
		for (int i = 0; i < 128; ++i) cudaEventCreate(&prof->ev[i]);
		cudaEventRecord(prof->ev[0], st);
	}
	const dim3 gt(P.ntiles, B);
	const int gb = (B + 127) / 128;
	bool base_forked = false;
	LAUNCH(TBK_K_MISC, (k_init_ctl<<<gb, 128, 0, st>>>(P, ws, meta, B)));
	if (tile_kernel == 0) LAUNCH(TBK_K_TILE_BASE, (k_tile_base<<<gt, TBK_NT, 0, st>>>(P, ws, cube, extra, mask)));
	else if (tile_kernel == 3 && extra) LAUNCH(TBK_K_TILE_BASE, (k_tile_base_w3<true, 2><<<gt, 64, 0, st>>>(P, ws, cube, extra, mask)));
	else if (tile_kernel == 3) LAUNCH(TBK_K_TILE_BASE, (k_tile_base_w3<false, 2><<<gt, 64, 0, st>>>(P, ws, cube, extra, mask)));
	else {
		const dim3 gz((P.ntiles + ZB_WARPS - 1) / ZB_WARPS, B);
		const int rt_cap = std::min(B * P.ntiles, std::max(B * P.ntiles / 8, 64));   // retry queue: up to an eighth of the meshes
		if (tile_kernel == 7) {
			// measured alternative: mesh staged in shared memory by TMA bulk copies
			static bool attr_set = false;
			if (!attr_set) {
				cudaFuncSetAttribute(k_tile_base_z<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZbStage));
				cudaFuncSetAttribute(k_tile_base_z<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZbStage));
				attr_set = true;
			}
			if (extra) LAUNCH(TBK_K_TILE_BASE, (k_tile_base_z<true, true, false><<<gz, 32 * ZB_WARPS, sizeof(ZbStage), st>>>(P, ws, cube, extra, mask, rt_cap)));
			else LAUNCH(TBK_K_TILE_BASE, (k_tile_base_z<false, true, false><<<gz, 32 * ZB_WARPS, sizeof(ZbStage), st>>>(P, ws, cube, extra, mask, rt_cap)));
		}
		else if (extra) LAUNCH(TBK_K_TILE_BASE, (k_tile_base_z<true, false, false><<<gz, 32 * ZB_WARPS, 0, st>>>(P, ws, cube, extra, mask, rt_cap)));
		else LAUNCH(TBK_K_TILE_BASE, (k_tile_base_z<false, false, false><<<gz, 32 * ZB_WARPS, 0, st>>>(P, ws, cube, extra, mask, rt_cap)));
		// the queued meshes (bucketed statistics) are only needed by k_mesh_finalize: side stream, joined before round 0's
		cudaStream_t fs = st;
		if (side && !prof) { fs = side->stream; cudaEventRecord(side->fork, st); cudaStreamWaitEvent(fs, side->fork, 0); base_forked = true; }
		{
			// second run of the zone kernel for the meshes whose first plan put a clip bound inside the bulk: one warp per queue
			// entry, warps beyond the queue length leave at once; the queue holds up to an eighth of the meshes, more go to the
			// bucketed path
			const dim3 gr(rt_cap, 1);
			if (gr.x > 0) {
				if (extra) LAUNCH(TBK_K_FALLBACK, (k_tile_base_z<true, false, true><<<gr, 32, 0, fs>>>(P, ws, cube, extra, mask, rt_cap)));
				else LAUNCH(TBK_K_FALLBACK, (k_tile_base_z<false, false, true><<<gr, 32, 0, fs>>>(P, ws, cube, extra, mask, rt_cap)));
			}
		}
		if (extra) LAUNCH(TBK_K_FALLBACK, (k_tile_base_fb<true><<<1184, 64, 0, fs>>>(P, ws, cube, extra)));
		else LAUNCH(TBK_K_FALLBACK, (k_tile_base_fb<false><<<1184, 64, 0, fs>>>(P, ws, cube, extra)));
		if (base_forked) cudaEventRecord(side->join, fs);
	}
	LAUNCH(TBK_K_MISC, (k_post_base<<<gb, 128, 0, st>>>(P, ws, status, B)));
	if (!launch_ok("base")) return TBK_ERR_CUDA;
	const size_t mesh_smem = mesh_finalize_smem(P.ntiles);
	for (int round = 0; round < P.bkgiters; ++round) {
		if (P.use_radial) {
			if (round > 0) {
				if (tile_kernel == 0) LAUNCH(TBK_K_ZP_MIN, (k_zp_min<<<gt, TBK_NT, 0, st>>>(P, ws, cube, mask)));
				else {
					const dim3 gz((P.ntiles + ZP_WARPS - 1) / ZP_WARPS, B);
					LAUNCH(TBK_K_ZP_MIN, (k_zp_bound<<<gz, 32 * ZP_WARPS, 0, st>>>(P, ws)));
					LAUNCH(TBK_K_ZP_MIN, (k_zp_exact<<<gz, 32 * ZP_WARPS, 0, st>>>(P, ws, cube, mask)));
				}
				LAUNCH(TBK_K_MISC, (k_set_zp<<<gb, 128, 0, st>>>(ws, B)));
			}
			if (tile_kernel == 0) LAUNCH(TBK_K_RING_GATHER, (k_ring_gather<<<dim3((P.nringpix + 255) / 256, B), 256, 0, st>>>(P, ws, cube, mask, round)));
			else if (tile_kernel == 3) LAUNCH(TBK_K_RING_GATHER, (k_ring_gather_t<<<dim3(P.n_ringtiles, B), 256, 0, st>>>(P, ws, cube, mask, round)));
			else LAUNCH(TBK_K_RING_GATHER, (k_ring_gather_d<<<dim3(P.n_ringtiles, B), 256, 0, st>>>(P, ws, cube, mask, round)));
			LAUNCH(TBK_K_RING_KDE, (k_ring_kde<<<dim3(B, P.nrings), TBK_KDE_NT, sizeof(KdeSmem), st>>>(P, ws)));
			LAUNCH(TBK_K_RADIAL_FIT, (k_radial_fit<<<B, 32, 0, st>>>(P, ws, status, round, B)));
			if (P.n_nonflat > 0) {
				if (tile_kernel == 0) LAUNCH(TBK_K_TILE_ROUND, (k_tile_round<<<dim3(P.n_nonflat, B), TBK_NT, 0, st>>>(P, ws, cube, mask)));
				else if (tile_kernel == 3) LAUNCH(TBK_K_TILE_ROUND, (k_tile_round_w<false><<<dim3(P.n_nonflat, B), 128, 0, st>>>(P, ws, cube, mask, round)));
				else {
					LAUNCH(TBK_K_TILE_ROUND, (k_tile_round_z<<<dim3(P.n_nonflat, B), 128, 0, st>>>(P, ws, cube, mask, round)));
					LAUNCH(TBK_K_TILE_ROUND, (k_tile_round_fin<<<dim3((P.n_nonflat + ZF_WARPS - 1) / ZF_WARPS, B), 32 * ZF_WARPS, 0, st>>>(P, ws, round)));
					// the queued meshes are only needed by this round's k_mesh_finalize: side stream, joined there
					cudaStream_t fs = st;
					if (side && !prof) { fs = side->stream; cudaEventRecord(side->fork, st); cudaStreamWaitEvent(fs, side->fork, 0); base_forked = true; }
					LAUNCH(TBK_K_FALLBACK, (k_tile_round_w<true><<<740, 128, 0, fs>>>(P, ws, cube, mask, round)));
					if (fs != st) cudaEventRecord(side->join, fs);
				}
			}
		}
		if (base_forked) { cudaStreamWaitEvent(st, side->join, 0); base_forked = false; }
		LAUNCH(TBK_K_MESH, (k_mesh_finalize<<<B, 1024, mesh_smem, st>>>(P, ws, status, round)));
		if (!launch_ok("round")) return TBK_ERR_CUDA;
	}
	{
		static const int final_minb = getenv("TBK_FINAL_MINB") ? atoi(getenv("TBK_FINAL_MINB")) : 3;
		const dim3 gf(P.ny * ((P.nx + TBK_FINAL_NM - 1) / TBK_FINAL_NM), B);
		if (final_minb != 4) LAUNCH(TBK_K_FINAL, (k_final<3><<<gf, TBK_NT, 0, st>>>(P, ws, bkg, mask)));
		else LAUNCH(TBK_K_FINAL, (k_final<4><<<gf, TBK_NT, 0, st>>>(P, ws, bkg, mask)));
	}
	if (!launch_ok("final")) return TBK_ERR_CUDA;
	if (prof) {
		cudaError_t e = cudaStreamSynchronize(st);
		for (int i = 0; i < TBK_K_COUNT; ++i) prof_ms[i] = 0.f;
		for (int i = 0; i < prof->n; ++i) {
			float ms = 0.f;
			cudaEventElapsedTime(&ms, prof->ev[i], prof->ev[i + 1]);
			prof_ms[prof->cls[i]] += ms;
		}
		for (int i = 0; i < 128; ++i) cudaEventDestroy(prof->ev[i]);
		if (e != cudaSuccess) { tbk_set_error("profiled fit: %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	}
	return TBK_OK;
}

int tbk_fit_configure(void)
{
	cudaError_t e = cudaFuncSetAttribute(k_ring_kde, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KdeSmem));
	if (e != cudaSuccess) { tbk_set_error("cudaFuncSetAttribute(k_ring_kde): %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	e = cudaFuncSetAttribute(k_mesh_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mesh_finalize_smem(4096));
	if (e != cudaSuccess) { tbk_set_error("cudaFuncSetAttribute(k_mesh_finalize): %s", cudaGetErrorString(e)); return TBK_ERR_CUDA; }
	return TBK_OK;
}
