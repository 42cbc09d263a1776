/*
 * tbk.h -- C ABI of the B200-native TASOC sky-background hot path ("tbk" = TESS background kernels).
 *
 * The reference (tasoc/photometry, pure Python) has no FFI for this path; its boundary is the
 * Python call ``photometry.backgrounds.fit_background(image, ...) -> (bkg, mask)``
 * (photometry/backgrounds.py:52-53,211) mapped over a process pool by
 * ``photometry.prepare.prepare_photometry`` (photometry/prepare.py:278-291), followed by the
 * time-smoothing loop (prepare.py:309-338) and the sumimage loop (prepare.py:347-470).
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 *
 * Conventions: every function returns 0 on success or a negative tbk_status; tbk_last_error()
 * returns a thread-local message.  No exceptions, no torch types.  All image pointers are DEVICE
 * pointers owned by the caller; every call is asynchronous on the given CUDA stream
 * (``stream`` is a ``cudaStream_t`` passed as ``void*``; NULL = default stream).
 */
#ifndef TBK_H
#define TBK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBK_VERSION 1
#define TBK_TILE 64           /* Background2D box size, backgrounds.py:200 */
#define TBK_MAX_ROUNDS 8      /* upper limit for bkgiters */
#define TBK_MAX_RINGS 128     /* upper limit for the number of radial rings */

typedef enum {
	TBK_OK = 0,
	TBK_ERR_INVALID = -1,     /* bad argument (reference: ValueError) */
	TBK_ERR_CUDA = -2,        /* CUDA runtime error */
	TBK_ERR_NOMEM = -3
} tbk_status;

/* Per-FFI header scalars that drive the manual excludes (pixel_flags.py:34-56) and the sumimage
 * gate (prepare.py:396-405,450). */
typedef struct {
	int32_t cadenceno;        /* FFIINDEX; INT32_MAX when absent (reference default: inf) */
	int32_t dquality;         /* DQUALITY */
	int32_t backapp;          /* BACKAPP: background already subtracted (prepare.py:419) */
	int32_t reserved;
	double  tstart;           /* TSTART */
	double  tstop;            /* TSTOP */
} tbk_ffi_meta;

/* Per-FFI diagnostics written by tbk_fit_batch (device memory, caller-owned). */
typedef struct {
	int32_t all_masked;                   /* backgrounds.py:101-102 early-out: bkg = NaN */
	int32_t no_good_mesh;                 /* photutils would raise ValueError: bkg = NaN */
	int32_t n_valid;                      /* number of unmasked pixels */
	int32_t rounds;                       /* rounds actually run */
	int32_t n_excluded[TBK_MAX_ROUNDS];   /* meshes filled by IDW per round */
	int32_t n_ring_valid[TBK_MAX_ROUNDS]; /* finite ring values after smoothing per round */
	int32_t radial_ok[TBK_MAX_ROUNDS];    /* 1 when the spline was built (backgrounds.py:188-191) */
	double  zeropoint[TBK_MAX_ROUNDS];    /* backgrounds.py:171 */
} tbk_ffi_status;

typedef struct tbk_plan tbk_plan;

/*
 * Plan = everything static for one (image shape, camera, ccd, parameter set): ring membership
 * tables (backgrounds.py:145-154), per-tile geometry, interpolation weights.
 *   H, W            image shape, multiples of TBK_TILE
 *   is_tess         0: ndarray input semantics (no radial component, one round; backgrounds.py:155-157)
 *   camera, ccd     1..4 when is_tess (unknown pair -> TBK_ERR_INVALID, backgrounds.py:139-140)
 *   xycen_override  NULL, or {x, y} replacing the camera-centre table (test hook)
 *   other args      the keyword arguments of fit_background (backgrounds.py:52-53)
 * Development switches, read from the environment when the plan is created (all variants give the same statistics and are
 * cross-checked by tests/test_gpu_parity.py::test_kernel_variants_agree):
 *   TBK_TILE_KERNEL = 6 (default) zone kernels, 0 CTA-per-mesh histogram kernels, 3 bucketed warp kernels,
 *                     7 zone kernels with the raw mesh staged by TMA bulk copies (the measured alternative, slower)
 *   TBK_FINAL_MINB  = 3 (default) | 4 resident CTAs per SM the final kernel is compiled for
 */
int tbk_plan_create(tbk_plan** plan, int H, int W, int is_tess, int camera, int ccd,
	double flux_cutoff, int bkgiters, double radial_cutoff, double radial_pixel_step,
	int radial_smooth, const double* xycen_override, int device);
int tbk_plan_destroy(tbk_plan* plan);

/* Number of rings / ring pixels of a plan (0 when the radial component is off). */
int tbk_plan_num_rings(const tbk_plan* plan);

/* Bytes of device scratch tbk_fit_batch needs for a batch of B FFIs. */
size_t tbk_workspace_bytes(const tbk_plan* plan, int B);

/*
 * fit_background over a device-resident batch (backgrounds.py:86-211 per FFI; the pool loop of
 * prepare.py:291).
 *   cube        const float [B, H, W]   science pixels (FFIImage.data, io.py:47)
 *   meta        const tbk_ffi_meta [B]  (device)
 *   extra_mask  const uint8 [B, H, W] or NULL; non-zero = masked (star-mask extension, OR-ed at
 *               the point of backgrounds.py:90)
 *   bkg_out     float [B, H, W]         img_bkg_radial + img_bkg_square, rounded to float32 (the
 *                                       reference casts to float32 at prepare.py:327-330)
 *   mask_out    uint8 [B, H, W]         1 = pixel not used (the returned ``mask``; stored as
 *                                       PixelQualityFlags.NotUsedForBackground, prepare.py:299)
 *   status      tbk_ffi_status [B]      (device) or NULL
 *   workspace   >= tbk_workspace_bytes(plan, B) bytes (device)
 */
int tbk_fit_batch(tbk_plan* plan, const float* cube, int B, const tbk_ffi_meta* meta,
	const uint8_t* extra_mask, float* bkg_out, uint8_t* mask_out, tbk_ffi_status* status,
	void* workspace, void* stream);

/*
 * Background time smoothing (prepare.py:317-335): out[k] = float32 nan-mean of
 * bkg[max(k-w,0) .. min(k+w, N-1)] accumulated in index order.  The local shard holds cadences
 * [0, n); halo_lo / halo_hi are the n_lo / n_hi (<= w) neighbouring frames owned by the previous /
 * next shard (NULL / 0 at the ends of the sector).
 */
int tbk_time_smooth(tbk_plan* plan, const float* bkg, int n, int w,
	const float* halo_lo, int n_lo, const float* halo_hi, int n_hi,
	float* bkg_smooth, void* stream);

/*
 * Final per-image loop (prepare.py:408-456) fused over the cadence axis:
 *   flags[k]  |= ManualExclude where pixel_manual_exclude (prepare.py:408-410)
 *   flux[k]    = cube[k] - bkg_smooth[k] (unless backapp), NaN where ManualExclude
 *   if (dquality & 4335) == 0: nimg += isfinite(flux); sum += nan->0(flux)   (prepare.py:450-453)
 *   used      += (flags & NotUsedForBackground) == 0                          (prepare.py:456)
 * flux_out may be NULL.  sum / nimg / used are accumulated into (caller zeroes them first).
 */
int tbk_sum_accumulate(tbk_plan* plan, const float* cube, const float* bkg_smooth,
	uint8_t* flags, const tbk_ffi_meta* meta, int n, float* flux_out,
	double* sum, int32_t* nimg, int32_t* used, void* stream);

/*
 * Transport format of the mask for host-resident results (the device-to-host link bounds the end-to-end path and the mask
 * is a fifth of the result bytes): tbk_pack_mask turns nbytes mask bytes (device, nbytes % 32 == 0) into nbytes / 8 bytes
 * of bits, bit (7 - j) of byte i = mask[8 i + j] != 0 (the order of numpy.packbits); tbk_unpack_mask_host expands
 * nbits_bytes such bytes into 0 / 1 bytes on the calling CPU thread (host pointers).
 */
int tbk_pack_mask(const uint8_t* mask, size_t nbytes, uint8_t* bits, void* stream);
int tbk_unpack_mask_host(const uint8_t* bits, size_t nbits_bytes, uint8_t* out);

/* prepare.py:459,468: sumimage = sum / nimg (0/0 -> NaN); pixels_used = used / numfiles > threshold. */
int tbk_sum_finalize(tbk_plan* plan, const double* sum, const int32_t* nimg, const int32_t* used,
	int numfiles, double threshold, double* sumimage, uint8_t* pixels_used, void* stream);

/*
 * Diagnostics of the most recent tbk_fit_batch on this plan (valid after the stream has been
 * synchronised).  Copies, for FFI ``b`` and round ``round``: the ring values after smoothing
 * (s2[nrings]), and the filtered low-resolution mesh (mesh[(H/64)*(W/64)]).  Either output may be NULL.
 */
int tbk_debug_fetch(tbk_plan* plan, const void* workspace, int B, int b, int round,
	double* s2, double* mesh);

/*
 * Ingest helper (io.py:46-48): decode B raw FITS image HDUs (big-endian float32, naxis2 rows of naxis1 pixels,
 * contiguous per FFI, already on the device) into the science crop ``[row0:row0+H, col0:col0+W]`` as native
 * float32 [B, H, W].  For TESS FFIs naxis1 = 2136, naxis2 = 2078, row0 = 0, col0 = 44, H = W = 2048.
 */
int tbk_decode_ffi_be(const uint8_t* raw, int B, int naxis1, int naxis2, int row0, int col0, int H, int W,
	float* cube_out, void* stream);

/* Kernel classes reported by tbk_fit_batch_profiled (milliseconds per class, summed over rounds). */
enum {
	TBK_K_TILE_BASE = 0,   /* mask build + sigma-clipped statistics of every mesh */
	TBK_K_TILE_ROUND,      /* sigma-clipped statistics of the meshes that see the radial gradient */
	TBK_K_ZP_MIN,          /* zeropoint pass of rounds >= 2 */
	TBK_K_RING_GATHER,     /* ring sample gather + log10 */
	TBK_K_RING_KDE,        /* KDE mode per ring */
	TBK_K_RADIAL_FIT,      /* moving median + spline */
	TBK_K_MESH,            /* mesh estimator, IDW, 3x3 median, prefilter */
	TBK_K_FINAL,           /* mesh-to-pixel interpolation + radial + write */
	TBK_K_FALLBACK,        /* bucketed statistics of the meshes the zone kernels queued (a side stream outside the profiled call) */
	TBK_K_MISC,            /* per-FFI bookkeeping kernels */
	TBK_K_COUNT
};

/* Same as tbk_fit_batch, but synchronises the stream and returns the device time of every kernel
 * class in ms[TBK_K_COUNT] (CUDA events between launches).  For measurement only. */
int tbk_fit_batch_profiled(tbk_plan* plan, const float* cube, int B, const tbk_ffi_meta* meta,
	const uint8_t* extra_mask, float* bkg_out, uint8_t* mask_out, tbk_ffi_status* status,
	void* workspace, void* stream, float* ms);

/*
 * Background-shenanigans detection (photometry/pixel_flags.py:61-79, photometry/prepare.py:514-622).  All pointers
 * are device pointers; no plan is needed.
 *
 * tbk_bkgshe_indicator: ind_out[k] = float32(median_filter(images[k] - sumimage, size=15, mode='reflect'))
 *   (pixel_flags.py:74-77; the float32 cast is the dtype of ``pixel_flags_individual``, prepare.py:537).
 *   sumimage may be NULL (SumImage=None).  Windows containing NaN give the median of their non-NaN values.
 * tbk_bkgshe_mean: mean[p] = (1 / ceil(n / 25)) * sum over blocks of 25 cadences, taken in ``order`` (int32[n], the
 *   shuffled indices of prepare.py:561-563), of nanmedian(block) with NaN -> 0 (prepare.py:556-576).  ``ind`` holds
 *   n images of ``npix`` pixels, ``stride`` elements apart (so a slab of rows can be reduced on its own).
 * tbk_bkgshe_flag: flags[k][p] = (flags[k][p] & ~bit) | (abs(ind[k][p] - mean[p]) > threshold ? bit : 0)
 *   (prepare.py:603-612; bit = PixelQualityFlags.BackgroundShenanigans = 4).
 */
int tbk_bkgshe_indicator(const float* images, const double* sumimage, int B, int H, int W, float* ind_out, void* stream);
int tbk_bkgshe_mean(const float* ind, size_t stride, size_t npix, int n, const int32_t* order, double* mean_out,
	void* stream);
int tbk_bkgshe_flag(const float* ind, const double* mean, int B, size_t npix, double threshold, int bit,
	uint8_t* flags, void* stream);

/*
 * Consumer side (photometry/BasePhotometry.py:720-751, _load_cube): cut the stamps of S targets out of a device-resident
 * stack [N, H, W] of 4-byte (images, images_err, backgrounds: float32) or 1-byte (pixel_flags) elements.
 * stamps: int32 [S][4] = (row0, row1, col0, col1) in stack coordinates (the caller subtracts PIXEL_OFFSET_ROW / _COLUMN),
 * half-open, inside the frame; out_offsets: int64 [S] element offsets into ``out``.  Stamp s is written as the
 * C-ordered array [row1-row0][col1-col0][N] (rows, cols, times) the reference builds.  All pointers are device pointers.
 */
int tbk_gather_stamps(const void* stack, int elem_bytes, int N, int H, int W, const int32_t* stamps,
	const int64_t* out_offsets, int S, void* out, void* stream);

/*
 * Catalog-driven star mask -- an EXTENSION: the reference accepts a ``catalog`` argument but does not use it yet
 * (photometry/backgrounds.py:64-65, TODO at :90).  stars: double [S][3] = (column, row, radius in pixels) in science-pixel
 * coordinates (device); every pixel with (x - column)^2 + (y - row)^2 <= radius^2 is set to 1 in ``mask`` (uint8 [H, W],
 * device, OR-ed into -- the caller zeroes it first).  The result is what tbk_fit_batch takes as ``extra_mask``.
 */
int tbk_star_mask(const double* stars, int S, int H, int W, uint8_t* mask, void* stream);

/*
 * Image movement kernels (photometry/image_motion.py:74-111 ``_prepare_flux``, :182-258 ``calc_kernel`` with
 * warpmode='translation'; driver photometry/prepare.py:678-698).  All pointers are device pointers; no plan is needed.
 *
 * tbk_motion_prepare: prepared[k] = float32(scharr(-1 + 2 (L - min L) / |max L - min L|)) with L = log10(images[k] -
 *   nanmin(images[k]) + 1), NaN -> 0 (skimage 0.19 ``scharr``: 3 x 3, mode='reflect').  scratch: 16 bytes per frame.
 * tbk_motion_ecc: ``cv2.findTransformECC(ref_prepared, prepared[k], eye(2, 3), MOTION_TRANSLATION, (COUNT | EPS, max_iter,
 *   eps), mask of ones, gaussFiltSize = 5)`` for every frame of the batch: out[k] = {dx, dy, rho, iterations}; dx = dy = NaN
 *   where OpenCV raises (the reference logs and stores NaN, image_motion.py:239-241).  The call synchronises the stream
 *   every few iterations to learn whether every frame has met its criterion.
 *   workspace: tbk_motion_workspace_bytes(B, H, W) bytes.
 */
int tbk_motion_prepare(const float* images, int B, int H, int W, float* prepared, void* scratch, void* stream);
size_t tbk_motion_workspace_bytes(int B, int H, int W);
int tbk_motion_ecc(const float* ref_prepared, const float* prepared, int B, int H, int W, int max_iter, double eps,
	void* workspace, double* out, void* stream);

/* Diagnostics: out[i] = the device log10 used for the ring samples (table-driven, see tbk_common.cuh) of in[i];
 * both device pointers.  Lets the tests bound its error against a host log10. */
int tbk_debug_log10(const double* in, double* out, int n, void* stream);

/*
 * Diagnostics, HOST execution, host pointers: the 10 nearest good meshes of every mesh position exactly as the IDW fill of
 * tbk_fit_batch chooses them -- the restatement of scipy.spatial.cKDTree(yx_good, leafsize=10).query(k=10) that photutils'
 * ShepardIDWInterpolator runs (photometry/backgrounds.py:200-205), tie order included.  good: uint8 [ny*nx], non-zero = good
 * mesh.  Outputs (any may be NULL): nbr_id / nbr_d2 int32 [ny*nx][10] (mesh id, squared distance; -1 past the number of good
 * meshes), idx_out int32 [ngood] (the tree's index permutation, cKDTree.indices), nodes_out int32 [<= 2 ngood + 2][4] =
 * (split_dim or -1, split, lesser | start, greater | end), nnodes_out int32 [1].  Lets the CPU tests compare with the real scipy.
 */
int tbk_debug_idw_neighbors(const uint8_t* good, int ny, int nx, int32_t* nbr_id, int32_t* nbr_d2,
	int32_t* idx_out, int32_t* nodes_out, int32_t* nnodes_out);

/* Number of kernels this library has launched in this process so far. */
unsigned long long tbk_launch_count(void);

/* Byte offsets of the workspace sections for a batch of B (diagnostics / tests):
 * offsets[0..8] = ctl, tile_base, tile_nf, coef, mesh_hist, s2_raw, s2_hist, ring_v, fallback counters (int32 [64]: meshes
 * the zone kernels handed to the bucketed kernels -- [0] raw-pixel statistics, [1 + round] residual statistics of a round);
 * sizes[0..2] = sizeof(FfiCtl), sizeof(TileStat), number of non-flat tiles. */
int tbk_workspace_layout(const tbk_plan* plan, int B, size_t* offsets, size_t* sizes);

const char* tbk_last_error(void);
int tbk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TBK_H */
