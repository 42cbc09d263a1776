// tbk_internal.h -- host-side declarations shared by the translation units of libtbk.
#pragma once
#include <cuda_runtime.h>
#include "../../include/tbk.h"

struct PlanDev;
struct Workspace;

void tbk_set_error(const char* fmt, ...);

// A side stream per caller stream: the fallback kernels of the zone statistics (a per cent of the meshes, latency bound) run
// there, forked / joined with events, while the caller's stream goes on with kernels that do not need their results.
struct TbkSide {
	cudaStream_t stream;
	cudaEvent_t fork, join;
};

// tbk_fit.cu
int tbk_fit_configure(void);
int tbk_launch_fit(const PlanDev& P, const Workspace& ws, const float* cube, int B,
	const tbk_ffi_meta* meta, const uint8_t* extra, float* bkg, uint8_t* mask,
	tbk_ffi_status* status, cudaStream_t st, float* prof_ms, int tile_kernel, const TbkSide* side);
unsigned long long tbk_launch_counter(void);

// tbk_prepare.cu
int tbk_launch_time_smooth(int H, int W, const float* bkg, int n, int w,
	const float* halo_lo, int n_lo, const float* halo_hi, int n_hi, float* out, cudaStream_t st);
int tbk_launch_sum_accumulate(const PlanDev& P, const float* cube, const float* bkg_smooth,
	uint8_t* flags, const tbk_ffi_meta* meta, int n, float* flux_out,
	double* sum, int32_t* nimg, int32_t* used, int* zero_flags, cudaStream_t st);
int tbk_launch_pack_mask(const uint8_t* mask, size_t nbytes, uint8_t* bits, cudaStream_t st);
int tbk_launch_sum_finalize(int H, int W, const double* sum, const int32_t* nimg, const int32_t* used,
	int numfiles, double threshold, double* sumimage, uint8_t* pixels_used, cudaStream_t st);
int tbk_launch_bkgshe_indicator(const float* images, const double* sum, int B, int H, int W, float* out, cudaStream_t st);
int tbk_launch_bkgshe_mean(const float* ind, size_t stride, size_t npix, int n, const int* order, double* mean, cudaStream_t st);
int tbk_launch_bkgshe_flag(const float* ind, const double* mean, int B, size_t npix, double threshold, int bit, uint8_t* flags, cudaStream_t st);
int tbk_launch_gather_stamps(const void* stack, int elem_bytes, int N, int H, int W, const int* stamps,
	const long long* offs, int S, int tiles_x, void* out, cudaStream_t st);
int tbk_launch_log10(const double* in, double* out, int n, cudaStream_t st);
int tbk_launch_decode(const uint8_t* raw, int B, int naxis1, int naxis2, int row0, int col0, int H, int W, float* out, cudaStream_t st);

// tbk_motion.cu
int tbk_launch_motion_prepare(const float* flux, int B, int H, int W, float* out, unsigned* scratch, cudaStream_t st);
size_t tbk_motion_workspace(int B, int H, int W);
int tbk_launch_motion_ecc(const float* ref_prepared, const float* prepared, int B, int H, int W, int max_iter, double eps,
	void* workspace, double* out, cudaStream_t st);
int tbk_launch_star_mask(const double* stars, int S, int H, int W, uint8_t* mask, cudaStream_t st);
