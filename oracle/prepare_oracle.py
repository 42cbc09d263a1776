"""
CPU oracle: the hot loops of ``photometry.prepare.prepare_photometry`` around ``fit_background``
(background time-smoothing and the sumimage accumulation), operating on in-memory stacks instead
of HDF5 groups.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative to
/root/reference.
"""
import numpy as np
from .backgrounds_oracle import FFIImageLite, fit_background, pixel_manual_exclude

# photometry/quality.py:123-124 -- TESSQualityFlags.DEFAULT_BITMASK
TESS_DEFAULT_BITMASK = 1 | 2 | 4 | 8 | 32 | 64 | 128 | 4096  # = 4335
# photometry/quality.py:161-166 -- PixelQualityFlags
PIXEL_NOT_USED_FOR_BACKGROUND = 1
PIXEL_MANUAL_EXCLUDE = 2


def _nanmean_f32_sequential(block):
	"""
	bottleneck.nanmean(block, axis=2) for a float32 block: float32 accumulator, values added in
	index order along the axis, divided by the non-NaN count; all-NaN -> NaN.
	"""
	asum = np.zeros(block.shape[:2], dtype='float32')
	count = np.zeros(block.shape[:2], dtype='int32')
	for i in range(block.shape[2]):
		v = block[:, :, i]
		ok = ~np.isnan(v)
		asum = np.where(ok, (asum + v).astype('float32'), asum)
		count += ok
	with np.errstate(invalid='ignore', divide='ignore'):
		out = (asum / count.astype('float32')).astype('float32')
	out[count == 0] = np.nan
	return out


def time_smooth_backgrounds(bkg_unsmoothed, time_smooth):
	"""
	photometry/prepare.py:317-335.  ``bkg_unsmoothed`` is [N, H, W] (float64 in the reference's
	temp file; the cast to float32 happens when the block is filled, prepare.py:327-330).
	Returns float32 [N, H, W].
	"""
	n = bkg_unsmoothed.shape[0]
	w = int(time_smooth) // 2
	out = np.empty(bkg_unsmoothed.shape, dtype='float32')
	for k in range(n):
		indx1 = max(k - w, 0)
		indx2 = min(k + w + 1, n)
		block = np.empty(bkg_unsmoothed.shape[1:] + (indx2 - indx1,), dtype='float32')
		for i, j in enumerate(range(indx1, indx2)):
			block[:, :, i] = bkg_unsmoothed[j]
		out[k] = _nanmean_f32_sequential(block)
	return out


def sumimage_accumulate(images, backgrounds, pixel_flags, quality, backapp=None,
	backgrounds_pixels_threshold=0.5):
	"""
	photometry/prepare.py:347-359, 413-470 for in-memory stacks.

	images      float32 [N, H, W]   science pixels (``img.data``)
	backgrounds float32 [N, H, W]   time-smoothed backgrounds (``backgrounds/NNNN``)
	pixel_flags uint8   [N, H, W]   bit 1 = NotUsedForBackground, bit 2 = ManualExclude
	quality     int32   [N]         DQUALITY per cadence
	backapp     bool    [N] or None header BACKAPP (background already applied)

	Returns dict(flux [N,H,W] f32, sumimage f64, nimg i32, used i32, backgrounds_pixels_used bool).
	"""
	n, H, W = images.shape
	SumImage = np.zeros((H, W), dtype='float64')
	Nimg = np.zeros((H, W), dtype='int32')
	Used = np.zeros((H, W), dtype='int32')
	flux = np.empty((n, H, W), dtype='float32')
	for k in range(n):
		flux0 = np.array(images[k], dtype='float32', copy=True)
		if backapp is None or not backapp[k]:
			flux0 -= backgrounds[k]  # float32 - float32 (prepare.py:419-420)
		excl = (pixel_flags[k] & PIXEL_MANUAL_EXCLUDE) != 0  # ~PixelQualityFlags.filter(...)
		flux0[excl] = np.nan
		flux[k] = flux0
		if (int(quality[k]) & TESS_DEFAULT_BITMASK) == 0:  # TESSQualityFlags.filter
			Nimg += np.isfinite(flux0)
			SumImage += np.where(np.isnan(flux0), np.float32(0), flux0)
		Used += ((pixel_flags[k] & PIXEL_NOT_USED_FOR_BACKGROUND) == 0)
	with np.errstate(invalid='ignore', divide='ignore'):
		SumImage = SumImage / Nimg
	used_bool = (Used / n) > backgrounds_pixels_threshold
	return dict(flux=flux, sumimage=SumImage, nimg=Nimg, used=Used, backgrounds_pixels_used=used_bool)


def prepare_stack(ffis, time_smooth, fit_kwargs=None, backgrounds_pixels_threshold=0.5):
	"""
	The [A] backgrounds + smoothing and [B] final per-image loops of prepare.py:265-470 for a list
	of :class:`FFIImageLite` (one sector/camera/CCD, time ordered).  Returns a dict with the same
	products the reference writes to HDF5.
	"""
	fit_kwargs = dict(fit_kwargs or {})
	n = len(ffis)
	H, W = ffis[0].shape
	bkg_us = np.empty((n, H, W), dtype='float64')
	flags = np.zeros((n, H, W), dtype='uint8')
	for k, img in enumerate(ffis):
		bck, mask = fit_background(img, **fit_kwargs)
		bkg_us[k] = bck
		flags[k] = np.where(mask, PIXEL_NOT_USED_FOR_BACKGROUND, 0).astype('uint8')  # prepare.py:299
	bkg = time_smooth_backgrounds(bkg_us, time_smooth)
	quality = np.array([int(img.header.get('DQUALITY', 0)) for img in ffis], dtype='int32')
	backapp = np.array([bool(img.header.get('BACKAPP', False)) for img in ffis])
	for k, img in enumerate(ffis):
		manexcl = pixel_manual_exclude(img)
		flags[k][manexcl] |= PIXEL_MANUAL_EXCLUDE  # prepare.py:408-410
	images = np.stack([np.asarray(img.data, dtype='float32') for img in ffis])
	out = sumimage_accumulate(images, bkg, flags, quality, backapp, backgrounds_pixels_threshold)
	out.update(backgrounds_unsmoothed=bkg_us, backgrounds=bkg, pixel_flags=flags, quality=quality)
	return out


__all__ = ['FFIImageLite', 'time_smooth_backgrounds', 'sumimage_accumulate', 'prepare_stack']
