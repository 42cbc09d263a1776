"""
Host mirror of ``photometry.pixel_flags.pixel_manual_exclude`` (photometry/pixel_flags.py:14-58).
The device path evaluates the same rules inside its kernels from the per-FFI header scalars; this
function exists for callers that want the mask on the host and for the parity tests.

``pixel_background_shenanigans`` is the drop-in for photometry/pixel_flags.py:61-79; it runs on the GPU through
``tbk_bkgshe_indicator`` (there is no CPU fallback).
"""
import numpy as np


def manual_exclude_rule(is_tess, camera, ccd, cadenceno, tstart, tstop):
	"""Return 'mars', 'earth' or None for the header scalars of one FFI."""
	if not is_tess:
		return None
	time = 0.5 * (tstart + tstop)
	if camera == 1 and ccd == 4 and (cadenceno <= 4724 or tstart <= 1325.881282301840):
		return 'mars'
	if camera == 1 and (11354 <= cadenceno <= 11366 or 1464.0158778 <= time <= 1464.265871):
		return 'earth'
	return None


def pixel_manual_exclude(img):
	mask = np.zeros(img.shape, dtype=bool)
	hdr = img.header
	rule = None
	if img.is_tess:
		rule = manual_exclude_rule(True, hdr['CAMERA'], hdr['CCD'], hdr.get('FFIINDEX', np.inf), hdr['TSTART'], hdr['TSTOP'])
	if rule == 'mars':
		mask[:, 1536:] = True
	elif rule == 'earth':
		mask[:, :] = True
	if img.is_tess and np.all(img.data == 0):
		mask[:, :] = True
	return mask


def pixel_background_shenanigans(img, SumImage=None):
	"""
	photometry/pixel_flags.py:61-79: ``median_filter(img - SumImage, size=15)``.  ``img`` is a 2-D array (or an object
	with ``.data``); returns a float64 array holding the float32-rounded values -- the precision the reference stores
	them with (``pixel_flags_individual`` is float32, prepare.py:537).  Windows that contain NaN give the median of
	their non-NaN values (scipy's result for them is unspecified).
	"""
	import torch
	from .shenanigans import shenanigans_indicator
	data = np.asarray(getattr(img, 'data', img) if not isinstance(img, np.ndarray) else img)
	if data.ndim != 2:
		raise ValueError("Input image must be a 2D ndarray.")
	dev = torch.device('cuda')
	d_sum = None
	if SumImage is not None:
		sm = np.asarray(SumImage, dtype='float64')
		if sm.shape != data.shape:
			raise ValueError("SumImage must have the shape of img")
		if data.dtype == np.float32:
			d_sum = torch.from_numpy(np.ascontiguousarray(sm)).to(dev)   # float64(img) - SumImage on the device
		else:
			data = data.astype('float64') - sm                             # other dtypes: difference in float64 here
	# rounding to float32 is monotone, so the median of the rounded values is the rounded median
	d_img = torch.from_numpy(np.ascontiguousarray(data.astype('float32', copy=False)))[None].to(dev)
	return shenanigans_indicator(d_img, d_sum)[0].cpu().numpy().astype('float64')
