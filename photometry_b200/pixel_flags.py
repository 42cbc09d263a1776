"""
Host mirror of ``photometry.pixel_flags.pixel_manual_exclude`` (photometry/pixel_flags.py:14-58).
The device path evaluates the same rules inside its kernels from the per-FFI header scalars; this
function exists for callers that want the mask on the host and for the parity tests.
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
