"""
Catalog-driven star masking -- an extension beyond the reference, default off.  ``fit_background(image, catalog=...)`` is
documented "currently not yet being used for anything" (photometry/backgrounds.py:64-65, TODO at :90); BASELINE.json's
north star asks for it, so the rule is defined here and restated identically in the oracle (``oracle.star_mask``):

    a star of TESS magnitude Tmag at science-pixel position (column, row) masks the disc of radius
        r(Tmag) = clip(RADIUS_AT_10 * 10 ** (-SLOPE * (Tmag - 10)), R_MIN, R_MAX)   pixels
    (a 10th-magnitude star ~ 4 px; three magnitudes brighter doubles it; fainter than ~Tmag 14.3 -> R_MIN).

The mask is OR-ed into the validity mask at the point of backgrounds.py:90 through ``extra_mask``.
"""
import ctypes as C
import numpy as np
import torch
from . import _lib

RADIUS_AT_10, SLOPE, R_MIN, R_MAX = 4.0, 0.1, 1.5, 40.0


def star_radius(tmag):
	return np.clip(RADIUS_AT_10 * 10.0 ** (-SLOPE * (np.asarray(tmag, dtype='float64') - 10.0)), R_MIN, R_MAX)


def star_mask(shape, catalog, device=None):
	"""
	uint8 CUDA tensor [H, W]: 1 where a catalog star's disc covers the pixel.  ``catalog``: array [S, 3] of
	(column, row, Tmag) in science-pixel coordinates (``PIXEL_OFFSET_COLUMN`` already subtracted).
	"""
	if not torch.cuda.is_available():
		raise _lib.TbkError("CUDA device required: photometry_b200 has no CPU fallback")
	lib = _lib.load()
	dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
	H, W = int(shape[0]), int(shape[1])
	mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
	cat = np.asarray(catalog, dtype='float64').reshape(-1, 3)
	if cat.shape[0] == 0:
		return mask
	stars = np.column_stack([cat[:, 0], cat[:, 1], star_radius(cat[:, 2])])
	stars_d = torch.from_numpy(np.ascontiguousarray(stars)).to(dev)
	_lib.check(lib.tbk_star_mask(C.c_void_p(stars_d.data_ptr()), int(stars.shape[0]), H, W, C.c_void_p(mask.data_ptr()),
		C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), 'tbk_star_mask')
	return mask
