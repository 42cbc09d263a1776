"""
CPU oracle: the stamp cut-outs of ``BasePhotometry._load_cube`` (photometry/BasePhotometry.py:720-751, datasource 'ffi',
no full-cube cache), with a list of per-cadence frames standing in for the HDF5 group.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).
"""
import numpy as np


def load_cube(frames, stamp, pixel_offset_row=0, pixel_offset_col=44):
	"""
	``cube[:, :, k] = hdf[group + '/%04d' % k][ir1:ir2, ic1:ic2]`` with ``ir = stamp - pixel_offset`` (:725-735);
	``frames`` is None when the group does not exist -> NaN cube (:736-737).  float32 like the reference (:733);
	the pixel-flags cube keeps the dtype of its frames (``pixelflags_cube``, :862-877).
	"""
	ir1, ir2 = stamp[0] - pixel_offset_row, stamp[1] - pixel_offset_row
	ic1, ic2 = stamp[2] - pixel_offset_col, stamp[3] - pixel_offset_col
	if frames is None:
		raise ValueError("number of cadences unknown without frames; pass an empty list with ntimes")
	ntimes = len(frames)
	dtype = 'float32' if np.asarray(frames[0]).dtype.kind == 'f' else np.asarray(frames[0]).dtype
	cube = np.empty((ir2 - ir1, ic2 - ic1, ntimes), dtype=dtype)
	for k in range(ntimes):
		cube[:, :, k] = np.asarray(frames[k])[ir1:ir2, ic1:ic2]
	return cube
