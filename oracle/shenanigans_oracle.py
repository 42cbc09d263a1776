"""
CPU oracle: background-shenanigans detection (photometry/pixel_flags.py:61-79 and the driver in
photometry/prepare.py:514-622), operating on in-memory stacks instead of HDF5 datasets.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative to /root/reference.

Pinned pieces: ``scipy.ndimage.median_filter`` is the real scipy (1.18.1 here, 1.7.3 pinned by the
reference; for windows without NaN the result is the 113th smallest of the 225 values in either),
``np.random.RandomState(0).shuffle`` is the frozen legacy generator behind ``np.random.seed(0);
np.random.shuffle`` (prepare.py:562-563).  ``bottleneck.nanmedian`` / ``replace`` are restated with
numpy (same semantics: even counts average the two middle values, all-NaN -> NaN).

Unspecified in the reference, and how this oracle pins it:
  * windows that contain a NaN: scipy's rank filter selects with ``<`` comparisons, so the value depends on the
    selection internals (observed: sometimes the +inf-ordering answer, sometimes the -inf one, sometimes NaN).
    ``indicator_stack`` returns scipy's value, and ``nan_affected`` marks those windows so tests can leave them out;
    the CUDA path defines them as the median of the non-NaN values.
  * stacks shorter than one block of 25: the reference's block buffer is ``np.empty`` (prepare.py:564) and its
    unused trailing slots enter the median; fresh large allocations are zero pages, so zeros are used here.
"""
import math
import numpy as np
from scipy.ndimage import median_filter, maximum_filter

PIXEL_BACKGROUND_SHENANIGANS = 4   # photometry/quality.py:165
BKGSHE_SIZE = 15                   # pixel_flags.py:77
BKGSHE_BLOCK = 25                  # prepare.py:560
BKGSHE_THRESHOLD = 40              # prepare.py:523


def pixel_background_shenanigans(img, SumImage=None):
	"""pixel_flags.py:61-79: median-filtered (15 x 15, scipy default mode='reflect') difference image, float64."""
	img = np.asarray(img)
	flux0 = (img - SumImage) if SumImage is not None else img
	return median_filter(flux0, size=BKGSHE_SIZE)


def indicator_stack(images, sumimage):
	"""
	prepare.py:531-549: one indicator image per cadence, stored as float32 (``pixel_flags_individual`` dtype).
	``images`` are the background-subtracted float32 frames (``images/NNNN``), ``sumimage`` float64.
	"""
	images = np.asarray(images)
	out = np.empty(images.shape, dtype='float32')
	for k in range(images.shape[0]):
		out[k] = pixel_background_shenanigans(images[k], sumimage).astype('float32')
	return out


def nan_affected(images, sumimage):
	"""True where the 15 x 15 window (reflect boundary) of a pixel contains a NaN of ``images[k] - sumimage``."""
	images = np.asarray(images)
	out = np.empty(images.shape, dtype=bool)
	for k in range(images.shape[0]):
		bad = np.isnan(images[k] - sumimage) if sumimage is not None else np.isnan(images[k])
		out[k] = maximum_filter(bad.astype('uint8'), size=BKGSHE_SIZE, mode='reflect').astype(bool)
	return out


def shuffled_order(numfiles):
	"""prepare.py:561-563: ``indicies = list(range(numfiles)); np.random.seed(0); np.random.shuffle(indicies)``."""
	idx = list(range(numfiles))
	np.random.RandomState(0).shuffle(idx)
	return np.asarray(idx, dtype='int32')


def mean_shenanigans(ind, block=BKGSHE_BLOCK):
	"""
	prepare.py:556-576: robust mean of the indicator images -- median over random blocks of ``block`` images,
	NaN -> 0, averaged over the blocks.  The block buffer is allocated once, so the last (partial) block still
	holds the trailing images of the block before it.
	"""
	ind = np.asarray(ind)
	numfiles, H, W = ind.shape
	order = shuffled_order(numfiles)
	mean = np.zeros((H, W), dtype='float64')
	buf = np.zeros((H, W, block), dtype='float64')
	for k in range(0, numfiles, block):
		for j, i in enumerate(order[k:k + block]):
			buf[:, :, j] = ind[i]
		with np.errstate(all='ignore'):
			import warnings
			with warnings.catch_warnings():
				warnings.simplefilter('ignore', RuntimeWarning)
				med = np.nanmedian(buf, axis=2)
		med[np.isnan(med)] = 0
		mean += med
	mean /= math.ceil(numfiles / block)
	return mean


def flag_shenanigans(ind, mean, pixel_flags, threshold=BKGSHE_THRESHOLD):
	"""
	prepare.py:581-612: clear the old bit, set it where ``abs(indicator - mean) > threshold``; uint8 flags.

	The reference's clearing step reads ``indx = (flags & PixelQualityFlags.BackgroundShenanigans != 0)`` (prepare.py:606).
	In Python ``&`` binds tighter than ``!=``, so this is ``(flags & 4) != 0``: exactly the pixels that carry the old bit,
	from which the bit is then subtracted (``np.array([0, 1, 4, 5, 2]) & 4 != 0 -> [F, F, T, T, F]``).  That is what is
	restated here (clear the bit, then set it where the indicator pops out).
	"""
	ind = np.asarray(ind)
	flags = np.array(pixel_flags, dtype='uint8', copy=True)
	for k in range(ind.shape[0]):
		with np.errstate(invalid='ignore'):
			bad = np.abs(ind[k] - mean) > threshold
		flags[k] &= np.uint8(255 - PIXEL_BACKGROUND_SHENANIGANS)
		flags[k][bad] |= np.uint8(PIXEL_BACKGROUND_SHENANIGANS)
	return flags


def background_shenanigans(images, sumimage, pixel_flags, threshold=BKGSHE_THRESHOLD, block=BKGSHE_BLOCK):
	"""The whole stage: returns (flags uint8 [N,H,W], mean_shenanigans float64 [H,W], indicator float32 [N,H,W])."""
	ind = indicator_stack(images, sumimage)
	mean = mean_shenanigans(ind, block)
	return flag_shenanigans(ind, mean, pixel_flags, threshold), mean, ind
