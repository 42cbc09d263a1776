"""
Shared definitions of the small parity cases (inputs are regenerated from seeds by
``tests/golden/make_golden.py`` and stored, together with the oracle's outputs, in tests/golden/*.npz).
"""
import numpy as np
from photometry_b200 import synth

TOL_REL = 1e-5   # BASELINE.json north_star: backgrounds within 1e-5 relative ...
TOL_ABS = 1e-3   # ... or 1e-3 e-/s absolute where the background is near zero


def in_tolerance(bkg, ref):
	bkg = np.asarray(bkg, dtype='float64')
	ref = np.asarray(ref, dtype='float64')
	with np.errstate(invalid='ignore'):
		ok = np.abs(bkg - ref) <= np.maximum(TOL_REL * np.abs(ref), TOL_ABS)
	return ok | (np.isnan(bkg) & np.isnan(ref))


def header(camera, ccd, k=0, cadenceno=9000, dquality=0, tstart=1400.0):
	t0 = tstart + k * (1800.0 / 86400.0)
	return dict(CAMERA=camera, CCD=ccd, TSTART=t0, TSTOP=t0 + 1800.0 / 86400.0, FFIINDEX=cadenceno + k, DQUALITY=dquality)


def case_nontess():
	rng = np.random.default_rng(101)
	img = (200 + 20 * rng.standard_normal((256, 320))).astype('float32')
	img += synth.synth_stack_numpy(1, 256, 320, seed=7, xycen=(1e5, 1e5), n_stars=300, noise=False, nan_frac=0, sky_level=0.0)[0]
	img[100:140, 200:260] += 5000
	img[5, 5] = np.nan
	img[6, 6] = np.inf
	img[7, 9] = -3
	img[30, 30] = 9e4
	img[31, 31] = -0.0
	return dict(kind='nontess', images=img[None], fit_kwargs={})


def case_tess_small():
	H, W = 384, 448
	xycen = (-30.0, 420.0)
	kw = dict(radial_cutoff=380, radial_pixel_step=15)
	stack = synth.synth_stack_numpy(2, H, W, seed=5, xycen=xycen, radial_cutoff=380.0, n_stars=700)
	return dict(kind='tess', images=stack, camera=1, ccd=2, xycen=xycen, fit_kwargs=kw,
		headers=[header(1, 2, k) for k in range(2)])


def case_mars():
	# camera 1 / ccd 4 before cadence 4724: columns >= 1536 are manually excluded (pixel_flags.py:44-46)
	# -> 8 mesh columns are excluded and filled by IDW.
	H, W = 256, 2048
	stack = synth.synth_stack_numpy(1, H, W, camera=1, ccd=4, seed=9, n_stars=2500)
	return dict(kind='tess', images=stack, camera=1, ccd=4, xycen=None, fit_kwargs={},
		headers=[header(1, 4, 0, cadenceno=4700, tstart=1330.0)])


def case_crowded():
	# star-mask extension: ~40 % of the pixels masked in blobs so meshes sit around the 50 % exclusion limit
	H, W = 384, 384
	xycen = (-20.0, 430.0)
	kw = dict(radial_cutoff=360, radial_pixel_step=12)
	stack = synth.synth_stack_numpy(1, H, W, seed=13, xycen=xycen, radial_cutoff=360.0, n_stars=6000,
		sky_level=1200.0, gradient=1.2)
	rng = np.random.default_rng(17)
	yy, xx = np.mgrid[0:H, 0:W]
	extra = np.zeros((H, W), dtype=bool)
	for _ in range(85):
		cy, cx, rad = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(6, 26)
		extra |= (yy - cy) ** 2 + (xx - cx) ** 2 < rad * rad
	extra[128:192, 64:128] |= rng.uniform(size=(64, 64)) < 0.55   # one mesh certainly excluded
	return dict(kind='tess', images=stack, camera=1, ccd=2, xycen=xycen, fit_kwargs=kw,
		headers=[header(1, 2, 0)], extra_mask=extra[None])


def case_prepare():
	# 6-cadence stack for time smoothing + sumimage; one cadence flagged by DQUALITY, one all-NaN column
	H, W = 256, 256
	xycen = (-25.0, 300.0)
	kw = dict(radial_cutoff=250, radial_pixel_step=15)
	stack = synth.synth_stack_numpy(6, H, W, seed=21, xycen=xycen, radial_cutoff=250.0, n_stars=300)
	stack[3, :, 17] = np.nan
	hdrs = [header(1, 2, k, dquality=(32 if k == 2 else 0)) for k in range(6)]
	return dict(kind='tess', images=stack, camera=1, ccd=2, xycen=xycen, fit_kwargs=kw, headers=hdrs, time_smooth=3)


CASES = {
	'nontess': case_nontess,
	'tess_small': case_tess_small,
	'mars': case_mars,
	'crowded': case_crowded,
	'prepare': case_prepare,
}


def case_shenanigans():
	"""
	Background-shenanigans stage: 30 background-subtracted frames (one full block of 25 + a partial one, so the stale
	slots of the reference's block buffer matter), a scattered-light blob in three cadences, a residual star field
	(what the median filter is there to remove), NaN-free so every window is specified by the reference.
	"""
	rng = np.random.default_rng(77)
	N, H, W = 30, 96, 112
	imgs = rng.normal(0.0, 6.0, (N, H, W)).astype('float32')
	stars = np.zeros((H, W), dtype='float32')
	for _ in range(60):
		stars[rng.integers(0, H), rng.integers(0, W)] += rng.uniform(50, 4000)
	imgs += stars * rng.uniform(0.97, 1.03, (N, 1, 1)).astype('float32')
	yy, xx = np.mgrid[0:H, 0:W]
	for k, amp in ((4, 120.0), (5, 90.0), (17, -70.0)):
		imgs[k] += (amp * np.exp(-(((yy - 30 - k) / 18.0) ** 2 + ((xx - 70 + k) / 25.0) ** 2))).astype('float32')
	sumimage = (stars + rng.normal(0, 0.5, (H, W))).astype('float64')
	flags = (rng.uniform(size=(N, H, W)) < 0.05).astype('uint8')          # NotUsedForBackground bits
	flags[9] |= 4                                                          # a stale shenanigans flag that must be cleared
	return dict(images=imgs, sumimage=sumimage, pixel_flags=flags)


def load_golden(path):
	"""Load a golden file and unpack the bit-packed masks."""
	g = dict(np.load(path))
	shape = tuple(g['shape'])
	g['mask'] = np.unpackbits(g['mask_bits'])[:int(np.prod(shape))].reshape(shape).astype(bool)
	return g


def images_digest(images):
	import hashlib
	return np.frombuffer(hashlib.sha256(np.ascontiguousarray(images).tobytes()).digest(), dtype='uint8')
