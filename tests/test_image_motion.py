"""
Image movement kernels (photometry/image_motion.py:74-111, 182-258).

CPU: the oracle's restatement of ``cv2.findTransformECC`` is pinned against the REAL OpenCV (importable here; the reference
pins opencv-python-headless 4.5.5) -- the warp is bit-equal, the converged translations agree to 1e-5 px.
GPU: ``photometry_b200.image_motion`` against the oracle and against the real cv2 on the same prepared images.
"""
import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import image_motion_oracle as imo

cv2 = pytest.importorskip('cv2')
CRIT = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10000, 1e-6)


def star_field(rng, H, W, nstars=150, sky=80.0):
	img = np.full((H, W), sky)
	yy, xx = np.mgrid[0:H, 0:W]
	for _ in range(nstars):
		y0, x0, f = rng.uniform(0, H), rng.uniform(0, W), 10 ** rng.uniform(2, 4.5)
		img += f * np.exp(-0.5 * ((yy - y0) ** 2 + (xx - x0) ** 2) / 0.9 ** 2) / (2 * np.pi * 0.81)
	return img


def frames(seed, H=160, W=200, n=4):
	rng = np.random.default_rng(seed)
	base = star_field(rng, H, W)
	out = [base + rng.normal(0, 3, base.shape)]
	shifts = [(0.0, 0.0)]
	for _ in range(n - 1):
		sh = rng.uniform(-0.6, 0.6, 2)
		out.append(ndi.shift(base, sh, order=3, mode='nearest') + rng.normal(0, 3, base.shape))
		shifts.append(tuple(sh))
	out = np.stack(out).astype('float32')
	out -= 80.0      # background-subtracted frames, like images/NNNN
	return out, shifts


def test_oracle_pieces_equal_cv2():
	rng = np.random.default_rng(1)
	img = ndi.gaussian_filter(rng.normal(0, 1, (90, 130)), 1.2).astype('float32')
	for tx, ty in ((0.37, -0.61), (-1.25, 2.03125), (0.0, 0.0), (3.999, -0.001)):
		M = np.array([[1, 0, tx], [0, 1, ty]], dtype='float32')
		ref = cv2.warpAffine(img, M, (130, 90), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP)
		assert np.array_equal(imo.warp_translation(img, np.float32(tx), np.float32(ty)), ref)
		m8 = np.ones(img.shape, 'uint8')
		refm = cv2.warpAffine(m8, M, (130, 90), flags=cv2.INTER_NEAREST + cv2.WARP_INVERSE_MAP)
		assert np.array_equal(imo.warp_translation(m8, np.float32(tx), np.float32(ty), nearest=True), refm)
	np.testing.assert_allclose(imo._gauss5(img), cv2.GaussianBlur(img, (5, 5), 0), atol=3e-7)
	g = cv2.GaussianBlur(img, (5, 5), 0)
	gx, gy = imo._gradients(g)
	np.testing.assert_allclose(gx, cv2.filter2D(g, -1, np.array([[-0.5, 0, 0.5]], dtype='float32')), atol=2e-7)
	np.testing.assert_allclose(gy, cv2.filter2D(g, -1, np.array([[-0.5], [0], [0.5]], dtype='float32')), atol=2e-7)


@pytest.mark.parametrize('seed', [3, 4, 5])
def test_oracle_ecc_equals_cv2(seed):
	imgs, shifts = frames(seed)
	ref = imo.prepare_flux(imgs[0])
	for k in range(1, imgs.shape[0]):
		p = imo.prepare_flux(imgs[k])
		cc, wm = cv2.findTransformECC(ref, p, np.eye(2, 3, dtype='float32'), cv2.MOTION_TRANSLATION, CRIT, np.ones(p.shape, 'uint8'), 5)
		rho, tx, ty = imo.find_transform_ecc_translation(ref, p)
		assert abs(tx - wm[0, 2]) < 2e-5 and abs(ty - wm[1, 2]) < 2e-5 and abs(rho - cc) < 1e-6
		# and the estimate is the motion that was put in (cv2 convention: warped(x) = image(x + t) matches the template, so t = the shift applied to the image)
		assert abs(tx - shifts[k][1]) < 0.05 and abs(ty - shifts[k][0]) < 0.05


def test_prepare_flux_properties():
	imgs, _ = frames(7, n=1)
	img = imgs[0].copy()
	img[5, 7] = np.nan
	p = imo.prepare_flux(img)
	assert p.dtype == np.float32 and p.shape == img.shape and np.isfinite(p).all()
	# a NaN pixel poisons its eight neighbours, which are then zeroed (image_motion.py:107-108); its own gradient does not
	# involve it (the Scharr kernels have a zero centre and scipy skips zero weights)
	blk = p[4:7, 6:9].copy(); centre = blk[1, 1]; blk[1, 1] = 0
	assert (blk == 0).all() and centre > 0
	assert p.max() <= np.sqrt(2) * 2 + 1e-6  # gradient magnitude of an image scaled to [-1, 1]


# ---- GPU ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_prepare_flux_matches_oracle():
	import torch
	from photometry_b200 import image_motion as im
	imgs, _ = frames(11)
	imgs[1, 20, 30] = np.nan
	got = im.prepare_flux(torch.from_numpy(imgs).cuda()).cpu().numpy()
	for k in range(imgs.shape[0]):
		ref = imo.prepare_flux(imgs[k])
		# float32 log10 / sqrt differ by an ulp between numpy and the device; the gradient image has values up to ~1
		np.testing.assert_allclose(got[k], ref, atol=3e-6)
		assert np.array_equal(got[k] == 0, ref == 0) or np.mean((got[k] == 0) != (ref == 0)) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('seed', [3, 21])
def test_gpu_ecc_matches_oracle_and_cv2(seed):
	import torch
	from photometry_b200 import image_motion as im
	imgs, shifts = frames(seed, H=256, W=320, n=6)
	imk = im.ImageMovementKernel(imgs[0])
	assert imk.n_params == 2 and imk.warpmode == 'translation'
	k, rho, iters = imk.calc_kernels(torch.from_numpy(imgs).cuda(), return_info=True)
	assert k.shape == (6, 2) and np.isfinite(k).all()
	ref = imo.prepare_flux(imgs[0])
	for j in range(imgs.shape[0]):
		p = imo.prepare_flux(imgs[j])
		cc, wm = cv2.findTransformECC(ref, p, np.eye(2, 3, dtype='float32'), cv2.MOTION_TRANSLATION, CRIT, np.ones(p.shape, 'uint8'), 5)
		# the criterion is 1e-6 on the correlation coefficient and the warp is quantised to 1/32 px: the converged translations of
		# two correct implementations agree to ~1e-3 px
		assert abs(k[j, 0] - wm[0, 2]) < 2e-3 and abs(k[j, 1] - wm[1, 2]) < 2e-3, (j, k[j], wm[:, 2])
		assert abs(rho[j] - cc) < 1e-4
		assert abs(k[j, 0] - shifts[j][1]) < 0.05 and abs(k[j, 1] - shifts[j][0]) < 0.05
	one = imk.calc_kernel(imgs[3])
	assert abs(one[0] - k[3, 0]) < 1e-6 and abs(one[1] - k[3, 1]) < 1e-6
	assert np.allclose(imk.apply_kernel([[10, 20], [30, 40]], k[3]), np.tile(k[3], (2, 1)))
