"""
CPU oracle: ``photometry.image_motion.ImageMovementKernel`` for ``warpmode='translation'``
(photometry/image_motion.py:74-111 ``_prepare_flux``, :182-258 ``calc_kernel``; driver photometry/prepare.py:678-698).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Two third-party pieces carry the arithmetic:

  * ``skimage.filters.scharr`` (scikit-image 0.19.2, absent here): ``sqrt((h**2 + v**2) / 2)`` of the two 3 x 3 Scharr
    convolutions (smoothing weights [3, 10, 3] / 16, edge weights [1, 0, -1]) through ``scipy.ndimage.convolve(mode='reflect')``
    -- restated with the real scipy;
  * ``cv2.findTransformECC`` (opencv 4.5.5 pinned; 4.13 is importable here): restated below step by step *and pinned against
    the real function* by tests/test_image_motion.py -- Gaussian 5-tap pre-filter, central-difference gradients, fixed-point
    bilinear ``warpAffine`` (coordinates rounded to 1/32 px), masked zero-mean correlation, Gauss-Newton update.
"""
import numpy as np
from scipy import ndimage as ndi

GAUSS5 = np.array([0.0625, 0.25, 0.375, 0.25, 0.0625], dtype='float32')   # cv::getGaussianKernel(5, sigma <= 0): fixed table


def scharr(image):
	"""skimage 0.19.2 ``scharr(image)`` for a 2-D float image (mask=None, axis=None, mode='reflect')."""
	image = np.asarray(image)
	ft = np.float32 if image.dtype == np.float32 else np.float64
	image = image.astype(ft)
	smooth = np.array([3, 10, 3]) / 16
	edge = np.array([1, 0, -1])
	out = np.zeros(image.shape, dtype=ft)
	for edge_dim in (0, 1):
		kernel = (edge.reshape(3, 1) * smooth.reshape(1, 3)) if edge_dim == 0 else (smooth.reshape(3, 1) * edge.reshape(1, 3))
		ax = ndi.convolve(image, kernel, mode='reflect')
		out += ax * ax
	return np.sqrt(out) / np.sqrt(2, dtype=ft)


def prepare_flux(flux):
	"""image_motion.py:74-111 (NumPy 1.21 casting: a float32 image stays float32 throughout)."""
	flux = np.asarray(flux)
	ft = flux.dtype.type if flux.dtype.kind == 'f' else np.float64
	with np.errstate(invalid='ignore', divide='ignore'):
		flux = np.log10(flux - np.nanmin(flux) + ft(1.0)).astype(ft)
		fmax = np.nanmax(flux)
		fmin = np.nanmin(flux)
		ran = np.abs(fmax - fmin)
		flux1 = (ft(-1) + ft(2) * ((flux - fmin) / ran)).astype(ft)
		flux1 = scharr(flux1)
	flux1[np.isnan(flux1)] = 0
	return np.asarray(flux1, dtype='float32')


# --------------------------------------------------------------------------------------------------
def _gauss5(img):
	"""cv::GaussianBlur(img, (5, 5), 0) on float32, BORDER_REFLECT_101."""
	img = np.asarray(img, dtype='float32')
	p = np.pad(img, 2, mode='reflect')
	rows = sum(GAUSS5[k] * p[:, k:k + img.shape[1]] for k in range(5)).astype('float32')
	return sum(GAUSS5[k] * rows[k:k + img.shape[0], :] for k in range(5)).astype('float32')


def _gradients(img):
	"""cv::filter2D with [-0.5, 0, 0.5] and its transpose (correlation), BORDER_REFLECT_101."""
	p = np.pad(img, 1, mode='reflect')
	gx = (0.5 * p[1:-1, 2:] - 0.5 * p[1:-1, :-2]).astype('float32')
	gy = (0.5 * p[2:, 1:-1] - 0.5 * p[:-2, 1:-1]).astype('float32')
	return gx, gy


def _round_half_even(v):
	return np.rint(v).astype(np.int64)


def warp_translation(src, tx, ty, nearest=False):
	"""
	cv::warpAffine(src, M = [[1, 0, tx], [0, 1, ty]], WARP_INVERSE_MAP, INTER_LINEAR or INTER_NEAREST, BORDER_CONSTANT 0):
	dst(x, y) = src(x + tx, y + ty) with the source coordinates in fixed point (AB_BITS = 10), rounded to 1/32 px for the
	bilinear table (float32 weights) or to the nearest pixel.
	"""
	H, W = src.shape
	AB = 1024
	xs = np.arange(W); ys = np.arange(H)
	adelta = _round_half_even(1.0 * xs * AB)
	bdelta = _round_half_even(0.0 * xs * AB)
	rd = AB // 2 if nearest else AB // 32 // 2
	X0 = _round_half_even((0.0 * ys + float(tx)) * AB) + rd
	Y0 = _round_half_even((1.0 * ys + float(ty)) * AB) + rd
	if nearest:
		X = (X0[:, None] + adelta[None, :]) >> 10
		Y = (Y0[:, None] + bdelta[None, :]) >> 10
		ok = (X >= 0) & (X < W) & (Y >= 0) & (Y < H)
		out = np.zeros((H, W), dtype=src.dtype)
		out[ok] = src[Y[ok], X[ok]]
		return out
	X = (X0[:, None] + adelta[None, :]) >> 5
	Y = (Y0[:, None] + bdelta[None, :]) >> 5
	sx, sy = X >> 5, Y >> 5
	fx = ((X & 31) / 32.0).astype('float32'); fy = ((Y & 31) / 32.0).astype('float32')
	w = [((1 - fy) * (1 - fx)).astype('float32'), ((1 - fy) * fx).astype('float32'), (fy * (1 - fx)).astype('float32'), (fy * fx).astype('float32')]

	def at(yy, xx):
		ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
		v = np.zeros((H, W), dtype='float32')
		v[ok] = src[yy[ok], xx[ok]]
		return v
	out = at(sy, sx) * w[0] + at(sy, sx + 1) * w[1] + at(sy + 1, sx) * w[2] + at(sy + 1, sx + 1) * w[3]
	return out.astype('float32')


def find_transform_ecc_translation(template, image, number_of_iterations=10000, termination_eps=1e-6, input_mask=None, trace=None):
	"""
	``cv2.findTransformECC(template, image, eye(2, 3), MOTION_TRANSLATION, criteria, inputMask, gaussFiltSize=5)``.
	Returns ``(rho, tx, ty)``; raises RuntimeError where OpenCV raises (no convergence / NaN).
	"""
	template = np.asarray(template, dtype='float32'); image = np.asarray(image, dtype='float32')
	tmpl = _gauss5(template)
	pre = np.ones(image.shape, 'uint8') if input_mask is None else (np.asarray(input_mask) > 0).astype('uint8')
	pmf = (_gauss5(pre.astype('float32')) * np.float32(0.5 / 0.95)).astype('float32')
	pre = np.rint(pmf).astype('uint8')           # "rounding conversion"
	pmf = pre.astype('float32')
	img = _gauss5(image)
	gx, gy = _gradients(img)
	gx = gx * pmf; gy = gy * pmf
	tx = ty = np.float32(0.0)                    # the warp matrix is float32
	rho, last_rho = -1.0, -termination_eps
	it = 0
	while it < number_of_iterations and abs(rho - last_rho) >= termination_eps:
		it += 1
		iw = warp_translation(img, tx, ty)
		gxw = warp_translation(gx, tx, ty); gyw = warp_translation(gy, tx, ty)
		m = warp_translation(pre, tx, ty, nearest=True) != 0
		n = int(m.sum())
		i64 = iw.astype('float64'); t64 = tmpl.astype('float64')
		img_mean = i64[m].mean(); tmp_mean = t64[m].mean()
		img_std = np.sqrt(max((i64[m] ** 2).mean() - img_mean ** 2, 0.0)); tmp_std = np.sqrt(max((t64[m] ** 2).mean() - tmp_mean ** 2, 0.0))
		iz = iw.copy(); iz[m] = (i64[m] - img_mean).astype('float32')                 # subtract(..., mask): untouched outside the mask
		tz = np.zeros_like(tmpl); tz[m] = (t64[m] - tmp_mean).astype('float32')
		tmp_norm = np.sqrt(n * tmp_std * tmp_std); img_norm = np.sqrt(n * img_std * img_std)
		gx64, gy64, iz64, tz64 = gxw.astype('float64'), gyw.astype('float64'), iz.astype('float64'), tz.astype('float64')
		hess = np.array([[np.sum(gx64 * gx64), np.sum(gx64 * gy64)], [np.sum(gx64 * gy64), np.sum(gy64 * gy64)]]).astype('float32')
		hinv = np.linalg.inv(hess.astype('float64')).astype('float32')
		corr = float(np.sum(tz64 * iz64))
		last_rho = rho
		rho = corr / (img_norm * tmp_norm)
		if np.isnan(rho):
			raise RuntimeError("NaN encountered.")
		ip = np.array([np.sum(gx64 * iz64), np.sum(gy64 * iz64)]).astype('float32')
		tp = np.array([np.sum(gx64 * tz64), np.sum(gy64 * tz64)]).astype('float32')
		iph = (hinv @ ip).astype('float32')
		lam_n = img_norm * img_norm - float(np.dot(ip.astype('float64'), iph.astype('float64')))
		lam_d = corr - float(np.dot(tp.astype('float64'), iph.astype('float64')))
		if lam_d <= 0.0:
			raise RuntimeError("The algorithm stopped before its convergence. The correlation is going to be minimized.")
		lam = lam_n / lam_d
		err = (np.float32(lam) * tz - iz).astype('float32')
		ep = np.array([np.sum(gx64 * err.astype('float64')), np.sum(gy64 * err.astype('float64'))]).astype('float32')
		dp = (hinv @ ep).astype('float32')
		tx = np.float32(tx + dp[0]); ty = np.float32(ty + dp[1])
		if trace is not None:
			trace.append((rho, float(tx), float(ty)))
	return rho, float(tx), float(ty)


def calc_kernel(image_ref, image, number_of_iterations=10000, termination_eps=1e-6):
	"""image_motion.py:182-258 for warpmode='translation': returns [dx, dy] (NaN, NaN when OpenCV raises)."""
	ref = prepare_flux(image_ref)
	img = prepare_flux(image)
	mask = np.isfinite(img).astype('uint8')
	try:
		_, tx, ty = find_transform_ecc_translation(ref, img, number_of_iterations, termination_eps, mask)
	except RuntimeError:
		return [np.nan, np.nan]
	return [tx, ty]
