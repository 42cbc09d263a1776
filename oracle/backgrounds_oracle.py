"""
CPU oracle: restatement of ``photometry.backgrounds.fit_background``.

TEST INFRASTRUCTURE ONLY -- never imported by the product path (see oracle/__init__.py).
PARITY UNPINNED at the third-party boundary (photutils/astropy/statsmodels/bottleneck are
restated here, scipy is called for real).  All citations are relative to /root/reference.

Floating-point policy (mirrors the reference running on its pinned NumPy 1.21):
  * round 1 of the radial component works on float32 pixels (``pix + zeropoint`` stays
    float32 under value-based casting) and takes ``log10`` in float32.  This oracle defines
    that float32 ``log10`` as the *correctly rounded* one (float64 log10 rounded to float32),
    which is what a <=0.5 ulp libm returns; the CUDA path does the same.
  * rounds 2..n work in float64 (``img0 - img_bkg_square`` is float64).
  * everything inside Background2D is float64 (astropy's fast sigma-clip returns float64).
"""
import warnings
import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline
from scipy.ndimage import generic_filter, zoom
from scipy.spatial import cKDTree
from scipy.stats import scoreatpercentile

# photometry/backgrounds.py:121-138 -- pixel coordinates of the camera centre per (camera, ccd).
XYCEN = {
	(1, 1): (2158.222313, 2099.523364),
	(1, 2): (-5.653058, 2098.018608),
	(1, 3): (2141.511437, 2099.868226),
	(1, 4): (-22.406442, 2100.116443),
	(2, 1): (2148.588316, 2094.033024),
	(2, 2): (-16.806140, 2095.810070),
	(2, 3): (2151.351646, 2105.747100),
	(2, 4): (-13.118570, 2105.982211),
	(3, 1): (2152.175481, 2092.337442),
	(3, 2): (-10.494413, 2093.108135),
	(3, 3): (2145.029218, 2107.883573),
	(3, 4): (-17.374782, 2105.296746),
	(4, 1): (2149.259760, 2091.433315),
	(4, 2): (-12.906931, 2093.350054),
	(4, 3): (2148.906766, 2110.730620),
	(4, 4): (-14.629676, 2111.341670),
}

PIXEL_OFFSET_COLUMN = 44  # photometry/io.py:47 (science pixels start at raw column 44)


# --------------------------------------------------------------------------------------------------
class FFIImageLite:
	"""
	Minimal stand-in for ``photometry.io.FFIImage`` (io.py:25-93): float32 science pixels,
	the merged header scalars the hot path reads, ``is_tess`` and ``mask = ~isfinite``.
	An ndarray input gives ``is_tess=False`` exactly like io.py:34-35.
	"""
	def __init__(self, data, header=None, is_tess=False):
		self.data = np.asarray(data)
		self.header = dict(header or {})
		self.is_tess = bool(is_tess)
		self.mask = ~np.isfinite(self.data)  # io.py:89
		self.shape = self.data.shape


# --------------------------------------------------------------------------------------------------
def pixel_manual_exclude(img):
	"""photometry/pixel_flags.py:14-58."""
	mask = np.zeros(img.shape, dtype=bool)
	hdr = img.header
	if img.is_tess:
		time = 0.5 * (hdr['TSTART'] + hdr['TSTOP'])
		cadenceno = hdr.get('FFIINDEX', np.inf)
	else:
		time = np.nan
		cadenceno = np.inf

	if img.is_tess and hdr['CAMERA'] == 1 and hdr['CCD'] == 4 \
		and (cadenceno <= 4724 or hdr['TSTART'] <= 1325.881282301840):
		mask[:, 1536:] = True  # Mars: register overflow (pixel_flags.py:44-46)
	elif img.is_tess and hdr['CAMERA'] == 1 \
		and (11354 <= cadenceno <= 11366 or 1464.0158778 <= time <= 1464.265871):
		mask[:, :] = True  # excessive Earth-shine (pixel_flags.py:48-50)

	if img.is_tess and np.all(img.data == 0):
		mask[:, :] = True  # whole image zero (pixel_flags.py:54-56)
	return mask


# --------------------------------------------------------------------------------------------------
def star_radius(tmag):
	"""Extension (photometry_b200/starmask.py): r = clip(4 * 10**(-0.1 (Tmag - 10)), 1.5, 40) pixels."""
	return np.clip(4.0 * 10.0 ** (-0.1 * (np.asarray(tmag, dtype='float64') - 10.0)), 1.5, 40.0)


def star_mask(shape, catalog):
	"""
	Extension beyond the reference (its ``catalog`` argument is unused, backgrounds.py:64-65, 90): boolean [H, W], True where
	(x - column)**2 + (y - row)**2 <= r(Tmag)**2 for a catalog row (column, row, Tmag) in science-pixel coordinates.
	"""
	H, W = shape
	mask = np.zeros((H, W), dtype=bool)
	cat = np.asarray(catalog, dtype='float64').reshape(-1, 3)
	for (sx, sy, tm), r in zip(cat, star_radius(cat[:, 2]) if cat.size else []):
		x0, x1 = max(0, int(np.ceil(sx - r))), min(W - 1, int(np.floor(sx + r)))
		y0, y1 = max(0, int(np.ceil(sy - r))), min(H - 1, int(np.floor(sy + r)))
		if x1 < x0 or y1 < y0:
			continue
		yy, xx = np.mgrid[y0:y1 + 1, x0:x1 + 1].astype('float64')
		mask[y0:y1 + 1, x0:x1 + 1] |= ((xx - sx) * (xx - sx) + (yy - sy) * (yy - sy)) <= r * r
	return mask


# --------------------------------------------------------------------------------------------------
def _nanmedian1d(x):
	x = np.asarray(x, dtype='float64')
	x = x[~np.isnan(x)]
	if x.size == 0:
		return np.nan
	return float(np.median(x))


def move_median_central(x, width_points):
	"""
	photometry/utilities.py:52-62 for a 1-D input: bottleneck ``move_median(x, w, min_count=1)``
	(trailing nan-aware window), rolled by ``-w//2+1``, then the first/last ``w//2+1`` points
	replaced by nan-medians of the growing edge windows.
	"""
	x = np.asarray(x, dtype='float64')
	n = x.shape[0]
	w = int(width_points)
	y = np.empty(n, dtype='float64')
	for i in range(n):  # bottleneck.move_median: window x[i-w+1 .. i], min_count=1
		y[i] = _nanmedian1d(x[max(0, i - w + 1):i + 1])
	y = np.roll(y, -w // 2 + 1)
	for k in range(w // 2 + 1):
		y[k] = _nanmedian1d(x[:(k + 2)])
		y[-(k + 1)] = _nanmedian1d(x[-(k + 2):])
	return y


# --------------------------------------------------------------------------------------------------
def kde_density(x, gridsize=2000):
	"""
	statsmodels 0.13.2 ``KDEUnivariate(x).fit(gridsize=...)`` with the defaults the reference
	uses (kernel='gau', bw='normal_reference', fft=True, cut=3, adjust=1): returns
	(density, support, bw).  Raises RuntimeError when the selected bandwidth is 0.
	"""
	x = np.asarray(x, dtype='float64')
	x = x[np.logical_and(x > -np.inf, x < np.inf)]
	nobs = x.shape[0]
	# bandwidths._select_sigma + bw_normal_reference
	with warnings.catch_warnings():
		warnings.simplefilter('ignore')
		iqr = (scoreatpercentile(x, 75) - scoreatpercentile(x, 25)) / 1.349
		std_dev = np.std(x, axis=0, ddof=1)
	sigma = np.minimum(std_dev, iqr) if iqr > 0 else std_dev
	bw = 1.0592238410488122 * sigma * nobs ** (-0.2)
	if bw == 0:
		raise RuntimeError("Selected KDE bandwidth is 0. Cannot estimate density.")
	M = int(2 ** np.ceil(np.log2(gridsize)))
	a = np.min(x) - 3 * bw
	b = np.max(x) + 3 * bw
	grid, delta = np.linspace(a, b, M, retstep=True)
	RANGE = b - a
	# linbin.fast_linbin
	dl = (b - a) / (M - 1)
	lxi = (x - a) / dl
	with np.errstate(invalid='ignore'):
		li = lxi.astype(np.int64)
	rem = lxi - li
	ok = (li > 1) & (li < M - 1)
	gcnts = np.zeros(M + 1, dtype='float64')
	gcnts += np.bincount(li[ok], weights=1.0 - rem[ok], minlength=M + 1)
	gcnts += np.bincount(li[ok] + 1, weights=rem[ok], minlength=M + 1)
	binned = gcnts[:M] / (delta * nobs)
	# forrt -> silverman_transform -> revrt  (== irfft(rfft(binned) * FAC))
	y = np.fft.rfft(binned, M) / M
	J = np.arange(M / 2 + 1)
	FAC1 = 2 * (np.pi * bw / RANGE) ** 2
	BC = 1 - 1.0 / 3 * (J * 1.0 / M * np.pi) ** 2
	FAC = np.exp(-(J ** 2 * FAC1)) / BC
	f = np.fft.irfft(y * FAC) * M
	return f, grid, bw


def reduce_mode(x):
	"""photometry/backgrounds.py:21-33."""
	if len(x) == 0:
		return np.nan
	x = np.asarray(x, dtype='float64')
	if x.shape[0] == 1:
		# std(ddof=1) of one sample is NaN -> bw NaN -> all-NaN density -> support[0] = NaN.
		return np.nan
	try:
		density, support, _ = kde_density(x, gridsize=2000)
	except RuntimeError:
		return float(np.median(x))
	return float(support[np.argmax(density)])


# --------------------------------------------------------------------------------------------------
def sigma_clip_bounds(rows, sigma=3.0, maxiters=5):
	"""
	astropy 5.1 ``SigmaClip(sigma, maxiters, cenfunc='median', stdfunc='std')`` fast C path
	(``stats/_fast_sigma_clip.c``) applied along axis 1 of ``rows`` (NaN = masked).
	Returns the last computed (lo, hi) per row as float64; rows with no valid value give NaN.
	"""
	buf = np.array(rows, dtype='float64', copy=True)
	nrow = buf.shape[0]
	lo = np.full(nrow, np.nan)
	hi = np.full(nrow, np.nan)
	count = np.sum(~np.isnan(buf), axis=1)
	active = np.flatnonzero(count > 0)
	for _ in range(maxiters):
		if active.size == 0:
			break
		sub = buf[active]
		n = count[active].astype('float64')
		mean = np.nansum(sub, axis=1) / n
		s = np.sort(sub, axis=1)  # NaNs sort last
		ni = count[active]
		ar = np.arange(active.size)
		median = 0.5 * (s[ar, (ni - 1) // 2] + s[ar, ni // 2])
		std = np.sqrt(np.nansum((mean[:, None] - sub) ** 2, axis=1) / n)
		lo_a = median - sigma * std
		hi_a = median + sigma * std
		lo[active] = lo_a
		hi[active] = hi_a
		with np.errstate(invalid='ignore'):
			out = (sub < lo_a[:, None]) | (sub > hi_a[:, None])
		sub[out] = np.nan
		buf[active] = sub
		new_count = np.sum(~np.isnan(sub), axis=1)
		changed = new_count != count[active]
		count[active] = new_count
		active = active[changed]
	return lo, hi


def sextractor_background(rows):
	"""
	photutils 1.3.0 ``SExtractorBackground.calc_background(data, axis=1)`` with
	``sigma_clip=None`` on NaN-filled float64 rows.  Returns (bkg, median, mean, std).
	"""
	with warnings.catch_warnings():
		warnings.simplefilter('ignore', category=RuntimeWarning)
		med = np.nanmedian(rows, axis=1)
		mean = np.nanmean(rows, axis=1)
		std = np.nanstd(rows, axis=1)
	bkg = 2.5 * med - 1.5 * mean
	bkg = np.where(std == 0, mean, bkg)
	idx = np.where(std != 0)
	with np.errstate(invalid='ignore'):
		cond = (np.abs(mean[idx] - med[idx]) / std[idx]) < 0.3
	bkg[idx] = np.where(cond, bkg[idx], med[idx])
	return bkg, med, mean, std


# photutils 1.3.0 ``ShepardIDWInterpolator.__init__(coordinates, values, weights=None, leafsize=10)`` builds
# ``cKDTree(coordinates, leafsize=leafsize)``; Background2D does not override it.  (scipy's own default is 16, which
# gives different leaves and therefore different tie decisions on the mesh lattice -- pinned by tests/test_kdtree.py.)
PHOTUTILS_IDW_LEAFSIZE = 10


def _idw_fill(good_yx, good_values, ny, nx, mode):
	"""
	photutils 1.3.0 ``Background2D._interpolate_meshes``: ShepardIDWInterpolator over the good
	meshes evaluated at every mesh position, n_neighbors=10, power=1, reg=0, conf_dist=1e-12.

	mode='ckdtree': neighbours exactly as ``scipy.spatial.cKDTree(yx, leafsize=10).query(k=10)`` returns them
	                (the reference; ties between equidistant lattice points are resolved by the tree's leaf
	                order and traversal order).  This is the default and what the CUDA path reproduces.
	mode='stable' : neighbours are the 10 smallest by (squared distance, good-mesh order) -- the rule the
	                round-1 CUDA path used; kept only to quantify how far that rule was from the reference
	                (tests/test_kdtree.py::test_stable_rule_deviation).
	"""
	coords = np.array([(iy, ix) for iy in range(ny) for ix in range(nx)], dtype='float64')
	k = 10
	npts = good_yx.shape[0]
	if mode == 'ckdtree':
		dist, idx = cKDTree(good_yx, leafsize=PHOTUTILS_IDW_LEAFSIZE).query(coords, k=k, eps=0.0)
	elif mode == 'stable':
		d2 = ((coords[:, None, :] - good_yx[None, :, :]) ** 2).sum(axis=2)
		kk = min(k, npts)
		order = np.argsort(d2, axis=1, kind='stable')[:, :kk]
		dist = np.sqrt(np.take_along_axis(d2, order, axis=1))
		idx = order
		if kk < k:
			dist = np.concatenate([dist, np.full((dist.shape[0], k - kk), np.inf)], axis=1)
			idx = np.concatenate([idx, np.full((idx.shape[0], k - kk), npts)], axis=1)
	else:
		raise ValueError(mode)
	out = np.zeros(coords.shape[0])
	for p in range(coords.shape[0]):
		valid = np.isfinite(dist[p])
		idk = idx[p][valid]
		dk = dist[p][valid]
		if dk.shape[0] == 0:
			out[p] = np.nan
			continue
		confused = dk <= 1e-12
		if np.any(confused):
			out[p] = good_values[idk[confused][0]]
			continue
		w = 1.0 / ((dk ** 1.0) + 0.0)
		wsum = np.sum(w)
		out[p] = np.sum(w * good_values[idk]) / wsum
	return out.reshape(ny, nx)


class Background2DOracle:
	"""
	photutils 1.3.0 ``Background2D(data, (box, box), filter_size=(3, 3), sigma_clip=SigmaClip(3, 5),
	bkg_estimator=SExtractorBackground, mask=mask, exclude_percentile=50)`` as called at
	photometry/backgrounds.py:200-205.  ``.background`` is the float64 full-resolution map.
	"""
	def __init__(self, data, mask, box=64, exclude_percentile=50.0, idw='ckdtree'):
		data = np.asarray(data)
		H, W = data.shape
		if H % box or W % box:
			raise ValueError("oracle supports only image sizes that are multiples of the box size")
		ny, nx = H // box, W // box
		npix = box * box
		d = data.astype('float64')
		d[np.asarray(mask, dtype=bool) | ~np.isfinite(d)] = np.nan
		rows = d.reshape(ny, box, nx, box).swapaxes(1, 2).reshape(ny * nx, npix)
		# sigma-clip every mesh once; bounds re-applied to the original row
		lo, hi = sigma_clip_bounds(rows, 3.0, 5)
		with np.errstate(invalid='ignore'):
			clipped = (rows < lo[:, None]) | (rows > hi[:, None])
		rows = rows.copy()
		rows[clipped] = np.nan
		nbad = np.sum(np.isnan(rows), axis=1)
		good = nbad <= (exclude_percentile / 100.0 * npix)
		if not np.any(good):
			raise ValueError("All meshes contain > %d masked pixels." % int(exclude_percentile / 100.0 * npix))
		stat, med, mean, std = sextractor_background(rows[good])
		self.clip_lo, self.clip_hi = lo, hi
		self.mesh_good = good.reshape(ny, nx)
		self.mesh_nbad = nbad.reshape(ny, nx)
		self.n_excluded = int(np.sum(~good))
		if self.n_excluded == 0:
			mesh = stat.reshape(ny, nx)
		else:
			gy, gx = np.divmod(np.flatnonzero(good), nx)
			good_yx = np.column_stack([gy, gx]).astype('float64')
			mesh = _idw_fill(good_yx, stat, ny, nx, idw)
		self.mesh_unfiltered = mesh
		# 3x3 nan-median with NaN padding (scipy for real)
		mesh = generic_filter(mesh, np.nanmedian, size=(3, 3), mode='constant', cval=np.nan)
		self.background_mesh = mesh
		# BkgZoomInterpolator(order=3, mode='reflect', grid_mode=True, clip=True)
		if np.ptp(mesh) == 0:
			bkg = np.zeros(data.shape, dtype=data.dtype) + data.dtype.type(np.min(mesh))
			bkg = bkg.astype('float64')
		else:
			bkg = zoom(mesh, (box, box), order=3, mode='reflect', cval=0.0, grid_mode=True)
			bkg = bkg[0:H, 0:W]
			np.clip(bkg, np.min(mesh), np.max(mesh), out=bkg)
		self.background = bkg


# --------------------------------------------------------------------------------------------------
def radial_geometry(shape, xycen, radial_cutoff, radial_pixel_step):
	"""photometry/backgrounds.py:145-154: r image, ring edges, ring centres."""
	xx, yy = np.meshgrid(
		np.arange(PIXEL_OFFSET_COLUMN, shape[1] + PIXEL_OFFSET_COLUMN, 1),
		np.arange(0, shape[0], 1))
	r = np.sqrt((xx - xycen[0]) ** 2 + (yy - xycen[1]) ** 2)
	radial_max = np.max(r) + radial_pixel_step
	bins = np.arange(radial_cutoff, radial_max, radial_pixel_step)
	bin_center = bins[1:] - radial_pixel_step / 2
	return r, bins, bin_center


def _log10_f32(a32):
	"""Correctly rounded float32 log10 (see module docstring)."""
	return np.log10(a32.astype('float64')).astype('float32')


def _ring_statistic(rvals, values, bins, discarded_bins):
	"""
	scipy.stats.binned_statistic(rvals, values, statistic=reduce_mode, bins=bins): bin i is
	edges[i] <= r < edges[i+1], r == edges[-1] joins the last bin; empty bins get
	reduce_mode([]) = NaN.  scipy also calls the statistic on the under/overflow bins and throws
	the result away -- reproduced only when ``discarded_bins`` is set (CPU-baseline timing).
	"""
	nb = len(bins) - 1
	idx = np.searchsorted(bins, rvals, side='right')
	idx[rvals == bins[-1]] -= 1
	out = np.full(nb, np.nan)
	order = np.argsort(idx, kind='stable')
	sidx = idx[order]
	sval = values[order]
	bounds = np.searchsorted(sidx, np.arange(0, nb + 3))
	for b in range(0, nb + 2):
		seg = sval[bounds[b]:bounds[b + 1]]
		if b == 0 or b == nb + 1:
			if discarded_bins and seg.size:
				reduce_mode(seg)
			continue
		out[b - 1] = reduce_mode(seg)
	return out


def fit_background(image, catalog=None, flux_cutoff=8e4, bkgiters=3, radial_cutoff=2400,
	radial_pixel_step=15, radial_smooth=3, *, extra_mask=None, xycen=None, idw='ckdtree',
	discarded_bins=False, diagnostics=None):
	"""
	photometry/backgrounds.py:52-211.  ``image`` is a 2-D ndarray (non-TESS path) or an
	:class:`FFIImageLite` (TESS path when ``is_tess``).  Returns ``(bkg float64, mask bool)``.

	Extensions beyond the reference (all default-off): ``extra_mask`` is OR-ed into the mask at
	the point of backgrounds.py:90 (star-mask extension); ``xycen`` overrides
	the camera-centre table so the radial path can be exercised on small images; ``idw='stable'``
	switches to the round-1 neighbour rule for comparison (see ``_idw_fill``); ``diagnostics`` (dict) receives intermediates.
	"""
	img0 = image if isinstance(image, FFIImageLite) else FFIImageLite(image)
	if img0.data.ndim != 2:
		raise ValueError("Input image must be either 2D ndarray or path to file.")
	hdr = img0.header
	data = img0.data
	diag = diagnostics if diagnostics is not None else {}

	mask = img0.mask.copy()
	mask |= ~np.isfinite(data)
	with np.errstate(invalid='ignore'):
		mask |= (data > flux_cutoff)
		mask |= (data < 0)
	if extra_mask is not None:
		mask |= np.asarray(extra_mask, dtype=bool)
	mask |= pixel_manual_exclude(img0)

	if np.all(mask):
		return np.full(data.shape, np.nan, dtype='float64'), mask

	use_radial = True
	if img0.is_tess:
		camera, ccd = hdr.get('CAMERA'), hdr.get('CCD')
		cen = xycen if xycen is not None else XYCEN.get((camera, ccd))
		if cen is None:
			raise ValueError(f"Invalid CAMERA or CCD in header: CAMERA={camera}, CCD={ccd}")
		r, bins, bin_center = radial_geometry(data.shape, cen, radial_cutoff, radial_pixel_step)
		if len(bins) < 2:
			raise ValueError("radial_cutoff leaves no radial bins inside the image")
	else:
		use_radial = False
		bkgiters = 1

	img_bkg_radial = None  # None stands for the reference's scalar/0-d zero
	img_bkg_square = None
	diag['rounds'] = []
	for _ in range(bkgiters):
		rd = {}
		if use_radial:
			if img_bkg_square is None:
				# float32 round: img0 - np.asarray(0) stays float32 (NumPy 1.21 casting)
				pix = data[~mask].astype('float32')
				zeropoint = 1.0 - np.float64(np.min(pix))
				logpix = _log10_f32(pix + np.float32(zeropoint))
			else:
				pix = (data.astype('float64') - img_bkg_square)[~mask]
				zeropoint = -np.min(pix) + 1.0
				logpix = np.log10(pix + zeropoint)
			s2 = _ring_statistic(r[~mask], logpix.astype('float64'), bins, discarded_bins)
			rd['s2_raw'] = s2.copy()
			if radial_smooth:
				s2 = move_median_central(s2, radial_smooth)
			rd['s2'] = s2.copy()
			rd['zeropoint'] = float(zeropoint)
			indx = ~np.isnan(s2)
			ngood = int(np.sum(indx))
			img_bkg_radial = None
			if ngood >= 3:
				try:
					intp = InterpolatedUnivariateSpline(bin_center[indx], s2[indx], k=3, ext=3)
					img_bkg_radial = 10 ** intp(r) - zeropoint
				except ValueError:
					img_bkg_radial = None  # backgrounds.py:192-194
			rd['radial_ok'] = img_bkg_radial is not None
		if img_bkg_radial is None:
			b2d_data = data
		else:
			b2d_data = data.astype('float64') - img_bkg_radial
		bkg = Background2DOracle(b2d_data, mask, box=64, exclude_percentile=50.0, idw=idw)
		img_bkg_square = bkg.background
		rd['mesh'] = bkg.background_mesh
		rd['mesh_unfiltered'] = bkg.mesh_unfiltered
		rd['mesh_good'] = bkg.mesh_good
		rd['n_excluded'] = bkg.n_excluded
		rd['mesh_nbad'] = bkg.mesh_nbad
		rd['clip_lo'], rd['clip_hi'] = bkg.clip_lo, bkg.clip_hi
		diag['rounds'].append(rd)

	if img_bkg_radial is None:
		img_bkg = img_bkg_square
	else:
		img_bkg = img_bkg_radial + img_bkg_square
	return np.asarray(img_bkg, dtype='float64'), mask
