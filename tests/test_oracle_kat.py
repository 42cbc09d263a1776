"""
CPU tests (no GPU): the oracle against every known answer the reference's own tests hold for this
path and against the committed golden vectors.
"""
import os
import numpy as np
import pytest
import oracle
from oracle import prepare_oracle
from cases import CASES, load_golden, images_digest, in_tolerance


def test_background_fakeimg():
	"""reference tests/test_background.py:36-54: constant image -> background 1000, nothing masked."""
	fakeimg = np.full([2048, 2048], 1000, dtype='float32')
	bck, mask = oracle.fit_background(fakeimg)
	assert bck.shape == fakeimg.shape and mask.shape == fakeimg.shape
	assert np.all(np.isfinite(bck))
	assert mask.dtype == 'bool'
	assert not np.any(mask)
	np.testing.assert_allclose(bck, 1000)


def test_move_median_central():
	"""reference tests/test_utilities.py:24-34."""
	x_1d = np.array([4, 2, 2, 0, 0, np.nan, 0, 2, 2, 4])
	np.testing.assert_allclose(oracle.move_median_central(x_1d, 3), [3, 2, 2, 0, 0, 0, 1, 2, 2, 3])


def _tess_img(data, **hdr):
	base = dict(CAMERA=2, CCD=1, TSTART=1500.0, TSTOP=1500.02, FFIINDEX=20000)
	base.update(hdr)
	return oracle.FFIImageLite(data, base, True)


def test_pixel_manual_exclude_mars():
	"""reference tests/test_pixel_flags.py:17-34."""
	img = _tess_img(np.ones((2048, 2048), dtype='float32'), CAMERA=1, CCD=4, FFIINDEX=4724)
	mask = oracle.pixel_manual_exclude(img)
	assert mask.dtype == 'bool' and mask.shape == img.shape
	assert np.all(mask[:, 1536:]) and not np.any(mask[:, :1536])


def test_pixel_manual_exclude_zero():
	"""reference tests/test_pixel_flags.py:37-52."""
	img = _tess_img(np.zeros((256, 256), dtype='float32'))
	assert np.all(oracle.pixel_manual_exclude(img))


def test_pixel_manual_exclude_earth():
	"""reference tests/test_pixel_flags.py:55-70."""
	img = _tess_img(np.ones((256, 256), dtype='float32'), CAMERA=1, FFIINDEX=11354)
	assert np.all(oracle.pixel_manual_exclude(img))
	img = _tess_img(np.ones((256, 256), dtype='float32'), CAMERA=1, FFIINDEX=11367)
	assert not np.any(oracle.pixel_manual_exclude(img))


def test_quality_bitmask():
	"""reference photometry/quality.py:123-124 / tests/test_quality.py."""
	assert prepare_oracle.TESS_DEFAULT_BITMASK == 4335
	assert prepare_oracle.PIXEL_NOT_USED_FOR_BACKGROUND == 1 and prepare_oracle.PIXEL_MANUAL_EXCLUDE == 2


def test_all_masked_returns_nan():
	"""photometry/backgrounds.py:101-102."""
	img = np.full((128, 128), np.nan, dtype='float32')
	bck, mask = oracle.fit_background(img)
	assert np.all(mask) and np.all(np.isnan(bck))


def test_invalid_camera():
	"""photometry/backgrounds.py:139-140."""
	with pytest.raises(ValueError):
		oracle.fit_background(_tess_img(np.ones((128, 128), dtype='float32'), CAMERA=5, CCD=1))


# ---- self-consistency of the restated third-party pieces -------------------------------------
def test_sigma_clip_matches_naive_loop():
	rng = np.random.default_rng(0)
	rows = rng.normal(100, 5, (6, 500))
	rows[:, :20] += rng.uniform(50, 5000, (6, 20))
	rows[2, 100:300] = np.nan
	rows[5, :] = np.nan
	lo, hi = oracle.sigma_clip_bounds(rows)
	for i in range(rows.shape[0]):
		buf = rows[i][~np.isnan(rows[i])]
		if buf.size == 0:
			assert np.isnan(lo[i]) and np.isnan(hi[i])
			continue
		it = 0
		while True:  # astropy/stats/_fast_sigma_clip.c
			mean = buf.sum() / buf.size
			med = np.median(buf)
			std = np.sqrt(((mean - buf) ** 2).sum() / buf.size)
			l, h = med - 3 * std, med + 3 * std
			new = buf[(buf >= l) & (buf <= h)]
			if new.size == buf.size:
				break
			buf = new
			it += 1
			if it >= 5:
				break
		np.testing.assert_allclose([lo[i], hi[i]], [l, h], rtol=1e-13)


def test_kde_matches_direct_gaussian_kde():
	rng = np.random.default_rng(1)
	x = np.concatenate([rng.normal(2.0, 0.02, 6000), rng.normal(2.3, 0.1, 500)])
	dens, grid, bw = oracle.kde_density(x)
	assert grid.shape == (2048,)
	np.testing.assert_allclose(np.trapezoid(dens, grid), 1.0, rtol=1e-6)
	direct = np.exp(-0.5 * ((grid[::16, None] - x[None, :]) / bw) ** 2).sum(1) / (x.size * bw * np.sqrt(2 * np.pi))
	assert np.max(np.abs(dens[::16] - direct)) < 1e-3 * dens.max()
	assert abs(oracle.reduce_mode(x) - 2.0) < 0.01
	assert np.isnan(oracle.reduce_mode(np.array([])))
	assert oracle.reduce_mode(np.full(10, 1.5)) == 1.5  # bandwidth 0 -> median


def test_time_smooth_and_sumimage_small():
	rng = np.random.default_rng(2)
	bkg = rng.normal(100, 1, (5, 8, 8))
	bkg[1, 2, 2] = np.nan
	sm = oracle.time_smooth_backgrounds(bkg, 3)
	assert sm.dtype == np.float32
	np.testing.assert_allclose(sm[0], np.nanmean(bkg[0:2].astype('float32'), axis=0), rtol=1e-6)
	np.testing.assert_allclose(sm[2], np.nanmean(bkg[1:4].astype('float32'), axis=0), rtol=1e-6)
	imgs = rng.normal(120, 3, (5, 8, 8)).astype('float32')
	flags = np.zeros((5, 8, 8), dtype='uint8')
	flags[0, 0, 0] = 1
	flags[4, 1, :] = 2
	q = np.array([0, 32, 0, 0, 0], dtype='int32')
	res = oracle.sumimage_accumulate(imgs, sm, flags, q)
	assert res['nimg'][3, 3] == 4 and res['nimg'][1, 0] == 3
	assert res['used'][0, 0] == 4 and res['used'][5, 5] == 5
	np.testing.assert_allclose(res['sumimage'][3, 3], np.mean([(imgs[k] - sm[k])[3, 3] for k in (0, 2, 3, 4)]), rtol=1e-6)


# ---- the committed golden vectors still describe this oracle ----------------------------------
@pytest.mark.parametrize('name', ['nontess', 'tess_small', 'crowded'])
def test_oracle_reproduces_golden(name, golden_dir):
	case = CASES[name]()
	g = load_golden(os.path.join(golden_dir, name + '.npz'))
	assert np.array_equal(images_digest(case['images']), g['images_sha256']), "synthetic inputs drifted; regenerate goldens"
	k = 0
	img = case['images'][k]
	extra = case['extra_mask'][k] if 'extra_mask' in case else None
	if case['kind'] == 'tess':
		b, m = oracle.fit_background(oracle.FFIImageLite(img, case['headers'][k], True), xycen=case['xycen'], extra_mask=extra, **case['fit_kwargs'])
	else:
		b, m = oracle.fit_background(img, extra_mask=extra)
	assert np.array_equal(m, g['mask'][k])
	assert in_tolerance(b, g['bkg'][k]).all()


# ---- further independent checks of the restated third-party pieces ------------------------------------
def test_sextractor_rule_hand_cases():
	"""photutils SExtractorBackground: 2.5 med - 1.5 mean; std == 0 -> mean; |mean - med| / std >= 0.3 -> med."""
	nan = np.nan
	rows = np.array([
		[1.0, 2.0, 3.0, 4.0, 10.0, nan],      # med 3, mean 4, std 3.16: |1| / 3.16 = 0.316 >= 0.3 -> med
		[1.0, 2.0, 3.0, 4.0, 5.5, nan],       # med 3, mean 3.1, std 1.56: 0.064 < 0.3 -> 2.5*3 - 1.5*3.1
		[7.0, 7.0, 7.0, nan, nan, nan],       # std 0 -> mean
	])
	bkg, med, mean, std = oracle.sextractor_background(rows)
	assert bkg[0] == 3.0 and med[0] == 3.0 and mean[0] == 4.0
	np.testing.assert_allclose(bkg[1], 2.5 * 3.0 - 1.5 * 3.1, rtol=1e-15)
	assert bkg[2] == 7.0 and std[2] == 0.0


def test_mesh_exclusion_threshold_is_inclusive_at_half():
	"""A mesh is kept iff its bad-pixel count (masked + clipped) is <= 50 % of 4096 (photutils exclude_percentile)."""
	rng = np.random.default_rng(4)
	data = (100 + rng.uniform(-1, 1, (64, 128))).astype('float64')   # uniform noise: 3 sigma never clips
	mask = np.zeros((64, 128), dtype=bool)
	mask.reshape(-1)[:0] = False
	mask[:32, :64] = True                  # exactly 2048 bad pixels in mesh 0 -> kept
	mask[:32, 64:] = True; mask[32, 64] = True   # 2049 in mesh 1 -> excluded, filled from mesh 0
	b = oracle.Background2DOracle(data, mask, box=64)
	assert b.mesh_good.tolist() == [[True, False]] and b.n_excluded == 1
	assert b.mesh_unfiltered[0, 1] == b.mesh_unfiltered[0, 0]       # single neighbour -> its value


def test_idw_tie_rules_agree_without_ties_and_differ_only_in_ties():
	from oracle.backgrounds_oracle import _idw_fill
	rng = np.random.default_rng(2)
	# jittered positions: all distances distinct -> the traversal-independent rule must equal scipy's cKDTree
	good = np.array([(iy, ix) for iy in range(6) for ix in range(7) if (iy, ix) not in [(2, 3), (2, 4), (3, 3)]], dtype='float64')
	jit = good + rng.uniform(-0.2, 0.2, good.shape)
	vals = rng.normal(100, 5, len(good))
	a = _idw_fill(jit, vals, 6, 7, 'ckdtree'); b = _idw_fill(jit, vals, 6, 7, 'stable')
	np.testing.assert_allclose(a, b, rtol=1e-14)
	# on the integer lattice the two rules pick different members of an equidistant shell at most
	a = _idw_fill(good, vals, 6, 7, 'ckdtree'); b = _idw_fill(good, vals, 6, 7, 'stable')
	gy, gx = good[:, 0].astype(int), good[:, 1].astype(int)
	np.testing.assert_array_equal(a[gy, gx], vals); np.testing.assert_array_equal(b[gy, gx], vals)   # good meshes keep their value
	assert np.all(np.abs(a - b) < 5 * 5)   # same shells, different tie members: bounded by the spread of the values


def test_zoom_restated_formula_matches_scipy():
	"""BkgZoomInterpolator: scipy zoom(order=3, mode='reflect', grid_mode=True) + clip; the CUDA kernels use the closed
	form u = (o + 0.5) / 64 - 0.5 with B-spline taps on prefiltered coefficients."""
	from scipy import ndimage
	rng = np.random.default_rng(6)
	mesh = rng.normal(100, 5, (5, 7))
	ref = ndimage.zoom(mesh, 64, order=3, mode='reflect', grid_mode=True)
	coef = ndimage.spline_filter(mesh, order=3, mode='reflect')    # not grid_mode aware: 'reflect' == half-sample symmetric
	def b3(t):
		t = np.abs(t)
		return np.where(t < 1, 2 / 3 - t * t + t ** 3 / 2, np.where(t < 2, (2 - t) ** 3 / 6, 0.0))
	def fold(i, n):
		i = np.where(i < 0, -i - 1, i)
		return np.where(i >= n, 2 * n - 1 - i, i)
	def axis_weights(n_out, n_in):
		u = (np.arange(n_out) + 0.5) / 64 - 0.5
		k0 = np.floor(u).astype(int) - 1
		idx = np.stack([fold(k0 + a, n_in) for a in range(4)], axis=1)
		w = np.stack([b3(u - (k0 + a)) for a in range(4)], axis=1)
		return idx, w
	iy, wy = axis_weights(5 * 64, 5); ix, wx = axis_weights(7 * 64, 7)
	out = np.einsum('ya,xb,yaxb->yx', wy, wx, coef[iy[:, :, None, None], ix[None, None, :, :]])
	np.testing.assert_allclose(out, ref, rtol=0, atol=2e-12)


def test_not_a_knot_spline_equivalence():
	"""InterpolatedUnivariateSpline(k=3) == not-a-knot cubic (what k_radial_fit solves); ext=3 clamps the abscissa."""
	from scipy.interpolate import InterpolatedUnivariateSpline, CubicSpline
	x = 2400 + 15 * np.arange(12) + 7.5
	y = np.log10(150 + 0.002 * (x - 2400) ** 2)
	s = InterpolatedUnivariateSpline(x, y, k=3, ext=3)
	c = CubicSpline(x, y, bc_type='not-a-knot')
	xx = np.linspace(x[0], x[-1], 500)
	np.testing.assert_allclose(s(xx), c(xx), rtol=0, atol=1e-13)
	assert s(x[0] - 100) == s(x[0]) and s(x[-1] + 100) == s(x[-1])
