"""
GPU parity tests (B200): the CUDA path, called through the Python host -> C ABI (include/tbk.h),
against (a) the committed golden vectors, (b) the live oracle on seeded inputs, (c) the reference's
own known answers and (d) size-independent properties at the full 2048 x 2048 size.

Tolerances (BASELINE.json north_star): masks exact; backgrounds within 1e-5 relative or 1e-3 e-/s.
"""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
	import photometry_b200 as pb
import oracle
from cases import CASES, load_golden, images_digest, in_tolerance, header


def _fit_case(case, **over):
	imgs = case['images']
	H, W = imgs.shape[1:]
	tess = case['kind'] == 'tess'
	kw = dict(case['fit_kwargs']); kw.update(over)
	fit = pb.BackgroundFitter((H, W), tess, case.get('camera', 0), case.get('ccd', 0), xycen=case.get('xycen'), **kw)
	cube = torch.from_numpy(imgs).cuda()
	meta = pb.meta_from_headers(case['headers']) if tess else None
	extra = torch.from_numpy(case['extra_mask']).cuda() if 'extra_mask' in case else None
	bkg, mask, status = fit.fit(cube, meta, extra)
	torch.cuda.synchronize()
	return fit, bkg.cpu().numpy(), mask.cpu().numpy().astype(bool), fit.status_to_numpy(status)


# ---- reference-authored known answers ---------------------------------------------------------
def test_background_fakeimg_public_api():
	"""reference tests/test_background.py:36-54 through the drop-in signature."""
	fakeimg = np.full([2048, 2048], 1000, dtype='float32')
	bck, mask = pb.fit_background(fakeimg)
	assert bck.shape == fakeimg.shape and mask.shape == fakeimg.shape
	assert bck.dtype == np.float64 and mask.dtype == bool
	assert np.all(np.isfinite(bck))
	assert not np.any(mask), "Nothing should be masked out"
	np.testing.assert_allclose(bck, 1000)


def test_manual_excludes_on_device():
	"""reference tests/test_pixel_flags.py:17-70 (Mars columns, whole-image zero, Earth-shine)."""
	H, W = 128, 2048
	rng = np.random.default_rng(0)
	img = (100 + rng.standard_normal((H, W))).astype('float32')
	fit = pb.BackgroundFitter((H, W), True, 1, 4)
	cube = torch.from_numpy(np.stack([img, img, np.zeros_like(img)])).cuda()
	hdrs = [header(1, 4, 0, cadenceno=4724, tstart=1330.0), header(1, 4, 0, cadenceno=11354, tstart=1400.0), header(1, 4, 0, cadenceno=20000, tstart=1500.0)]
	bkg, mask, st = fit.fit(cube, pb.meta_from_headers(hdrs))
	mask = mask.cpu().numpy().astype(bool); st = fit.status_to_numpy(st)
	assert np.all(mask[0][:, 1536:]) and not np.any(mask[0][:, :1536])
	assert np.all(mask[1]) and st['all_masked'][1] == 1 and torch.isnan(bkg[1]).all()
	assert np.all(mask[2]) and st['all_masked'][2] == 1 and torch.isnan(bkg[2]).all()
	assert st['all_masked'][0] == 0 and torch.isfinite(bkg[0]).all()


def test_error_conventions():
	with pytest.raises(ValueError):  # backgrounds.py:139-140
		pb.BackgroundFitter((128, 128), True, 5, 1)
	with pytest.raises(ValueError):  # io.py:84
		pb.fit_background(12345)
	with pytest.raises(ValueError):
		pb.BackgroundFitter((100, 128))
	bck, mask = pb.fit_background(np.full((128, 128), np.nan, dtype='float32'))  # backgrounds.py:101-102
	assert np.all(mask) and np.all(np.isnan(bck))


# ---- golden vectors -----------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['nontess', 'tess_small', 'mars', 'crowded', 'prepare'])
def test_fit_matches_golden(name, golden_dir):
	case = CASES[name]()
	g = load_golden(os.path.join(golden_dir, name + '.npz'))
	assert np.array_equal(images_digest(case['images']), g['images_sha256']), "synthetic inputs drifted; regenerate goldens"
	fit, bkg, mask, st = _fit_case(case)
	n = bkg.shape[0]
	assert np.array_equal(mask, g['mask']), "mask must be identical"
	ok = in_tolerance(bkg, g['bkg'])
	assert ok.all(), f"{(~ok).sum()} pixels outside tolerance"
	for k in range(n):
		rounds = g['mesh'].shape[1]
		assert st['rounds'][k] == rounds
		for r in range(rounds):
			s2, mesh = fit.debug_fetch(k, r)
			np.testing.assert_allclose(mesh, g['mesh'][k, r], rtol=1e-8)
			assert st['n_excluded'][k][r] == g['n_excluded'][k, r]
			if 's2' in g:
				assert np.array_equal(np.isnan(s2), np.isnan(g['s2'][k, r]))
				np.testing.assert_allclose(s2, g['s2'][k, r], rtol=0, atol=1e-9)  # no KDE argmax flip
				np.testing.assert_allclose(st['zeropoint'][k][r], g['zeropoint'][k, r], rtol=1e-9)


# ---- live oracle on fresh seeds -----------------------------------------------------------------
@pytest.mark.parametrize('seed', [31, 32, 33])
def test_fit_matches_oracle_fresh_seed(seed):
	from photometry_b200 import synth
	H, W = 320, 384
	xycen = (-15.0 - seed, 400.0)
	kw = dict(radial_cutoff=330, radial_pixel_step=15)
	stack = synth.synth_stack_numpy(1, H, W, seed=seed, xycen=xycen, radial_cutoff=330.0, n_stars=900, sky_level=80.0 + 40 * (seed % 3))
	case = dict(kind='tess', images=stack, camera=1, ccd=2, xycen=xycen, fit_kwargs=kw, headers=[header(1, 2, 0)])
	_, bkg, mask, st = _fit_case(case)
	rb, rm = oracle.fit_background(oracle.FFIImageLite(stack[0], case['headers'][0], True), xycen=xycen, **kw)
	assert np.array_equal(mask[0], rm)
	assert in_tolerance(bkg[0], rb).all()


def test_parameter_variants_match_oracle():
	"""bkgiters / radial_smooth / flux_cutoff are honoured like the reference keywords."""
	case = CASES['tess_small']()
	img, hdr = case['images'][0], case['headers'][0]
	for over in (dict(bkgiters=1), dict(bkgiters=2, radial_smooth=0), dict(flux_cutoff=500.0, radial_smooth=5)):
		c1 = dict(case); c1['images'] = case['images'][:1]; c1['headers'] = [hdr]
		_, bkg, mask, st = _fit_case(c1, **over)
		kw = dict(case['fit_kwargs']); kw.update(over)
		rb, rm = oracle.fit_background(oracle.FFIImageLite(img, hdr, True), xycen=case['xycen'], **kw)
		assert np.array_equal(mask[0], rm)
		assert in_tolerance(bkg[0], rb).all(), over


def test_mesh_exclusion_boundary():
	"""A mesh with exactly 2048 bad pixels is kept, 2049 is excluded (exclude_percentile=50)."""
	rng = np.random.default_rng(3)
	img = (300 + 5 * rng.standard_normal((256, 256))).astype('float32')
	extra = np.zeros((256, 256), dtype=bool)
	extra[0:32, 0:64] = True            # mesh (0,0): exactly 2048 masked -> clip may push it over
	extra[64:96, 64:128] = True; extra[96, 64] = True  # mesh (1,1): 2049 masked -> excluded
	extra[128:192, 128:192] = True      # mesh (2,2): fully masked
	fit = pb.BackgroundFitter((256, 256))
	bkg, mask, st = fit.fit(torch.from_numpy(img[None]).cuda(), None, torch.from_numpy(extra[None]).cuda())
	d = {}
	rb, rm = oracle.fit_background(img, extra_mask=extra, diagnostics=d)
	assert fit.status_to_numpy(st)['n_excluded'][0][0] == d['rounds'][0]['n_excluded'] >= 2
	assert np.array_equal(mask[0].cpu().numpy().astype(bool), rm)
	assert in_tolerance(bkg[0].cpu().numpy(), rb).all()


# ---- prepare-stage loops ------------------------------------------------------------------------
def test_prepare_stack_matches_golden(golden_dir):
	case = CASES['prepare']()
	g = load_golden(os.path.join(golden_dir, 'prepare.npz'))
	imgs = case['images']
	n, H, W = imgs.shape
	fit = pb.BackgroundFitter((H, W), True, case['camera'], case['ccd'], xycen=case['xycen'], **case['fit_kwargs'])
	res = pb.prepare_stack(fit, torch.from_numpy(imgs).cuda(), pb.meta_from_headers(case['headers']), time_smooth=case['time_smooth'], chunk=4)
	torch.cuda.synchronize()
	assert res.numfiles == n
	assert in_tolerance(res.backgrounds.cpu().numpy(), g['prep_backgrounds']).all()
	flux = res.images.cpu().numpy()
	assert np.array_equal(np.isnan(flux), np.isnan(g['prep_flux']))
	assert np.nanmax(np.abs(flux - g['prep_flux'])) <= 1e-3 + 1e-5 * np.nanmax(np.abs(g['prep_flux']))
	np.testing.assert_array_equal(res.nimg.cpu().numpy(), g['prep_nimg'])
	np.testing.assert_array_equal(res.used.cpu().numpy(), g['prep_used'])
	assert in_tolerance(res.sumimage.cpu().numpy(), g['prep_sumimage']).all()
	used_bits = np.packbits(res.backgrounds_pixels_used.cpu().numpy().astype(bool))
	np.testing.assert_array_equal(used_bits, g['prep_pixels_used_bits'])
	man = np.packbits((res.pixel_flags.cpu().numpy() & 2) != 0)
	np.testing.assert_array_equal(man, g['prep_pixel_flags_manexcl_bits'])


def test_time_smooth_bit_exact_and_sharded():
	"""Same float32 accumulation order as bottleneck.nanmean; halos reproduce the unsharded result."""
	rng = np.random.default_rng(4)
	n, H, W = 12, 64, 128
	bkg = rng.normal(100, 2, (n, H, W)).astype('float32')
	bkg[5, 3, 3] = np.nan; bkg[4:7, 9, 9] = np.nan
	fit = pb.BackgroundFitter((H, W))
	t = torch.from_numpy(bkg).cuda()
	for w in (1, 2, 4):   # 1 and 4 walk the cadence axis with the window in registers, 2 takes the frame-parallel kernel
		ref = oracle.time_smooth_backgrounds(bkg, 2 * w + 1)
		full = fit.time_smooth(t, w).cpu().numpy()
		np.testing.assert_array_equal(full, ref)
		lo = fit.time_smooth(t[:6], w, None, t[6:6 + w]).cpu().numpy()
		hi = fit.time_smooth(t[6:], w, t[6 - w:6], None).cpu().numpy()
		np.testing.assert_array_equal(np.concatenate([lo, hi]), ref)


@pytest.mark.gpu
def test_time_smooth_long_stack_all_windows():
	"""Stacks longer than one segment of the cadence walk, every window of the TESS cadences (w = 1, 4, 13), short halos."""
	rng = np.random.default_rng(41)
	n, H, W = 300, 64, 64
	bkg = rng.normal(100, 2, (n, H, W)).astype('float32')
	bkg[rng.uniform(size=bkg.shape) < 0.01] = np.nan
	bkg[100:140, 2, 5] = np.nan   # a pixel whose whole window is NaN for a while
	fit = pb.BackgroundFitter((H, W))
	t = torch.from_numpy(bkg).cuda()
	for w in (1, 4, 13):
		ref = oracle.time_smooth_backgrounds(bkg, 2 * w + 1)
		np.testing.assert_array_equal(fit.time_smooth(t, w).cpu().numpy(), ref)
		cut = 131
		lo = fit.time_smooth(t[:cut], w, None, t[cut:cut + w]).cpu().numpy()
		hi = fit.time_smooth(t[cut:], w, t[cut - w:cut], None).cpu().numpy()
		np.testing.assert_array_equal(np.concatenate([lo, hi]), ref)
		# halos shorter than the window (a neighbour that holds only a few frames)
		short = fit.time_smooth(t[:5], w, None, t[5:7]).cpu().numpy()
		ref_short = oracle.time_smooth_backgrounds(bkg[:7], 2 * w + 1)[:5]
		np.testing.assert_array_equal(short, ref_short)


def test_sum_accumulate_matches_oracle():
	rng = np.random.default_rng(6)
	n, H, W = 9, 64, 2048
	imgs = rng.normal(130, 3, (n, H, W)).astype('float32')
	imgs[2, 5, 5] = np.nan; imgs[3, 6, 6] = np.inf
	imgs[7] = 0  # whole image zero -> ManualExclude everywhere (pixel_flags.py:54-56)
	bkg = rng.normal(100, 1, (n, H, W)).astype('float32')
	flags = (rng.uniform(size=(n, H, W)) < 0.2).astype('uint8')
	hdrs = [header(1, 4, k, cadenceno=4720, dquality=(4 if k == 1 else 0), tstart=1330.0) for k in range(n)]  # k <= 4: Mars columns
	hdrs[8]['BACKAPP'] = True
	fit = pb.BackgroundFitter((H, W), True, 1, 4)
	ffis = [oracle.FFIImageLite(imgs[k], hdrs[k], True) for k in range(n)]
	fl_ref = flags.copy()
	for k in range(n):
		fl_ref[k][oracle.pixel_manual_exclude(ffis[k])] |= 2
	ref = oracle.sumimage_accumulate(imgs, bkg, fl_ref, np.array([h['DQUALITY'] for h in hdrs], dtype='int32'),
		backapp=np.array([bool(h.get('BACKAPP', False)) for h in hdrs]))
	d = lambda a: torch.from_numpy(a).cuda()
	s = torch.zeros((H, W), dtype=torch.float64, device='cuda'); ni = torch.zeros((H, W), dtype=torch.int32, device='cuda'); us = torch.zeros_like(ni)
	fl = d(flags.copy()); flux = torch.empty((n, H, W), dtype=torch.float32, device='cuda')
	fit.sum_accumulate(d(imgs), d(bkg), fl, pb.meta_from_headers(hdrs), s, ni, us, flux_out=flux)
	sumimage, used = fit.sum_finalize(s, ni, us, n, 0.5)
	np.testing.assert_array_equal(fl.cpu().numpy(), fl_ref)
	np.testing.assert_array_equal(flux.cpu().numpy(), ref['flux'])
	np.testing.assert_array_equal(ni.cpu().numpy(), ref['nimg'])
	np.testing.assert_array_equal(us.cpu().numpy(), ref['used'])
	np.testing.assert_array_equal(used.cpu().numpy().astype(bool), ref['backgrounds_pixels_used'])
	got = sumimage.cpu().numpy()
	assert np.array_equal(np.isnan(got), np.isnan(ref['sumimage']))
	fin = np.isfinite(ref['sumimage'])
	np.testing.assert_allclose(got[fin], ref['sumimage'][fin], rtol=1e-12)
	assert np.array_equal(got[~fin & ~np.isnan(got)], ref['sumimage'][~fin & ~np.isnan(got)])


# ---- full size (BASELINE configs): oracle spot-check + size-independent properties ---------------
@pytest.fixture(scope='module')
def full_stack():
	from photometry_b200 import synth
	n = 6
	cube = synth.synth_stack_torch(n, 2048, 2048, torch.device('cuda'), camera=1, ccd=2, seed=20260117 + 1)
	hdrs = [header(1, 2, k) for k in range(n)]
	fit = pb.BackgroundFitter((2048, 2048), True, 1, 2)
	bkg, mask, st = fit.fit(cube, pb.meta_from_headers(hdrs))
	torch.cuda.synchronize()
	return cube, hdrs, fit, bkg, mask, fit.status_to_numpy(st)


def test_full_size_matches_oracle(full_stack):
	cube, hdrs, fit, bkg, mask, st = full_stack
	for k in (0, 5):
		rb, rm = oracle.fit_background(oracle.FFIImageLite(cube[k].cpu().numpy(), hdrs[k], True))
		assert np.array_equal(mask[k].cpu().numpy().astype(bool), rm)
		ok = in_tolerance(bkg[k].cpu().numpy(), rb)
		assert ok.all(), f"FFI {k}: {(~ok).sum()} pixels outside tolerance"


def test_full_size_properties(full_stack):
	cube, hdrs, fit, bkg, mask, st = full_stack
	assert (st['rounds'] == 3).all() and (st['radial_ok'][:, :3] == 1).all() and (st['all_masked'] == 0).all()
	assert torch.isfinite(bkg).all()
	# mask definition (backgrounds.py:91-94) recomputed independently
	ref_mask = ~torch.isfinite(cube) | (cube > 8e4) | (cube < 0)
	assert torch.equal(mask.bool(), ref_mask)
	# batch independence and determinism: a single-FFI launch gives the same bits as the batched one
	b1, m1, _ = fit.fit(cube[3:4].clone(), pb.meta_from_headers(hdrs[3:4]))
	assert torch.equal(m1[0], mask[3])
	assert torch.allclose(b1[0], bkg[3], rtol=1e-6, atol=0)
	# masked pixels never influence the result: overwrite them with other masked values
	alt = cube[2:3].clone()
	alt[0][mask[2].bool()] = float('nan')
	b2, m2, _ = fit.fit(alt, pb.meta_from_headers(hdrs[2:3]))
	assert torch.equal(m2[0], mask[2]) and torch.allclose(b2[0], bkg[2], rtol=1e-6, atol=0)
	# the background is smooth at mesh scale: bounded by the unmasked data range, no NaN leakage
	assert float(bkg.min()) > 0 and float(bkg.max()) < 8e4


def test_repeat_run_bit_identical(full_stack):
	"""
	Nothing in the chain depends on timing: the fallback and retry queues are filled in arbitrary order but every mesh is
	evaluated on its own, the KDE bins fixed-point sums, the tail / zone lists are ballot-ordered.  Two runs give the same bits,
	on the caller's stream alone and with the chunks spread over several streams.
	"""
	cube, hdrs, fit, bkg, mask, st = full_stack
	meta = pb.meta_from_headers(hdrs)
	b1, m1, _ = fit.fit(cube, meta)
	assert torch.equal(b1, bkg) and torch.equal(m1, mask)
	b2 = torch.empty_like(cube); m2 = torch.empty(cube.shape, dtype=torch.uint8, device=cube.device)
	fit.fit_stack(cube, fit.meta_to_device(meta), b2, m2, chunk=2, nstreams=3)
	torch.cuda.synchronize()
	assert torch.equal(b2, bkg) and torch.equal(m2, mask)


# ---- independent implementations agree --------------------------------------------------------
VARIANTS = ('0', '3', '6', '7')


def test_kernel_variants_agree(monkeypatch):
	"""
	TBK_TILE_KERNEL selects independent implementations of the mesh statistics / zeropoint / ring gather
	(0: generic CTA-per-mesh kernels with iterated histogram selection and full passes; 3: block-cooperative bucketed
	kernels with the keys staged in shared memory; 6: the zone kernels -- bulk moments + tail / zone lists, one warp per mesh --
	with the bucketed kernels as their fallback, the default).  They must give the same statistics bit for bit in the median and to
	rounding in mean / std, and the same backgrounds.
	"""
	case = CASES['tess_small']()
	imgs = case['images']
	H, W = imgs.shape[1:]
	cube = torch.from_numpy(imgs).cuda()
	meta = pb.meta_from_headers(case['headers'])
	res = {}
	for variant in VARIANTS:
		monkeypatch.setenv('TBK_TILE_KERNEL', variant)
		fit = pb.BackgroundFitter((H, W), True, case['camera'], case['ccd'], xycen=case['xycen'], **case['fit_kwargs'])
		bkg, mask, st = fit.fit(cube, meta)
		ws = fit.debug_workspace()
		res[variant] = (bkg.cpu().numpy(), mask.cpu().numpy(), ws['tile_base'].copy(), ws['tile_nf'].copy(), fit.status_to_numpy(st).copy())
	ref = res['0']
	for variant in VARIANTS[1:]:
		got = res[variant]
		assert np.array_equal(got[1], ref[1])
		for idx, (a, b) in enumerate(((got[2], ref[2]), (got[3], ref[3]))):
			assert np.array_equal(a['nfin'], b['nfin'])
			ok = b['nfin'] > 0
			if idx == 0:
				# raw float32 pixels: the medians are exact order statistics of identical inputs
				assert np.array_equal(a['med'][ok], b['med'][ok])
			else:
				# residuals x - radial: the zeropoint (hence radial) differs at rounding level between variants, and the zone
				# kernels evaluate the radial profile from degree-6 Taylor pieces (relative error < 1e-13 of the profile value)
				np.testing.assert_allclose(a['med'][ok], b['med'][ok], rtol=1e-10, atol=5e-10)
			np.testing.assert_allclose(a['mean'][ok], b['mean'][ok], rtol=1e-10 if idx else 1e-11, atol=5e-10 if idx else 1e-11)
			np.testing.assert_allclose(a['std'][ok], b['std'][ok], rtol=1e-9)
		np.testing.assert_allclose(got[4]['zeropoint'], ref[4]['zeropoint'], rtol=1e-12)
		assert in_tolerance(got[0], ref[0]).all()
		np.testing.assert_allclose(got[0], ref[0], rtol=1e-6)


def test_fit_stack_streams_match_single_launch():
	"""Chunked multi-stream driver == one launch over the whole stack."""
	case = CASES['prepare']()
	imgs = case['images']
	n, H, W = imgs.shape
	fit = pb.BackgroundFitter((H, W), True, case['camera'], case['ccd'], xycen=case['xycen'], **case['fit_kwargs'])
	cube = torch.from_numpy(imgs).cuda()
	meta = pb.meta_from_headers(case['headers'])
	b0, m0, _ = fit.fit(cube, meta)
	b1 = torch.empty_like(cube); m1 = torch.empty(cube.shape, dtype=torch.uint8, device='cuda')
	fit.fit_stack(cube, meta, b1, m1, chunk=2, nstreams=3)
	torch.cuda.synchronize()
	assert torch.equal(m0, m1)
	assert torch.allclose(b0, b1, rtol=1e-6, atol=0)
	# and through host buffers (the end-to-end path)
	hb = torch.empty((n, H, W), dtype=torch.float32).pin_memory(); hm = torch.empty((n, H, W), dtype=torch.uint8).pin_memory()
	for pack in (True, False):   # mask sent as bits and expanded by host threads / sent as bytes
		hb.zero_(); hm.fill_(7)
		h2d, d2h = pb.fit_stack_host(fit, torch.from_numpy(imgs).pin_memory(), meta, hb, hm, chunk=4, pack_mask=pack)
		torch.cuda.synchronize()
		assert torch.equal(hm, m0.cpu()) and torch.allclose(hb, b0.cpu(), rtol=1e-6, atol=0)
		assert h2d == n * H * W * 4 and d2h == n * H * W * 4 + (n * H * W // 8 if pack else n * H * W)


def test_mask_pack_unpack_round_trip():
	"""tbk_pack_mask (device) / tbk_unpack_mask_host: the order of numpy.packbits, any non-zero byte counts as set."""
	import ctypes as C
	from photometry_b200 import _lib
	lib = _lib.load()
	rng = np.random.default_rng(8)
	m = (rng.uniform(size=64 * 4096) < 0.3).astype('uint8') * rng.integers(1, 255, 64 * 4096).astype('uint8')
	md = torch.from_numpy(m).cuda()
	bits = torch.empty(m.size // 8, dtype=torch.uint8, device='cuda')
	_lib.check(lib.tbk_pack_mask(C.c_void_p(md.data_ptr()), m.size, C.c_void_p(bits.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'tbk_pack_mask')
	got = bits.cpu().numpy()
	np.testing.assert_array_equal(got, np.packbits(m != 0))
	out = np.empty(m.size, dtype='uint8')
	_lib.check(lib.tbk_unpack_mask_host(got.ctypes.data_as(C.c_void_p), got.size, out.ctypes.data_as(C.c_void_p)), 'tbk_unpack_mask_host')
	np.testing.assert_array_equal(out, (m != 0).astype('uint8'))


# ---- edge cases: shapes, degenerate rings, ties, heavy masking -----------------------------------
def _compare_with_oracle(img, hdr=None, xycen=None, extra=None, **kw):
	tess = hdr is not None
	H, W = img.shape
	fit = pb.BackgroundFitter((H, W), tess, hdr['CAMERA'] if tess else 0, hdr['CCD'] if tess else 0, xycen=xycen, **kw)
	cube = torch.from_numpy(img[None].copy()).cuda()
	ex = torch.from_numpy(extra[None].copy()).cuda() if extra is not None else None
	bkg, mask, st = fit.fit(cube, pb.meta_from_headers([hdr]) if tess else None, ex)
	torch.cuda.synchronize()
	st = fit.status_to_numpy(st)[0]
	d = {}
	if tess:
		rb, rm = oracle.fit_background(oracle.FFIImageLite(img, hdr, True), xycen=xycen, extra_mask=extra, diagnostics=d, **kw)
	else:
		rb, rm = oracle.fit_background(img, extra_mask=extra, diagnostics=d, **kw)
	assert np.array_equal(mask[0].cpu().numpy().astype(bool), rm)
	ok = in_tolerance(bkg[0].cpu().numpy(), rb)
	assert ok.all(), f"{(~ok).sum()} pixels outside tolerance"
	return st, d


@pytest.mark.parametrize('shape', [(64, 64), (64, 192), (192, 128)])
def test_small_and_non_square_shapes(shape):
	rng = np.random.default_rng(shape[0] + shape[1])
	img = (120 + 0.05 * np.arange(shape[1])[None, :] + 6 * rng.standard_normal(shape)).astype('float32')
	img[10:14, 20:26] += 3000
	_compare_with_oracle(img)


def test_quantised_pixels_with_many_ties():
	"""Integer-valued pixels: hundreds of equal values per mesh exercise the large-bin exact selection."""
	rng = np.random.default_rng(8)
	img = np.round(100 + 3 * rng.standard_normal((256, 256))).astype('float32')
	img[:64, :64] = 77.0                   # a constant mesh
	img[64:128, :64] = np.round(50 + 0.4 * rng.standard_normal((64, 64)))  # only 3-4 distinct values
	img[200:230, 100:140] += 800
	_compare_with_oracle(img)
	hdr = header(1, 2, 0)
	_compare_with_oracle(img, hdr, xycen=(-20.0, 300.0), radial_cutoff=230, radial_pixel_step=15)


def test_heavily_contaminated_meshes():
	"""Meshes where 45 % of the pixels are bright: clipping runs all five iterations on a bimodal mesh."""
	rng = np.random.default_rng(9)
	img = (300 + 10 * rng.standard_normal((256, 256))).astype('float32')
	sel = rng.uniform(size=(256, 256)) < 0.45
	img[sel] += rng.uniform(50, 4000, sel.sum()).astype('float32')
	_compare_with_oracle(img)


def test_radial_profile_degenerate_cases():
	"""Too few finite rings -> no radial component (backgrounds.py:192-197); rings with 0 / 1 samples -> NaN."""
	rng = np.random.default_rng(10)
	H, W = 192, 192
	img = (150 + 5 * rng.standard_normal((H, W))).astype('float32')
	hdr = header(2, 3, 0)
	# (a) only three rings exist: the spline needs four points ("m must be > k") -> radial = 0 every round
	st, d = _compare_with_oracle(img, hdr, xycen=(-10.0, 230.0), radial_cutoff=295, radial_pixel_step=15)
	assert (st['radial_ok'][:3] == 0).all() and not any(r['radial_ok'] for r in d['rounds'])
	# (b) several rings masked out entirely by the extra mask, one ring left with a single pixel
	xycen = (-10.0, 230.0)
	r, bins, cen = oracle.radial_geometry((H, W), xycen, 200, 10)
	extra = (r >= bins[3]) & (r < bins[6])
	ring7 = np.argwhere((r >= bins[7]) & (r < bins[8]))
	extra[(r >= bins[7]) & (r < bins[8])] = True
	extra[ring7[0][0], ring7[0][1]] = False
	st, d = _compare_with_oracle(img, hdr, xycen=xycen, extra=extra, radial_cutoff=200, radial_pixel_step=10)
	assert st['radial_ok'][0] == 1 and np.isnan(d['rounds'][0]['s2_raw'][7]) and np.isnan(d['rounds'][0]['s2_raw'][4])


@pytest.mark.parametrize('kw', [dict(bkgiters=5), dict(radial_smooth=7, radial_pixel_step=8), dict(flux_cutoff=250.0)])
def test_more_parameter_variants(kw):
	case = CASES['tess_small']()
	kw2 = dict(case['fit_kwargs']); kw2.update(kw)
	_compare_with_oracle(case['images'][1], case['headers'][1], xycen=case['xycen'], **kw2)


@pytest.mark.parametrize('step', [45, 60])
def test_wide_rings_take_the_unstaged_kde_path(step):
	"""
	Rings of more than 16,384 pixels do not fit the 16-bit bin staging of k_ring_kde and go through the sweep that
	recomputes the bins (radial_pixel_step = 45 at full size: up to 57 k pixels per ring); with step 60 the first
	rings exceed 65,535 samples and the linear binning switches from the fixed-point counters to float64 adds.
	Small rings elsewhere in this file cover the staged / fixed-point paths.
	"""
	from photometry_b200 import synth
	img = synth.synth_stack_numpy(1, 2048, 2048, camera=1, ccd=2, seed=314)[0]
	hdr = header(1, 2, 0)
	fit = pb.BackgroundFitter((2048, 2048), True, 1, 2, radial_pixel_step=step)
	assert fit.nrings < 20
	st, d = _compare_with_oracle(img, hdr, radial_pixel_step=step)
	assert st['rounds'] == 3 and (st['radial_ok'][:3] == 1).all()


@pytest.mark.parametrize('seed', list(range(40, 48)))
def test_randomised_fields(seed):
	"""Random backgrounds / star densities / NaN and mask fractions, TESS and non-TESS, against the live oracle."""
	rng = np.random.default_rng(seed)
	H, W = int(rng.choice([128, 192, 256])), int(rng.choice([192, 256, 320]))
	level = float(rng.uniform(20, 3000))
	img = level * (1 + 0.3 * rng.uniform(-1, 1) * np.linspace(-1, 1, W)[None, :] + 0.2 * rng.uniform(-1, 1) * np.linspace(-1, 1, H)[:, None])
	img = img + rng.standard_normal((H, W)) * np.sqrt(img + 4)
	nstar = int(rng.integers(0, 400))
	ys, xs = rng.integers(0, H, nstar), rng.integers(0, W, nstar)
	img[ys, xs] += 10 ** rng.uniform(1, 5.2, nstar)
	img[rng.uniform(size=(H, W)) < rng.choice([0, 1e-3, 0.05])] = np.nan
	img = img.astype('float32')
	extra = (rng.uniform(size=(H, W)) < rng.choice([0, 0.2, 0.45])) if seed % 2 else None
	if seed % 3 == 0:
		_compare_with_oracle(img, extra=extra)
	else:
		xycen = (-float(rng.uniform(5, 40)), float(H + rng.uniform(10, 60)))
		rmax = np.hypot(W + 44 - xycen[0], xycen[1])
		_compare_with_oracle(img, header(int(rng.integers(1, 5)), int(rng.integers(1, 4)), 0), xycen=xycen, extra=extra,
			radial_cutoff=float(0.75 * rmax), radial_pixel_step=float(rng.choice([8, 12, 15])), radial_smooth=int(rng.choice([0, 3, 5])))


# ---- input side: FITS(.gz) -> device cube ----------------------------------------------------------
def test_load_ffi_stack_from_fits(tmp_path):
	"""Device-side big-endian decode + science crop equals the host decode of io.FFIImage (io.py:46-52)."""
	import gzip
	from test_host_logic import _hdu
	from photometry_b200.io import FFIImage
	rng = np.random.default_rng(12)
	paths = []
	for k in range(3):
		raw = rng.normal(100 + k, 3, (2078, 2136)).astype('float32')
		raw[5, 50] = np.nan
		prim = _hdu([('SIMPLE', True), ('BITPIX', 8), ('NAXIS', 0), ('EXTEND', True), ('TELESCOP', 'TESS'), ('CAMERA', 2), ('CCD', 3)])
		ext = [('XTENSION', 'IMAGE'), ('BITPIX', -32), ('NAXIS', 2), ('NAXIS1', 2136), ('NAXIS2', 2078), ('PCOUNT', 0), ('GCOUNT', 1),
			('TSTART', 1400.0 + 0.02 * k), ('TSTOP', 1400.02 + 0.02 * k), ('FFIINDEX', 9000 + k), ('DQUALITY', 0)]
		blob = prim + _hdu(ext, raw) + _hdu(ext[:7], np.ones((2078, 2136), dtype='float32'))
		path = str(tmp_path / f'tess2018{k:09d}-s0001-2-3-0120-s_ffic.fits') + ('.gz' if k != 1 else '')
		with (gzip.open(path, 'wb', compresslevel=1) if path.endswith('.gz') else open(path, 'wb')) as fid:
			fid.write(blob)
		paths.append(path)
	cube, headers = pb.load_ffi_stack(paths, threads=2, batch=2)
	torch.cuda.synchronize()
	assert cube.shape == (3, 2048, 2048) and cube.dtype == torch.float32
	for k, path in enumerate(paths):
		ref = FFIImage(path)
		assert ref.is_tess
		np.testing.assert_array_equal(cube[k].cpu().numpy(), ref.data)   # NaN == NaN by position in assert_array_equal
		assert headers[k]['FFIINDEX'] == 9000 + k and headers[k]['CAMERA'] == 2 and headers[k]['CCD'] == 3
	# and the stack runs through the fit
	fit = pb.BackgroundFitter((2048, 2048), True, 2, 3)
	bkg, mask, st = fit.fit(cube, pb.meta_from_headers(headers))
	assert torch.isfinite(bkg).all() and int(mask[0, 5, 6]) == 1 and int(mask.sum()) == 3


# ---- device math --------------------------------------------------------------------------------
def test_device_log10():
	"""
	The ring samples use a table-driven float64 log10 (tbk_common.cuh:tbk_log10).  Bound its error against
	numpy's log10 (itself < 1 ulp): at most 2 ulp apart anywhere on the range the path produces (pix + zp >= 1,
	up to the flux cutoff scale) and on wide-range / special arguments, and exact at the powers of ten it can hit.
	"""
	import ctypes as C
	from photometry_b200 import _lib
	lib = _lib.load()
	rng = np.random.default_rng(5)
	x = np.concatenate([
		1.0 + rng.uniform(0, 1e5, 400000),                 # the working range
		np.exp(rng.uniform(np.log(0.5), np.log(2.0), 200000)),   # around 1 (every table interval)
		np.exp(rng.uniform(-200, 200, 100000)),            # wide range
		np.nextafter(2.0 ** np.arange(-20, 21), np.inf), np.nextafter(2.0 ** np.arange(-20, 21), 0),
		10.0 ** np.arange(0, 16), [1.0, 0.6875, 1.375, np.nextafter(1.375, 0), 5e-324, 0.0, -1.0, np.inf, np.nan]])
	xin = torch.from_numpy(x).cuda()
	out = torch.empty_like(xin)
	rc = lib.tbk_debug_log10(C.c_void_p(xin.data_ptr()), C.c_void_p(out.data_ptr()), x.size, None)
	assert rc == 0
	torch.cuda.synchronize()
	got = out.cpu().numpy()
	with np.errstate(all='ignore'):
		ref = np.log10(x)
	fin = np.isfinite(ref)
	assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~fin & ~np.isnan(ref)], ref[~fin & ~np.isnan(ref)])
	ulp = np.spacing(np.abs(ref[fin]))
	err = np.abs(got[fin] - ref[fin]) / ulp
	near_zero = np.abs(ref[fin]) < 1e-3          # next to 1 the absolute error is what matters (|log10| -> 0)
	assert err[~near_zero].max() <= 2.0, err[~near_zero].max()
	assert np.abs(got[fin] - ref[fin])[near_zero].max() < 1e-18
	p10 = 10.0 ** np.arange(0, 16)
	sel = np.isin(x, p10)
	assert np.abs(got[sel] - np.round(ref[sel])).max() <= 4.5e-16 * 15


# ---- the prepare driver on the GPU engine -----------------------------------------------------------
def test_driver_gpu_products_equal_prepare_stack(tmp_path):
	"""
	prepare_photometry (FITS files -> resumable product store, GpuEngine) writes exactly what prepare_stack computes for
	the same stack held on the device: smoothed backgrounds, flags, images, image errors, sum image, used-pixel map.
	"""
	import gzip
	from test_host_logic import _hdu
	from photometry_b200 import prepare_driver, synth
	from photometry_b200.store import open_store
	n = 5
	stack = synth.synth_stack_numpy(n, 2048, 2048, camera=1, ccd=4, seed=3, n_stars=4000)
	paths = []
	for k in range(n):
		raw = np.full((2078, 2136), 50.0, dtype='float32')
		raw[0:2048, 44:2092] = stack[k]
		err = np.full((2078, 2136), 2.0 + k, dtype='float32')
		prim = _hdu([('SIMPLE', True), ('BITPIX', 8), ('NAXIS', 0), ('EXTEND', True), ('TELESCOP', 'TESS'), ('CAMERA', 1), ('CCD', 4), ('DATA_REL', 1)])
		ext = [('XTENSION', 'IMAGE'), ('BITPIX', -32), ('NAXIS', 2), ('NAXIS1', 2136), ('NAXIS2', 2078), ('PCOUNT', 0), ('GCOUNT', 1),
			('TSTART', 1330.0 + 0.02 * k), ('TSTOP', 1330.02 + 0.02 * k), ('FFIINDEX', 4700 + k), ('DQUALITY', 32 if k == 2 else 0), ('BARYCORR', 0.002)]
		path = str(tmp_path / f'tess2018{k:09d}-s0001-1-4-0120-s_ffic.fits.gz')
		with gzip.open(path, 'wb', compresslevel=1) as fid:
			fid.write(prim + _hdu(ext, raw) + _hdu(ext[:7], err))
		paths.append(path)
	out = prepare_driver.prepare_photometry(str(tmp_path), sectors=1, cameras=1, ccds=4, store_backend='npy', batch=2)
	assert len(out) == 1
	# the same stack through the device-resident path (camera 1 / ccd 4 before cadence 4724: Mars columns -> ManualExclude + IDW fill)
	hdrs = [dict(CAMERA=1, CCD=4, TSTART=1330.0 + 0.02 * k, TSTOP=1330.02 + 0.02 * k, FFIINDEX=4700 + k, DQUALITY=32 if k == 2 else 0) for k in range(n)]
	fit = pb.BackgroundFitter((2048, 2048), True, 1, 4)
	res = pb.prepare_stack(fit, torch.from_numpy(stack).cuda(), pb.meta_from_headers(hdrs), time_smooth=3, chunk=2)
	with open_store(out[0], 'a', 'npy') as hdf:
		for k in range(n):
			name = f'{k:04d}'
			assert np.array_equal(np.asarray(hdf['backgrounds'][name]), res.backgrounds[k].cpu().numpy(), equal_nan=True)
			assert np.array_equal(np.asarray(hdf['pixel_flags'][name]), res.pixel_flags[k].cpu().numpy())
			assert np.array_equal(np.asarray(hdf['images'][name]), res.images[k].cpu().numpy(), equal_nan=True)
			e = np.asarray(hdf['images_err'][name])
			assert np.isnan(e[:, 1536:]).all() and (e[:, :1536] == 2.0 + k).all()          # ManualExclude -> NaN (prepare.py:423-425)
		assert np.allclose(np.asarray(hdf['sumimage']), res.sumimage.cpu().numpy(), rtol=1e-12, equal_nan=True)
		assert np.array_equal(np.asarray(hdf['backgrounds_pixels_used']).astype(np.uint8), res.backgrounds_pixels_used.cpu().numpy())
		assert list(np.asarray(hdf['cadenceno'])) == [4700 + k for k in range(n)] and int(np.asarray(hdf['quality'])[2]) == 32
		assert hdf.require_group('images').attrs['CAMERA'] == 1 and hdf.require_group('images').attrs['CADENCE'] == 1800


# ---- catalog-driven star mask (extension) ------------------------------------------------------------
def test_star_mask_equals_oracle_and_feeds_the_fit():
	from photometry_b200.starmask import star_mask, star_radius
	rng = np.random.default_rng(8)
	H, W = 256, 320
	cat = np.column_stack([rng.uniform(-5, W + 5, 400), rng.uniform(-5, H + 5, 400), rng.uniform(4, 16, 400)])
	got = star_mask((H, W), cat).cpu().numpy().astype(bool)
	ref = oracle.star_mask((H, W), cat)
	assert np.array_equal(got, ref) and 0.02 < ref.mean() < 0.6
	assert np.allclose(star_radius([10.0, 7.0, 20.0, 0.0]), [4.0, 4.0 * 10 ** 0.3, 1.5, 40.0])
	assert np.array_equal(oracle.star_radius([10.0, 7.0, 20.0, 0.0]), star_radius([10.0, 7.0, 20.0, 0.0]))
	# through the drop-in: fit_background(image, catalog) == oracle with the same mask as extra_mask
	img = (300 + 20 * rng.standard_normal((H, W))).astype('float32')
	for sx, sy, tm in cat[:60]:
		if 0 <= sx < W and 0 <= sy < H:
			img[int(sy), int(sx)] += 5e4 * 10 ** (-0.4 * (tm - 6))
	bkg, mask = pb.fit_background(img, catalog=cat)
	rb, rm = oracle.fit_background(img, extra_mask=ref)
	assert np.array_equal(mask, rm) and in_tolerance(bkg, rb).all()
	bkg0, mask0 = pb.fit_background(img)                 # default: the catalog is not used, as in the reference
	assert not mask0.any() or mask0.sum() < mask.sum()
