"""
Background-shenanigans stage (photometry/pixel_flags.py:61-79, photometry/prepare.py:514-622).

CPU tests: the oracle against an independent brute-force restatement and hand-computed answers, the golden vector,
the host logic (shuffle, slabs, the two-rank re-sharding over gloo).  GPU tests (marked): the CUDA path through the
C ABI against the oracle -- indicator images bit-exact (float32), the robust mean bit-exact (float64, same summation
order), flags exact.
"""
import os
import socket
import numpy as np
import pytest
import torch
import oracle
from oracle import shenanigans_oracle as so
from cases import case_shenanigans, images_digest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'shenanigans.npz')


def brute_indicator(img, sumimage=None):
	"""
	Independent restatement: reflect-pad, take every 15 x 15 window, nan-median of the float32-rounded differences
	(float64 mean of the two middle values for an even count), float32 result.
	"""
	d = np.asarray(img, dtype='float64')
	if sumimage is not None:
		d = d - sumimage
	d = d.astype('float32')
	pad = np.pad(d, 7, mode='symmetric')       # numpy 'symmetric' == scipy 'reflect' (half-sample)
	win = np.lib.stride_tricks.sliding_window_view(pad, (15, 15)).reshape(d.shape + (225,))
	import warnings
	with warnings.catch_warnings():
		warnings.simplefilter('ignore', RuntimeWarning)
		return np.nanmedian(win.astype('float64'), axis=2).astype('float32')


# ---- oracle -----------------------------------------------------------------------------------------
def test_oracle_median_filter_matches_brute_force():
	rng = np.random.default_rng(1)
	img = rng.normal(0, 10, (40, 53)).astype('float32')
	sm = rng.normal(0, 1, (40, 53))
	ind = so.indicator_stack(img[None], sm)[0]
	assert ind.dtype == np.float32
	assert np.array_equal(ind, brute_indicator(img, sm))
	# SumImage=None and the plain function (float64 out, pixel_flags.py:74-79)
	f = so.pixel_background_shenanigans(img)
	assert f.dtype == np.float32 or f.dtype == np.float64
	assert np.array_equal(f.astype('float32'), brute_indicator(img))
	# constant image -> constant
	assert np.all(so.pixel_background_shenanigans(np.full((20, 20), 3.5)) == 3.5)


def test_shuffled_order_is_the_legacy_generator():
	# np.random.seed(0); np.random.shuffle(list(range(10))) -- value fixed by NumPy's frozen legacy stream
	assert so.shuffled_order(10).tolist() == [2, 8, 4, 9, 1, 6, 7, 3, 0, 5]
	state = np.random.get_state()
	np.random.seed(0); ref = list(range(1340)); np.random.shuffle(ref)
	np.random.set_state(state)
	assert so.shuffled_order(1340).tolist() == ref
	from photometry_b200.shenanigans import shuffled_order
	assert shuffled_order(1340).tolist() == ref and shuffled_order(10).dtype == np.int32


def test_mean_shenanigans_block_semantics():
	# 27 images of one pixel with value = cadence index: blocks are order[0:25] and order[25:27]; the second block
	# still holds order[2:25] of the first in its trailing slots (buffer allocated once, prepare.py:564-569)
	n = 27
	ind = np.arange(n, dtype='float32').reshape(n, 1, 1)
	order = so.shuffled_order(n)
	b0 = np.median(order[:25].astype('float64'))
	b1 = np.median(np.concatenate([order[25:27], order[2:25]]).astype('float64'))
	assert so.mean_shenanigans(ind)[0, 0] == (b0 + b1) / 2
	# shorter than a block: unused slots are zeros; NaN-only pixels count as 0
	ind = np.full((3, 1, 2), 8.0, dtype='float32'); ind[:, 0, 1] = np.nan
	m = so.mean_shenanigans(ind)
	assert m[0, 0] == 0.0 and m[0, 1] == 0.0     # median of (8, 8, 8, 22 zeros) = 0; nanmedian of 22 zeros + 3 NaN = 0
	ind = np.full((13, 1, 1), 8.0, dtype='float32')
	assert so.mean_shenanigans(ind)[0, 0] == 8.0  # 13 of 25 slots -> the median is 8


def test_flagging_rule():
	ind = np.array([[[0.0, 50.0, -50.0, 40.0, np.nan]]], dtype='float32')
	mean = np.array([[0.0, 5.0, -5.0, 0.0, 0.0]])
	flags = np.array([[[4, 1, 2, 7, 5]]], dtype='uint8')
	out = so.flag_shenanigans(ind, mean, flags)
	assert out.dtype == np.uint8 and out.tolist() == [[[0, 5, 6, 3, 1]]]   # strict '>' (prepare.py:603), old bit cleared


def test_oracle_matches_golden():
	case = case_shenanigans()
	g = np.load(GOLDEN)
	assert np.array_equal(images_digest(case['images']), g['images_sha256'])
	flags, mean, ind = so.background_shenanigans(case['images'], case['sumimage'], case['pixel_flags'])
	assert np.array_equal(ind[[0, 4, 17, 29]], g['indicator'])
	assert np.array_equal(mean, g['mean'])
	assert np.array_equal(np.packbits((flags & 4) != 0), g['flag_bits'])
	assert np.array_equal(np.packbits((flags & 3) != 0), g['flags_other'])     # other bits untouched
	n_flagged = ((flags & 4) != 0).sum(axis=(1, 2))
	assert n_flagged[4] > 500 and n_flagged[5] > 200 and n_flagged[17] > 100 and n_flagged[9] == 0 and n_flagged[0] == 0


# ---- host logic / two ranks over gloo -----------------------------------------------------------------
def _free_port():
	s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
	return port


def _slab_worker(rank, world, port, out_dir):
	import torch.distributed as dist
	os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
	dist.init_process_group('gloo', rank=rank, world_size=world)
	from photometry_b200.prepare import shard_bounds
	from photometry_b200.shenanigans import cadence_to_row_slabs
	n, H, W = 7, 9, 5
	full = torch.arange(n * H * W, dtype=torch.float32).reshape(n, H, W)
	lo, hi = shard_bounds(n, world, rank)
	slab = cadence_to_row_slabs(full[lo:hi].clone())
	np.save(os.path.join(out_dir, f'slab_{rank}.npy'), slab.numpy())
	dist.destroy_process_group()


def test_two_rank_cadence_to_row_slabs(tmp_path):
	import torch.multiprocessing as mp
	from photometry_b200.shenanigans import slab_bounds
	world = 2
	mp.spawn(_slab_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
	n, H, W = 7, 9, 5
	full = np.arange(n * H * W, dtype='float32').reshape(n, H, W)
	for r in range(world):
		lo, hi = slab_bounds(H, world, r)
		assert np.array_equal(np.load(tmp_path / f'slab_{r}.npy'), full[:, lo:hi])


# ---- CUDA path ---------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize('shape', [(96, 112), (15, 15), (7, 200), (131, 64), (70, 129)])
def test_indicator_bit_exact(shape):
	import photometry_b200 as pb
	rng = np.random.default_rng(sum(shape))
	imgs = rng.normal(0, 20, (3,) + shape).astype('float32')
	imgs[1] = np.round(imgs[1])                    # many ties
	sm = rng.normal(0, 3, shape)
	got = pb.shenanigans_indicator(torch.from_numpy(imgs).cuda(), torch.from_numpy(sm).cuda()).cpu().numpy()
	assert np.array_equal(got, so.indicator_stack(imgs, sm))
	got = pb.shenanigans_indicator(torch.from_numpy(imgs).cuda()).cpu().numpy()
	assert np.array_equal(got, so.indicator_stack(imgs, None))


@gpu
def test_indicator_nan_windows_and_drop_in():
	import photometry_b200 as pb
	rng = np.random.default_rng(9)
	img = rng.normal(0, 20, (90, 140)).astype('float32')
	img[rng.integers(0, 90, 40), rng.integers(0, 140, 40)] = np.nan
	img[:, 100:] = np.nan                           # a manually excluded region (prepare.py:414)
	img[40:60, 20:40] = np.nan                      # a hole larger than the window -> all-NaN windows
	sm = rng.normal(0, 3, (90, 140)); sm[3, 3] = np.nan
	got = pb.shenanigans_indicator(torch.from_numpy(img[None]).cuda(), torch.from_numpy(sm).cuda())[0].cpu().numpy()
	ref = brute_indicator(img, sm)
	assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(got).any()
	assert np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
	# windows without NaN agree with the reference's scipy call
	clean = ~so.nan_affected(img[None], sm)[0]
	assert clean.any() and np.array_equal(got[clean], so.indicator_stack(img[None], sm)[0][clean])
	# drop-in function (pixel_flags.py:61-79): float64 out, float32 and float64 input
	ok = rng.normal(0, 20, (64, 80))
	off = rng.normal(0, 2, (64, 80))
	f = pb.pixel_background_shenanigans(ok, SumImage=off)
	assert f.dtype == np.float64 and np.array_equal(f.astype('float32'), so.pixel_background_shenanigans(ok, off).astype('float32'))
	f32 = pb.pixel_background_shenanigans(ok.astype('float32'))
	assert np.array_equal(f32.astype('float32'), so.pixel_background_shenanigans(ok.astype('float32')))
	with pytest.raises(ValueError):
		pb.pixel_background_shenanigans(np.zeros((3, 4, 5)))


@gpu
@pytest.mark.parametrize('n', [3, 25, 27, 30, 61])
def test_mean_and_flags_exact(n):
	import photometry_b200 as pb
	rng = np.random.default_rng(n)
	ind = rng.normal(0, 30, (n, 20, 33)).astype('float32')
	ind[rng.uniform(size=ind.shape) < 0.05] = np.nan
	ind[:, 5, 5] = np.nan
	ind[:, 6, :] = np.round(ind[:, 6, :] / 10) * 10     # ties
	d_ind = torch.from_numpy(ind).cuda()
	mean = pb.mean_shenanigans(d_ind)
	ref_mean = so.mean_shenanigans(ind)
	assert np.array_equal(mean.cpu().numpy(), ref_mean)
	flags = (rng.uniform(size=ind.shape) < 0.3).astype('uint8') * 7
	d_flags = torch.from_numpy(flags).cuda()
	pb.flag_shenanigans(d_ind, mean, d_flags, threshold=40)
	assert np.array_equal(d_flags.cpu().numpy(), so.flag_shenanigans(ind, ref_mean, flags, 40))


@gpu
def test_stage_matches_golden_and_oracle():
	import photometry_b200 as pb
	case = case_shenanigans()
	g = np.load(GOLDEN)
	flags = torch.from_numpy(case['pixel_flags'].copy()).cuda()
	mean, ind = pb.background_shenanigans(torch.from_numpy(case['images']).cuda(), torch.from_numpy(case['sumimage']).cuda(),
		flags, return_indicator=True)
	assert np.array_equal(ind.cpu().numpy()[[0, 4, 17, 29]], g['indicator'])
	assert np.array_equal(mean.cpu().numpy(), g['mean'])
	f = flags.cpu().numpy()
	assert np.array_equal(np.packbits((f & 4) != 0), g['flag_bits']) and np.array_equal(np.packbits((f & 3) != 0), g['flags_other'])


@gpu
def test_full_size_indicator():
	"""One 2048 x 2048 frame against scipy, plus properties that hold at any size: negation symmetry, a constant offset
	in SumImage moves the indicator by exactly that constant for dyadic values, a constant image stays constant."""
	import photometry_b200 as pb
	g = torch.Generator(device='cuda'); g.manual_seed(5)
	img = torch.randn((2, 2048, 2048), device='cuda', generator=g) * 25
	img[1] = torch.round(img[1])
	ind = pb.shenanigans_indicator(img)
	assert torch.equal(pb.shenanigans_indicator(-img), -ind)
	shifted = pb.shenanigans_indicator(img[1:], torch.full((2048, 2048), -4.0, dtype=torch.float64, device='cuda'))
	assert torch.equal(shifted[0], ind[1] + 4.0)
	assert torch.equal(pb.shenanigans_indicator(torch.full((1, 2048, 2048), 7.25, device='cuda')), torch.full((1, 2048, 2048), 7.25, device='cuda'))
	ref = so.pixel_background_shenanigans(img[0].cpu().numpy())
	assert np.array_equal(ind[0].cpu().numpy(), ref.astype('float32'))
