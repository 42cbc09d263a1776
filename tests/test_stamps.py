"""
Consumer-side cube loads (photometry/BasePhotometry.py:720-751): StampServer / tbk_gather_stamps against the literal
restatement in oracle/cube_oracle.py.  Pure data movement: every cube must be bit-identical.
"""
import numpy as np
import pytest
import torch
import oracle


def test_oracle_load_cube_layout():
	frames = [np.arange(6 * 50, dtype='float32').reshape(6, 50) + 1000 * k for k in range(3)]
	cube = oracle.load_cube(frames, (2, 5, 46, 50))      # CCD columns 46..50 -> array columns 2..6 (offset 44)
	assert cube.shape == (3, 4, 3) and cube.dtype == np.float32
	assert cube[1, 2, 2] == frames[2][3, 4] and cube[0, 0, 0] == frames[0][2, 2]
	assert np.array_equal(cube, np.stack([f[2:5, 2:6] for f in frames], axis=2))


def test_stamp_server_rejects_host_tensors():
	import photometry_b200 as pb
	with pytest.raises(ValueError):
		pb.StampServer(images=torch.zeros((2, 8, 8)))
	with pytest.raises(ValueError):
		pb.StampServer()


@pytest.mark.gpu
def test_cubes_bit_identical():
	import photometry_b200 as pb
	rng = np.random.default_rng(3)
	N, H, W = 75, 96, 160
	images = rng.normal(0, 50, (N, H, W)).astype('float32'); images[3, 10, 10] = np.nan
	bkg = rng.normal(100, 5, (N, H, W)).astype('float32')
	flags = rng.integers(0, 8, (N, H, W)).astype('uint8')
	srv = pb.StampServer(images=torch.from_numpy(images).cuda(), backgrounds=torch.from_numpy(bkg).cuda(),
		pixel_flags=torch.from_numpy(flags).cuda())
	stamps = [(0, 11, 44, 55), (85, 96, 44 + 149, 44 + 160), (20, 21, 44 + 7, 44 + 8), (0, 96, 44, 204), (5, 45, 44 + 30, 44 + 101),
		(33, 50, 44 + 64, 44 + 96)]
	stamps += [tuple(int(v) for v in (r, r + h, 44 + c, 44 + c + w)) for r, h, c, w in
		zip(rng.integers(0, 60, 20), rng.integers(1, 36, 20), rng.integers(0, 100, 20), rng.integers(1, 60, 20))]
	for group, stack in (('images', images), ('backgrounds', bkg), ('pixel_flags', flags)):
		cubes = srv.load_cubes(stamps, group)
		for st, cube in zip(stamps, cubes):
			ref = oracle.load_cube(list(stack), st)
			got = cube.cpu().numpy()
			assert got.shape == ref.shape and got.dtype == ref.dtype
			assert np.array_equal(got, ref, equal_nan=(group == 'images'))
	one = srv.images_cube(stamps[4]).cpu().numpy()
	assert np.array_equal(one, oracle.load_cube(list(images), stamps[4]), equal_nan=True)
	assert srv.pixelflags_cube(stamps[1]).dtype == torch.uint8
	# a group that does not exist -> NaN cube (BasePhotometry.py:736-737)
	miss = srv.images_err_cube(stamps[0])
	assert miss.shape == (11, 11, N) and bool(torch.isnan(miss).all())
	for bad in [(0, 0, 44, 50), (-1, 5, 44, 50), (0, 97, 44, 50), (0, 5, 43, 50), (0, 5, 44, 205), (5, 3, 44, 50)]:
		with pytest.raises(ValueError):
			srv.load_cube(bad)


@pytest.mark.gpu
def test_many_stamps_full_size():
	"""2,000 stamps over a 2048 x 2048 x 40 stack: every cube equals the strided slice of the stack."""
	import photometry_b200 as pb
	g = torch.Generator(device='cuda'); g.manual_seed(1)
	stack = torch.randn((40, 2048, 2048), device='cuda', generator=g)
	srv = pb.StampServer(images=stack)
	rng = np.random.default_rng(8)
	r0 = rng.integers(0, 2048 - 40, 2000); c0 = rng.integers(0, 2048 - 40, 2000)
	hh = rng.integers(5, 40, 2000); ww = rng.integers(5, 40, 2000)
	stamps = np.stack([r0, r0 + hh, c0 + 44, c0 + ww + 44], axis=1)
	cubes = srv.load_cubes(stamps)
	for i in (0, 1, 999, 1999) + tuple(rng.integers(0, 2000, 40)):
		ref = stack[:, r0[i]:r0[i] + hh[i], c0[i]:c0[i] + ww[i]].permute(1, 2, 0)
		assert torch.equal(cubes[i], ref)
