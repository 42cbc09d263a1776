"""
CPU tests of the kd-tree restatement behind the mesh IDW fill (photometry_b200/csrc/tbk_kdtree.cuh), run on the host
through ``tbk_debug_idw_neighbors`` and compared with the REAL ``scipy.spatial.cKDTree`` -- the library photutils'
``ShepardIDWInterpolator`` calls (photometry/backgrounds.py:200-205).  The same header is compiled into
``k_mesh_finalize``; the GPU tests then check the filled meshes against the oracle (which calls scipy).
"""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from photometry_b200 import _lib
import oracle
from oracle.backgrounds_oracle import _idw_fill, PHOTUTILS_IDW_LEAFSIZE
from cases import CASES, in_tolerance


def ours(good):
	lib = _lib.load()
	ny, nx = good.shape
	g = np.ascontiguousarray(good.astype(np.uint8).ravel())
	ng = int(g.sum())
	nid = np.empty((ny * nx, 10), np.int32)
	nd2 = np.empty((ny * nx, 10), np.int32)
	idx = np.empty(max(ng, 1), np.int32)
	nodes = np.empty((2 * ng + 2, 4), np.int32)
	nn = np.zeros(1, np.int32)
	_lib.check(lib.tbk_debug_idw_neighbors(g.ctypes.data, ny, nx, nid.ctypes.data, nd2.ctypes.data, idx.ctypes.data,
		nodes.ctypes.data, nn.ctypes.data), 'tbk_debug_idw_neighbors')
	return nid, nd2, idx[:ng], nodes[:nn[0]]


def scipy_ref(good, leafsize=PHOTUTILS_IDW_LEAFSIZE):
	ny, nx = good.shape
	flat = np.flatnonzero(good.ravel())
	gy, gx = np.divmod(flat, nx)
	tree = cKDTree(np.column_stack([gy, gx]).astype('float64'), leafsize=leafsize)
	pos = np.column_stack(np.unravel_index(np.arange(ny * nx), (ny, nx))).astype('float64')
	d, i = tree.query(pos, k=10, eps=0.0)
	ids = np.where(i < len(flat), flat[np.minimum(i, len(flat) - 1)], -1)
	return ids, d, tree


def patterns():
	rng = np.random.default_rng(20261017)
	shapes = [(32, 32), (8, 8), (6, 6), (16, 16), (4, 8), (32, 32), (64, 64), (3, 40), (1, 25)]
	for trial in range(180):
		ny, nx = shapes[trial % len(shapes)]
		mode = trial % 6
		good = np.ones((ny, nx), bool)
		if mode == 0:
			good[:, (3 * nx) // 4:] = False                      # the Mars pattern: the last quarter of the mesh columns
		elif mode == 1:
			good = rng.random((ny, nx)) > 0.3
		elif mode == 2:
			good = rng.random((ny, nx)) > 0.7
		elif mode == 3:
			good = rng.random((ny, nx)) > 0.1
			y0, x0 = rng.integers(0, max(ny - 3, 1)), rng.integers(0, max(nx - 3, 1))
			good[y0:y0 + 4, x0:x0 + 4] = False                   # a hole of 4 x 4 meshes
		elif mode == 4:
			good = rng.random((ny, nx)) > 0.95                    # fewer than ten good meshes happens here
		else:
			good[rng.integers(0, ny)] = False                    # a whole mesh row
			good[:, rng.integers(0, nx)] = False
		if good.sum() == 0:
			good[ny // 2, nx // 2] = True
		yield good


def test_tree_and_neighbours_equal_scipy():
	"""Index permutation (hence every leaf's content and order), neighbour ids in query order, and distances."""
	n = 0
	for good in patterns():
		nid, nd2, idx, nodes = ours(good)
		ids, d, tree = scipy_ref(good)
		assert np.array_equal(idx, tree.indices)
		assert np.array_equal(nid, ids)
		dd = np.where(nd2 >= 0, np.sqrt(np.maximum(nd2, 0).astype('float64')), np.inf)
		assert np.array_equal(dd, d)
		n += 1
	assert n == 180


def test_tree_structure_equals_scipy():
	good = np.ones((32, 32), bool)
	good[:, 24:] = False
	_, _, _, nodes = ours(good)

	def walk(node, k):
		dim, split, a, b = nodes[k]
		assert dim == node.split_dim
		if dim == -1:
			assert (a, b) == (node.start_idx, node.end_idx)
			return
		assert split == node.split
		walk(node.lesser, a)
		walk(node.greater, b)
	walk(scipy_ref(good)[2].tree, 0)


def test_leafsize_matters_and_is_photutils_value():
	"""
	The tie decisions depend on the leaf size: scipy's default (16) and photutils' (10) pick different neighbour SETS on
	the Mars lattice, so the oracle must pass photutils' value -- and the CUDA path must restate a leafsize-10 tree.
	"""
	good = np.ones((32, 32), bool)
	good[:, 24:] = False
	ids10, _, _ = scipy_ref(good, 10)
	ids16, _, _ = scipy_ref(good, 16)
	differ = sum(set(a) != set(b) for a, b in zip(ids10, ids16))
	assert differ > 100
	assert PHOTUTILS_IDW_LEAFSIZE == 10
	nid, _, _, _ = ours(good)
	assert np.array_equal(nid, ids10)


def test_idw_fill_equals_oracle_fill():
	"""Shepard sums over the restated neighbours reproduce the oracle's (scipy-driven) fill to rounding."""
	rng = np.random.default_rng(3)
	for good in list(patterns())[:40]:
		ny, nx = good.shape
		flat = np.flatnonzero(good.ravel())
		vals = rng.normal(100, 5, flat.size)
		gy, gx = np.divmod(flat, nx)
		ref = _idw_fill(np.column_stack([gy, gx]).astype('float64'), vals, ny, nx, 'ckdtree')
		nid, nd2, _, _ = ours(good)
		full = np.full(ny * nx, np.nan)
		full[flat] = vals
		got = np.empty(ny * nx)
		for t in range(ny * nx):
			m = nid[t] >= 0
			if nd2[t][0] == 0:
				got[t] = full[nid[t][0]]
				continue
			w = 1.0 / np.sqrt(nd2[t][m].astype('float64'))
			got[t] = np.sum(w * full[nid[t][m]]) / np.sum(w)
		np.testing.assert_allclose(got.reshape(ny, nx), ref, rtol=1e-14)


@pytest.mark.parametrize('name', ['mars', 'crowded'])
def test_stable_rule_deviation(name):
	"""
	Quantifies what the round-1 neighbour rule "(distance, mesh index)" cost against the reference's cKDTree order: a large
	fraction of the unmasked pixels leaves the 1e-5 / 1e-3 tolerance on the cases with excluded meshes.  (The CUDA path no
	longer uses that rule; this test documents why.)
	"""
	case = CASES[name]()
	img = case['images'][0]
	extra = case['extra_mask'][0] if 'extra_mask' in case else None
	ffi = oracle.FFIImageLite(img, case['headers'][0], True)
	res = {}
	for mode in ('ckdtree', 'stable'):
		res[mode] = oracle.fit_background(ffi, xycen=case['xycen'], extra_mask=extra, idw=mode, **case['fit_kwargs'])
	mask = res['ckdtree'][1]
	assert np.array_equal(mask, res['stable'][1])
	ok = in_tolerance(res['stable'][0], res['ckdtree'][0])
	frac_out = 1.0 - ok[~mask].mean()
	print(f"{name}: {100 * frac_out:.1f} % of the unmasked pixels outside tolerance under the 'stable' rule")
	assert frac_out > 0.2
