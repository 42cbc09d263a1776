"""CPU tests (no GPU): C-ABI surface, host-side packing, sharding arithmetic, FITS reader, loud failure."""
import ctypes
import os
import re
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
	import __graft_entry__
	__graft_entry__.build()
	from photometry_b200 import _lib
	hdr = open(os.path.join(ROOT, 'include', 'tbk.h')).read()
	declared = set(re.findall(r'\b(tbk_[a-z0-9_]+)\s*\(', hdr))
	assert {'tbk_plan_create', 'tbk_fit_batch', 'tbk_time_smooth', 'tbk_sum_accumulate', 'tbk_sum_finalize'} <= declared
	lib = ctypes.CDLL(_lib.LIBPATH)
	for name in declared:
		assert hasattr(lib, name), f"{name} declared in include/tbk.h but not exported"
	assert set(_lib.SIGNATURES) == declared
	assert _lib.load().tbk_version() == 1


def test_struct_layouts_match_header():
	from photometry_b200 import _lib
	assert _lib.META_DTYPE.itemsize == ctypes.sizeof(_lib.FFIMeta) == 32
	assert _lib.STATUS_DTYPE.itemsize == 4 * 4 + 3 * 8 * 4 + 8 * 8


def test_meta_packing():
	from photometry_b200 import meta_from_headers
	m = meta_from_headers([dict(TSTART=1.0, TSTOP=2.0, FFIINDEX=4700, DQUALITY=32), dict(TSTART=3.0, TSTOP=4.0, BACKAPP=True)])
	assert m['cadenceno'][0] == 4700 and m['cadenceno'][1] == 2**31 - 1
	assert m['dquality'][0] == 32 and m['backapp'][1] == 1 and m['tstop'][1] == 4.0


def test_manual_exclude_rule_and_quality():
	from photometry_b200.pixel_flags import manual_exclude_rule
	from photometry_b200 import TESSQualityFlags, PixelQualityFlags
	assert manual_exclude_rule(True, 1, 4, 4724, 1400.0, 1400.02) == 'mars'
	assert manual_exclude_rule(True, 1, 4, 9000, 1325.0, 1325.02) == 'mars'
	assert manual_exclude_rule(True, 1, 2, 11354, 1400.0, 1400.02) == 'earth'
	assert manual_exclude_rule(True, 1, 2, 9000, 1464.1, 1464.12) == 'earth'
	assert manual_exclude_rule(True, 2, 4, 4000, 1400.0, 1400.02) is None
	assert manual_exclude_rule(False, 1, 4, 4000, 1400.0, 1400.02) is None
	assert TESSQualityFlags.DEFAULT_BITMASK == 4335
	assert TESSQualityFlags.filter(16) and not TESSQualityFlags.filter(32)
	assert PixelQualityFlags.filter(np.array([0, 1, 2, 3])).tolist() == [True, True, False, False]


def test_shard_bounds_cover_axis():
	from photometry_b200.prepare import shard_bounds
	for n in (7, 1340, 4000):
		for ws in (1, 2, 4, 8):
			b = [shard_bounds(n, ws, r) for r in range(ws)]
			assert b[0][0] == 0 and b[-1][1] == n
			assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
			assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def _card(key, val):
	if isinstance(val, bool):
		v = 'T' if val else 'F'
		return f"{key:<8}= {v:>20}".ljust(80)
	if isinstance(val, str):
		q = "'" + val.ljust(8) + "'"
		return f"{key:<8}= {q}".ljust(80)
	return f"{key:<8}= {val!r:>20}".ljust(80)


def _hdu(cards, data=None):
	txt = ''.join(_card(k, v) for k, v in cards) + 'END'.ljust(80)
	txt = txt.ljust((len(txt) + 2879) // 2880 * 2880)
	out = txt.encode('ascii')
	if data is not None:
		raw = data.astype('>f4').tobytes()
		out += raw + b'\0' * ((-len(raw)) % 2880)
	return out


def test_fits_reader_tess_layout(tmp_path):
	"""io.py:46-52: TESS detection, science crop [0:2048, 44:2092], merged headers."""
	from photometry_b200.io import FFIImage
	rng = np.random.default_rng(0)
	raw = rng.normal(100, 1, (2078, 2136)).astype('float32')
	err = np.ones((2078, 2136), dtype='float32')
	prim = _hdu([('SIMPLE', True), ('BITPIX', 8), ('NAXIS', 0), ('EXTEND', True), ('TELESCOP', 'TESS'), ('CAMERA', 1), ('CCD', 4)])
	ext = [('XTENSION', 'IMAGE'), ('BITPIX', -32), ('NAXIS', 2), ('NAXIS1', 2136), ('NAXIS2', 2078), ('PCOUNT', 0), ('GCOUNT', 1),
		('TSTART', 1325.5), ('TSTOP', 1325.52), ('FFIINDEX', 4710), ('DQUALITY', 0)]
	path = str(tmp_path / 'tess-s0001-1-4-0120-s_ffic.fits')
	with open(path, 'wb') as fid:
		fid.write(prim + _hdu(ext, raw) + _hdu(ext[:7], err))
	img = FFIImage(path)
	assert img.is_tess and img.shape == (2048, 2048) and img.data.dtype == np.float32
	np.testing.assert_array_equal(img.data, raw[0:2048, 44:2092])
	assert img.header['CAMERA'] == 1 and img.header['CCD'] == 4 and img.header['FFIINDEX'] == 4710
	with pytest.raises(ValueError):
		FFIImage(12345)
	arr = FFIImage(raw[:64, :64])
	assert not arr.is_tess and arr.header == {}


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_product_path_fails_loudly_without_gpu():
	"""No CPU fallback: without a CUDA device the public entry point raises instead of computing."""
	import photometry_b200 as pb
	from photometry_b200._lib import TbkError
	with pytest.raises((TbkError, RuntimeError)):
		pb.fit_background(np.full((128, 128), 1000, dtype='float32'))


def test_abi_rejects_bad_arguments_without_touching_the_gpu():
	"""Every entry point validates its arguments first and reports TBK_ERR_INVALID (-1) with a message: no CUDA call is made."""
	import ctypes as C
	from photometry_b200 import _lib
	lib = _lib.load()
	null = None
	one = C.c_void_p(16)   # a non-null, aligned dummy; never dereferenced because another argument is invalid
	plan = C.c_void_p()
	assert lib.tbk_plan_create(C.byref(plan), 100, 64, 1, 1, 2, 8e4, 3, 2400.0, 15.0, 3, None, 0) == -1      # H not a multiple of 64
	assert b'tbk_plan_create' in lib.tbk_last_error() or len(lib.tbk_last_error()) > 0
	assert lib.tbk_plan_create(C.byref(plan), 2048, 2048, 1, 9, 9, 8e4, 3, 2400.0, 15.0, 3, None, 0) == -1    # unknown camera / ccd
	assert lib.tbk_fit_batch(null, one, 1, one, null, one, one, one, one, null) == -1
	assert lib.tbk_time_smooth(null, one, 1, 1, null, 0, null, 0, one, null) == -1
	assert lib.tbk_bkgshe_indicator(null, null, 1, 64, 64, one, null) == -1
	assert lib.tbk_bkgshe_indicator(one, null, 0, 64, 64, one, null) == -1
	assert lib.tbk_bkgshe_mean(one, 10, 20, 3, one, one, null) == -1            # stride < npix
	assert lib.tbk_bkgshe_flag(one, one, 1, 64, 40.0, 0, one, null) == -1       # bit out of range
	assert lib.tbk_gather_stamps(one, 2, 4, 64, 64, one, one, 1, one, null) == -1   # element size must be 1 or 4
	assert lib.tbk_decode_ffi_be(one, 1, 2136, 2078, 0, 44, 2048, 2050, one, null) == -1   # W % 4
	assert lib.tbk_debug_log10(null, one, 4, null) == -1
	assert lib.tbk_debug_fetch(null, one, 1, 0, 0, null, null) == -1
	assert b'tbk_debug_fetch' in lib.tbk_last_error()


def test_host_mask_unpack_matches_numpy():
	"""tbk_unpack_mask_host (the host half of the bit-packed mask transport) is numpy.unpackbits; no GPU involved."""
	import ctypes as C
	from photometry_b200 import _lib
	lib = _lib.load()
	rng = np.random.default_rng(3)
	for nbytes in (1, 7, 64, 4096 * 3 + 5):
		bits = rng.integers(0, 256, nbytes).astype('uint8')
		out = np.full(nbytes * 8 + 8, 9, dtype='uint8')
		rc = lib.tbk_unpack_mask_host(bits.ctypes.data_as(C.c_void_p), nbytes, out.ctypes.data_as(C.c_void_p))
		assert rc == 0
		np.testing.assert_array_equal(out[:nbytes * 8], np.unpackbits(bits))
		assert (out[nbytes * 8:] == 9).all()   # nothing written past the end
	assert lib.tbk_unpack_mask_host(None, 8, None) != 0
