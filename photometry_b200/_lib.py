"""
ctypes binding of libtbk.so (include/tbk.h).  There is no CPU fallback: if the library is missing
or a call fails, an exception is raised.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get('TBK_LIBPATH') or os.path.join(_HERE, 'lib', 'libtbk.so')   # override: development builds only

TBK_MAX_ROUNDS = 8
KERNEL_CLASSES = ('tile_base', 'tile_round', 'zp_min', 'ring_gather', 'ring_kde', 'radial_fit', 'mesh', 'final', 'fallback', 'misc')


class TbkError(RuntimeError):
	pass


class FFIMeta(C.Structure):
	"""tbk_ffi_meta"""
	_fields_ = [('cadenceno', C.c_int32), ('dquality', C.c_int32), ('backapp', C.c_int32),
		('reserved', C.c_int32), ('tstart', C.c_double), ('tstop', C.c_double)]


META_DTYPE = np.dtype([('cadenceno', '<i4'), ('dquality', '<i4'), ('backapp', '<i4'), ('reserved', '<i4'),
	('tstart', '<f8'), ('tstop', '<f8')])
STATUS_DTYPE = np.dtype([('all_masked', '<i4'), ('no_good_mesh', '<i4'), ('n_valid', '<i4'), ('rounds', '<i4'),
	('n_excluded', '<i4', (TBK_MAX_ROUNDS,)), ('n_ring_valid', '<i4', (TBK_MAX_ROUNDS,)),
	('radial_ok', '<i4', (TBK_MAX_ROUNDS,)), ('zeropoint', '<f8', (TBK_MAX_ROUNDS,))])

TILESTAT_DTYPE = np.dtype([('mean', '<f8'), ('med', '<f8'), ('std', '<f8'), ('nfin', '<i4'), ('pad', '<i4')])
CTL_DTYPE = np.dtype([('min_bits', '<u4'), ('any_nonzero', '<i4'), ('n_valid', '<i4'), ('mars', '<i4'), ('earth', '<i4'),
	('all_masked', '<i4'), ('no_good_mesh', '<i4'), ('radial_ok', '<i4'), ('npts', '<i4'), ('mesh_const', '<i4'), ('kde_fallbacks', '<i4'), ('pad0', '<i4'),
	('min_key', '<u8'), ('min_ub', '<u8'), ('zp', '<f8'), ('c_flat', '<f8'), ('x0', '<f8'), ('xlast', '<f8'), ('mesh_min', '<f8'), ('mesh_max', '<f8'),
	('kx', '<f8', (128,)), ('pp', '<f8', (128, 4)), ('seg', '<f8', (128, 6)), ('seg_of_ring', '<i2', (128,))])

# name -> (restype, argtypes); every symbol include/tbk.h declares
_p = C.c_void_p
SIGNATURES = {
	'tbk_plan_create': (C.c_int, [C.POINTER(_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
		C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.c_int]),
	'tbk_plan_destroy': (C.c_int, [_p]),
	'tbk_plan_num_rings': (C.c_int, [_p]),
	'tbk_workspace_bytes': (C.c_size_t, [_p, C.c_int]),
	'tbk_fit_batch': (C.c_int, [_p, _p, C.c_int, _p, _p, _p, _p, _p, _p, _p]),
	'tbk_time_smooth': (C.c_int, [_p, _p, C.c_int, C.c_int, _p, C.c_int, _p, C.c_int, _p, _p]),
	'tbk_sum_accumulate': (C.c_int, [_p, _p, _p, _p, _p, C.c_int, _p, _p, _p, _p, _p]),
	'tbk_pack_mask': (C.c_int, [_p, C.c_size_t, _p, _p]),
	'tbk_unpack_mask_host': (C.c_int, [_p, C.c_size_t, _p]),
	'tbk_sum_finalize': (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_double, _p, _p, _p]),
	'tbk_debug_fetch': (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p, _p]),
	'tbk_bkgshe_indicator': (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p, _p]),
	'tbk_bkgshe_mean': (C.c_int, [_p, C.c_size_t, C.c_size_t, C.c_int, _p, _p, _p]),
	'tbk_bkgshe_flag': (C.c_int, [_p, _p, C.c_int, C.c_size_t, C.c_double, C.c_int, _p, _p]),
	'tbk_gather_stamps': (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, C.c_int, _p, _p]),
	'tbk_debug_log10': (C.c_int, [_p, _p, C.c_int, _p]),
	'tbk_decode_ffi_be': (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p]),
	'tbk_fit_batch_profiled': (C.c_int, [_p, _p, C.c_int, _p, _p, _p, _p, _p, _p, _p, C.POINTER(C.c_float)]),
	'tbk_launch_count': (C.c_ulonglong, []),
	'tbk_workspace_layout': (C.c_int, [_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
	'tbk_debug_idw_neighbors': (C.c_int, [_p, C.c_int, C.c_int, _p, _p, _p, _p, _p]),
	'tbk_star_mask': (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p]),
	'tbk_motion_prepare': (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p, _p]),
	'tbk_motion_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
	'tbk_motion_ecc': (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _p, _p, _p]),
	'tbk_last_error': (C.c_char_p, []),
	'tbk_version': (C.c_int, []),
}

_lib = None


def load():
	"""Load libtbk.so (raises TbkError when it has not been built)."""
	global _lib
	if _lib is not None:
		return _lib
	if not os.path.exists(LIBPATH):
		raise TbkError(f"{LIBPATH} not found: run `python -m photometry_b200.build` (there is no CPU fallback)")
	lib = C.CDLL(LIBPATH)
	for name, (res, args) in SIGNATURES.items():
		fn = getattr(lib, name)
		fn.restype = res
		fn.argtypes = args
	_lib = lib
	return lib


def check(rc, what):
	if rc != 0:
		msg = load().tbk_last_error().decode('utf-8', 'replace')
		if rc == -1:
			raise ValueError(msg)
		raise TbkError(f"{what} failed ({rc}): {msg}")
