"""
Drop-in for ``photometry.backgrounds`` (photometry/backgrounds.py:52-211) on one B200.

``fit_background(image, ...) -> (bkg, mask)`` keeps the reference signature and return types.  The work
is done by :class:`BackgroundFitter`, which owns a ``tbk_plan`` (include/tbk.h) and runs the batched
CUDA path on a device-resident FFI cube.  PyTorch is used for device memory and streams only.
"""
import ctypes as C
import functools
import numpy as np
import torch
from . import _lib
from ._lib import META_DTYPE, STATUS_DTYPE, check
from .io import FFIImage

INT32_MAX = 2**31 - 1


def make_meta(n, cadenceno=None, dquality=None, backapp=None, tstart=None, tstop=None):
	"""Pack per-FFI header scalars into the ``tbk_ffi_meta`` layout (numpy structured array)."""
	m = np.zeros(n, dtype=META_DTYPE)
	if cadenceno is None:
		m['cadenceno'] = INT32_MAX  # header without FFIINDEX: the reference uses inf (pixel_flags.py:36)
	else:
		cad = np.asarray(cadenceno, dtype='float64')
		cad = np.where(np.isfinite(cad), cad, INT32_MAX)
		m['cadenceno'] = np.clip(cad, -INT32_MAX, INT32_MAX).astype('int32')
	m['dquality'] = 0 if dquality is None else dquality
	m['backapp'] = 0 if backapp is None else backapp
	m['tstart'] = np.nan if tstart is None else tstart
	m['tstop'] = np.nan if tstop is None else tstop
	return m


def meta_from_headers(headers):
	"""``tbk_ffi_meta`` array from a list of FFI headers (dict-like)."""
	n = len(headers)
	return make_meta(n,
		cadenceno=np.array([h.get('FFIINDEX', np.inf) for h in headers], dtype='float64'),
		dquality=np.array([int(h.get('DQUALITY', 0)) for h in headers]),
		backapp=np.array([1 if h.get('BACKAPP', False) else 0 for h in headers]),
		tstart=np.array([h.get('TSTART', np.nan) for h in headers], dtype='float64'),
		tstop=np.array([h.get('TSTOP', np.nan) for h in headers], dtype='float64'))


def _ptr(t):
	return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class BackgroundFitter:
	"""
	Batched ``fit_background`` for one (image shape, camera, ccd, parameter set) on one GPU.

	Parameters mirror the keyword arguments of the reference function; ``is_tess=False`` gives the
	ndarray-input semantics (no radial component, a single round).  ``xycen`` overrides the
	camera-centre table (test hook for small images).
	"""
	def __init__(self, shape, is_tess=False, camera=0, ccd=0, flux_cutoff=8e4, bkgiters=3,
		radial_cutoff=2400, radial_pixel_step=15, radial_smooth=3, xycen=None, device=None):
		if not torch.cuda.is_available():
			raise _lib.TbkError("CUDA device required: photometry_b200 has no CPU fallback")
		self.lib = _lib.load()
		self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
		self.H, self.W = int(shape[0]), int(shape[1])
		self.is_tess = bool(is_tess)
		self.bkgiters = int(bkgiters) if self.is_tess else 1
		self._plan = C.c_void_p(0)
		xy = (C.c_double * 2)(*xycen) if xycen is not None else None
		check(self.lib.tbk_plan_create(C.byref(self._plan), self.H, self.W, int(self.is_tess),
			int(camera or 0), int(ccd or 0), float(flux_cutoff), int(bkgiters), float(radial_cutoff),
			float(radial_pixel_step), int(radial_smooth or 0), xy, self.device.index), 'tbk_plan_create')
		self.nrings = self.lib.tbk_plan_num_rings(self._plan)
		self.ntiles = (self.H // 64) * (self.W // 64)
		self._ws = {}
		self._last = None

	def close(self):
		if getattr(self, '_plan', None) is not None and self._plan.value:
			self.lib.tbk_plan_destroy(self._plan)
			self._plan = C.c_void_p(0)

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass

	# ------------------------------------------------------------------------------------------
	def workspace(self, B, slot=0):
		"""Scratch buffer for a batch of B; ``slot`` selects independent buffers for concurrent streams."""
		ws = self._ws.get((B, slot))
		if ws is None:
			nbytes = self.lib.tbk_workspace_bytes(self._plan, B)
			ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
			off = (-ws.data_ptr()) % 256
			ws = ws[off:off + nbytes]
			self._ws[(B, slot)] = ws
		return ws

	def fit_stack(self, cube, meta, bkg_out, mask_out, chunk=64, extra_mask=None, status_out=None, nstreams=2):
		"""
		Fit a whole device-resident stack in chunks of ``chunk`` FFIs, alternating between ``nstreams`` CUDA
		streams (each with its own scratch) so that the small latency-bound kernels of one chunk overlap the
		throughput kernels of another.  The current stream waits for all of them at the end.
		"""
		n = cube.shape[0]
		meta_d = meta if isinstance(meta, torch.Tensor) else self.meta_to_device(meta)
		isz, ssz = META_DTYPE.itemsize, STATUS_DTYPE.itemsize
		cur = torch.cuda.current_stream(self.device)
		if not hasattr(self, '_streams') or len(self._streams) < nstreams:
			self._streams = [torch.cuda.Stream(self.device) for _ in range(nstreams)]
		for s in self._streams[:nstreams]:
			s.wait_stream(cur)
		for idx, a in enumerate(range(0, n, chunk)):
			b = min(a + chunk, n)
			k = idx % nstreams
			with torch.cuda.stream(self._streams[k]):
				self.fit(cube[a:b], meta_d[a * isz:b * isz], None if extra_mask is None else extra_mask[a:b],
					bkg_out=bkg_out[a:b], mask_out=mask_out[a:b],
					status_out=None if status_out is None else status_out[a * ssz:b * ssz], slot=k)
		for s in self._streams[:nstreams]:
			cur.wait_stream(s)

	def meta_to_device(self, meta):
		meta = np.ascontiguousarray(meta, dtype=META_DTYPE)
		return torch.from_numpy(meta.view(np.uint8).copy()).to(self.device, non_blocking=True)

	def fit(self, cube, meta=None, extra_mask=None, bkg_out=None, mask_out=None, status_out=None, profile=None, slot=0):
		"""
		Fit a device-resident batch.  ``cube`` float32 cuda tensor [B, H, W]; ``meta`` a
		``tbk_ffi_meta`` numpy array or an already uploaded uint8 tensor; ``extra_mask`` optional
		uint8/bool cuda tensor [B, H, W].  Returns ``(bkg float32, mask uint8, status uint8-tensor)``,
		all on the device; asynchronous on the current stream.
		"""
		if cube.dtype != torch.float32 or not cube.is_cuda or cube.dim() != 3 or tuple(cube.shape[1:]) != (self.H, self.W):
			raise ValueError(f"cube must be a float32 CUDA tensor of shape [B, {self.H}, {self.W}]")
		cube = cube.contiguous()
		B = cube.shape[0]
		if meta is None:
			meta = make_meta(B)
		meta_d = meta if isinstance(meta, torch.Tensor) else self.meta_to_device(meta)
		if meta_d.numel() != B * META_DTYPE.itemsize:
			raise ValueError("meta must hold one tbk_ffi_meta per FFI")
		if extra_mask is not None:
			extra_mask = extra_mask.view(torch.uint8) if extra_mask.dtype == torch.bool else extra_mask
			extra_mask = extra_mask.contiguous()
			if extra_mask.shape != cube.shape or extra_mask.dtype != torch.uint8:
				raise ValueError("extra_mask must be uint8/bool with the shape of cube")
		bkg = bkg_out if bkg_out is not None else torch.empty_like(cube)
		mask = mask_out if mask_out is not None else torch.empty(cube.shape, dtype=torch.uint8, device=self.device)
		status = status_out if status_out is not None else torch.empty(B * STATUS_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
		ws = self.workspace(B, slot)
		stream = torch.cuda.current_stream(self.device).cuda_stream
		if profile is not None:
			# measurement aid: synchronising variant that returns ms per kernel class in ``profile`` (dict)
			ms = (C.c_float * len(_lib.KERNEL_CLASSES))()
			check(self.lib.tbk_fit_batch_profiled(self._plan, _ptr(cube), B, _ptr(meta_d), _ptr(extra_mask), _ptr(bkg),
				_ptr(mask), _ptr(status), _ptr(ws), C.c_void_p(stream), ms), 'tbk_fit_batch_profiled')
			for name, v in zip(_lib.KERNEL_CLASSES, ms):
				profile[name] = profile.get(name, 0.0) + float(v)
		else:
			check(self.lib.tbk_fit_batch(self._plan, _ptr(cube), B, _ptr(meta_d), _ptr(extra_mask), _ptr(bkg),
				_ptr(mask), _ptr(status), _ptr(ws), C.c_void_p(stream)), 'tbk_fit_batch')
		self._last = (ws, B)
		return bkg, mask, status

	@staticmethod
	def status_to_numpy(status):
		return status.cpu().numpy().view(STATUS_DTYPE)

	def debug_fetch(self, b, round):
		"""(s2[nrings], mesh[ny, nx]) of FFI ``b`` / ``round`` of the most recent fit (synchronises)."""
		ws, B = self._last
		torch.cuda.synchronize(self.device)
		s2 = np.full(max(self.nrings, 1), np.nan)
		mesh = np.empty(self.ntiles)
		check(self.lib.tbk_debug_fetch(self._plan, _ptr(ws), B, b, round,
			s2.ctypes.data_as(C.c_void_p), mesh.ctypes.data_as(C.c_void_p)), 'tbk_debug_fetch')
		return s2[:self.nrings], mesh.reshape(self.H // 64, self.W // 64)

	def debug_workspace(self):
		"""Raw per-FFI control blocks and per-tile statistics of the most recent fit (synchronises)."""
		ws, B = self._last
		torch.cuda.synchronize(self.device)
		offs = (C.c_size_t * 9)(); sizes = (C.c_size_t * 3)()
		check(self.lib.tbk_workspace_layout(self._plan, B, offs, sizes), 'tbk_workspace_layout')
		raw = ws.cpu().numpy()
		assert sizes[0] == _lib.CTL_DTYPE.itemsize and sizes[1] == _lib.TILESTAT_DTYPE.itemsize
		nnf = int(sizes[2])
		ctl = raw[offs[0]:offs[0] + B * sizes[0]].view(_lib.CTL_DTYPE)
		base = raw[offs[1]:offs[1] + B * self.ntiles * sizes[1]].view(_lib.TILESTAT_DTYPE).reshape(B, self.ntiles)
		nf = raw[offs[2]:offs[2] + B * nnf * sizes[1]].view(_lib.TILESTAT_DTYPE).reshape(B, nnf)
		coef = raw[offs[3]:offs[3] + B * self.ntiles * 8].view('<f8').reshape(B, self.ntiles)
		nr = max(self.nrings, 1)
		s2_raw = raw[offs[5]:offs[5] + B * nr * 8].view('<f8').reshape(B, nr)[:, :self.nrings]
		fallbacks = raw[offs[8]:offs[8] + 64 * 4].view('<i4').copy()
		return dict(ctl=ctl, tile_base=base, tile_nf=nf, coef=coef, s2_raw=s2_raw, fallbacks=fallbacks)

	# ------------------------------------------------------------------------------------------
	def time_smooth(self, bkg, w, halo_lo=None, halo_hi=None, out=None):
		"""prepare.py:317-335 on a device shard; halos are the neighbouring shards' edge frames."""
		n = bkg.shape[0]
		out = out if out is not None else torch.empty_like(bkg)
		n_lo = 0 if halo_lo is None else halo_lo.shape[0]
		n_hi = 0 if halo_hi is None else halo_hi.shape[0]
		stream = torch.cuda.current_stream(self.device).cuda_stream
		check(self.lib.tbk_time_smooth(self._plan, _ptr(bkg.contiguous()), n, int(w),
			_ptr(halo_lo.contiguous() if n_lo else None), n_lo, _ptr(halo_hi.contiguous() if n_hi else None), n_hi,
			_ptr(out), C.c_void_p(stream)), 'tbk_time_smooth')
		return out

	def sum_accumulate(self, cube, bkg_smooth, flags, meta, sum_, nimg, used, flux_out=None):
		"""prepare.py:408-456 fused over the cadence axis; accumulates into sum_/nimg/used."""
		meta_d = meta if isinstance(meta, torch.Tensor) else self.meta_to_device(meta)
		stream = torch.cuda.current_stream(self.device).cuda_stream
		check(self.lib.tbk_sum_accumulate(self._plan, _ptr(cube), _ptr(bkg_smooth), _ptr(flags), _ptr(meta_d),
			cube.shape[0], _ptr(flux_out), _ptr(sum_), _ptr(nimg), _ptr(used), C.c_void_p(stream)), 'tbk_sum_accumulate')

	def sum_finalize(self, sum_, nimg, used, numfiles, threshold=0.5):
		"""prepare.py:459,468."""
		sumimage = torch.empty_like(sum_)
		pixels_used = torch.empty(sum_.shape, dtype=torch.uint8, device=self.device)
		stream = torch.cuda.current_stream(self.device).cuda_stream
		check(self.lib.tbk_sum_finalize(self._plan, _ptr(sum_), _ptr(nimg), _ptr(used), int(numfiles), float(threshold),
			_ptr(sumimage), _ptr(pixels_used), C.c_void_p(stream)), 'tbk_sum_finalize')
		return sumimage, pixels_used


@functools.lru_cache(maxsize=8)
def _cached_fitter(shape, is_tess, camera, ccd, flux_cutoff, bkgiters, radial_cutoff, radial_pixel_step, radial_smooth, device):
	return BackgroundFitter(shape, is_tess, camera, ccd, flux_cutoff, bkgiters, radial_cutoff, radial_pixel_step, radial_smooth, device=device)


def fit_background(image, catalog=None, flux_cutoff=8e4, bkgiters=3, radial_cutoff=2400,
	radial_pixel_step=15, radial_smooth=3):
	"""
	Estimate background in Full Frame Image -- same signature and return values as
	``photometry.backgrounds.fit_background`` (backgrounds.py:52-211).

	Parameters:
		image (ndarray, str or :class:`FFIImage`): 2D image or path to a FITS(.gz)/NPY file.
		catalog: ``None`` (default) reproduces the reference, where the argument is accepted but unused (backgrounds.py:64-65).
			Extension: an array [S, 3] of (column, row, Tmag) in science-pixel coordinates masks a disc around every star
			(:mod:`photometry_b200.starmask`; OR-ed into the mask at the point of backgrounds.py:90).
		flux_cutoff, bkgiters, radial_cutoff, radial_pixel_step, radial_smooth: see the reference.

	Returns:
		tuple: ``(bkg float64 ndarray, mask bool ndarray)``; ``mask`` is True where the pixel was
		not used.  ``bkg`` is all-NaN when every pixel is masked.

	Raises:
		ValueError: bad input type, or unknown CAMERA/CCD in a TESS header.
	"""
	img0 = image if isinstance(image, FFIImage) else FFIImage(image)
	data = np.ascontiguousarray(img0.data, dtype='float32')
	if data.ndim != 2:
		raise ValueError("Input image must be either 2D ndarray or path to file.")
	hdr = img0.header
	camera = hdr.get('CAMERA') if img0.is_tess else 0
	ccd = hdr.get('CCD') if img0.is_tess else 0
	if img0.is_tess and not (isinstance(camera, (int, np.integer)) and isinstance(ccd, (int, np.integer))):
		raise ValueError(f"Invalid CAMERA or CCD in header: CAMERA={camera}, CCD={ccd}")
	fitter = _cached_fitter(data.shape, img0.is_tess, int(camera or 0), int(ccd or 0), float(flux_cutoff), int(bkgiters),
		float(radial_cutoff), float(radial_pixel_step), int(radial_smooth or 0), torch.cuda.current_device())
	meta = meta_from_headers([hdr]) if img0.is_tess else make_meta(1)
	# one set of pinned staging buffers per cached fitter, reused across calls (the reference calls this once per FFI)
	stg = getattr(fitter, '_staging', None)
	if stg is None:
		stg = fitter._staging = (torch.empty(data.shape, dtype=torch.float32).pin_memory(), torch.empty(data.shape, dtype=torch.float32).pin_memory(),
			torch.empty(data.shape, dtype=torch.uint8).pin_memory())
	stg[0].copy_(torch.from_numpy(data))
	cube = stg[0].to(fitter.device, non_blocking=True).unsqueeze(0)
	extra = None
	if catalog is not None:
		from .starmask import star_mask
		extra = star_mask(data.shape, catalog, device=fitter.device.index).unsqueeze(0)
	bkg, mask, status = fitter.fit(cube, meta, extra_mask=extra)
	stg[1].copy_(bkg[0], non_blocking=True)
	stg[2].copy_(mask[0], non_blocking=True)
	st = fitter.status_to_numpy(status)[0]       # synchronises
	bkg_h = stg[1].numpy().astype('float64')
	mask_h = stg[2].numpy().astype(bool)
	if st['no_good_mesh']:
		# photutils: "All meshes contain > 2048 masked pixels" (uncaught in the reference)
		raise ValueError("All meshes contain > 2048 (50.0 percent per mesh) masked pixels. Please check your data or increase \"exclude_percentile\".")
	return bkg_h, mask_h
