"""
Consumer-side cube loads: the stamp cut-outs ``BasePhotometry._load_cube`` builds for every target
(photometry/BasePhotometry.py:720-751, properties ``images_cube`` / ``images_err_cube`` / ``backgrounds_cube`` /
``pixelflags_cube`` :754-877) served from the device-resident stacks of :class:`photometry_b200.SectorResult`
instead of one strided HDF5 read per cadence.
"""
import ctypes as C
import numpy as np
import torch
from . import _lib
from ._lib import check


class StampServer:
	"""
	Holds the stacks of one sector / camera / CCD on the device (``[N, H, W]``, time ordered) under the names of the
	HDF5 groups the reference reads (``images``, ``images_err``, ``backgrounds``, ``pixel_flags``).

	``stamp`` arguments are ``(row_min, row_max, col_min, col_max)`` in CCD pixel coordinates exactly as
	``BasePhotometry._stamp``; ``pixel_offset_row`` / ``pixel_offset_col`` are the ``PIXEL_OFFSET_ROW`` /
	``PIXEL_OFFSET_COLUMN`` attributes of the HDF5 file (0 and 44 for TESS FFIs, prepare.py:371-372).
	"""
	def __init__(self, pixel_offset_row=0, pixel_offset_col=44, **stacks):
		self.pixel_offset_row = int(pixel_offset_row)
		self.pixel_offset_col = int(pixel_offset_col)
		self.stacks = {}
		shape = None
		for name, t in stacks.items():
			if t is None:
				continue
			if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dim() != 3 or t.dtype not in (torch.float32, torch.uint8):
				raise ValueError(f"{name}: stacks must be float32 or uint8 CUDA tensors of shape [N, H, W]")
			if shape is not None and tuple(t.shape) != shape:
				raise ValueError("all stacks must have the same shape")
			shape = tuple(t.shape)
			self.stacks[name] = t.contiguous()
		if shape is None:
			raise ValueError("no stack given")
		self.N, self.H, self.W = shape
		self.lib = _lib.load()

	def _slices(self, stamps):
		st = np.asarray(stamps, dtype='int64').reshape(-1, 4).copy()
		st[:, 0:2] -= self.pixel_offset_row
		st[:, 2:4] -= self.pixel_offset_col
		if (st[:, 0] < 0).any() or (st[:, 1] > self.H).any() or (st[:, 2] < 0).any() or (st[:, 3] > self.W).any() \
			or (st[:, 1] <= st[:, 0]).any() or (st[:, 3] <= st[:, 2]).any():
			raise ValueError("stamp outside the frame or empty")
		return st.astype('int32')

	def load_cubes(self, stamps, hdf_group='images', views=True):
		"""
		Cubes ``(rows, cols, times)`` of several stamps in one launch; returns a list of device tensors (views of one
		buffer).  A missing group gives NaN cubes, like the reference (BasePhotometry.py:736-737).  ``views=False``
		returns ``(flat buffer, element offsets, shapes)`` instead (no per-stamp Python objects: thousands of stamps).
		"""
		st = self._slices(stamps)
		sizes = (st[:, 1] - st[:, 0]).astype('int64') * (st[:, 3] - st[:, 2]) * self.N
		offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype('int64')
		shapes = [(int(s[1] - s[0]), int(s[3] - s[2]), self.N) for s in st]
		stack = self.stacks.get(hdf_group)
		dev = next(iter(self.stacks.values())).device
		if stack is None:
			return [torch.full(sh, float('nan'), dtype=torch.float32, device=dev) for sh in shapes]
		out = torch.empty(int(sizes.sum()), dtype=stack.dtype, device=dev)
		d_st = torch.from_numpy(st).to(dev)
		d_off = torch.from_numpy(offs).to(dev)
		stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
		for i in range(0, len(st), 65535):
			j = min(i + 65535, len(st))
			check(self.lib.tbk_gather_stamps(C.c_void_p(stack.data_ptr()), stack.element_size(), self.N, self.H, self.W,
				C.c_void_p(d_st[i:j].data_ptr()), C.c_void_p(d_off[i:j].data_ptr()), j - i, C.c_void_p(out.data_ptr()), stream),
				'tbk_gather_stamps')
		if not views:
			return out, offs, shapes
		return [out[o:o + n].view(sh) for o, n, sh in zip(offs.tolist(), sizes.tolist(), shapes)]

	def load_cube(self, stamp, hdf_group='images'):
		"""``BasePhotometry._load_cube(hdf_group=...)`` for one stamp."""
		return self.load_cubes([stamp], hdf_group)[0]

	# the reference's property names
	def images_cube(self, stamp): return self.load_cube(stamp, 'images')
	def images_err_cube(self, stamp): return self.load_cube(stamp, 'images_err')
	def backgrounds_cube(self, stamp): return self.load_cube(stamp, 'backgrounds')
	def pixelflags_cube(self, stamp): return self.load_cube(stamp, 'pixel_flags')
