"""
Image movement kernels on the device (photometry/image_motion.py:26-258, ``ImageMovementKernel`` with
``warpmode='translation'`` -- the mode ``prepare_photometry`` uses, photometry/prepare.py:681): for every frame of a
device-resident ``images`` stack, the (dx, dy) translation against a reference frame by ECC maximisation.

``ImageMovementKernel(image_ref).calc_kernel(image)`` keeps the reference's call shape for one image;
``calc_kernels(images)`` is the batched form the prepare stage uses (one launch chain for the whole stack).
"""
import ctypes as C
import numpy as np
import torch
from . import _lib


def _ptr(t):
	return C.c_void_p(t.data_ptr())


def prepare_flux(images):
	"""``ImageMovementKernel._prepare_flux`` (image_motion.py:74-111) for a float32 CUDA stack [B, H, W] (or one image [H, W])."""
	if not torch.cuda.is_available():
		raise _lib.TbkError("CUDA device required: photometry_b200 has no CPU fallback")
	lib = _lib.load()
	single = images.dim() == 2
	x = (images[None] if single else images).contiguous()
	if x.dtype != torch.float32 or not x.is_cuda:
		raise ValueError("images must be a float32 CUDA tensor")
	B, H, W = x.shape
	out = torch.empty_like(x)
	scratch = torch.empty(4 * B, dtype=torch.int32, device=x.device)
	_lib.check(lib.tbk_motion_prepare(_ptr(x), B, H, W, _ptr(out), _ptr(scratch), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), 'tbk_motion_prepare')
	return out[0] if single else out


class ImageMovementKernel:
	"""
	Translation-only movement kernel (``warpmode='translation'``, ``n_params = 2``).  ``image_ref`` is the reference frame
	(``images[ref_frame]``, prepare.py:681); it is prepared once, like the reference does in ``__init__``.
	"""
	n_params = 2
	warpmode = 'translation'

	def __init__(self, image_ref, warpmode='translation', device=None):
		if warpmode != 'translation':
			raise NotImplementedError("only warpmode='translation' (the mode prepare_photometry uses) runs on the device")
		if not torch.cuda.is_available():
			raise _lib.TbkError("CUDA device required: photometry_b200 has no CPU fallback")
		self.lib = _lib.load()
		ref = image_ref if isinstance(image_ref, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image_ref, dtype='float32'))
		self.device = ref.device if ref.is_cuda else torch.device('cuda', torch.cuda.current_device() if device is None else device)
		self.image_ref = prepare_flux(ref.to(self.device, torch.float32))
		self._ws = {}

	def calc_kernels(self, images, number_of_iterations=10000, termination_eps=1e-6, batch=32, return_info=False):
		"""
		``calc_kernel`` (image_motion.py:182-258) for every frame of ``images`` (float32 [N, H, W]; CUDA tensor, or host
		array / tensor that is uploaded batch by batch).  Returns float64 [N, 2] = (dx, dy) per frame, NaN where OpenCV would
		raise; with ``return_info`` also the final correlation coefficient and the iteration count.
		"""
		H, W = self.image_ref.shape
		n = images.shape[0]
		out = np.empty((n, 4), dtype='float64')
		stream = torch.cuda.current_stream(self.device)
		for a in range(0, n, batch):
			b = min(a + batch, n)
			x = images[a:b]
			if not isinstance(x, torch.Tensor):
				x = torch.from_numpy(np.ascontiguousarray(x, dtype='float32'))
			x = x.to(self.device, torch.float32)
			m = b - a
			prepared = prepare_flux(x)
			ws = self._ws.get(m)
			if ws is None:
				nbytes = self.lib.tbk_motion_workspace_bytes(m, H, W)
				raw = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
				off = (-raw.data_ptr()) % 256
				ws = self._ws[m] = raw[off:off + nbytes]
			res = torch.empty((m, 4), dtype=torch.float64, device=self.device)
			_lib.check(self.lib.tbk_motion_ecc(_ptr(self.image_ref), _ptr(prepared), m, H, W, int(number_of_iterations), float(termination_eps),
				_ptr(ws), _ptr(res), C.c_void_p(stream.cuda_stream)), 'tbk_motion_ecc')
			out[a:b] = res.cpu().numpy()
		return (out[:, :2], out[:, 2], out[:, 3].astype(int)) if return_info else out[:, :2]

	def calc_kernel(self, image, number_of_iterations=10000, termination_eps=1e-6):
		"""One image -> ``[dx, dy]`` (the reference's return value for warpmode='translation')."""
		img = image if isinstance(image, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(image, dtype='float32'))
		k = self.calc_kernels(img[None], number_of_iterations, termination_eps)
		return [float(k[0, 0]), float(k[0, 1])]

	def apply_kernel(self, xy, kernel):
		"""image_motion.py:113-176 for a translation kernel: every position moves by (dx, dy)."""
		xy = np.atleast_2d(np.asarray(xy, dtype='float64'))
		return np.broadcast_to(np.asarray(kernel, dtype='float64')[:2], xy.shape).copy()
