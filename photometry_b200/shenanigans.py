"""
Background-shenanigans detection on device-resident stacks: the stage that follows the sumimage in
``photometry.prepare.prepare_photometry`` (photometry/prepare.py:514-622) and its per-image function
``photometry.pixel_flags.pixel_background_shenanigans`` (photometry/pixel_flags.py:61-79).

  indicator[k] = float32(median_filter(images[k] - SumImage, size=15))          tbk_bkgshe_indicator
  mean         = robust mean over shuffled blocks of 25 indicator images          tbk_bkgshe_mean
  flags[k]    |= BackgroundShenanigans where abs(indicator[k] - mean) > 40         tbk_bkgshe_flag

Multi-GPU: the indicator and the flagging are per cadence (ranks keep their cadence shards; rank 0's sumimage is
broadcast).  The robust mean needs every cadence of a pixel, so the indicator stack is re-sharded from cadence blocks
to row slabs (one all-to-all exchange), reduced per slab, and the slabs of the mean image are broadcast back.
"""
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist
from . import _lib
from ._lib import check
from .quality import PixelQualityFlags

BKGSHE_BLOCK = 25        # prepare.py:560
BKGSHE_THRESHOLD = 40    # prepare.py:523


def _ptr(t):
	return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
	return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(t, dtype, what):
	if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != dtype:
		raise ValueError(f"{what} must be a {dtype} CUDA tensor")
	return t.contiguous()


def shuffled_order(numfiles):
	"""``indicies = list(range(numfiles)); np.random.seed(0); np.random.shuffle(indicies)`` (prepare.py:561-563)."""
	idx = list(range(int(numfiles)))
	np.random.RandomState(0).shuffle(idx)   # the frozen legacy generator behind np.random.seed / shuffle
	return np.asarray(idx, dtype='int32')


def shenanigans_indicator(images, sumimage=None, out=None):
	"""Indicator images of a device stack ``images`` float32 [n, H, W]; ``sumimage`` float64 [H, W] or None."""
	images = _need_cuda(images, torch.float32, 'images')
	if images.dim() != 3:
		raise ValueError("images must have shape [n, H, W]")
	n, H, W = images.shape
	if sumimage is not None:
		sumimage = _need_cuda(sumimage, torch.float64, 'sumimage')
		if tuple(sumimage.shape) != (H, W):
			raise ValueError("sumimage must have the shape of one image")
	out = out if out is not None else torch.empty_like(images)
	lib = _lib.load()
	for i in range(0, n, 65535):
		j = min(i + 65535, n)
		check(lib.tbk_bkgshe_indicator(_ptr(images[i:j]), _ptr(sumimage), j - i, H, W, _ptr(out[i:j]), _stream(images)),
			'tbk_bkgshe_indicator')
	return out


def mean_shenanigans(ind, order=None):
	"""Robust mean image (float64) of an indicator stack [n, ...pixels]; ``order`` defaults to the reference's shuffle."""
	ind = _need_cuda(ind, torch.float32, 'ind')
	n = ind.shape[0]
	npix = ind[0].numel()
	order = shuffled_order(n) if order is None else np.ascontiguousarray(order, dtype='int32')
	if order.shape != (n,):
		raise ValueError("order must hold one index per image")
	order_d = torch.from_numpy(order).to(ind.device)
	mean = torch.empty(tuple(ind.shape[1:]), dtype=torch.float64, device=ind.device)
	check(_lib.load().tbk_bkgshe_mean(_ptr(ind), npix, npix, n, _ptr(order_d), _ptr(mean), _stream(ind)), 'tbk_bkgshe_mean')
	return mean


def flag_shenanigans(ind, mean, pixel_flags, threshold=BKGSHE_THRESHOLD):
	"""Clear and set ``PixelQualityFlags.BackgroundShenanigans`` in ``pixel_flags`` (uint8 [n, H, W], in place)."""
	ind = _need_cuda(ind, torch.float32, 'ind')
	mean = _need_cuda(mean, torch.float64, 'mean')
	if not isinstance(pixel_flags, torch.Tensor) or not pixel_flags.is_cuda or pixel_flags.dtype != torch.uint8 \
		or not pixel_flags.is_contiguous() or pixel_flags.shape != ind.shape:
		raise ValueError("pixel_flags must be a contiguous uint8 CUDA tensor with the shape of ind")
	n = ind.shape[0]
	npix = ind[0].numel()
	lib = _lib.load()
	for i in range(0, n, 65535):
		j = min(i + 65535, n)
		check(lib.tbk_bkgshe_flag(_ptr(ind[i:j]), _ptr(mean), j - i, npix, float(threshold),
			int(PixelQualityFlags.BackgroundShenanigans), _ptr(pixel_flags[i:j]), _stream(ind)), 'tbk_bkgshe_flag')
	return pixel_flags


def slab_bounds(H, world_size, rank):
	"""Row slab [lo, hi) of ``rank`` for the robust mean."""
	return (rank * H) // world_size, ((rank + 1) * H) // world_size


def cadence_to_row_slabs(ind, group=None):
	"""
	Re-shard an indicator stack from cadence blocks (this rank: [n_local, H, W]) to row slabs: returns
	``[numfiles, rows_of_this_rank, W]`` in global cadence order.  Works on CUDA (NCCL) and CPU (gloo) tensors.
	"""
	rank, world = dist.get_rank(group), dist.get_world_size(group)
	n_local, H, W = ind.shape
	counts = [torch.zeros(1, dtype=torch.int64, device=ind.device) for _ in range(world)]
	dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=ind.device), group=group)
	counts = [int(c.item()) for c in counts]
	lo, hi = slab_bounds(H, world, rank)
	# ONE all_to_all_single: the send buffer holds, rank after rank, this shard's rows of that rank's slab; the chunk
	# received from rank r is cadence block r of the slab, which is contiguous in [numfiles, rows, W].
	slab = torch.empty((sum(counts), hi - lo, W), dtype=ind.dtype, device=ind.device)
	send = torch.empty(n_local * H * W, dtype=ind.dtype, device=ind.device)
	in_splits, out_splits, off = [], [], 0
	for r in range(world):
		rlo, rhi = slab_bounds(H, world, r)
		m = n_local * (rhi - rlo) * W
		if m:
			send[off:off + m].view(n_local, rhi - rlo, W).copy_(ind[:, rlo:rhi])
		in_splits.append(m); off += m
		out_splits.append(counts[r] * (hi - lo) * W)
	dist.all_to_all_single(slab.view(-1), send, out_splits, in_splits, group=group)
	return slab


def background_shenanigans(images, sumimage, pixel_flags, threshold=BKGSHE_THRESHOLD, group=None, return_indicator=False):
	"""
	The whole stage for this rank's cadence shard (prepare.py:514-622).

	images       float32 CUDA tensor [n_local, H, W]: background-subtracted frames (``images/NNNN``)
	sumimage     float64 [H, W] on rank 0 (``SectorResult.sumimage``); other ranks may pass None
	pixel_flags  uint8 [n_local, H, W], updated in place
	Returns ``mean_shenanigans`` (float64 [H, W], on every rank), plus the indicator stack if asked.
	"""
	distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
	n, H, W = images.shape
	if distributed:
		if sumimage is None:
			sumimage = torch.empty((H, W), dtype=torch.float64, device=images.device)
		src = dist.get_global_rank(group, 0) if group is not None else 0
		dist.broadcast(sumimage, src=src, group=group)
	ind = shenanigans_indicator(images, sumimage)
	if not distributed:
		mean = mean_shenanigans(ind)
	else:
		rank, world = dist.get_rank(group), dist.get_world_size(group)
		slab = cadence_to_row_slabs(ind, group)
		mean = torch.empty((H, W), dtype=torch.float64, device=images.device)
		lo, hi = slab_bounds(H, world, rank)
		if hi > lo:
			mean[lo:hi] = mean_shenanigans(slab)
		for r in range(world):
			rlo, rhi = slab_bounds(H, world, r)
			if rhi > rlo:
				part = mean[rlo:rhi].contiguous()
				dist.broadcast(part, src=(dist.get_global_rank(group, r) if group is not None else r), group=group)
				mean[rlo:rhi] = part
	flag_shenanigans(ind, mean, pixel_flags, threshold)
	return (mean, ind) if return_indicator else mean
