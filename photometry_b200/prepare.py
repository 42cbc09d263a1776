"""
The hot loops of ``photometry.prepare.prepare_photometry`` around ``fit_background`` on device-resident
FFI stacks (photometry/prepare.py:265-470): per-FFI background fit, background time smoothing, the
final per-image loop and the sumimage / Nimg / UsedInBackgrounds accumulation.

Multi-GPU: one process per GPU; each rank owns a contiguous block of the cadence axis
(:func:`shard_bounds`).  The only exchanges are (1) ``w = time_smooth // 2`` edge frames with each
neighbour for the smoothing window and (2) one sum-reduce of the three accumulators to rank 0.
"""
from dataclasses import dataclass
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
	"""Contiguous cadence block [lo, hi) of ``rank`` (GPU g gets [g*N/G, (g+1)*N/G))."""
	return (rank * n) // world_size, ((rank + 1) * n) // world_size


def exchange_halos(frames, w, group=None):
	"""
	Exchange the ``w`` edge frames of ``frames`` [n_local, H, W] with the previous / next rank.
	Returns ``(halo_lo, halo_hi)`` (None at the ends of the sector or when not distributed).
	Works on CUDA tensors (NCCL) and CPU tensors (gloo).
	"""
	if w <= 0 or not (dist.is_available() and dist.is_initialized()):
		return None, None
	rank, world = dist.get_rank(group), dist.get_world_size(group)
	if world == 1:
		return None, None
	if frames.shape[0] < w:
		raise ValueError(f"shard of {frames.shape[0]} cadences is shorter than the smoothing half-width {w}")
	ops = []
	halo_lo = halo_hi = None
	to_global = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
	if rank > 0:
		halo_lo = torch.empty((w,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
		ops.append(dist.P2POp(dist.isend, frames[:w].contiguous(), to_global(rank - 1), group))
		ops.append(dist.P2POp(dist.irecv, halo_lo, to_global(rank - 1), group))
	if rank < world - 1:
		halo_hi = torch.empty((w,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
		ops.append(dist.P2POp(dist.isend, frames[-w:].contiguous(), to_global(rank + 1), group))
		ops.append(dist.P2POp(dist.irecv, halo_hi, to_global(rank + 1), group))
	for req in dist.batch_isend_irecv(ops):
		req.wait()
	return halo_lo, halo_hi


def reduce_accumulators(sum_, nimg, used, n_local, group=None):
	"""
	Sum-reduce SumImage / Nimg / UsedInBackgrounds to rank 0 and return the global file count: ONE reduce of one packed
	float64 buffer [SumImage | Nimg | Used | count] -- the counters are integers far below 2**53, so their float64 sums are
	exact -- followed by one small broadcast of the count (every rank needs ``numfiles``).
	"""
	if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
		return n_local
	dst = dist.get_global_rank(group, 0) if group is not None else 0
	npix = sum_.numel()
	pack = torch.empty(3 * npix + 1, dtype=torch.float64, device=sum_.device)
	pack[:npix] = sum_.reshape(-1)
	pack[npix:2 * npix] = nimg.reshape(-1)
	pack[2 * npix:3 * npix] = used.reshape(-1)
	pack[3 * npix] = float(n_local)
	dist.reduce(pack, dst=dst, op=dist.ReduceOp.SUM, group=group)
	cnt = pack[3 * npix:].clone()
	dist.broadcast(cnt, src=dst, group=group)
	if dist.get_rank(group) == 0:
		sum_.copy_(pack[:npix].view_as(sum_))
		nimg.copy_(pack[npix:2 * npix].view_as(nimg))       # float64 -> int32, exact
		used.copy_(pack[2 * npix:3 * npix].view_as(used))
	return int(round(float(cnt.item())))


@dataclass
class SectorResult:
	"""Device-resident products of one (shard of a) sector/camera/CCD stack; names follow the HDF5 layout."""
	backgrounds_unsmoothed: torch.Tensor  # float32 [n, H, W]  (temp file group of the reference)
	backgrounds: torch.Tensor             # float32 [n, H, W]  backgrounds/NNNN
	pixel_flags: torch.Tensor             # uint8   [n, H, W]  pixel_flags/NNNN
	images: torch.Tensor                  # float32 [n, H, W]  images/NNNN (background subtracted) or None
	sumimage: torch.Tensor                # float64 [H, W]     (rank 0; None elsewhere)
	backgrounds_pixels_used: torch.Tensor # uint8   [H, W]     (rank 0; None elsewhere)
	nimg: torch.Tensor                    # int32   [H, W]     reduced on rank 0
	used: torch.Tensor                    # int32   [H, W]     reduced on rank 0
	status: np.ndarray                    # tbk_ffi_status per local FFI
	numfiles: int                         # global number of cadences


def check_status(status, first_cadence=0):
	"""
	The reference aborts when a frame has no usable mesh: photutils raises ``ValueError`` from ``Background2D`` (every
	mesh has more than ``exclude_percentile`` masked pixels) and ``prepare_photometry`` lets it propagate.  The batched
	kernels report that condition per FFI in ``tbk_ffi_status.no_good_mesh`` and write a NaN background; this raises the
	same exception type, naming the cadences, so that such a frame can never be smoothed over or accumulated silently.
	"""
	bad = np.flatnonzero(np.asarray(status['no_good_mesh']) != 0)
	if bad.size:
		raise ValueError("All meshes contain > 2048 (50.0 percent per mesh) masked pixels in cadence(s) %s of this stack: "
			"fit_background cannot estimate a background (photutils.Background2D raises here in the reference)"
			% ', '.join(str(int(first_cadence + k)) for k in bad[:20]) + (' ...' if bad.size > 20 else ''))


def prepare_stack(fitter, cube, meta, time_smooth=3, extra_mask=None, chunk=8, keep_images=True,
	backgrounds_pixels_threshold=0.5, group=None, timings=None, nstreams=1):
	"""
	Run prepare.py:265-470 for this rank's shard.

	fitter      :class:`photometry_b200.BackgroundFitter` for the stack's shape / camera / ccd
	cube        float32 CUDA tensor [n_local, H, W], time ordered
	meta        ``tbk_ffi_meta`` numpy array [n_local]
	time_smooth smoothing window in cadences (prepare.py:258: 3 at 1800 s, 9 at 600 s)
	chunk       FFIs per ``tbk_fit_batch`` launch;  nstreams > 1 runs the chunks on alternating CUDA streams
	timings     optional dict: receives the device time (ms) of the halo exchange and of the reduce ('halo_ms', 'reduce_ms')
	"""
	from ._lib import STATUS_DTYPE
	n = cube.shape[0]
	dev = cube.device
	meta = np.ascontiguousarray(meta)
	meta_d = fitter.meta_to_device(meta)
	isz = meta.dtype.itemsize
	bkg_us = torch.empty_like(cube)
	flags = torch.empty(cube.shape, dtype=torch.uint8, device=dev)
	status = torch.empty(n * STATUS_DTYPE.itemsize, dtype=torch.uint8, device=dev)
	ssz = STATUS_DTYPE.itemsize
	if nstreams > 1:
		fitter.fit_stack(cube, meta_d, bkg_us, flags, chunk=chunk, extra_mask=extra_mask, status_out=status, nstreams=nstreams)
	else:
		for i in range(0, n, chunk):
			j = min(i + chunk, n)
			fitter.fit(cube[i:j], meta_d[i * isz:j * isz], None if extra_mask is None else extra_mask[i:j],
				bkg_out=bkg_us[i:j], mask_out=flags[i:j], status_out=status[i * ssz:j * ssz])
	w = int(time_smooth) // 2
	ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timings is not None else None
	if ev: ev[0].record()
	halo_lo, halo_hi = exchange_halos(bkg_us, w, group)
	if ev: ev[1].record()
	bkg = fitter.time_smooth(bkg_us, w, halo_lo, halo_hi)
	H, W = cube.shape[1:]
	sum_ = torch.zeros((H, W), dtype=torch.float64, device=dev)
	nimg = torch.zeros((H, W), dtype=torch.int32, device=dev)
	used = torch.zeros((H, W), dtype=torch.int32, device=dev)
	images = torch.empty_like(cube) if keep_images else None
	fitter.sum_accumulate(cube, bkg, flags, meta_d, sum_, nimg, used, flux_out=images)
	if ev: ev[2].record()
	numfiles = reduce_accumulators(sum_, nimg, used, n, group)
	if ev:
		ev[3].record()
		torch.cuda.synchronize(dev)
		timings['halo_ms'] = ev[0].elapsed_time(ev[1])
		timings['reduce_ms'] = ev[2].elapsed_time(ev[3])
	is_root = not (dist.is_available() and dist.is_initialized()) or dist.get_rank(group) == 0
	sumimage = pixels_used = None
	if is_root:
		sumimage, pixels_used = fitter.sum_finalize(sum_, nimg, used, numfiles, backgrounds_pixels_threshold)
	status_np = status.cpu().numpy().view(STATUS_DTYPE)
	check_status(status_np)
	return SectorResult(bkg_us, bkg, flags, images, sumimage, pixels_used, nimg, used, status_np, numfiles)


def fit_stack_host(fitter, host_cube, meta, out_bkg, out_mask, chunk=16, nbuf=3, extra_mask=None, pack_mask=True, unpack_threads=6):
	"""
	End-to-end ``fit_background`` over a HOST-resident stack (the pool loop of prepare.py:291 with the
	FFIs already decoded): pinned host cube -> device -> fit -> pinned host results, pipelined over
	``nbuf`` device staging buffers and three streams (H2D, compute, D2H).

	host_cube  float32 pinned CPU tensor [n, H, W];  out_bkg float32 / out_mask uint8 pinned CPU tensors
	pack_mask  the device-to-host link bounds this path (21 MB out against 16.8 MB in per FFI) and the mask is a fifth of the
	           result bytes: it crosses the link as bits (``tbk_pack_mask``) and a few host threads expand it into ``out_mask``
	           (``tbk_unpack_mask_host``) while the next chunks are in flight.  The arrays the caller gets are the same.
	Returns the number of bytes copied (h2d, d2h).  Raises ``ValueError`` (like the reference) when a frame has no usable mesh.
	"""
	import ctypes as C
	from concurrent.futures import ThreadPoolExecutor
	from ._lib import STATUS_DTYPE, check
	n, H, W = host_cube.shape
	dev = fitter.device
	ssz = STATUS_DTYPE.itemsize
	status = torch.empty(n * ssz, dtype=torch.uint8, device=dev)
	meta_d = fitter.meta_to_device(np.ascontiguousarray(meta))
	isz = meta.dtype.itemsize
	pack_mask = bool(pack_mask) and (H * W) % 32 == 0
	s_in, s_c, s_out = (torch.cuda.Stream(dev) for _ in range(3))
	ins = [torch.empty((chunk, H, W), dtype=torch.float32, device=dev) for _ in range(nbuf)]
	bks = [torch.empty((chunk, H, W), dtype=torch.float32, device=dev) for _ in range(nbuf)]
	mks = [torch.empty((chunk, H, W), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
	nbits = H * W // 8
	if pack_mask:
		bits_d = [torch.empty((chunk, nbits), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
		bits_h = [torch.empty((chunk, nbits), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
		pool = ThreadPoolExecutor(max_workers=max(1, int(unpack_threads)))
		pending = [[] for _ in range(nbuf)]     # unpack jobs still reading bits_h[k]
		out_ptr = out_mask.data_ptr()
		lib = fitter.lib

		def unpack(k, j, frame, ev):
			ev.synchronize()                      # the chunk's bits have landed in bits_h[k]
			check(lib.tbk_unpack_mask_host(C.c_void_p(bits_h[k].data_ptr() + j * nbits), nbits,
				C.c_void_p(out_ptr + frame * H * W)), 'tbk_unpack_mask_host')
	ev_in = [torch.cuda.Event() for _ in range(nbuf)]
	ev_c = [torch.cuda.Event() for _ in range(nbuf)]
	ev_out = [torch.cuda.Event() for _ in range(nbuf)]
	cur = torch.cuda.current_stream(dev)
	for s in (s_in, s_c, s_out):
		s.wait_stream(cur)
	try:
		for idx, a in enumerate(range(0, n, chunk)):
			b = min(a + chunk, n)
			m = b - a
			k = idx % nbuf
			if idx >= nbuf:
				s_in.wait_event(ev_c[k])     # staging input free once its fit has run
				s_c.wait_event(ev_out[k])    # result buffers free once copied out
			with torch.cuda.stream(s_in):
				ins[k][:m].copy_(host_cube[a:b], non_blocking=True)
				ev_in[k].record(s_in)
			s_c.wait_event(ev_in[k])
			with torch.cuda.stream(s_c):
				fitter.fit(ins[k][:m], meta_d[a * isz:b * isz], None if extra_mask is None else extra_mask[a:b].to(dev, non_blocking=True),
					bkg_out=bks[k][:m], mask_out=mks[k][:m], status_out=status[a * ssz:b * ssz])
				if pack_mask:
					check(fitter.lib.tbk_pack_mask(C.c_void_p(mks[k].data_ptr()), m * H * W, C.c_void_p(bits_d[k].data_ptr()),
						C.c_void_p(s_c.cuda_stream)), 'tbk_pack_mask')
				ev_c[k].record(s_c)
			s_out.wait_event(ev_c[k])
			if pack_mask:
				for f in pending[k]:
					f.result()               # the previous chunk that used bits_h[k] has been expanded
			with torch.cuda.stream(s_out):
				out_bkg[a:b].copy_(bks[k][:m], non_blocking=True)
				if pack_mask:
					bits_h[k][:m].copy_(bits_d[k][:m], non_blocking=True)
				else:
					out_mask[a:b].copy_(mks[k][:m], non_blocking=True)
				ev_out[k].record(s_out)
			if pack_mask:
				ev = torch.cuda.Event()
				ev.record(s_out)
				pending[k] = [pool.submit(unpack, k, j, a + j, ev) for j in range(m)]
		for s in (s_in, s_c, s_out):
			cur.wait_stream(s)
		check_status(status.cpu().numpy().view(STATUS_DTYPE))   # synchronises: the device results are complete
		if pack_mask:
			for fl in pending:
				for f in fl:
					f.result()
	finally:
		if pack_mask:
			pool.shutdown(wait=True)
	return n * H * W * 4, n * H * W * 4 + (n * nbits if pack_mask else n * H * W)
