"""
Input side of the hot path (SURVEY 8f "next" #2): TESS FFI FITS(.gz) files -> device-resident float32 cube.

The reference decodes every file twice on the host (gunzip + FITS parse + big-endian -> native + crop,
photometry/io.py:37-52, called from backgrounds.py:86 and again from prepare.py:379).  Here the host only
inflates the file and parses the header cards (thread pool; zlib releases the GIL); the raw big-endian image HDU
goes to the device through a pinned staging buffer and ``tbk_decode_ffi_be`` byte-swaps and crops
``[0:2048, 44:2092]`` straight into the cube.
"""
import ctypes as C
from collections import deque
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import torch
from . import _lib
from .io import read_ffi_raw


def decode_ffi_be(raw_dev, B, naxis1, naxis2, out, row0=0, col0=44):
	"""Device decode of B raw image HDUs (uint8 CUDA tensor) into ``out`` float32 [B, H, W]."""
	lib = _lib.load()
	H, W = out.shape[1:]
	stream = torch.cuda.current_stream(out.device).cuda_stream
	_lib.check(lib.tbk_decode_ffi_be(C.c_void_p(raw_dev.data_ptr()), int(B), int(naxis1), int(naxis2), int(row0), int(col0),
		int(H), int(W), C.c_void_p(out.data_ptr()), C.c_void_p(stream)), 'tbk_decode_ffi_be')
	return out


def _read_hdu(path, with_err=False):
	"""Worker: inflate + parse one file; only the image HDU's own bytes are kept (not the whole inflated file)."""
	if with_err:
		hdr, raw, n1, n2, raw_err = read_ffi_raw(path, with_err=True)
		return hdr, np.frombuffer(raw, dtype=np.uint8).copy(), n1, n2, np.frombuffer(raw_err, dtype=np.uint8).copy()
	hdr, raw, n1, n2 = read_ffi_raw(path)
	return hdr, np.frombuffer(raw, dtype=np.uint8).copy(), n1, n2, None


def load_ffi_stack(paths, device=None, threads=8, batch=8, with_err=False):
	"""
	Read the FFIs in ``paths`` (time ordered) into a CUDA tensor float32 [N, 2048, 2048].
	Returns ``(cube, headers)``; ``photometry_b200.meta_from_headers(headers)`` gives the per-FFI meta.  With ``with_err``
	the uncertainty HDU (io.py:48) is decoded the same way and ``(cube, headers, err_cube)`` is returned.
	At most ``2 * batch`` decoded files (17.7 MB per HDU) are held on the host at any time.
	"""
	if not torch.cuda.is_available():
		raise _lib.TbkError("CUDA device required: photometry_b200 has no CPU fallback")
	device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
	n = len(paths)
	if n == 0:
		raise ValueError("load_ffi_stack: no files given")
	H = W = 2048
	cube = torch.empty((n, H, W), dtype=torch.float32, device=device)
	err_cube = torch.empty((n, H, W), dtype=torch.float32, device=device) if with_err else None
	headers = [None] * n
	hdu_bytes = 2136 * 2078 * 4
	nh = 2 if with_err else 1
	stage = [torch.empty((nh * batch, hdu_bytes), dtype=torch.uint8).pin_memory() for _ in range(2)]
	stage_dev = [torch.empty((nh * batch, hdu_bytes), dtype=torch.uint8, device=device) for _ in range(2)]
	done = [torch.cuda.Event(), torch.cuda.Event()]
	with ThreadPoolExecutor(max_workers=threads) as pool:
		# a bounded window of reads in flight, consumed in file order
		pending = deque()
		submitted = 0

		def refill():
			nonlocal submitted
			while submitted < n and len(pending) < 2 * batch:
				pending.append(pool.submit(_read_hdu, paths[submitted], with_err))
				submitted += 1
		refill()
		for bi, a in enumerate(range(0, n, batch)):
			b = min(a + batch, n)
			k = bi & 1
			if bi >= 2:
				done[k].synchronize()   # the staging buffer must have been consumed
			for j in range(a, b):
				hdr, raw, n1, n2, raw_err = pending.popleft().result()
				refill()
				if (n1, n2) != (2136, 2078):
					raise ValueError(f"{paths[j]}: unexpected image size {n1} x {n2}")
				headers[j] = hdr
				stage[k].numpy()[j - a, :] = raw
				if with_err:
					stage[k].numpy()[batch + j - a, :] = raw_err
			m = b - a
			stage_dev[k][:m].copy_(stage[k][:m], non_blocking=True)
			decode_ffi_be(stage_dev[k], m, 2136, 2078, cube[a:b])
			if with_err:
				stage_dev[k][batch:batch + m].copy_(stage[k][batch:batch + m], non_blocking=True)
				decode_ffi_be(stage_dev[k][batch:], m, 2136, 2078, err_cube[a:b])
			done[k].record()
	return (cube, headers, err_cube) if with_err else (cube, headers)
