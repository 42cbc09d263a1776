"""Device time of the smoothing / accumulation kernels of the prepare stage (GPU box, one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import photometry_b200 as pb
from photometry_b200 import synth
dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cube = synth.synth_stack_torch(n, 2048, 2048, dev, camera=1, ccd=2, seed=5)
hdrs = [dict(CAMERA=1, CCD=2, TSTART=1400.0 + k * 0.0208, TSTOP=1400.0208 + k * 0.0208, FFIINDEX=9000 + k) for k in range(n)]
meta = pb.meta_from_headers(hdrs)
fit = pb.BackgroundFitter((2048, 2048), True, 1, 2)
meta_d = fit.meta_to_device(meta)
bkg_us = torch.empty_like(cube); flags = torch.empty(cube.shape, dtype=torch.uint8, device=dev)
fit.fit_stack(cube, meta_d, bkg_us, flags, chunk=64, nstreams=4)
torch.cuda.synchronize()
def timed(f, reps=3):
	f(); torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(reps): f()
	e1.record(); torch.cuda.synchronize()
	return e0.elapsed_time(e1) / reps
for w in (1, 4, 13):
	ms = timed(lambda: fit.time_smooth(bkg_us, w, None, None))
	gb = n * 2048 * 2048 * 4 * 2 / 1e9
	print(f"time_smooth w={w}: {ms:.2f} ms = {ms / n * 1e3:.1f} us/FFI, {gb / ms * 1e3:.0f} GB/s of (1 read + 1 write) algorithmic bytes", flush=True)
bkg = fit.time_smooth(bkg_us, 1, None, None)
H, W = 2048, 2048
for keep in (False, True):
	images = torch.empty_like(cube) if keep else None
	def acc():
		s = torch.zeros((H, W), dtype=torch.float64, device=dev); ni = torch.zeros((H, W), dtype=torch.int32, device=dev); u = torch.zeros((H, W), dtype=torch.int32, device=dev)
		fit.sum_accumulate(cube, bkg, flags, meta_d, s, ni, u, flux_out=images)
	ms = timed(acc)
	gb = n * 2048 * 2048 * (4 + 4 + 1 + (4 if keep else 0)) / 1e9
	print(f"sum_accumulate keep_images={keep}: {ms:.2f} ms = {ms / n * 1e3:.1f} us/FFI, {gb / ms * 1e3:.0f} GB/s algorithmic", flush=True)
ms = timed(lambda: fit.fit_stack(cube, meta_d, bkg_us, flags, chunk=64, nstreams=4))
print(f"fit_stack: {ms:.2f} ms = {ms / n * 1e3:.1f} us/FFI", flush=True)
for ts in (3, 9):
	ms = timed(lambda: pb.prepare_stack(fit, cube, meta, time_smooth=ts, chunk=64, keep_images=False, nstreams=4), reps=2)
	print(f"prepare_stack time_smooth={ts} keep_images=False: {ms:.2f} ms = {ms / n * 1e3:.1f} us/FFI", flush=True)
