"""Throughput and per-class times on the crowded synthetic configuration (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import photometry_b200 as pb
from photometry_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device('cuda:0')
cube = synth.synth_stack_torch(n, 2048, 2048, dev, camera=2, ccd=3, seed=20260118, n_stars=400000, sky_level=1400.0, gradient=1.3)
hdrs = [dict(CAMERA=2, CCD=3, TSTART=1400.0 + k * 0.0208, TSTOP=1400.0208 + k * 0.0208, FFIINDEX=9000 + k) for k in range(n)]
fit = pb.BackgroundFitter((2048, 2048), True, 2, 3)
meta = pb.meta_from_headers(hdrs); meta_d = fit.meta_to_device(meta)
bk = torch.empty_like(cube); mk = torch.empty(cube.shape, dtype=torch.uint8, device=dev)
prof = {}
for i in range(0, n, 64):
	fit.fit(cube[i:i + 64], meta[i:i + 64], bkg_out=bk[i:i + 64], mask_out=mk[i:i + 64], profile=prof)
print("crowded per-FFI us: " + ' '.join(f"{k}={v / n * 1e3:.1f}" for k, v in prof.items()), flush=True)
for ns in (1, 4):
	fit.fit_stack(cube, meta_d, bk, mk, chunk=64, nstreams=ns); torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(3): fit.fit_stack(cube, meta_d, bk, mk, chunk=64, nstreams=ns)
	e1.record(); torch.cuda.synchronize()
	print(f"crowded fit_stack nstreams={ns}: {3 * n / e0.elapsed_time(e1) * 1e3:.0f} FFIs/s", flush=True)
