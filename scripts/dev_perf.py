"""Per-kernel-class timing of the fit on 64 full-size FFIs for a few chunk sizes (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import photometry_b200 as pb
from photometry_b200 import synth
n = 512
dev = torch.device('cuda:0')
MARS = '--mars' in sys.argv   # camera 1 / ccd 4 before cadence 4724: 8 mesh columns excluded -> IDW fill in every frame
NOSTACK = '--no-stack' in sys.argv
sys.argv = [a for a in sys.argv if not a.startswith('--')]
cam, ccd, cad0 = (1, 4, 1000) if MARS else (1, 2, 9000)
cube = synth.synth_stack_torch(n, 2048, 2048, dev, camera=cam, ccd=ccd, seed=20260118)
hdrs = [dict(CAMERA=cam, CCD=ccd, TSTART=1400.0 + k * 0.0208, TSTOP=1400.0208 + k * 0.0208, FFIINDEX=cad0 + k) for k in range(n)]
fit = pb.BackgroundFitter((2048, 2048), True, cam, ccd)
meta = pb.meta_from_headers(hdrs)
bk = torch.empty_like(cube); mk = torch.empty(cube.shape, dtype=torch.uint8, device=dev)
for chunk in [int(a) for a in sys.argv[1:]] or [16, 64]:
	for rep in range(2):
		prof = {}
		torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
		for i in range(0, n, chunk):
			fit.fit(cube[i:i + chunk], meta[i:i + chunk], bkg_out=bk[i:i + chunk], mask_out=mk[i:i + chunk])
		e1.record(); torch.cuda.synchronize()
		ms = e0.elapsed_time(e1)
		for i in range(0, n, chunk):
			fit.fit(cube[i:i + chunk], meta[i:i + chunk], bkg_out=bk[i:i + chunk], mask_out=mk[i:i + chunk], profile=prof)
	print(f"chunk={chunk}: {n / ms * 1e3:.0f} FFIs/s ({ms / n * 1e3:.1f} us/FFI)  per-FFI us: " + ' '.join(f"{k}={v / n * 1e3:.1f}" for k, v in prof.items()), flush=True)


for ns, chunk in (() if NOSTACK else ((2, 64), (3, 64), (4, 64), (4, 96), (4, 128), (6, 64), (6, 32), (8, 32), (8, 64))):
	for rep in range(2):
		torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
		fit.fit_stack(cube, meta, bk, mk, chunk=chunk, nstreams=ns)
		e1.record(); torch.cuda.synchronize()
	ms = e0.elapsed_time(e1)
	print(f"fit_stack nstreams={ns} chunk={chunk}: {n / ms * 1e3:.0f} FFIs/s", flush=True)
