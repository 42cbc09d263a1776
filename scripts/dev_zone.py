"""Fallback counts of the zone kernels on full-size synthetic FFIs (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import photometry_b200 as pb
from photometry_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda:0')
for name, cam, ccd, kw in (('config 2', 1, 2, {}), ('crowded', 2, 3, dict(n_stars=400000, sky_level=1400.0, gradient=1.3))):
	cube = synth.synth_stack_torch(n, 2048, 2048, dev, camera=cam, ccd=ccd, seed=20260118, **kw)
	hdrs = [dict(CAMERA=cam, CCD=ccd, TSTART=1400.0 + k * 0.0208, TSTOP=1400.0208 + k * 0.0208, FFIINDEX=9000 + k) for k in range(n)]
	fit = pb.BackgroundFitter((2048, 2048), True, cam, ccd)
	fit.fit(cube, pb.meta_from_headers(hdrs))
	fb = fit.debug_workspace()['fallbacks']
	print(f"{name}: {n} FFIs, fallback counters {fb[:5].tolist()} of {n * 1024} meshes ({100.0 * fb[0] / (n * 1024):.2f} % raw-pixel)", flush=True)
	why = ('-', 'sample', 'range', 'list', 'sort', 'rank', 'bound', 'empty')
	print("   raw-pixel reasons: " + ' '.join(f"{why[i]}={fb[16 + i]}" for i in range(1, 8)) + "   residual reasons (3 rounds): " + ' '.join(f"{why[i]}={fb[24 + i]}" for i in range(1, 8)), flush=True)
