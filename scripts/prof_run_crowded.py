"""Short workload for ncu on the crowded synthetic configuration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import photometry_b200 as pb
from photometry_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device('cuda:0')
cube = synth.synth_stack_torch(n, 2048, 2048, dev, camera=2, ccd=3, seed=20260118, n_stars=400000, sky_level=1400.0, gradient=1.3)
hdrs = [dict(CAMERA=2, CCD=3, TSTART=1400.0 + k * 0.0208, TSTOP=1400.0208 + k * 0.0208, FFIINDEX=9000 + k) for k in range(n)]
fit = pb.BackgroundFitter((2048, 2048), True, 2, 3)
meta = pb.meta_from_headers(hdrs)
for _ in range(2):
	fit.fit(cube, meta)
	torch.cuda.synchronize()
print('done')
