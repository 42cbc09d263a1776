import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import photometry_b200 as pb
from photometry_b200 import synth
n = 8
cube = synth.synth_stack_torch(n, 2048, 2048, torch.device('cuda'), camera=1, ccd=2, seed=20260118)
hdrs = [dict(CAMERA=1, CCD=2, TSTART=1400.0 + k * 0.0208, TSTOP=1400.0208 + k * 0.0208, FFIINDEX=9000 + k) for k in range(n)]
fit = pb.BackgroundFitter((2048, 2048), True, 1, 2)
fit.fit(cube, pb.meta_from_headers(hdrs))
w = fit.debug_workspace()
print('kde fallbacks per FFI (of 4 ranks x 38 rings x 3 rounds = 456):', w['ctl']['kde_fallbacks'])
