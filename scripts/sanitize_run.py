"""Small workload for compute-sanitizer: every kernel of the fit + prepare path on a 256x320 TESS stack."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import photometry_b200 as pb
from cases import CASES
case = CASES['prepare']()
imgs = case['images'][:4]
n, H, W = imgs.shape
fit = pb.BackgroundFitter((H, W), True, case['camera'], case['ccd'], xycen=case['xycen'], **case['fit_kwargs'])
res = pb.prepare_stack(fit, torch.from_numpy(imgs).cuda(), pb.meta_from_headers(case['headers'][:4]), time_smooth=3, chunk=2)
torch.cuda.synchronize()
flags = res.pixel_flags.clone()
mean = pb.background_shenanigans(res.images, res.sumimage, flags)
torch.cuda.synchronize()
print('ok', float(res.sumimage.nanmean()), float(mean.abs().max()))
