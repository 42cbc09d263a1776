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
# IDW fill through the kd-tree (Mars columns excluded), star mask, image movement kernels
mars = CASES['mars']()
fitm = pb.BackgroundFitter(mars['images'].shape[1:], True, mars['camera'], mars['ccd'])
bk, mk, st = fitm.fit(torch.from_numpy(mars['images']).cuda(), pb.meta_from_headers(mars['headers']))
from photometry_b200.starmask import star_mask
from photometry_b200.image_motion import ImageMovementKernel
sm = star_mask((H, W), np.array([[10.5, 20.2, 8.0], [200.0, 100.0, 12.0], [-3.0, 5.0, 6.0]]))
imk = ImageMovementKernel(res.images[0])
kern = imk.calc_kernels(res.images[:3], number_of_iterations=20)
torch.cuda.synchronize()
print('ok2', int(fitm.status_to_numpy(st)[0]['n_excluded'][0]), int(sm.sum()), kern.shape)
# host-resident stack: pinned copies, bit-packed mask transport, host-side expansion; long time-smoothing windows
hb = torch.empty((n, H, W), dtype=torch.float32).pin_memory(); hm = torch.empty((n, H, W), dtype=torch.uint8).pin_memory()
pb.fit_stack_host(fit, torch.from_numpy(imgs).pin_memory(), pb.meta_from_headers(case['headers'][:4]), hb, hm, chunk=2)
for w in (4, 13):
	fit.time_smooth(res.backgrounds, w)
torch.cuda.synchronize()
print('ok3', int(hm.sum()))
