"""Compare the statistics of the tile-kernel variants on a small case (GPU box): where do they differ?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import photometry_b200 as pb
from cases import CASES
name = sys.argv[1] if len(sys.argv) > 1 else 'tess_small'
case = CASES[name]()
imgs = case['images']; H, W = imgs.shape[1:]
cube = torch.from_numpy(imgs).cuda(); meta = pb.meta_from_headers(case['headers'])
res = {}
for v in ('0', '6'):
	os.environ['TBK_TILE_KERNEL'] = v
	fit = pb.BackgroundFitter((H, W), True, case['camera'], case['ccd'], xycen=case['xycen'], **case['fit_kwargs'])
	ex = torch.from_numpy(case['extra_mask']).cuda() if 'extra_mask' in case else None
	bkg, mask, st = fit.fit(cube, meta, extra_mask=ex)
	ws = fit.debug_workspace()
	res[v] = (ws['tile_base'].copy(), ws['tile_nf'].copy(), ws['fallbacks'].copy(), bkg.cpu().numpy())
print('fallbacks', res['6'][2][:5])
for idx, nm in ((0, 'tile_base'), (1, 'tile_nf')):
	a, b = res['6'][idx], res['0'][idx]
	bad = np.argwhere(a['nfin'] != b['nfin'])
	print(nm, 'shape', a.shape, 'nfin mismatches', len(bad))
	for k, t in bad[:6]:
		print('  ', k, t, 'z:', a[k, t], ' ref:', b[k, t])
	ok = (a['nfin'] == b['nfin']) & (b['nfin'] > 0)
	if ok.any():
		print('   max |med diff|', np.nanmax(np.abs(a['med'][ok] - b['med'][ok])), ' max |mean diff|', np.nanmax(np.abs(a['mean'][ok] - b['mean'][ok])),
			' max rel std diff', np.nanmax(np.abs(a['std'][ok] / b['std'][ok] - 1)))
print('bkg max abs diff', np.nanmax(np.abs(res['6'][3] - res['0'][3])))
