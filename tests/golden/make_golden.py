"""
Regenerate tests/golden/*.npz: inputs of the small parity cases (tests/cases.py) and the oracle's
outputs for them.  Run from the repo root:  python tests/golden/make_golden.py

The reference itself cannot be imported in the build container (astropy/photutils/statsmodels/
bottleneck/h5py absent), so these vectors pin the ORACLE (numpy 2.3 / scipy 1.18),
not the reference; the four known answers the reference's own tests hold are checked separately in
tests/test_oracle_kat.py.
"""
import os
import sys
import hashlib
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle  # noqa: E402
from cases import CASES  # noqa: E402


def run_case(name, case):
	imgs = case['images']
	n = imgs.shape[0]
	# inputs are regenerated from seeds (tests/cases.py); only their digest is stored
	out = dict(images_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(imgs).tobytes()).digest(), dtype='uint8'))
	bkgs, masks, s2s, meshes, zps, nexcl = [], [], [], [], [], []
	for k in range(n):
		img = imgs[k]
		diag = {}
		extra = case['extra_mask'][k] if 'extra_mask' in case else None
		if case['kind'] == 'tess':
			ffi = oracle.FFIImageLite(img, case['headers'][k], True)
			b, m = oracle.fit_background(ffi, xycen=case['xycen'], extra_mask=extra, diagnostics=diag, **case['fit_kwargs'])
		else:
			b, m = oracle.fit_background(img, extra_mask=extra, diagnostics=diag, **case['fit_kwargs'])
		bkgs.append(b); masks.append(m)
		rounds = diag.get('rounds', [])
		meshes.append(np.stack([r['mesh'] for r in rounds]) if rounds else np.zeros((0,)))
		nexcl.append([r['n_excluded'] for r in rounds])
		if rounds and 's2' in rounds[0]:
			s2s.append(np.stack([r['s2'] for r in rounds]))
			zps.append([r['zeropoint'] for r in rounds])
	out['bkg'] = np.stack(bkgs).astype('float32')   # the CUDA path returns float32 as well
	out['mask_bits'] = np.packbits(np.stack(masks))
	out['shape'] = np.array(imgs.shape)
	out['mesh'] = np.stack(meshes)
	out['n_excluded'] = np.array(nexcl)
	if s2s:
		out['s2'] = np.stack(s2s)
		out['zeropoint'] = np.array(zps)
	if 'time_smooth' in case:
		ffis = [oracle.FFIImageLite(imgs[k], case['headers'][k], True) for k in range(n)]
		res = oracle.prepare_stack(ffis, case['time_smooth'], fit_kwargs=dict(xycen=case['xycen'], **case['fit_kwargs']))
		for key in ('backgrounds', 'flux', 'sumimage', 'nimg', 'used'):
			out['prep_' + key] = res[key]
		out['prep_pixel_flags_manexcl_bits'] = np.packbits((res['pixel_flags'] & 2) != 0)
		out['prep_pixels_used_bits'] = np.packbits(res['backgrounds_pixels_used'])
	return out


def run_shenanigans(case):
	flags, mean, ind = oracle.background_shenanigans(case['images'], case['sumimage'], case['pixel_flags'])
	return dict(images_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(case['images']).tobytes()).digest(), dtype='uint8'),
		shape=np.array(case['images'].shape), indicator=ind[[0, 4, 17, 29]], mean=mean,
		flag_bits=np.packbits((flags & 4) != 0), flags_other=np.packbits((flags & 3) != 0), order=oracle.shuffled_order(case['images'].shape[0]))


if __name__ == '__main__':
	here = os.path.dirname(os.path.abspath(__file__))
	only = sys.argv[1:]
	if not only or 'shenanigans' in only:
		from cases import case_shenanigans
		res = run_shenanigans(case_shenanigans())
		path = os.path.join(here, 'shenanigans.npz')
		np.savez_compressed(path, **res)
		print('shenanigans', {k: v.shape for k, v in res.items()}, '%.1f KiB' % (os.path.getsize(path) / 1024))
	for name, fn in CASES.items():
		if only and name not in only:
			continue
		res = run_case(name, fn())
		path = os.path.join(here, name + '.npz')
		np.savez_compressed(path, **res)
		print(name, {k: v.shape for k, v in res.items()}, '%.1f KiB' % (os.path.getsize(path) / 1024))
