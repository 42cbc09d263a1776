"""
Parity report per SURVEY.md section 8d: for each synthetic configuration, fit a set of full-size FFIs on the
GPU and with the oracle, and report

  * mask equality, % of pixels inside the tolerance  |d| <= max(1e-5 |ref|, 1e-3), % of FFIs fully inside;
  * KDE-argmax flips: rings whose (smoothed, per round) and raw (last round) mode differs by more than 1e-9 dex --
    one grid cell is ~6e-4 dex, rounding noise is ~1e-13;
  * sigma-clip membership differences: sum over meshes of |n_kept(GPU) - n_kept(oracle)| in the last round;
  * IDW-affected meshes (excluded meshes filled by the Shepard interpolator), GPU and oracle.

Runs on the GPU box (python tests/tools/parity_report.py [--out FILE]); test infrastructure, not product code.
"""
import argparse, os, sys, time
import multiprocessing as mp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

BASE_SEED = 20260117


def _hdr(camera, ccd, k, cadenceno):
	return dict(CAMERA=camera, CCD=ccd, TSTART=1400.0 + 0.0208 * k, TSTOP=1400.0208 + 0.0208 * k, FFIINDEX=cadenceno, DQUALITY=0)


def _worker(job):
	import oracle
	img, hdr, extra, g = job
	d = {}
	t0 = time.time()
	rb, rm = oracle.fit_background(oracle.FFIImageLite(img, hdr, True), extra_mask=extra, diagnostics=d)
	sec = time.time() - t0
	gb = g['bkg'].astype('float64')
	tol = np.maximum(1e-5 * np.abs(rb), 1e-3)
	with np.errstate(invalid='ignore'):
		diff = np.abs(gb - rb)
		ok = (diff <= tol) | (np.isnan(gb) & np.isnan(rb))
	out = dict(mask_equal=bool(np.array_equal(g['mask'].astype(bool), rm)), n_pix=ok.size, n_ok=int(ok.sum()),
		max_abs=float(np.nanmax(diff)) if np.isfinite(diff).any() else 0.0,
		max_rel=float(np.nanmax(diff / np.maximum(np.abs(rb), 1e-30))) if np.isfinite(diff).any() else 0.0, oracle_sec=sec)
	flips = 0; rings = 0
	for rnd, rd in enumerate(d['rounds']):
		if 's2' not in rd:
			continue
		a, b = g['s2'][rnd], rd['s2']
		both = ~(np.isnan(a) & np.isnan(b))
		with np.errstate(invalid='ignore'):
			bad = both & ~(np.abs(a - b) <= 1e-9)
		flips += int(bad.sum()); rings += int(both.sum())
	last = d['rounds'][-1]
	raw_flips = 0
	if 's2_raw' in last:
		a, b = g['s2_raw'], last['s2_raw']
		both = ~(np.isnan(a) & np.isnan(b))
		with np.errstate(invalid='ignore'):
			raw_flips = int((both & ~(np.abs(a - b) <= 1e-9)).sum())
	kept_ref = (4096 - last['mesh_nbad']).ravel()
	good = last['mesh_good'].ravel()
	# meshes with no finite pixel at all report 0 kept on both sides
	out.update(kde_flips=flips, kde_rings=rings, kde_raw_flips_last=raw_flips,
		clip_diff=int(np.abs(kept_ref[good] - g['nfin'][good]).sum()), clip_meshes=int(good.sum()),
		idw_ref=int(last['n_excluded']), idw_gpu=int(g['n_excluded']),
		zp_rel=max(abs(g['zp'][r] - d['rounds'][r]['zeropoint']) / abs(d['rounds'][r]['zeropoint']) for r in range(len(d['rounds'])) if 'zeropoint' in d['rounds'][r]))
	return out


def run_config(name, pool, camera, ccd, n, cadence0, seed, synth_kw, extra_frac=0.0, log=print):
	import torch
	import photometry_b200 as pb
	from photometry_b200 import synth
	dev = torch.device('cuda:0')
	H = W = 2048
	cube = synth.synth_stack_torch(n, H, W, dev, camera=camera, ccd=ccd, seed=seed, **synth_kw)
	hdrs = [_hdr(camera, ccd, k, cadence0 + k) for k in range(n)]
	extra = None
	if extra_frac > 0:
		# star-mask extension: blobs covering ~extra_frac of the CCD so many meshes sit at the 50 % exclusion limit
		rng = np.random.default_rng(seed + 99)
		yy, xx = np.mgrid[0:H, 0:W].astype('float32')
		em = np.zeros((H, W), dtype=bool)
		while em.mean() < extra_frac:
			cy, cx, rad = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(10, 70)
			em |= (yy - cy) ** 2 + (xx - cx) ** 2 < rad * rad
		extra = np.broadcast_to(em, (n, H, W)).copy()
	fit = pb.BackgroundFitter((H, W), True, camera, ccd)
	ex_t = None if extra is None else torch.from_numpy(extra).to(dev)
	bkg, mask, st = fit.fit(cube, pb.meta_from_headers(hdrs), extra_mask=ex_t)
	torch.cuda.synchronize()
	stn = fit.status_to_numpy(st)
	dbg = fit.debug_workspace()
	rounds = int(stn[0]['rounds'])
	# kept pixels per mesh in the last round: meshes that see a non-constant radial component are re-clipped
	# every round (tile_nf), the others once on the raw pixels (tile_base)
	r = np.hypot(np.arange(W)[None, :] + 44 - synth.camera_centre(camera, ccd)[0], np.arange(H)[:, None] - synth.camera_centre(camera, ccd)[1])
	corner_max = np.maximum.reduce([r[0::64, 0::64], r[0::64, 63::64], r[63::64, 0::64], r[63::64, 63::64]]).ravel()
	nonflat = np.flatnonzero(corner_max > 2400.0 + 7.5)
	assert len(nonflat) == dbg['tile_nf'].shape[1], (len(nonflat), dbg['tile_nf'].shape)
	jobs = []
	for k in range(n):
		nfin = dbg['tile_base'][k]['nfin'].copy()
		if stn[k]['radial_ok'][rounds - 1]:
			nfin[nonflat] = dbg['tile_nf'][k]['nfin']
		g = dict(bkg=bkg[k].cpu().numpy(), mask=mask[k].cpu().numpy(), s2=[fit.debug_fetch(k, rnd)[0] for rnd in range(rounds)],
			s2_raw=dbg['s2_raw'][k].copy(), nfin=nfin, n_excluded=stn[k]['n_excluded'][rounds - 1], zp=[float(z) for z in stn[k]['zeropoint'][:rounds]])
		jobs.append((cube[k].cpu().numpy(), hdrs[k], None if extra is None else extra[k], g))
	res = pool.map(_worker, jobs)
	npix = sum(x['n_pix'] for x in res); nok = sum(x['n_ok'] for x in res)
	log(f"{name}: camera {camera} ccd {ccd}, {n} FFIs of 2048x2048, seed {seed}" + (f", extra mask {extra[0].mean() * 100:.1f} % of pixels" if extra is not None else ""))
	log(f"  mask exact            : {sum(x['mask_equal'] for x in res)}/{n} FFIs   (masked fraction {float(mask.float().mean()) * 100:.2f} %)")
	log(f"  pixels in tolerance   : {100.0 * nok / npix:.6f} %  ({npix - nok} outside);  FFIs fully inside: {sum(x['n_ok'] == x['n_pix'] for x in res)}/{n}")
	log(f"  max |d|, max rel      : {max(x['max_abs'] for x in res):.3e}, {max(x['max_rel'] for x in res):.3e}   (tolerance 1e-3 abs or 1e-5 rel)")
	log(f"  KDE argmax flips      : {sum(x['kde_flips'] for x in res)} of {sum(x['kde_rings'] for x in res)} ring values (all rounds, after smoothing); "
		f"{sum(x['kde_raw_flips_last'] for x in res)} raw in the last round")
	log(f"  sigma-clip membership : {sum(x['clip_diff'] for x in res)} pixels differ over {sum(x['clip_meshes'] for x in res)} kept meshes (last round)")
	log(f"  IDW-filled meshes     : GPU {sum(x['idw_gpu'] for x in res)}, oracle {sum(x['idw_ref'] for x in res)}  (of {n * 1024})")
	log(f"  zeropoint max rel diff: {max(x['zp_rel'] for x in res):.2e};  oracle {np.mean([x['oracle_sec'] for x in res]):.1f} s/FFI")
	return res


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--out', default=None)
	ap.add_argument('--n', type=int, default=16)
	ap.add_argument('--procs', type=int, default=min(os.cpu_count() or 1, 16))
	ap.add_argument('--seed-offset', type=int, default=0, help='added to the survey seeds (a second, independent sample)')
	args = ap.parse_args()
	lines = []

	def log(s):
		print(s, flush=True); lines.append(s)
	import torch
	log(f"Parity report, GPU path vs oracle (idw='ckdtree': scipy.spatial.cKDTree(leafsize=10), the reference's neighbour order), {torch.cuda.get_device_name(0)}, seed offset {args.seed_offset}")
	global BASE_SEED
	BASE_SEED += args.seed_offset
	with mp.get_context('spawn').Pool(args.procs) as pool:
		# config 1: the reference's CPU-runnable case substituted by synthetic camera 1 / CCD 4, cadences 4697-4700 (Mars exclude)
		run_config('config 1 (mars)', pool, 1, 4, 4, 4697, BASE_SEED + 0, {}, log=log)
		run_config('config 2 (sector, w=1)', pool, 1, 2, args.n, 9000, BASE_SEED + 1, {}, log=log)
		run_config('config 3 (camera 4 ccd 1)', pool, 4, 1, args.n, 20000, BASE_SEED + 2, {}, log=log)
		run_config('config 5 (crowded)', pool, 2, 3, max(args.n // 2, 1), 30000, BASE_SEED + 4,
			dict(n_stars=400000, sky_level=1400.0, gradient=1.3), extra_frac=0.40, log=log)
	if args.out:
		os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
		open(args.out, 'w').write("\n".join(lines) + "\n")


if __name__ == '__main__':
	main()
