"""Developer smoke check run on the GPU box: parity vs the oracle on a few cases + first timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import oracle
import photometry_b200 as pb
from photometry_b200 import synth


def cmp(name, bkg, ref, mask, rmask):
	tol = np.maximum(1e-5 * np.abs(ref), 1e-3)
	d = np.abs(bkg - ref)
	ok = (d <= tol) | (np.isnan(bkg) & np.isnan(ref))
	print(f"[{name}] mask_equal={np.array_equal(mask, rmask)} in_tol={ok.mean()*100:.4f}% max_abs={np.nanmax(d):.3e} "
		f"max_rel={np.nanmax(d/np.maximum(np.abs(ref),1e-30)):.3e}", flush=True)


def main():
	dev = torch.device('cuda:0')
	print(torch.cuda.get_device_name(0))
	# 1. constant KAT
	img = np.full((2048, 2048), 1000, dtype='float32')
	b, m = pb.fit_background(img)
	print('const: max|b-1000| =', np.abs(b - 1000).max(), 'mask any', m.any(), b.dtype, m.dtype)
	# 2. non-TESS noisy 512
	rng = np.random.default_rng(3)
	img = (200 + 20 * rng.standard_normal((512, 512))).astype('float32')
	img[100:140, 200:260] += 5000
	img[5, 5] = np.nan; img[7, 9] = -3; img[300, 300] = 9e4
	b, m = pb.fit_background(img)
	rb, rm = oracle.fit_background(img)
	cmp('nontess512', b, rb, m, rm)
	# 3. TESS small with radial
	H = W = 512
	xycen = (-30.0, 560.0)
	stack = synth.synth_stack_numpy(2, H, W, seed=5, xycen=xycen, radial_cutoff=500.0, n_stars=800)
	fit = pb.BackgroundFitter((H, W), True, 1, 2, radial_cutoff=500, radial_pixel_step=15, xycen=xycen)
	hdrs = [dict(CAMERA=1, CCD=2, TSTART=1400.0 + 0.02 * k, TSTOP=1400.02 + 0.02 * k, FFIINDEX=9000 + k) for k in range(2)]
	cube = torch.from_numpy(stack).to(dev)
	bk, mk, st = fit.fit(cube, pb.meta_from_headers(hdrs))
	torch.cuda.synchronize()
	stn = fit.status_to_numpy(st)
	for k in range(2):
		d = {}
		rb, rm = oracle.fit_background(oracle.FFIImageLite(stack[k], hdrs[k], True), radial_cutoff=500, xycen=xycen, diagnostics=d)
		cmp(f'tess512[{k}]', bk[k].cpu().numpy().astype('float64'), rb, mk[k].cpu().numpy().astype(bool), rm)
		for rnd in range(3):
			s2, mesh = fit.debug_fetch(k, rnd)
			rs2 = d['rounds'][rnd]['s2']
			print(f"   round {rnd}: zp gpu={stn[k]['zeropoint'][rnd]:.9f} ref={d['rounds'][rnd]['zeropoint']:.9f} "
				f"s2 maxdiff={np.nanmax(np.abs(s2 - rs2)):.3e} nan_eq={np.array_equal(np.isnan(s2), np.isnan(rs2))} "
				f"mesh maxrel={np.max(np.abs(mesh - d['rounds'][rnd]['mesh'])/np.abs(d['rounds'][rnd]['mesh'])):.3e}", flush=True)
	# 4. full-size TESS, 2 FFIs parity + timing on 16
	H = W = 2048
	n = 16
	cube = synth.synth_stack_torch(n, H, W, dev, camera=1, ccd=2, seed=11)
	fit = pb.BackgroundFitter((H, W), True, 1, 2)
	hdrs = [dict(CAMERA=1, CCD=2, TSTART=1400.0 + 0.02 * k, TSTOP=1400.02 + 0.02 * k, FFIINDEX=9000 + k) for k in range(n)]
	meta = pb.meta_from_headers(hdrs)
	for chunk in (16, 8):
		bk = torch.empty_like(cube); mk = torch.empty(cube.shape, dtype=torch.uint8, device=dev)
		for it in range(3):
			torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
			e0.record()
			for i in range(0, n, chunk):
				fit.fit(cube[i:i + chunk], meta[i:i + chunk], bkg_out=bk[i:i + chunk], mask_out=mk[i:i + chunk])
			e1.record(); torch.cuda.synchronize()
			ms = e0.elapsed_time(e1)
			print(f"full-size fit chunk={chunk}: {ms:.2f} ms for {n} FFIs -> {n / ms * 1e3:.1f} FFIs/s", flush=True)
	for k in (0, 7):
		t = time.time()
		rb, rm = oracle.fit_background(oracle.FFIImageLite(cube[k].cpu().numpy(), hdrs[k], True))
		print(f'oracle {time.time()-t:.1f}s')
		cmp(f'tess2048[{k}]', bk[k].cpu().numpy().astype('float64'), rb, mk[k].cpu().numpy().astype(bool), rm)


if __name__ == '__main__':
	main()
