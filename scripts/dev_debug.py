import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, oracle
import photometry_b200 as pb
from photometry_b200 import synth
np.set_printoptions(linewidth=200, precision=6)
dev = torch.device('cuda:0')
H = W = 512
xycen = (-30.0, 560.0)
stack = synth.synth_stack_numpy(1, H, W, seed=5, xycen=xycen, radial_cutoff=500.0, n_stars=800)
fit = pb.BackgroundFitter((H, W), True, 1, 2, radial_cutoff=500, radial_pixel_step=15, bkgiters=1, xycen=xycen)
hdrs = [dict(CAMERA=1, CCD=2, TSTART=1400.0, TSTOP=1400.02, FFIINDEX=9000)]
cube = torch.from_numpy(stack).to(dev)
bk, mk, st = fit.fit(cube, pb.meta_from_headers(hdrs))
d = {}
rb, rm = oracle.fit_background(oracle.FFIImageLite(stack[0], hdrs[0], True), radial_cutoff=500, xycen=xycen, bkgiters=1, diagnostics=d)
w = fit.debug_workspace()
c = w['ctl'][0]
print('ctl', {k: c[k] for k in c.dtype.names if k not in ('kx', 'pp', 'seg_of_ring')})
print('kx', c['kx'][:c['npts']])
print('pp', c['pp'][:3])
print('seg', c['seg_of_ring'][:40])
r, bins, cen = oracle.radial_geometry((H, W), xycen, 500, 15)
from scipy.interpolate import InterpolatedUnivariateSpline
s2 = d['rounds'][0]['s2']; ok = ~np.isnan(s2)
intp = InterpolatedUnivariateSpline(cen[ok], s2[ok], k=3, ext=3)
radial = 10**intp(r) - d['rounds'][0]['zeropoint']
dd = stack[0].astype('float64') - radial
dd[rm] = np.nan
rows = dd.reshape(8, 64, 8, 64).swapaxes(1, 2).reshape(64, 4096)
lo, hi = d['rounds'][0]['clip_lo'], d['rounds'][0]['clip_hi']
rows2 = rows.copy(); rows2[(rows < lo[:, None]) | (rows > hi[:, None])] = np.nan
omean = np.nanmean(rows2, 1); omed = np.nanmedian(rows2, 1); ostd = np.nanstd(rows2, 1); on = np.sum(~np.isnan(rows2), 1)
print('nonflat slots', w['tile_nf'].shape)
print('gpu nf[:8]', w['tile_nf'][0][:8])
print('ora    [:8]', list(zip(omean[:8], omed[:8], ostd[:8], on[:8])))
s2g, meshg = fit.debug_fetch(0, 0)
print('mesh gpu', meshg[:2]); print('mesh ora', d['rounds'][0]['mesh'][:2])
# emulate radial from ctl
kx = c['kx'][:c['npts']]; pp = c['pp']; seg = c['seg_of_ring']
t = np.minimum(r, c['xlast'])
i = np.clip(np.floor((t - 507.5) / 15).astype(int), 0, fit.nrings - 1)
s = seg[i].astype(int)
u = t - kx[s]
y = pp[s, 0] + u * (pp[s, 1] + u * (pp[s, 2] + u * pp[s, 3]))
rad_emul = np.where(r <= c['x0'], c['c_flat'], 10**y - c['zp'])
print('emul radial vs oracle maxdiff', np.abs(rad_emul - radial).max())
bg = bk[0].cpu().numpy().astype('float64')
err = bg - rb
print('bkg err: max', np.nanmax(np.abs(err)))
idx = np.argsort(-np.abs(err).ravel())[:10]
for q in idx:
	yy, xx = divmod(q, W)
	print(yy, xx, 'r=%.2f' % r[yy, xx], 'gpu', bg[yy, xx], 'ora', rb[yy, xx], 'radial', radial[yy, xx])
for rr in (510, 600, 700, 800):
	sel = np.abs(r - rr) < 1
	print(rr, 'median err', np.median(err[sel]), 'max', np.abs(err[sel]).max())
print('---- spline check')
ky = s2g[~np.isnan(s2g)]
m = len(kx)
from scipy.interpolate import CubicSpline
cs = CubicSpline(kx, ky, bc_type='not-a-knot')
ppy = cs.c.T[:, ::-1]  # [piece][c0..c3]
for i in range(m - 1):
	print(i, pp[i], ppy[i], np.abs(pp[i] - ppy[i]).max())
