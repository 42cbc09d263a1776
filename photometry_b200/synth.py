"""
Synthetic FFI stacks: smooth sky + corner glow beyond the radial cutoff +
stars (pixel-integrated Gaussians, bleed columns above the flux cutoff) + noise + sparse NaNs.

``synth_stack_numpy`` is used by the parity tests (host arrays, oracle-sized); ``synth_stack_torch``
generates the same model directly on the device for the benchmark (there is no dataset to download).
"""
import math
import numpy as np

# photometry/backgrounds.py:121-138 (camera centre per camera/ccd), used to place the corner glow
_XYCEN = {
	(1, 1): (2158.222313, 2099.523364), (1, 2): (-5.653058, 2098.018608),
	(1, 3): (2141.511437, 2099.868226), (1, 4): (-22.406442, 2100.116443),
	(2, 1): (2148.588316, 2094.033024), (2, 2): (-16.806140, 2095.810070),
	(2, 3): (2151.351646, 2105.747100), (2, 4): (-13.118570, 2105.982211),
	(3, 1): (2152.175481, 2092.337442), (3, 2): (-10.494413, 2093.108135),
	(3, 3): (2145.029218, 2107.883573), (3, 4): (-17.374782, 2105.296746),
	(4, 1): (2149.259760, 2091.433315), (4, 2): (-12.906931, 2093.350054),
	(4, 3): (2148.906766, 2110.730620), (4, 4): (-14.629676, 2111.341670),
}


def camera_centre(camera, ccd):
	return _XYCEN[(camera, ccd)]


def _star_table(rng, H, W, n_stars):
	"""Positions and fluxes: N(<m) ~ 10**(0.3 m), flux = 10**(-0.4 (Tmag - 20.44)) (utilities.mag2flux)."""
	x = rng.uniform(0, W, n_stars)
	y = rng.uniform(0, H, n_stars)
	u = rng.uniform(0, 1, n_stars)
	tmag = 16.0 + np.log10(np.maximum(u, 1e-9)) / 0.3  # bright tail down to ~Tmag 4 at 20k stars
	tmag = np.clip(tmag, 3.0, 16.0)
	flux = 10 ** (-0.4 * (tmag - 20.44))
	return x, y, flux


def _erf_cdf(z):
	return 0.5 * (1.0 + np.vectorize(math.erf)(z / math.sqrt(2.0)))


def synth_stack_numpy(n, H, W, camera=1, ccd=2, seed=0, n_stars=None, xycen=None, radial_cutoff=2400.0,
	sky_level=150.0, gradient=0.5, glow_frac=0.3, nan_frac=1e-4, noise=True):
	"""Return float32 [n, H, W] (numpy).  Deterministic for a given seed."""
	rng = np.random.default_rng(seed)
	xc, yc = xycen if xycen is not None else camera_centre(camera, ccd)
	yy, xx = np.mgrid[0:H, 0:W].astype('float64')
	r = np.hypot(xx + 44 - xc, yy - yc)
	rmax = r.max()
	u, v = xx / W - 0.5, yy / H - 0.5
	shape = 1.0 + gradient * (0.6 * u + 0.4 * v) + 0.3 * gradient * (u * u - v * v + u * v)
	if n_stars is None:
		n_stars = int(20000 * H * W / 2048.0 ** 2)
	sx, sy, sf = _star_table(rng, H, W, n_stars)
	stars = np.zeros((H, W))
	half = 3
	for x0, y0, f in zip(sx, sy, sf):
		ix, iy = int(x0), int(y0)
		xs = np.arange(max(ix - half, 0), min(ix + half + 1, W))
		ys = np.arange(max(iy - half, 0), min(iy + half + 1, H))
		if xs.size == 0 or ys.size == 0:
			continue
		px = _erf_cdf((xs + 1 - x0) / 0.8) - _erf_cdf((xs - x0) / 0.8)
		py = _erf_cdf((ys + 1 - y0) / 0.8) - _erf_cdf((ys - y0) / 0.8)
		stars[np.ix_(ys, xs)] += f * np.outer(py, px)
		if f > 8e4:  # bleed column
			stars[max(iy - 40, 0):min(iy + 40, H), ix] += 1e5
	out = np.empty((n, H, W), dtype='float32')
	for t in range(n):
		amp = sky_level * (1.0 + 0.25 * math.sin(2 * math.pi * t / (13.7 * 48)))
		sky = amp * shape
		glow = np.where(r > radial_cutoff, glow_frac * sky * ((r - radial_cutoff) / max(rmax - radial_cutoff, 1.0)) ** 2, 0.0)
		img = sky + glow + stars
		if noise:
			img = img + rng.standard_normal((H, W)) * np.sqrt(img + 4.0)
		if nan_frac > 0:
			k = max(int(nan_frac * H * W), 1)
			img[rng.integers(0, H, k), rng.integers(0, W, k)] = np.nan
		out[t] = img.astype('float32')
	return out


def synth_stack_torch(n, H, W, device, camera=1, ccd=2, seed=0, n_stars=None, radial_cutoff=2400.0,
	sky_level=150.0, gradient=0.5, glow_frac=0.3, nan_frac=1e-4, out=None):
	"""Same model generated on the device; float32 [n, H, W] torch tensor."""
	import torch
	g = torch.Generator(device=device)
	g.manual_seed(int(seed))
	xc, yc = camera_centre(camera, ccd)
	yy = torch.arange(H, device=device, dtype=torch.float32).view(H, 1)
	xx = torch.arange(W, device=device, dtype=torch.float32).view(1, W)
	r = torch.hypot(xx + 44 - xc, yy - yc)
	rmax = float(r.max())
	u, v = xx / W - 0.5, yy / H - 0.5
	shape = 1.0 + gradient * (0.6 * u + 0.4 * v) + 0.3 * gradient * (u * u - v * v + u * v)
	if n_stars is None:
		n_stars = int(20000 * H * W / 2048.0 ** 2)
	rng = np.random.default_rng(seed)
	sx, sy, sf = _star_table(rng, H, W, n_stars)
	sx_t = torch.tensor(sx, device=device, dtype=torch.float32)
	sy_t = torch.tensor(sy, device=device, dtype=torch.float32)
	sf_t = torch.tensor(sf, device=device, dtype=torch.float32)
	stars = torch.zeros(H * W, device=device, dtype=torch.float32)
	ix, iy = sx_t.floor().long(), sy_t.floor().long()
	s2 = 0.8 * math.sqrt(2.0)
	for dy in range(-3, 4):
		py = 0.5 * (torch.erf((iy + dy + 1 - sy_t) / s2) - torch.erf((iy + dy - sy_t) / s2))
		for dx in range(-3, 4):
			px = 0.5 * (torch.erf((ix + dx + 1 - sx_t) / s2) - torch.erf((ix + dx - sx_t) / s2))
			yy_i, xx_i = iy + dy, ix + dx
			ok = (yy_i >= 0) & (yy_i < H) & (xx_i >= 0) & (xx_i < W)
			stars.index_put_(((yy_i * W + xx_i)[ok],), (sf_t * py * px)[ok], accumulate=True)
	stars = stars.view(H, W)
	for k in np.flatnonzero(sf > 8e4):
		stars[max(int(sy[k]) - 40, 0):min(int(sy[k]) + 40, H), int(sx[k])] += 1e5
	glow_shape = torch.where(r > radial_cutoff, glow_frac * ((r - radial_cutoff) / max(rmax - radial_cutoff, 1.0)) ** 2, torch.zeros_like(r))
	if out is None:
		out = torch.empty((n, H, W), device=device, dtype=torch.float32)
	nk = max(int(nan_frac * H * W), 1) if nan_frac > 0 else 0
	for t in range(n):
		amp = sky_level * (1.0 + 0.25 * math.sin(2 * math.pi * t / (13.7 * 48)))
		img = amp * shape * (1.0 + glow_shape) + stars
		img = img + torch.randn((H, W), device=device, generator=g) * torch.sqrt(img + 4.0)
		if nk:
			idx = torch.randint(0, H * W, (nk,), device=device, generator=g)
			img.view(-1)[idx] = float('nan')
		out[t] = img
	return out
