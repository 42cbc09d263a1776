"""
``prepare_photometry`` -- the prepare-stage driver around the CUDA hot path, shaped like
``photometry.prepare.prepare_photometry`` (photometry/prepare.py:79-706): for every (sector, camera, CCD) it finds the FFI
files, fits the backgrounds, smooths them in time, writes the background-subtracted images and accumulates the sum image,
producing the same product names / dtypes (``photometry_b200.store``) so the file is what ``BasePhotometry``, ``todolist`` and
``run_ffimovie`` open.

What replaces what:

  reference (prepare.py)                                     here
  ---------------------------------------------------------  ------------------------------------------------------------
  Pool.imap(fit_background, files[k:])        :278-302       engine.fit on batches of ``batch`` files (tbk_fit_batch)
  backgrounds_unsmoothed/NNNN in *.tmp.hdf5   :265-271, 296  same: a temporary store next to the product file
  smoothing loop, skip existing frames        :309-338       engine.smooth on windows of the unsmoothed store (tbk_time_smooth)
  final per-image loop                        :347-470       engine.accumulate on batches (tbk_sum_accumulate), then finalize
  resume by dataset presence                  :265, 273,     identical rules: ``len(backgrounds) < numfiles``,
                                               289-290, 321,  ``len(pixel_flags) < numfiles`` -> restart at last + 1,
                                               347            existing ``backgrounds/NNNN`` / ``images/NNNN`` are kept
  "... %f sec/image" log lines                :307, 338, 505 same text, logger ``photometry_b200.prepare``

The arithmetic lives in an *engine*: :class:`GpuEngine` (the product; raises without a CUDA device -- there is no CPU
fallback) or any object with the same five methods (the CPU tests drive the resume logic with a small stand-in).
Not done here (other stages of the reference, out of this path's scope): catalog download, WCS validation (astropy), the
time-offset fixes of early data releases, TPF quality transfer.
"""
import logging
import os
import re
from timeit import default_timer
import numpy as np

from .quality import TESSQualityFlags, PixelQualityFlags
from .store import open_store

CADENCE_TIME_SMOOTH = {1800: 3, 600: 9, 200: 27}    # prepare.py:258 ({1800: 3, 600: 9}); 200 s keeps the 5,400-s window (extension)


def sector_cadence(sector):
	"""FFI cadence in seconds (photometry/data/sectors.json: 1800 s up to sector 26, 600 s to 54, 200 s from 55)."""
	return 1800 if sector < 27 else (600 if sector < 55 else 200)


def find_ffi_files(rootdir, sector=None, camera=None, ccd=None):
	"""photometry/io.py:122-166: recursive search, sorted by file name (i.e. by time)."""
	sector_str = r'\d{4}' if sector is None else f'{sector:04d}'
	cam = r'\d' if camera is None else str(camera)
	cc = r'\d' if ccd is None else str(ccd)
	regexp = re.compile(r'^tess\d+-s(?P<sector>' + sector_str + ')-(?P<camera>' + cam + r')-(?P<ccd>' + cc + r')-\d{4}-[xsab]_ffic\.fits(\.gz)?$')
	matches = []
	for root, _, filenames in os.walk(rootdir, followlinks=True):
		matches += [os.path.join(root, f) for f in filenames if regexp.match(f)]
	matches.sort(key=os.path.basename)
	return matches


# --------------------------------------------------------------------------------------------------
class GpuEngine:
	"""The CUDA engine: FITS -> device cube (ingest), tbk_fit_batch, tbk_time_smooth, tbk_sum_accumulate / finalize."""
	def __init__(self, device=None):
		import torch
		from . import _lib
		if not torch.cuda.is_available():
			raise _lib.TbkError("CUDA device required: photometry_b200 has no CPU fallback")
		self.torch = torch
		self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
		self.fitter = None
		self.acc = None

	def configure(self, shape, camera, ccd, **fit_kwargs):
		from .backgrounds import BackgroundFitter
		self.fitter = BackgroundFitter(shape, True, camera, ccd, device=self.device.index, **fit_kwargs)
		self.shape = shape
		self.acc = None

	def load(self, files, with_err=False):
		from .ingest import load_ffi_stack
		return load_ffi_stack(files, device=self.device, with_err=with_err)

	def fit(self, cube, headers):
		"""-> (unsmoothed backgrounds float32 [B,H,W], NotUsedForBackground flags uint8 [B,H,W]) as host arrays."""
		from .backgrounds import meta_from_headers
		from .prepare import check_status
		from ._lib import STATUS_DTYPE
		bkg, mask, status = self.fitter.fit(cube, meta_from_headers(headers))
		check_status(status.cpu().numpy().view(STATUS_DTYPE))
		return bkg.cpu().numpy(), mask.cpu().numpy()

	def smooth(self, block, w, first, last):
		"""nan-mean over [k - w, k + w] (clipped to the block) for k in [first, last) of ``block`` float32 [m,H,W]."""
		t = self.torch.from_numpy(np.ascontiguousarray(block)).to(self.device)
		out = self.fitter.time_smooth(t, w)
		return out[first:last].cpu().numpy()

	def accumulate(self, cube, err, bkg, flags, headers):
		"""prepare.py:408-456 for one batch: -> (images, images_err, flags) host arrays; the accumulators stay on the device."""
		torch = self.torch
		from .backgrounds import meta_from_headers
		H, W = self.shape
		if self.acc is None:
			self.acc = (torch.zeros((H, W), dtype=torch.float64, device=self.device), torch.zeros((H, W), dtype=torch.int32, device=self.device),
				torch.zeros((H, W), dtype=torch.int32, device=self.device))
		fl = torch.from_numpy(np.ascontiguousarray(flags)).to(self.device)
		bk = torch.from_numpy(np.ascontiguousarray(bkg)).to(self.device)
		flux = torch.empty_like(cube)
		self.fitter.sum_accumulate(cube, bk, fl, meta_from_headers(headers), *self.acc, flux_out=flux)
		img_err = None
		if err is not None:
			img_err = err.clone()
			img_err[(fl & PixelQualityFlags.ManualExclude) != 0] = float('nan')      # prepare.py:423-425
			img_err = img_err.cpu().numpy()
		return flux.cpu().numpy(), img_err, fl.cpu().numpy()

	def finalize(self, numfiles, threshold):
		sumimage, used = self.fitter.sum_finalize(*self.acc, numfiles, threshold)
		return sumimage.cpu().numpy(), used.cpu().numpy().astype(bool)


# --------------------------------------------------------------------------------------------------
def prepare_photometry(input_folder=None, sectors=None, cameras=None, ccds=None, calc_movement_kernel=False,
	backgrounds_pixels_threshold=0.5, output_file=None, *, engine=None, store_backend=None, batch=32):
	"""
	Same parameters as the reference (prepare.py:79-114) plus ``engine`` (default :class:`GpuEngine`), ``store_backend``
	('h5' | 'npy' | None = h5py when importable) and ``batch`` (files per kernel launch).  Returns the list of product paths.
	"""
	logger = logging.getLogger('photometry_b200.prepare')
	if input_folder is None:
		input_folder = os.environ.get('TESSPHOT_INPUT', os.path.join(os.path.dirname(__file__), 'tests', 'input'))
	if not os.path.isdir(input_folder):
		raise NotADirectoryError("The given path does not exist or is not a directory")
	if calc_movement_kernel:
		logger.warning("calc_movement_kernel: the ECC movement kernels are computed by photometry_b200.image_motion, not by this driver")
	as_tuple = lambda v, default=None: default if v is None else (tuple(v) if hasattr(v, '__iter__') else (v,))
	cameras = as_tuple(cameras, (1, 2, 3, 4))
	ccds = as_tuple(ccds, (1, 2, 3, 4))
	sectors = as_tuple(sectors)
	if sectors is None:
		found = set()
		for fname in find_ffi_files(input_folder):
			found.add(int(re.match(r'^tess.+-s(\d+)-.+\.fits', os.path.basename(fname)).group(1)))
		sectors = tuple(sorted(found))
	if not sectors:
		logger.error("No sectors were found")
		return []
	if engine is None:
		engine = GpuEngine()
	products = []
	for sector in sectors:
		for camera in cameras:
			for ccd in ccds:
				files = find_ffi_files(input_folder, sector=sector, camera=camera, ccd=ccd)
				numfiles = len(files)
				if numfiles == 0:
					continue
				logger.info("Running SECTOR=%d, CAMERA=%d, CCD=%d", sector, camera, ccd)
				logger.info("Number of files: %d", numfiles)
				if output_file is None:
					hdf_file = os.path.join(input_folder, f'sector{sector:03d}_camera{camera:d}_ccd{ccd:d}.hdf5')
				else:
					hdf_file = os.path.abspath(output_file)
					if not hdf_file.endswith('.hdf5'):
						hdf_file += '.hdf5'
				_prepare_one(engine, files, sector, camera, ccd, hdf_file, backgrounds_pixels_threshold, store_backend, batch, logger)
				products.append(hdf_file)
	return products


def _prepare_one(engine, files, sector, camera, ccd, hdf_file, threshold, store_backend, batch, logger):
	numfiles = len(files)
	tic_total = default_timer()
	cadence = sector_cadence(sector)
	with open_store(hdf_file, 'a', store_backend) as hdf:
		images = hdf.require_group('images')
		images_err = hdf.require_group('images_err')
		backgrounds = hdf.require_group('backgrounds')
		pixel_flags = hdf.require_group('pixel_flags')
		hdf.require_group('wcs')
		# background parameters are persisted in, and re-read from, backgrounds.attrs (prepare.py:258-263, 311-316)
		time_smooth = int(backgrounds.attrs.get('time_smooth', CADENCE_TIME_SMOOTH[cadence]))
		fit_kwargs = dict(flux_cutoff=float(backgrounds.attrs.get('flux_cutoff', 8e4)), bkgiters=int(backgrounds.attrs.get('bkgiters', 3)),
			radial_cutoff=float(backgrounds.attrs.get('radial_cutoff', 2400)), radial_pixel_step=float(backgrounds.attrs.get('radial_pixel_step', 15)),
			radial_smooth=int(backgrounds.attrs.get('radial_smooth', 3)))
		engine.configure((2048, 2048), camera, ccd, **fit_kwargs)

		# ---- [A] backgrounds (prepare.py:265-345)
		if len(backgrounds) < numfiles:
			tmp = open_store(hdf_file.replace('.hdf5', '.tmp.hdf5'), 'a', store_backend)
			try:
				bck_us = tmp.require_group('backgrounds_unsmoothed')
				if len(pixel_flags) < numfiles:
					logger.info('Calculating backgrounds...')
					tic = default_timer()
					keys = pixel_flags.keys()
					last = -1 if not keys else int(keys[-1])                               # prepare.py:289
					for a in range(last + 1, numfiles, batch):
						b = min(a + batch, numfiles)
						cube, headers = engine.load(files[a:b])[:2]
						bck, mask = engine.fit(cube, headers)
						for k in range(a, b):
							name = f'{k:04d}'
							bck_us.create_dataset(name, bck[k - a])
							pixel_flags.create_dataset(name, np.where(mask[k - a] != 0, PixelQualityFlags.NotUsedForBackground, 0).astype('uint8'))
						hdf.flush(); tmp.flush()
					logger.info("Background estimation: %f sec/image", (default_timer() - tic) / max(numfiles - last, 1))
				logger.info('Smoothing backgrounds in time...')
				backgrounds.attrs['time_smooth'] = time_smooth
				for key, val in fit_kwargs.items():
					backgrounds.attrs[key] = val
				w = time_smooth // 2
				tic = default_timer()
				k = 0
				while k < numfiles:
					if f'{k:04d}' in backgrounds:                                            # prepare.py:321
						k += 1
						continue
					k1 = k
					while k1 < numfiles and k1 - k < batch and f'{k1:04d}' not in backgrounds:
						k1 += 1
					i1, i2 = max(k - w, 0), min(k1 + w, numfiles)
					block = np.stack([np.asarray(bck_us[f'{i:04d}'], dtype='float32') for i in range(i1, i2)])
					# frames outside [i1, i2) cannot contribute to cadences [k, k1): the window of cadence j is [j - w, j + w]
					sm = engine.smooth(block, w, k - i1, k1 - i1)
					for j in range(k, k1):
						backgrounds.create_dataset(f'{j:04d}', sm[j - k])
					k = k1
				logger.info("Background smoothing: %f sec/image", (default_timer() - tic) / numfiles)
				hdf.flush()
			finally:
				tmp.close()
			tmp.remove()                                                                    # prepare.py:344-345

		# ---- [B] final per-image loop (prepare.py:347-505)
		if len(images) < numfiles or 'sumimage' not in hdf or 'backgrounds_pixels_used' not in hdf or 'time_start' not in hdf:
			logger.info('Final processing of individual images...')
			tic = default_timer()
			time = np.empty(numfiles, 'float64'); timecorr = np.empty(numfiles, 'float32')
			time_start = np.empty(numfiles, 'float64'); time_stop = np.empty(numfiles, 'float64')
			cadenceno = np.empty(numfiles, 'int32'); quality = np.empty(numfiles, 'int32')
			attributes = dict.fromkeys(('CAMERA', 'CCD', 'DATA_REL', 'PROCVER', 'NUM_FRM', 'NREADOUT', 'CRMITEN', 'CRBLKSZ', 'CRSPOC'))
			hdf.set_dataset('imagespaths', np.array([os.path.basename(f).rstrip('.gz').encode('ascii', 'strict') for f in files]))
			for a in range(0, numfiles, batch):
				b = min(a + batch, numfiles)
				cube, headers, err = engine.load(files[a:b], with_err=True)
				for k, hdr in zip(range(a, b), headers):
					if k == 0:
						for key in attributes:
							attributes[key] = hdr.get(key)
					else:
						for key, value in attributes.items():
							if hdr.get(key) != value:
								logger.error("%04d: %s is not constant! (%s, %s)", k, key, value, hdr.get(key))
					time_start[k] = hdr['TSTART']; time_stop[k] = hdr['TSTOP']
					time[k] = 0.5 * (hdr['TSTART'] + hdr['TSTOP'])
					timecorr[k] = hdr.get('BARYCORR', 0); quality[k] = hdr.get('DQUALITY', 0)
					if 'FFIINDEX' not in hdr:
						raise RuntimeError("Could not determine CADENCENO for TESS data")
					cadenceno[k] = hdr['FFIINDEX']
				bkg = np.stack([np.asarray(backgrounds[f'{k:04d}'], dtype='float32') for k in range(a, b)])
				flg = np.stack([np.asarray(pixel_flags[f'{k:04d}'], dtype='uint8') for k in range(a, b)])
				flux, flux_err, flg_new = engine.accumulate(cube, err, bkg, flg, headers)
				for k in range(a, b):
					name = f'{k:04d}'
					if not np.array_equal(flg_new[k - a], flg[k - a]):
						pixel_flags.update(name, flg_new[k - a])                                 # ManualExclude bits, prepare.py:408-410
					if name not in images:
						images.create_dataset(name, flux[k - a])
						if flux_err is not None:
							images_err.create_dataset(name, flux_err[k - a])
			sumimage, pixels_used = engine.finalize(numfiles, threshold)
			if 'backgrounds_pixels_used' not in hdf:
				hdf.set_dataset('backgrounds_pixels_used', pixels_used, attrs={'threshold': threshold}, chunks=(64, 64), dtype='bool')
			images.attrs['SECTOR'] = sector
			images.attrs['CADENCE'] = cadence
			for key, value in attributes.items():
				if value is not None:
					images.attrs[key] = value
			images.attrs['PIXEL_OFFSET_ROW'] = 0
			images.attrs['PIXEL_OFFSET_COLUMN'] = 44
			for name, data in (('sumimage', sumimage), ('time', time), ('timecorr', timecorr), ('time_start', time_start),
				('time_stop', time_stop), ('cadenceno', cadenceno), ('quality', quality)):
				hdf.set_dataset(name, data)
			hdf.flush()
			logger.info("Individual image processing: %f sec/image", (default_timer() - tic) / numfiles)
	logger.info("Total: %f sec/image", (default_timer() - tic_total) / numfiles)
