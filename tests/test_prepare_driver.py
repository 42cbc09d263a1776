"""
The prepare driver (photometry_b200/prepare_driver.py) against the reference's semantics (photometry/prepare.py:249-505):
product names / dtypes, the batched smoothing windows, and the resume rules.  The arithmetic is injected: here a small
stand-in engine built on the CPU oracle (test infrastructure) -- the product engine is ``GpuEngine`` and is exercised by
``test_driver_gpu_products_equal_prepare_stack`` on the GPU box.
"""
import os
import logging
import numpy as np
import pytest

import oracle
from oracle import prepare_oracle
from photometry_b200 import prepare_driver as pd
from photometry_b200.store import open_store, have_h5py


class OracleEngine:
	"""Stand-in engine: files are .npy sidecars of tiny images; background = per-frame median (enough to test the plumbing)."""
	def __init__(self):
		self.fit_calls = []
		self.acc = None

	def configure(self, shape, camera, ccd, **kw):
		self.kw = kw
		self.acc = None

	def load(self, files, with_err=False):
		cube = np.stack([np.load(f + '.npy') for f in files])
		hdrs = [dict(np.load(f + '.hdr.npy', allow_pickle=True).item()) for f in files]
		return (cube, hdrs, cube * 0 + 1) if with_err else (cube, hdrs)

	def fit(self, cube, headers):
		self.fit_calls.append([h['FFIINDEX'] for h in headers])
		bkg = np.stack([np.full(im.shape, np.nanmedian(im), 'float32') for im in cube])
		return bkg, (cube > 400).astype('uint8')

	def smooth(self, block, w, first, last):
		return prepare_oracle.time_smooth_backgrounds(block, 2 * w + 1)[first:last]

	def accumulate(self, cube, err, bkg, flags, headers):
		q = np.array([h.get('DQUALITY', 0) for h in headers], 'int32')
		res = prepare_oracle.sumimage_accumulate(cube, bkg, flags, q)
		if self.acc is None:
			self.acc = [np.zeros(cube.shape[1:]), np.zeros(cube.shape[1:], 'int32'), np.zeros(cube.shape[1:], 'int32')]
		self.acc[0] += np.nan_to_num(res['sumimage'] * res['nimg'], nan=0.0); self.acc[1] += res['nimg']; self.acc[2] += res['used']
		return res['flux'], err, flags

	def finalize(self, numfiles, threshold):
		with np.errstate(invalid='ignore', divide='ignore'):
			return self.acc[0] / self.acc[1], (self.acc[2] / numfiles) > threshold


def make_sector(tmp_path, n=11, shape=(16, 24)):
	rng = np.random.default_rng(5)
	files = []
	for k in range(n):
		name = os.path.join(tmp_path, f'tess20182061{k:05d}-s0001-1-2-0120-s_ffic.fits.gz')
		open(name, 'wb').close()
		img = (100 + 10 * rng.standard_normal(shape) + 3 * k).astype('float32')
		img[3, 4] = 900
		np.save(name + '.npy', img)
		np.save(name + '.hdr.npy', dict(CAMERA=1, CCD=2, TSTART=1400 + 0.02 * k, TSTOP=1400.02 + 0.02 * k, FFIINDEX=9000 + k,
			DQUALITY=32 if k == 4 else 0, BARYCORR=0.001, DATA_REL=1), allow_pickle=True)
		files.append(name)
	return files


@pytest.mark.parametrize('backend', ['npy'] + (['h5'] if have_h5py() else []))
def test_products_and_resume(tmp_path, backend, caplog):
	tmp_path = str(tmp_path)
	files = make_sector(tmp_path)
	n = len(files)
	assert pd.find_ffi_files(tmp_path, sector=1, camera=1, ccd=2) == files
	assert pd.find_ffi_files(tmp_path, sector=2) == []
	eng = OracleEngine()
	caplog.set_level(logging.INFO, logger='photometry_b200.prepare')
	out = pd.prepare_photometry(tmp_path, sectors=1, cameras=1, ccds=2, engine=eng, store_backend=backend, batch=4)
	assert out == [os.path.join(tmp_path, 'sector001_camera1_ccd2.hdf5')]
	assert 'sec/image' in caplog.text and 'Background estimation' in caplog.text      # the reference's log lines
	# whole-stack reference through the oracle loops
	cube = np.stack([np.load(f + '.npy') for f in files])
	bkg_us = np.stack([np.full(im.shape, np.nanmedian(im), 'float32') for im in cube])
	bkg = prepare_oracle.time_smooth_backgrounds(bkg_us, 3)
	flags = np.where(cube > 400, 1, 0).astype('uint8')
	q = np.array([32 if k == 4 else 0 for k in range(n)], 'int32')
	ref = prepare_oracle.sumimage_accumulate(cube, bkg, flags, q)
	with open_store(out[0], 'a', backend) as hdf:
		for grp in ('images', 'images_err', 'backgrounds', 'pixel_flags'):
			assert hdf.require_group(grp).keys() == [f'{k:04d}' for k in range(n)]          # tests/test_prepare.py:34-86
		for k in range(n):
			assert np.array_equal(np.asarray(hdf['backgrounds'][f'{k:04d}']), bkg[k])
			assert np.asarray(hdf['backgrounds'][f'{k:04d}']).dtype == np.float32
			assert np.asarray(hdf['pixel_flags'][f'{k:04d}']).dtype == np.uint8
			np.testing.assert_array_equal(np.asarray(hdf['images'][f'{k:04d}']), ref['flux'][k])
		np.testing.assert_allclose(np.asarray(hdf['sumimage']), ref['sumimage'], rtol=1e-12)
		assert np.array_equal(np.asarray(hdf['backgrounds_pixels_used']).astype(bool), ref['backgrounds_pixels_used'])
		assert np.array_equal(np.asarray(hdf['cadenceno']), 9000 + np.arange(n))
		assert np.asarray(hdf['time']).shape == (n,) and np.asarray(hdf['quality'])[4] == 32
		at = hdf.require_group('images').attrs
		assert at['SECTOR'] == 1 and at['CADENCE'] == 1800 and at['CAMERA'] == 1 and at['CCD'] == 2 and at['PIXEL_OFFSET_COLUMN'] == 44
		assert hdf.require_group('backgrounds').attrs['time_smooth'] == 3
	assert not os.path.exists(out[0].replace('.hdf5', '.tmp.hdf5')) and not os.path.exists(out[0].replace('.hdf5', '.tmp.hdf5') + '.d')
	assert eng.fit_calls == [[9000, 9001, 9002, 9003], [9004, 9005, 9006, 9007], [9008, 9009, 9010]]

	# ---- a finished file is left alone
	eng2 = OracleEngine()
	pd.prepare_photometry(tmp_path, sectors=1, cameras=1, ccds=2, engine=eng2, store_backend=backend, batch=4)
	assert eng2.fit_calls == [] and eng2.acc is None


def test_resume_after_crash_in_background_stage(tmp_path):
	"""prepare.py:265-273, 289-290: fits restart after the last existing pixel_flags member; prepare.py:321: smoothed frames are kept."""
	tmp_path = str(tmp_path)
	files = make_sector(tmp_path)

	class Crash(Exception):
		pass

	class CrashingEngine(OracleEngine):
		def fit(self, cube, headers):
			if len(self.fit_calls) == 1:
				raise Crash()
			return super().fit(cube, headers)
	with pytest.raises(Crash):
		pd.prepare_photometry(tmp_path, sectors=1, cameras=1, ccds=2, engine=CrashingEngine(), store_backend='npy', batch=4)
	eng = OracleEngine()
	out = pd.prepare_photometry(tmp_path, sectors=1, cameras=1, ccds=2, engine=eng, store_backend='npy', batch=4)
	assert eng.fit_calls == [[9004, 9005, 9006, 9007], [9008, 9009, 9010]]       # 0..3 were on disk already
	cube = np.stack([np.load(f + '.npy') for f in files])
	bkg = prepare_oracle.time_smooth_backgrounds(np.stack([np.full(im.shape, np.nanmedian(im), 'float32') for im in cube]), 3)
	with open_store(out[0], 'a', 'npy') as hdf:
		for k in range(len(files)):
			assert np.array_equal(np.asarray(hdf['backgrounds'][f'{k:04d}']), bkg[k])


def test_cadence_table_and_errors(tmp_path):
	assert [pd.sector_cadence(s) for s in (1, 26, 27, 54, 55, 60)] == [1800, 1800, 600, 600, 200, 200]
	assert pd.CADENCE_TIME_SMOOTH == {1800: 3, 600: 9, 200: 27}
	with pytest.raises(NotADirectoryError):
		pd.prepare_photometry(os.path.join(str(tmp_path), 'nope'), engine=OracleEngine())
	assert pd.prepare_photometry(str(tmp_path), engine=OracleEngine()) == []       # no sectors found
