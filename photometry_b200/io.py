"""
Host-side image container mirroring ``photometry.io.FFIImage`` (photometry/io.py:25-93) for the
hot path: float32 science pixels, the merged header scalars, ``is_tess`` and ``mask``.

A path may be a ``.npy`` file or a FITS(.gz) file; FITS is decoded by a small pure-numpy reader
(primary + image extensions, no astropy).  WCS, smear rows and uncertainties beyond the science
crop are outside the hot path.
"""
import gzip
import numpy as np

_BLOCK = 2880


def _parse_card(card):
	key = card[:8].strip()
	if card[8:10] != '= ' or key in ('COMMENT', 'HISTORY', ''):
		return key, None
	body = card[10:]
	s = body.lstrip()
	if s.startswith("'"):
		end = 1
		while True:  # '' is an escaped quote
			end = s.find("'", end)
			if end < 0:
				return key, s[1:].rstrip()
			if s[end:end + 2] == "''":
				end += 2
				continue
			break
		return key, s[1:end].replace("''", "'").rstrip()
	val = body.split('/', 1)[0].strip()
	if val == 'T':
		return key, True
	if val == 'F':
		return key, False
	try:
		return key, int(val)
	except ValueError:
		pass
	try:
		return key, float(val.replace('D', 'E'))
	except ValueError:
		return key, val


def backfill_ffiindex(hdr):
	"""
	io.py:56-67: files from before sector 6 carry no FFIINDEX; the cadence number is extrapolated linearly from the
	mid-exposure time (30-min cadence files only).  Modifies and returns ``hdr``.
	"""
	if 'FFIINDEX' not in hdr and hdr['EXPOSURE'] * 86400 > 1000:
		time = 0.5 * (hdr['TSTART'] + hdr['TSTOP'])
		timecorr = hdr.get('BARYCORR', 0)
		first_time = 0.5 * (1325.317007851970 + 1325.337841177751) - 3.9072474e-03
		timedelt = 1800 / 86400
		offset = 4697 - first_time / timedelt
		hdr['FFIINDEX'] = np.round((time - timecorr) / timedelt + offset)
	return hdr


def scan_fits(buf):
	"""Yield (header dict, data offset, shape, bitpix) for every HDU of an in-memory FITS file."""
	pos = 0
	while pos + _BLOCK <= len(buf):
		hdr = {}
		done = False
		while not done:
			block = bytes(buf[pos:pos + _BLOCK]).decode('ascii', 'replace')
			pos += _BLOCK
			for i in range(0, _BLOCK, 80):
				card = block[i:i + 80]
				if card.startswith('END') and card[3:].strip() == '':
					done = True
					break
				key, val = _parse_card(card)
				if val is not None and key not in hdr:
					hdr[key] = val
			if pos >= len(buf) and not done:
				raise ValueError("truncated FITS header")
		naxis = int(hdr.get('NAXIS', 0))
		shape = [int(hdr[f'NAXIS{i}']) for i in range(naxis, 0, -1)]
		bitpix = int(hdr.get('BITPIX', 8))
		nbytes = abs(bitpix) // 8 * int(np.prod(shape)) if naxis else 0
		nbytes = (nbytes + int(hdr.get('PCOUNT', 0))) * int(hdr.get('GCOUNT', 1)) if naxis else 0
		yield hdr, pos, shape, bitpix
		pos += (nbytes + _BLOCK - 1) // _BLOCK * _BLOCK


def read_ffi_raw(path, with_err=False):
	"""
	Read a TESS FFI FITS(.gz) file WITHOUT decoding the pixels: returns
	``(merged header, raw big-endian bytes of the image HDU, naxis1, naxis2)`` for the device-side decode
	(:func:`photometry_b200.ingest.load_ffi_stack`); with ``with_err`` a fifth item, the raw bytes of the uncertainty
	HDU (``hdu[2]``, io.py:48).  Raises ValueError for non-TESS files (io.py:46).
	"""
	opener = gzip.open if str(path).endswith('.gz') else open
	with opener(path, 'rb') as fid:
		buf = fid.read()
	hdus = []
	for item in scan_fits(buf):
		hdus.append(item)
		if len(hdus) == (3 if with_err else 2):
			break
	if len(hdus) < 2:
		raise ValueError(f"{path}: no image extension")
	hdr0, (hdr1, off, shape, bitpix) = hdus[0][0], hdus[1]
	if hdr0.get('TELESCOP') != 'TESS' or hdr1.get('NAXIS1') != 2136 or hdr1.get('NAXIS2') != 2078 or bitpix != -32:
		raise ValueError(f"{path}: not a TESS full-frame image")
	hdr = backfill_ffiindex({**hdr0, **hdr1})
	n1, n2 = shape[1], shape[0]
	if with_err:
		if len(hdus) < 3 or hdus[2][2] != shape or hdus[2][3] != -32:
			raise ValueError(f"{path}: no uncertainty extension of the image's shape")
		off2 = hdus[2][1]
		return hdr, memoryview(buf)[off:off + 4 * n1 * n2], n1, n2, memoryview(buf)[off2:off2 + 4 * n1 * n2]
	return hdr, memoryview(buf)[off:off + 4 * n1 * n2], n1, n2


def read_fits_hdus(path):
	"""Return [(header dict, ndarray or None), ...] for the primary HDU and image extensions."""
	opener = gzip.open if str(path).endswith('.gz') else open
	with opener(path, 'rb') as fid:
		buf = fid.read()
	hdus = []
	for hdr, pos, shape, bitpix in scan_fits(buf):
		data = None
		if shape and hdr.get('XTENSION', 'IMAGE').strip() == 'IMAGE':
			dt = {8: 'u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}[bitpix]
			data = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape)), offset=pos).reshape(shape)
			if 'BSCALE' in hdr or 'BZERO' in hdr:
				data = data * hdr.get('BSCALE', 1) + hdr.get('BZERO', 0)
		hdus.append((hdr, data))
	return hdus


class FFIImage:
	"""
	``FFIImage(path_or_ndarray)`` -- same input rules as photometry/io.py:34-84:
	ndarray -> ``is_tess=False`` and an empty header; FITS path -> TESS detection by
	``TELESCOP == 'TESS'`` and raw size 2136 x 2078, science crop ``[0:2048, 44:2092]`` as float32,
	merged primary + extension-1 headers, ``FFIINDEX`` back-filled for pre-sector-6 files.
	``header`` may be passed explicitly together with an ndarray to build a TESS image in memory.
	"""
	def __init__(self, path, header=None, is_tess=None):
		self.is_tess = False
		self.uncertainty = None
		hdr = {}
		if isinstance(path, np.ndarray):
			data = path
			if header is not None:
				hdr = dict(header)
				self.is_tess = bool(is_tess) if is_tess is not None else False
		elif isinstance(path, str):
			if path.endswith('.npy'):
				data = np.load(path)
			else:
				hdus = read_fits_hdus(path)
				hdr0, _ = hdus[0]
				if hdr0.get('TELESCOP') == 'TESS' and len(hdus) > 2 and hdus[1][0].get('NAXIS1') == 2136 \
					and hdus[1][0].get('NAXIS2') == 2078:
					data = np.asarray(hdus[1][1][0:2048, 44:2092], dtype='float32')
					self.uncertainty = np.asarray(hdus[2][1][0:2048, 44:2092], dtype='float32')
					self.is_tess = True
					hdr = backfill_ffiindex({**hdr0, **hdus[1][0]})
				else:
					data = np.asarray(hdus[0][1], dtype='float32')
					if len(hdus) > 1 and hdus[1][1] is not None:
						self.uncertainty = np.asarray(hdus[1][1], dtype='float32')
					hdr = dict(hdr0)
		else:
			raise ValueError("Input image must be either 2D ndarray or path to file.")
		self.data = data
		self.header = hdr
		self.mask = ~np.isfinite(data)
		self.shape = data.shape
