"""
On-disk products of the prepare stage with the layout ``photometry.prepare.prepare_photometry`` writes
(photometry/prepare.py:249-257, 296-300, 335, 428-447, 467-502) and ``BasePhotometry`` / ``todolist`` read back:

  groups   images/ images_err/ backgrounds/ pixel_flags/ wcs/   with members '%04d'
  datasets sumimage time timecorr time_start time_stop cadenceno quality backgrounds_pixels_used imagespaths
  attrs    images.attrs{SECTOR, CADENCE, CAMERA, CCD, ...}, backgrounds.attrs{time_smooth, flux_cutoff, ...},
           pixel_flags.attrs{bkgshe_done}, backgrounds_pixels_used.attrs{threshold}

Two interchangeable back ends behind one small interface:

  * ``H5Store``  -- a real HDF5 file through h5py with the reference's dataset options (image datasets in (64, 64)
    chunks with lzf + shuffle + fletcher32, prepare.py:136-141) -- used whenever h5py is importable, so the file is what
    the rest of the reference pipeline opens;
  * ``NpyStore`` -- a directory tree ``<file>.d/<group>/<NNNN>.npy`` + ``attrs.json`` with the same names, dtypes and
    resume behaviour, for machines without h5py (this build image has none).

Resume works as in the reference: what exists is kept, and the driver decides per '%04d' member what is left to do.
"""
import json
import os
import shutil
import numpy as np

IMG_CHUNKS = (64, 64)                                                        # prepare.py:141
H5_ARGS = dict(compression='lzf', shuffle=True, fletcher32=True)            # prepare.py:136-140


def have_h5py():
	try:
		import h5py  # noqa: F401
		return True
	except ImportError:
		return False


def open_store(path, mode='a', backend=None):
	"""Open (create) the product file ``path``; ``backend`` = 'h5' | 'npy' | None (h5py when importable)."""
	if backend is None:
		backend = 'h5' if have_h5py() else 'npy'
	if backend == 'h5':
		return H5Store(path, mode)
	if backend == 'npy':
		return NpyStore(path, mode)
	raise ValueError(f"unknown store backend {backend!r}")


class _Attrs:
	"""dict-like attributes persisted by the owning store."""
	def __init__(self, store, key):
		self._s, self._k = store, key

	def _d(self):
		return self._s._attrs.setdefault(self._k, {})

	def get(self, name, default=None):
		return self._d().get(name, default)

	def __getitem__(self, name):
		return self._d()[name]

	def __setitem__(self, name, value):
		if isinstance(value, (np.generic,)):
			value = value.item()
		self._d()[name] = value
		self._s._save_attrs()

	def __contains__(self, name):
		return name in self._d()

	def items(self):
		return self._d().items()


class _NpyGroup:
	def __init__(self, store, name):
		self._s, self.name = store, name
		self.path = os.path.join(store.root, name)
		os.makedirs(self.path, exist_ok=True)
		self.attrs = _Attrs(store, name)

	def keys(self):
		return sorted(f[:-4] for f in os.listdir(self.path) if f.endswith('.npy'))

	def __len__(self):
		return len(self.keys())

	def __contains__(self, key):
		return os.path.exists(os.path.join(self.path, key + '.npy'))

	def __getitem__(self, key):
		return np.load(os.path.join(self.path, key + '.npy'), mmap_mode='r')

	def create_dataset(self, key, data, **_):
		# write-then-rename: a crash never leaves a half-written member that resume would mistake for a finished one
		tmp = os.path.join(self.path, key + '.tmp')
		with open(tmp, 'wb') as fid:
			np.save(fid, np.asarray(data))
		os.replace(tmp, os.path.join(self.path, key + '.npy'))

	def update(self, key, data):
		self.create_dataset(key, data)


class NpyStore:
	"""Directory-backed store (see module docstring)."""
	backend = 'npy'

	def __init__(self, path, mode='a'):
		self.path = path
		self.root = path + '.d'
		if mode == 'w' and os.path.isdir(self.root):
			shutil.rmtree(self.root)
		os.makedirs(self.root, exist_ok=True)
		self._attrfile = os.path.join(self.root, 'attrs.json')
		self._attrs = {}
		if os.path.exists(self._attrfile):
			with open(self._attrfile) as fid:
				self._attrs = json.load(fid)
		self._groups = {}

	def _save_attrs(self):
		tmp = self._attrfile + '.tmp'
		with open(tmp, 'w') as fid:
			json.dump(self._attrs, fid, indent=1, sort_keys=True)
		os.replace(tmp, self._attrfile)

	def require_group(self, name):
		if name not in self._groups:
			self._groups[name] = _NpyGroup(self, name)
		return self._groups[name]

	def __contains__(self, name):
		return os.path.exists(os.path.join(self.root, name + '.npy')) or os.path.isdir(os.path.join(self.root, name))

	def __getitem__(self, name):
		if os.path.isdir(os.path.join(self.root, name)):
			return self.require_group(name)
		return np.load(os.path.join(self.root, name + '.npy'), mmap_mode='r')

	def set_dataset(self, name, data, attrs=None, **_):
		"""Create or replace a top-level dataset (the reference deletes and re-creates them, prepare.py:489-502)."""
		tmp = os.path.join(self.root, name + '.tmp')
		with open(tmp, 'wb') as fid:
			np.save(fid, np.asarray(data))
		os.replace(tmp, os.path.join(self.root, name + '.npy'))
		if attrs:
			self._attrs.setdefault(name, {}).update(attrs)
			self._save_attrs()

	def dataset_attrs(self, name):
		return dict(self._attrs.get(name, {}))

	def flush(self):
		self._save_attrs()

	def close(self):
		self.flush()

	def remove(self):
		shutil.rmtree(self.root, ignore_errors=True)

	def __enter__(self):
		return self

	def __exit__(self, *exc):
		self.close()


class _H5Group:
	def __init__(self, grp, image_like):
		self._g = grp
		self._image_like = image_like
		self.attrs = grp.attrs

	def keys(self):
		return sorted(self._g.keys())

	def __len__(self):
		return len(self._g)

	def __contains__(self, key):
		return key in self._g

	def __getitem__(self, key):
		return self._g[key]

	def create_dataset(self, key, data, **kw):
		data = np.asarray(data)
		if self._image_like and data.ndim == 2:
			self._g.create_dataset(key, data=data, chunks=IMG_CHUNKS, **H5_ARGS)      # prepare.py:300, 335, 428-429
		else:
			self._g.create_dataset(key, data=data, **kw)

	def update(self, key, data):
		self._g[key][...] = data


class H5Store:
	"""h5py-backed store writing the reference's HDF5 layout and filters."""
	backend = 'h5'

	def __init__(self, path, mode='a'):
		import h5py
		self.path = path
		self._h = h5py.File(path, mode, libver='latest')                           # prepare.py:249

	def require_group(self, name):
		return _H5Group(self._h.require_group(name), image_like=name not in ('wcs',))

	def __contains__(self, name):
		return name in self._h

	def __getitem__(self, name):
		import h5py
		obj = self._h[name]
		return _H5Group(obj, True) if isinstance(obj, h5py.Group) else obj

	def set_dataset(self, name, data, attrs=None, chunks=None, dtype=None):
		if name in self._h:
			del self._h[name]
		kw = dict(H5_ARGS)
		if chunks is not None:
			kw['chunks'] = chunks
		if dtype is not None:
			kw['dtype'] = dtype
		ds = self._h.create_dataset(name, data=np.asarray(data), **kw)
		for k, v in (attrs or {}).items():
			ds.attrs[k] = v

	def dataset_attrs(self, name):
		return dict(self._h[name].attrs)

	def flush(self):
		self._h.flush()

	def close(self):
		self._h.close()

	def remove(self):
		self.close()
		if os.path.exists(self.path):
			os.remove(self.path)

	def __enter__(self):
		return self

	def __exit__(self, *exc):
		self.close()
