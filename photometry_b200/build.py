"""
Build libtbk.so (the C-ABI CUDA library, include/tbk.h) in-tree for sm_100a.

Usage: ``python -m photometry_b200.build [--force]``.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIBPATH = os.path.join(LIBDIR, 'libtbk.so')
SOURCES = ['tbk_api.cu', 'tbk_fit.cu', 'tbk_prepare.cu', 'tbk_shenanigans.cu']
NVCC_FLAGS = [
	'-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
	'-Xcompiler', '-fPIC', '-shared', '-diag-suppress', '177',
]


def _stale():
	if not os.path.exists(LIBPATH):
		return True
	t = os.path.getmtime(LIBPATH)
	deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
	deps.append(os.path.join(HERE, '..', 'include', 'tbk.h'))
	return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
	"""Compile csrc/*.cu into lib/libtbk.so; returns the library path."""
	if not force and not _stale():
		return LIBPATH
	os.makedirs(LIBDIR, exist_ok=True)
	nvcc = os.environ.get('NVCC', 'nvcc')
	extra = os.environ.get('TBK_NVCC_FLAGS', '').split()
	cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIBPATH] + [os.path.join(CSRC, s) for s in SOURCES]
	res = subprocess.run(cmd, capture_output=True, text=True)
	if res.returncode != 0:
		raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
	if verbose:
		print(res.stderr)
	return LIBPATH


if __name__ == '__main__':
	print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
