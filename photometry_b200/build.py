"""
Build libtbk.so (the C-ABI CUDA library, include/tbk.h) in-tree for sm_100a.

Usage: ``python -m photometry_b200.build [--force] [-v]``.  nvcc cross-compiles without a GPU.
Each translation unit is compiled to an object file (in parallel, only when stale) and the objects are linked.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(LIBDIR, 'obj')
LIBPATH = os.path.join(LIBDIR, 'libtbk.so')
SOURCES = ['tbk_api.cu', 'tbk_fit.cu', 'tbk_prepare.cu', 'tbk_shenanigans.cu', 'tbk_motion.cu']
NVCC_FLAGS = [
	'-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
	'-Xcompiler', '-fPIC', '-diag-suppress', '177',
]


def _headers():
	deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
	deps.append(os.path.join(HERE, '..', 'include', 'tbk.h'))
	return deps


def _newer(target, deps):
	if not os.path.exists(target):
		return True
	t = os.path.getmtime(target)
	return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
	"""Compile csrc/*.cu into lib/libtbk.so; returns the library path."""
	os.makedirs(OBJDIR, exist_ok=True)
	nvcc = os.environ.get('NVCC', 'nvcc')
	extra = os.environ.get('TBK_NVCC_FLAGS', '').split()
	flagfile = os.path.join(OBJDIR, 'flags.txt')
	flags = ' '.join(NVCC_FLAGS + extra)
	if not os.path.exists(flagfile) or open(flagfile).read() != flags:
		force = True
	hdrs = _headers()
	jobs = []
	for s in SOURCES:
		src = os.path.join(CSRC, s)
		obj = os.path.join(OBJDIR, s[:-3] + '.o')
		if force or _newer(obj, [src] + hdrs):
			jobs.append([nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src])

	def run(cmd):
		res = subprocess.run(cmd, capture_output=True, text=True)
		if res.returncode != 0:
			raise RuntimeError("nvcc failed:\n" + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
		return res.stderr

	if jobs:
		with ThreadPoolExecutor(len(jobs)) as pool:
			for err in pool.map(run, jobs):
				if verbose:
					print(err)
	objs = [os.path.join(OBJDIR, s[:-3] + '.o') for s in SOURCES]
	if jobs or _newer(LIBPATH, objs):
		run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIBPATH] + objs)
		open(flagfile, 'w').write(flags)
	return LIBPATH


if __name__ == '__main__':
	print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
