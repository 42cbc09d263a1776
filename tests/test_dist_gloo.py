"""
CPU tests of the N > 1 host path with the gloo backend (world_size 2): cadence sharding, halo
exchange for the smoothing window and the accumulator reduce, checked against the single-process oracle.
"""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
	s = socket.socket()
	s.bind(('127.0.0.1', 0))
	port = s.getsockname()[1]
	s.close()
	return port


def _worker(rank, world, port, n, w, out_dir):
	os.environ['MASTER_ADDR'] = '127.0.0.1'
	os.environ['MASTER_PORT'] = str(port)
	dist.init_process_group('gloo', rank=rank, world_size=world)
	import oracle
	from photometry_b200.prepare import shard_bounds, exchange_halos, reduce_accumulators
	rng = np.random.default_rng(5)
	bkg = rng.normal(100, 2, (n, 16, 16)).astype('float32')
	bkg[3, 1, 1] = np.nan
	imgs = rng.normal(130, 3, (n, 16, 16)).astype('float32')
	flags = (rng.uniform(size=(n, 16, 16)) < 0.1).astype('uint8')
	quality = np.where(np.arange(n) % 4 == 1, 32, 0).astype('int32')
	lo, hi = shard_bounds(n, world, rank)
	local = torch.from_numpy(bkg[lo:hi].copy())
	h_lo, h_hi = exchange_halos(local, w)
	parts = [t for t in (h_lo, local, h_hi) if t is not None]
	ext = torch.cat(parts).numpy()
	sm = oracle.time_smooth_backgrounds(ext, 2 * w + 1)
	off = 0 if h_lo is None else h_lo.shape[0]
	sm_local = sm[off:off + (hi - lo)]
	# edges of the sector: the extended stack must not have borrowed frames that do not exist
	res = oracle.sumimage_accumulate(imgs[lo:hi], sm_local, flags[lo:hi], quality[lo:hi])
	s = torch.from_numpy(np.where(np.isnan(res['sumimage']), 0, res['sumimage'] * res['nimg']))
	ni = torch.from_numpy(res['nimg'].copy())
	us = torch.from_numpy(res['used'].copy())
	total = reduce_accumulators(s, ni, us, hi - lo)
	np.save(os.path.join(out_dir, f'sm_{rank}.npy'), sm_local)
	if rank == 0:
		np.savez(os.path.join(out_dir, 'root.npz'), sum=s.numpy(), nimg=ni.numpy(), used=us.numpy(), total=total)
	dist.destroy_process_group()


def test_two_rank_halo_and_reduce(tmp_path):
	import oracle
	n, w, world = 11, 2, 2
	port = _free_port()
	mp.spawn(_worker, args=(world, port, n, w, str(tmp_path)), nprocs=world, join=True)
	rng = np.random.default_rng(5)
	bkg = rng.normal(100, 2, (n, 16, 16)).astype('float32')
	bkg[3, 1, 1] = np.nan
	imgs = rng.normal(130, 3, (n, 16, 16)).astype('float32')
	flags = (rng.uniform(size=(n, 16, 16)) < 0.1).astype('uint8')
	quality = np.where(np.arange(n) % 4 == 1, 32, 0).astype('int32')
	ref_sm = oracle.time_smooth_backgrounds(bkg, 2 * w + 1)
	got = np.concatenate([np.load(tmp_path / f'sm_{r}.npy') for r in range(world)])
	np.testing.assert_array_equal(got, ref_sm)  # bit-exact: same float32 accumulation order
	ref = oracle.sumimage_accumulate(imgs, ref_sm, flags, quality)
	root = np.load(tmp_path / 'root.npz')
	assert int(root['total']) == n
	np.testing.assert_array_equal(root['nimg'], ref['nimg'])
	np.testing.assert_array_equal(root['used'], ref['used'])
	np.testing.assert_allclose(root['sum'] / root['nimg'], ref['sumimage'], rtol=1e-12)
