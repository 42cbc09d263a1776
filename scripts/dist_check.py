"""
Multi-GPU parity check (run with torchrun --nproc-per-node N): every rank runs prepare_stack on its contiguous
cadence shard (halo exchange + NCCL reduce inside); rank 0 also runs the whole stack alone and compares.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import photometry_b200 as pb
from photometry_b200 import synth
from photometry_b200.prepare import shard_bounds

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n, H, W = 24, 512, 512
xycen = (-30.0, 560.0)
stack = synth.synth_stack_numpy(n, H, W, seed=77, xycen=xycen, radial_cutoff=500.0, n_stars=600)
hdrs = [dict(CAMERA=1, CCD=2, TSTART=1400.0 + 0.02 * k, TSTOP=1400.02 + 0.02 * k, FFIINDEX=9000 + k, DQUALITY=(32 if k % 5 == 2 else 0)) for k in range(n)]
meta = pb.meta_from_headers(hdrs)
fit = pb.BackgroundFitter((H, W), True, 1, 2, radial_cutoff=500, xycen=xycen, device=local)
for ts in (3, 9):
	lo, hi = shard_bounds(n, world, rank)
	res = pb.prepare_stack(fit, torch.from_numpy(stack[lo:hi]).cuda(), meta[lo:hi], time_smooth=ts, chunk=5)
	torch.cuda.synchronize()
	parts = [None] * world
	dist.all_gather_object(parts, res.backgrounds.cpu().numpy())
	if rank == 0:
		saved = os.environ.pop('WORLD_SIZE')  # single-process reference on this rank only
		full_fit = pb.BackgroundFitter((H, W), True, 1, 2, radial_cutoff=500, xycen=xycen, device=local)
		# run without the process group: emulate by calling the pieces directly
		cube = torch.from_numpy(stack).cuda()
		bk, mk, st = full_fit.fit(cube, meta)
		sm = full_fit.time_smooth(bk, ts // 2)
		s = torch.zeros((H, W), dtype=torch.float64, device='cuda'); ni = torch.zeros((H, W), dtype=torch.int32, device='cuda'); us = torch.zeros_like(ni)
		full_fit.sum_accumulate(cube, sm, mk.clone(), meta, s, ni, us)
		sumimage, used = full_fit.sum_finalize(s, ni, us, n, 0.5)
		os.environ['WORLD_SIZE'] = saved
		got = np.concatenate(parts)
		assert np.array_equal(got, sm.cpu().numpy()), "sharded smoothing differs"
		assert res.numfiles == n
		assert torch.equal(res.nimg, ni) and torch.equal(res.used, us)
		assert torch.allclose(res.sumimage, sumimage, rtol=1e-12, equal_nan=True)
		assert torch.equal(res.backgrounds_pixels_used, used)
		print(f"time_smooth={ts}: {world}-rank sharded prepare_stack == single-GPU result (backgrounds bit-equal, sumimage rtol 1e-12)", flush=True)
# ---- background shenanigans: cadence shards -> row slabs over NCCL must reproduce the single-GPU stage
g = torch.Generator().manual_seed(11)
n2, H2, W2 = 61, 130, 192
imgs = torch.randn((n2, H2, W2), generator=g) * 6
imgs[7, 30:80, 40:150] += 90
imgs[40, 10:60, 100:190] -= 70
imgs[3, 5, 5] = float('nan')
sumimage = torch.randn((H2, W2), generator=g, dtype=torch.float64)
flags0 = (torch.rand((n2, H2, W2), generator=g) < 0.1).to(torch.uint8) * 3
lo, hi = shard_bounds(n2, world, rank)
fl = flags0[lo:hi].clone().cuda()
mean = pb.background_shenanigans(imgs[lo:hi].cuda(), sumimage.cuda() if rank == 0 else None, fl)
torch.cuda.synchronize()
parts = [None] * world
dist.all_gather_object(parts, fl.cpu().numpy())
means = [None] * world
dist.all_gather_object(means, mean.cpu().numpy())
if rank == 0:
	# single-GPU reference through the non-distributed pieces
	ind = pb.shenanigans_indicator(imgs.cuda(), sumimage.cuda())
	m1 = pb.mean_shenanigans(ind)
	f1 = flags0.clone().cuda()
	pb.flag_shenanigans(ind, m1, f1)
	assert all(np.array_equal(mm, m1.cpu().numpy()) for mm in means), "sharded mean_shenanigans differs"
	assert np.array_equal(np.concatenate(parts), f1.cpu().numpy()), "sharded shenanigans flags differ"
	print(f"background shenanigans: {world}-rank result == single-GPU result (mean image bit-equal on every rank, flags exact; "
		f"{int((f1 & 4).bool().sum())} pixels flagged)", flush=True)
dist.barrier()
dist.destroy_process_group()
