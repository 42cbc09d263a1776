"""prepare_stack phase timing under torchrun (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import photometry_b200 as pb
from photometry_b200 import synth
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 256
cube = synth.synth_stack_torch(n, 2048, 2048, dev, camera=1, ccd=2, seed=5 + rank)
hdrs = [dict(CAMERA=1, CCD=2, TSTART=1400.0 + k * 0.0208, TSTOP=1400.0208 + k * 0.0208, FFIINDEX=9000 + k) for k in range(n)]
meta = pb.meta_from_headers(hdrs)
fit = pb.BackgroundFitter((2048, 2048), True, 1, 2, device=local)
for keep, ns in ((True, 1), (True, 1), (False, 1), (False, 4), (True, 4)):
	torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
	tm = {}
	t0 = time.perf_counter()
	res = pb.prepare_stack(fit, cube, meta, time_smooth=3, chunk=64, keep_images=keep, timings=tm, nstreams=ns)
	torch.cuda.synchronize()
	dt = time.perf_counter() - t0
	print(f"rank {rank} keep_images={keep} nstreams={ns}: {dt * 1e3:.1f} ms wall, halo {tm['halo_ms']:.2f} ms, reduce {tm['reduce_ms']:.2f} ms", flush=True)
	del res
dist.barrier(); dist.destroy_process_group()
