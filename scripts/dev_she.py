"""Timing of the background-shenanigans kernels on full-size frames (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import photometry_b200 as pb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(3)
imgs = torch.randn((n, 2048, 2048), device=dev, generator=g) * 8
imgs[:, 500:600, 700:900] += 60 * torch.rand((n, 1, 1), device=dev, generator=g)
sm = torch.randn((2048, 2048), device=dev, generator=g, dtype=torch.float64)
flags = torch.zeros((n, 2048, 2048), dtype=torch.uint8, device=dev)
ind = torch.empty_like(imgs)


def timed(f, reps=3):
	best = 1e30
	for _ in range(reps):
		torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
		e0.record(); r = f(); e1.record(); torch.cuda.synchronize()
		best = min(best, e0.elapsed_time(e1))
	return best, r


ms, _ = timed(lambda: pb.shenanigans_indicator(imgs, sm, out=ind))
print(f"indicator: {ms / n * 1e3:.1f} us/FFI ({n / ms * 1e3:.0f} FFIs/s)")
ms2, mean = timed(lambda: pb.mean_shenanigans(ind))
print(f"mean: {ms2:.2f} ms for {n} FFIs ({ms2 / n * 1e3:.1f} us/FFI)")
ms3, _ = timed(lambda: pb.flag_shenanigans(ind, mean, flags))
print(f"flag: {ms3 / n * 1e3:.1f} us/FFI; flagged fraction {float((flags & 4).bool().float().mean()):.4f}")
print(f"stage: {(ms + ms2 + ms3) / n * 1e3:.1f} us/FFI -> {n / (ms + ms2 + ms3) * 1e3:.0f} FFIs/s")
