"""Copy-only ceiling of the end-to-end path with the mask sent as bytes (4 MiB / FFI) or as bits (0.5 MiB / FFI)."""
import torch, time
dev = torch.device('cuda:0')
n, H, W, chunk = 256, 2048, 2048, 16
hin = torch.empty((n, H, W), dtype=torch.float32).pin_memory(); hb = torch.empty((n, H, W), dtype=torch.float32).pin_memory()
for label, mbytes in (('mask as bytes', H * W), ('mask as bits', H * W // 8)):
	hm = torch.empty((n, mbytes), dtype=torch.uint8).pin_memory()
	din = [torch.empty((chunk, H, W), dtype=torch.float32, device=dev) for _ in range(3)]
	db = [torch.empty((chunk, H, W), dtype=torch.float32, device=dev) for _ in range(3)]
	dm = [torch.empty((chunk, mbytes), dtype=torch.uint8, device=dev) for _ in range(3)]
	s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
	def run():
		for i, a in enumerate(range(0, n, chunk)):
			k = i % 3
			with torch.cuda.stream(s_in): din[k].copy_(hin[a:a + chunk], non_blocking=True)
			with torch.cuda.stream(s_out):
				hb[a:a + chunk].copy_(db[k], non_blocking=True); hm[a:a + chunk].copy_(dm[k], non_blocking=True)
		torch.cuda.synchronize()
	run()
	t0 = time.perf_counter(); run(); run(); dt = (time.perf_counter() - t0) / 2
	gb = n * (H * W * 8 + mbytes) / 1e9
	print(f"{label}: {n / dt:.0f} FFIs/s, {gb / dt:.1f} GB/s both directions", flush=True)
