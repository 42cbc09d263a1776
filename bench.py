#!/usr/bin/env python3
"""
bench.py -- FFIs/sec background-fitted (2048 x 2048) on N B200 GPUs, and % of the HBM roofline.

Workload (BASELINE.json configs[1]): a synthetic single-CCD 30-min-cadence sector, 2048 x 2048 x 1,340
FFIs per GPU (camera 1, ccd 2, TESS path: 3 rounds of radial + mesh background).  One "step" = one
pass of ``fit_background`` over the whole device-resident 1,340-FFI cube (22.5 GB, far larger than the
126 MB L2, so every step streams from HBM).  With N > 1 every rank fits its own 1,340-FFI stack
(cadence shards need no exchange inside the fit), i.e. weak scaling.

Keys beyond the base contract: ``roofline`` (dominant kernel, CUDA-event timed), ``cpu_baseline``
(the CPU oracle under the reference's spawn-Pool driver on this box's cores), ``kernel_ms`` (device
time per kernel class for one step), ``prepare_path`` (fit + time smoothing + sumimage accumulation
[+ NCCL reduce], FFIs/s), ``shenanigans_path`` (the background-shenanigans stage on the same frames, FFIs/s;
algorithmic bytes per FFI = 2048^2 x 18: image read, indicator write + two re-reads, flags read + write).

``--impl reference`` times the reference's CPU path (restated oracle; the reference itself cannot be
imported here) with all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_FFI = 2048 * 2048 * (4 + 4 + 1)  # image read + background write + mask write
H = W = 2048
CAMERA, CCD = 1, 2
SEED = 20260117 + 1


def reference_probe():
	"""
	Can the real reference (tasoc/photometry) run here?  It needs astropy, photutils, statsmodels, bottleneck, h5py and its own
	package on sys.path (PHOTOMETRY_REFERENCE, default /root/reference); the GPU box has no /root/reference, so this normally
	reports False with the missing module list and the CPU arm times the oracle restatement (kind "port").
	"""
	import importlib.util
	missing = [m for m in ('astropy', 'photutils', 'statsmodels', 'bottleneck', 'h5py', 'scipy') if importlib.util.find_spec(m) is None]
	ref_dir = os.environ.get('PHOTOMETRY_REFERENCE', '/root/reference')
	have_pkg = os.path.isdir(os.path.join(ref_dir, 'photometry'))
	if not have_pkg:
		missing.append('photometry (reference package)')
	return {"reference_importable": not missing, "missing": missing}


def load_peaks():
	path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
	if os.path.exists(path):
		with open(path) as fid:
			return float(json.load(fid)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
	return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
	"""nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""
	Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

	def __init__(self, index):
		self.index = index
		self.lines = []            # (monotonic time of arrival, csv line)
		self.proc = None
		self.t0 = self.t1 = None

	def wait_first(self, timeout=5.0):
		"""nvidia-smi needs a few hundred ms to print its first sample: wait for it before the region of interest."""
		end = time.monotonic() + timeout
		while self.proc is not None and not self.lines and time.monotonic() < end:
			time.sleep(0.02)

	def mark_begin(self):
		self.t0 = time.monotonic()

	def mark_end(self):
		self.t1 = time.monotonic()

	def start(self):
		try:
			self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
				'--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
			self.thread = threading.Thread(target=self._read, daemon=True)
			self.thread.start()
		except OSError:
			self.proc = None

	def _read(self):
		for line in self.proc.stdout:
			self.lines.append((time.monotonic(), line.strip()))

	def stop(self):
		if self.proc is None:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
		self.proc.terminate()
		try:
			self.proc.wait(timeout=5)
		except Exception:
			self.proc.kill()
		sm, mx, reasons = [], [], set()
		names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
		# samples that arrived while the GPU was inside the marked region (a sample describes the ~100 ms before it)
		lines = [ln for t, ln in self.lines if self.t0 is None or (t >= self.t0 and (self.t1 is None or t <= self.t1 + 0.1))]
		for ln in lines:
			f = [x.strip() for x in ln.split(',')]
			if len(f) < 7:
				continue
			try:
				sm.append(float(f[0])); mx.append(float(f[1]))
			except ValueError:
				continue
			for nm, v in zip(names, f[3:7]):
				if v.lower().startswith('active'):
					reasons.add(nm)
		sm.sort()
		return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
			"samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
def _oracle_worker(args):
	img, hdr = args
	import oracle
	bkg, mask = oracle.fit_background(oracle.FFIImageLite(img, hdr, True), discarded_bins=True)
	return float(bkg[0, 0]), int(mask.sum())


def cpu_reference_run(images, headers, cores):
	"""The reference's driver shape (prepare.py:184-199, 291): spawn Pool(cores).imap over the FFIs."""
	import multiprocessing
	ctx = multiprocessing.get_context('spawn')
	items = list(zip(images, headers))
	with ctx.Pool(cores) as pool:
		list(pool.imap(_oracle_worker, items[:cores]))  # warm-up: interpreter start + imports
		t0 = time.perf_counter()
		list(pool.imap(_oracle_worker, items))
		dt = time.perf_counter() - t0
	return len(items) / dt, dt


def make_headers(n):
	return [dict(CAMERA=CAMERA, CCD=CCD, TSTART=1400.0 + k * 1800 / 86400, TSTOP=1400.0 + (k + 1) * 1800 / 86400,
		FFIINDEX=9000 + k, DQUALITY=(32 if k % 37 == 5 else 0)) for k in range(n)]


def host_sample(n):
	"""n synthetic FFIs of the benchmark workload as host arrays (generated on the GPU when there is one)."""
	import torch
	from photometry_b200 import synth
	if torch.cuda.is_available():
		cube = synth.synth_stack_torch(n, H, W, torch.device('cuda'), camera=CAMERA, ccd=CCD, seed=SEED)
		return cube.cpu().numpy()
	return synth.synth_stack_numpy(n, H, W, camera=CAMERA, ccd=CCD, seed=SEED)


def run_reference(args):
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	cores = os.cpu_count() or 1
	per_step = max(cores, 8) if args.ref_ffis is None else args.ref_ffis
	imgs = host_sample(per_step)
	hdrs = make_headers(per_step)
	import multiprocessing
	ctx = multiprocessing.get_context('spawn')
	items = list(zip(list(imgs), hdrs))
	with ctx.Pool(cores) as pool:
		for _ in range(max(args.warmup, 1) if args.warmup else 0):
			list(pool.imap(_oracle_worker, items[:cores]))
		t0 = time.perf_counter()
		for _ in range(args.steps):
			list(pool.imap(_oracle_worker, items))
		dt = time.perf_counter() - t0
	value = per_step * args.steps / dt
	sample = f"{per_step} FFIs per step x {args.steps} steps of the same synthetic 2048x2048 workload, spawn Pool({cores}).imap, in-memory float32 inputs"
	line = {
		"impl": "reference", "metric": "FFIs/sec background-fitted (2048x2048)", "value": value, "unit": "FFIs/s",
		"n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
		"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": "synthetic single CCD 2048x2048 x 1,340 FFIs (30-min cadence sector), bounded sample", "sample_ffis_per_step": per_step},
		"cpu_baseline": {"value": value, "unit": "FFIs/s", "cores": cores, "kind": "port", "sample": sample},
		"e2e": {"value": value, "unit": "FFIs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}
	line.update(reference_probe())
	print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def sharded_parity(fit_cls, dev, world, rank, dist):
	"""
	Sharded-vs-single check carried in the bench line: a small replicated stack (24 cadences of 512 x 512) is run through
	prepare_stack on every rank's cadence shard (halo exchange + packed reduce) and, on rank 0, through the same kernels
	unsharded.  Returns booleans (rank 0) for time_smooth 9 and 27.
	"""
	import numpy as np
	import torch
	import photometry_b200 as pb
	from photometry_b200 import synth
	from photometry_b200.prepare import shard_bounds
	n, Hs, Ws = 8 * max(world, 4), 512, 512
	xycen = (-30.0, 560.0)
	stack = synth.synth_stack_numpy(n, Hs, Ws, seed=77, xycen=xycen, radial_cutoff=500.0, n_stars=600)
	hdrs = [dict(CAMERA=1, CCD=2, TSTART=1400.0 + 0.02 * k, TSTOP=1400.02 + 0.02 * k, FFIINDEX=9000 + k, DQUALITY=(32 if k % 5 == 2 else 0)) for k in range(n)]
	meta = pb.meta_from_headers(hdrs)
	fit = fit_cls((Hs, Ws), True, 1, 2, radial_cutoff=500, xycen=xycen, device=dev.index)
	out = {}
	for ts in (9, 27):
		if n // world < ts // 2:
			continue
		lo, hi = shard_bounds(n, world, rank)
		res = pb.prepare_stack(fit, torch.from_numpy(stack[lo:hi]).to(dev), meta[lo:hi], time_smooth=ts, chunk=8)
		parts = [torch.empty((shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0], Hs, Ws), dtype=torch.float32, device=dev) for r in range(world)]
		dist.all_gather(parts, res.backgrounds.contiguous())
		if rank == 0:
			cube = torch.from_numpy(stack).to(dev)
			bk, mk, st = fit.fit(cube, meta)
			sm = fit.time_smooth(bk, ts // 2)
			s = torch.zeros((Hs, Ws), dtype=torch.float64, device=dev); ni = torch.zeros((Hs, Ws), dtype=torch.int32, device=dev); us = torch.zeros_like(ni)
			fit.sum_accumulate(cube, sm, mk.clone(), meta, s, ni, us)
			sumimage, used = fit.sum_finalize(s, ni, us, n, 0.5)
			out[f"time_smooth_{ts}"] = {
				"backgrounds_bit_equal": bool(torch.equal(torch.cat(parts), sm)),
				"nimg_used_exact": bool(torch.equal(res.nimg, ni) and torch.equal(res.used, us) and torch.equal(res.backgrounds_pixels_used, used)),
				"sumimage_rtol_1e-12": bool(torch.allclose(res.sumimage, sumimage, rtol=1e-12, equal_nan=True)),
				"numfiles": int(res.numfiles)}
	return out


def sector_path(pb, synth, dist, dev, rank, world, total_ffis, time_smooth, camera, ccd, seed, chunk, streams, barrier, label):
	"""
	One synthetic sector of ``total_ffis`` cadences sharded contiguously over the ranks (strong scaling): fit + halo exchange
	(w = time_smooth // 2) + time smoothing + sumimage accumulation + packed NCCL reduce + finalize.  Returns the dict for the
	bench line (rank 0) -- FFIs/s over the max-over-ranks device time, and the halo / reduce milliseconds.
	"""
	import torch
	from photometry_b200.prepare import shard_bounds
	lo, hi = shard_bounds(total_ffis, world, rank)
	n_loc = hi - lo
	cube = synth.synth_stack_torch(n_loc, H, W, dev, camera=camera, ccd=ccd, seed=seed + rank)
	hdrs = [dict(CAMERA=camera, CCD=ccd, TSTART=1400.0 + k * 600 / 86400, TSTOP=1400.0 + (k + 1) * 600 / 86400,
		FFIINDEX=20000 + k, DQUALITY=(32 if k % 37 == 5 else 0)) for k in range(lo, hi)]
	meta = pb.meta_from_headers(hdrs)
	fit = pb.BackgroundFitter((H, W), True, camera, ccd, device=dev.index)
	tm = {}
	pb.prepare_stack(fit, cube[:min(n_loc, 2 * chunk)], meta[:min(n_loc, 2 * chunk)], time_smooth=time_smooth, chunk=chunk, keep_images=False, nstreams=streams)
	barrier()
	g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	g0.record()
	res = pb.prepare_stack(fit, cube, meta, time_smooth=time_smooth, chunk=chunk, keep_images=False, timings=tm, nstreams=streams)
	g1.record()
	barrier()
	ms = g0.elapsed_time(g1)
	t = torch.tensor([ms, tm.get('halo_ms', 0.0), tm.get('reduce_ms', 0.0)], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	ms, halo_ms, reduce_ms = (float(x) for x in t.tolist())
	out = {"value": total_ffis / (ms * 1e-3), "unit": "FFIs/s", "scaling": "strong", "total_ffis": total_ffis, "ffis_per_gpu": n_loc,
		"time_smooth": time_smooth, "halo_w": time_smooth // 2, "ms": ms, "halo_ms": halo_ms, "reduce_ms": reduce_ms,
		"limits": "fit kernels" if max(halo_ms, reduce_ms) < 0.1 * ms else ("halo exchange" if halo_ms > reduce_ms else "reduce"),
		"numfiles": int(res.numfiles), "workload": label}
	del res, cube
	torch.cuda.empty_cache()
	return out


def run_b200(args):
	import numpy as np
	import torch
	import torch.distributed as dist
	import photometry_b200 as pb
	from photometry_b200 import synth, _lib
	import __graft_entry__
	rank = int(os.environ.get('RANK', '0'))
	world = int(os.environ.get('WORLD_SIZE', '1'))
	local_rank = int(os.environ.get('LOCAL_RANK', '0'))
	if not torch.cuda.is_available():
		raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
	torch.cuda.set_device(local_rank)
	dev = torch.device('cuda', local_rank)
	from photometry_b200 import affinity
	aff = affinity.bind_to_gpu(local_rank) if not args.no_affinity else {"bound": False}
	if rank == 0:
		__graft_entry__.build()
	if world > 1:
		os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
		dist.init_process_group('nccl', device_id=dev)
		dist.barrier()
	n = args.ffis
	chunk = args.chunk
	cube = synth.synth_stack_torch(n, H, W, dev, camera=CAMERA, ccd=CCD, seed=SEED + rank)
	hdrs = make_headers(n)
	meta = pb.meta_from_headers(hdrs)
	fit = pb.BackgroundFitter((H, W), True, CAMERA, CCD, device=local_rank)
	meta_d = fit.meta_to_device(meta)
	isz = meta.dtype.itemsize
	bkg = torch.empty_like(cube)
	mask = torch.empty(cube.shape, dtype=torch.uint8, device=dev)
	lib = _lib.load()

	def step():
		fit.fit_stack(cube, meta_d, bkg, mask, chunk=chunk, nstreams=args.streams)

	def barrier():
		torch.cuda.synchronize(dev)
		if world > 1:
			dist.barrier()
			torch.cuda.synchronize(dev)

	sampler = ClockSampler(local_rank)
	if rank == 0:
		sampler.start()
	for _ in range(args.warmup):
		step()
	barrier()
	if rank == 0:
		sampler.wait_first()
		sampler.mark_begin()
	l0 = lib.tbk_launch_count()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(args.steps):
		step()
	e1.record()
	barrier()
	launches = int(lib.tbk_launch_count() - l0)
	ms = e0.elapsed_time(e1)
	clocks = None
	if rank == 0:
		sampler.mark_end()
		window = "timed region"
		if sum(1 for t, _ in sampler.lines if t >= sampler.t0) < 2 and world == 1:
			# a timed region shorter than two 100-ms sampling periods: keep the same load running (untimed) for the sampler
			end = time.monotonic() + 0.6
			while time.monotonic() < end:
				step()
				torch.cuda.synchronize(dev)
			sampler.mark_end()
			window = "timed region + 0.6 s of the same load (untimed)"
		clocks = sampler.stop()
		clocks["window"] = window
	if world > 1:
		t = torch.tensor([ms], dtype=torch.float64, device=dev)
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		ms = float(t.item())
		lt = torch.tensor([launches], dtype=torch.int64, device=dev)
		dist.all_reduce(lt, op=dist.ReduceOp.SUM)
		launches = int(lt.item())
	value = world * n * args.steps / (ms * 1e-3)

	# ---- per-kernel device times for one step (events between launches) -> roofline of the dominant kernel
	prof = {}
	for a in range(0, n, chunk):
		b = min(a + chunk, n)
		fit.fit(cube[a:b], meta_d[a * isz:b * isz], bkg_out=bkg[a:b], mask_out=mask[a:b], profile=prof)
	ncalls = (n + chunk - 1) // chunk
	dom = max((k for k in prof if k not in ('misc', 'fallback')), key=lambda k: prof[k])
	# launches per tbk_fit_batch and class (3 rounds): zone statistics of the raw pixels; per round producer + finish for the
	# residuals; 'fallback' = the bucketed kernels for the queued meshes (1 + 3; on a side stream outside the profiled call)
	launches_per_call = {'tile_base': 1, 'tile_round': 6, 'fallback': 5, 'zp_min': 4, 'ring_gather': 3, 'ring_kde': 3, 'radial_fit': 3, 'mesh': 3, 'final': 1}
	dom_launch_ms = prof[dom] / (ncalls * launches_per_call.get(dom, 1))
	peak, peak_src = load_peaks()
	# The fit is a chain of ~33 kernel launches per batch and no single kernel dominates (the largest is < 30 % of
	# the step), so the roofline is stated for the whole chain: algorithmic bytes of one tbk_fit_batch launch
	# (37,748,736 B x FFIs per launch) over the summed device time of its kernels (CUDA events between the
	# launches, tbk_fit_batch_profiled).  The dominant kernel is reported beside it with its share of the step.
	kern_total_ms = sum(prof.values())
	launch_ms = kern_total_ms / ncalls
	achieved = ALGO_BYTES_PER_FFI * min(chunk, n) / (launch_ms * 1e-3) / 1e9
	traffic = None
	dom_traffic = None
	try:
		with open(os.path.join(ROOT, 'profiles', 'r02_ncu_fit_summary.json')) as fid:
			prof_json = json.load(fid)
		traffic = prof_json['dram_bytes_per_ffi'] * min(chunk, n)
		for kname, e in prof_json['kernels'].items():
			if kname.split('<')[0].startswith('k_' + dom):
				dom_traffic = (e['dram_read_bytes'] + e['dram_write_bytes']) / e['launches'] / prof_json['ffis_per_launch'] * min(chunk, n)
	except (OSError, KeyError, ValueError, AttributeError, TypeError, ZeroDivisionError):
		pass
	launches_per_batch = launches // max(args.steps * ncalls, 1)
	roofline = {"bound": "hbm", "kernel": f"tbk_fit_batch (chain of {launches_per_batch} launches; sum of kernel device times)",
		"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
		"traffic_source": "profiles/r02_ncu_fit_summary.json (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum over the chain, scaled to this launch size)",
		"peak_source": peak_src, "launch_ms": launch_ms, "ffis_per_launch": min(chunk, n),
		"limiter": "instruction issue of the exact sigma-clip statistics and the KDE sweeps, not HBM (see DESIGN.md)",
		"dominant_kernel": {"name": "k_" + dom, "share_of_step": prof[dom] / kern_total_ms, "launch_ms": dom_launch_ms,
			"launches_per_batch": launches_per_call.get(dom, 1), "traffic": dom_traffic},
		"kernel_shares": {k: round(v / kern_total_ms, 4) for k, v in prof.items()},
		"with_streams_achieved": value / world * ALGO_BYTES_PER_FFI / 1e9, "with_streams_frac": value / world * ALGO_BYTES_PER_FFI / 1e9 / peak}

	# ---- end to end: pinned host stack -> device -> fit -> pinned host results (same metric)
	ne = min(args.e2e_ffis, n, max(64, 2048 // world))   # keep the pinned host footprint bounded when 8 ranks share one host
	host_in = torch.empty((ne, H, W), dtype=torch.float32).pin_memory()
	host_in.copy_(cube[:ne])
	host_bkg = torch.empty((ne, H, W), dtype=torch.float32).pin_memory()
	host_mask = torch.empty((ne, H, W), dtype=torch.uint8).pin_memory()
	def e2e_step():
		return pb.fit_stack_host(fit, host_in, meta[:ne], host_bkg, host_mask, chunk=args.e2e_chunk, pack_mask=os.environ.get('TBK_E2E_PACK', '1') != '0')
	e2e_step()
	barrier()
	esteps = max(1, min(args.steps, 3))
	f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	# wall clock between two synchronisations: the step ends when the host threads have expanded the last mask, which no
	# CUDA event sees
	t_e0 = time.perf_counter()
	for _ in range(esteps):
		h2d, d2h = e2e_step()
	torch.cuda.synchronize(dev)
	ems = (time.perf_counter() - t_e0) * 1e3
	barrier()
	if world > 1:
		t = torch.tensor([ems], dtype=torch.float64, device=dev)
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		ems = float(t.item())
	e2e_value = world * ne * esteps / (ems * 1e-3)
	# spot-check the transferred result against the resident one
	assert torch.allclose(host_bkg[ne - 1], bkg[ne - 1].cpu(), rtol=1e-6, equal_nan=True)
	assert torch.equal(host_mask[ne - 1], mask[ne - 1].cpu())
	# copy-only ceiling at this N: the same pinned buffers, the same bytes in both directions (the mask crosses as bits), no kernels
	bits_dev = torch.empty((ne, H * W // 8), dtype=torch.uint8, device=dev)
	host_bits = torch.empty((ne, H * W // 8), dtype=torch.uint8).pin_memory()
	def copy_step():
		s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
		ck = args.e2e_chunk
		for a in range(0, ne, ck):
			b = min(a + ck, ne)
			with torch.cuda.stream(s_in):
				cube[a:b].copy_(host_in[a:b], non_blocking=True)
			with torch.cuda.stream(s_out):
				host_bkg[a:b].copy_(bkg[a:b], non_blocking=True)
				host_bits[a:b].copy_(bits_dev[a:b], non_blocking=True)
		torch.cuda.current_stream(dev).wait_stream(s_in); torch.cuda.current_stream(dev).wait_stream(s_out)
	copy_step()
	barrier()
	f0.record()
	for _ in range(esteps):
		copy_step()
	f1.record()
	barrier()
	cms = f0.elapsed_time(f1)
	if world > 1:
		t = torch.tensor([cms], dtype=torch.float64, device=dev)
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		cms = float(t.item())
	copy_ceiling = world * ne * esteps / (cms * 1e-3)
	del host_in, host_bkg, host_mask, host_bits, bits_dev

	# ---- prepare path: fit + time smoothing + sumimage accumulation (+ NCCL reduce)
	prep = shen = stamps_path = None
	if args.prepare:
		np_ = min(n, args.prepare_ffis)
		barrier()
		g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		pb.prepare_stack(fit, cube[:np_], meta[:np_], time_smooth=3, chunk=chunk, keep_images=True, nstreams=args.streams)
		barrier()
		# three timed calls, the median is reported: a call that has to grow the allocator's pool pays ~10 ms of cudaMalloc
		runs = []
		for _ in range(3):
			res = None
			ptm = {}
			barrier()
			g0.record()
			res = pb.prepare_stack(fit, cube[:np_], meta[:np_], time_smooth=3, chunk=chunk, keep_images=True, timings=ptm, nstreams=args.streams)
			g1.record()
			barrier()
			runs.append((g0.elapsed_time(g1), ptm))
		runs.sort(key=lambda r: r[0])
		pms, ptm = runs[1]
		if world > 1:
			t = torch.tensor([pms], dtype=torch.float64, device=dev)
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
			pms = float(t.item())
		prep = {"value": world * np_ / (pms * 1e-3), "unit": "FFIs/s", "ffis_per_gpu": np_, "numfiles": res.numfiles,
			"stages": "fit + time_smooth(w=1) + sum_accumulate + reduce + finalize", "ms": pms, "halo_ms": ptm.get('halo_ms'), "reduce_ms": ptm.get('reduce_ms')}
		# ---- background shenanigans (prepare.py:514-622) on the same frames: 15 x 15 median indicator per cadence,
		# robust mean over shuffled blocks (cadence -> row-slab exchange when N > 1), flagging
		flags_copy = res.pixel_flags.clone()
		pb.background_shenanigans(res.images, res.sumimage, flags_copy)
		barrier()
		g0.record()
		pb.background_shenanigans(res.images, res.sumimage, flags_copy)
		g1.record()
		barrier()
		sms = g0.elapsed_time(g1)
		if world > 1:
			t = torch.tensor([sms], dtype=torch.float64, device=dev)
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
			sms = float(t.item())
		shen = {"value": world * np_ / (sms * 1e-3), "unit": "FFIs/s", "ffis_per_gpu": np_,
			"stages": "median-filter indicator + robust mean (blocks of 25) + flagging",
			"algorithmic_bytes_per_ffi": 2048 * 2048 * 18, "hbm_frac": world * np_ / (sms * 1e-3) * 2048 * 2048 * 18 / (peak * 1e9 * world)}
		# ---- consumer-side cube loads (BasePhotometry._load_cube): 2,000 stamps of 15 x 15 pixels over the frames
		srv = pb.StampServer(images=res.images)
		rs = np.random.default_rng(5)
		r0 = rs.integers(0, H - 15, 2000); c0 = rs.integers(0, W - 15, 2000)
		stamps_arr = np.stack([r0, r0 + 15, c0 + 44, c0 + 15 + 44], axis=1)
		srv.load_cubes(stamps_arr, views=False)
		barrier()
		g0.record()
		cubes = srv.load_cubes(stamps_arr, views=False)
		g1.record()
		barrier()
		gms = g0.elapsed_time(g1)
		if world > 1:
			t = torch.tensor([gms], dtype=torch.float64, device=dev)
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
			gms = float(t.item())
		moved = 2000 * 15 * 15 * np_ * 4 * 2
		stamps_path = {"value": world * 2000 / (gms * 1e-3), "unit": "stamps/s", "stamps": 2000, "stamp": "15x15", "cadences": np_,
			"achieved_gbs": moved / (gms * 1e-3) / 1e9, "hbm_frac": moved / (gms * 1e-3) / 1e9 / peak,
			"note": "bytes = cube elements read + written; reads are 60-byte row segments (32-byte sectors)"}
		del res, flags_copy, cubes, srv

	cpu_sample = cube[:min(max(os.cpu_count() or 1, 8), 64)].cpu().numpy() if (rank == 0 and world == 1 and not args.no_cpu) else None
	# ---- north-star multi-GPU configurations (BASELINE.json configs[2], configs[3]), bounded; the resident benchmark cube
	# is released first (a 4,000-FFI sector needs the memory)
	config3 = config4 = parity = None
	if args.configs:
		del cube, bkg, mask
		torch.cuda.empty_cache()
		if world > 1:
			parity = sharded_parity(pb.BackgroundFitter, dev, world, rank, dist)
			# config 3: one 10-min-cadence sector of 4,000 FFIs, camera 4 ccd 1, time_smooth = 9 (w = 4), strong scaling
			config3 = sector_path(pb, synth, dist, dev, rank, world, args.config3_ffis, 9, 4, 1, SEED + 100, chunk, args.streams, barrier,
				"synthetic 10-min cadence sector, 2048x2048 x %d FFIs sharded by cadence over %d GPUs" % (args.config3_ffis, world))
		# config 4: 200-s cadence camera, 4 CCDs one after the other, time_smooth = 27 (w = 13: the reference's 5,400-s window),
		# per-CCD reduce; bounded to config4_ffis cadences per GPU and CCD
		c4 = []
		for ccd4 in (1, 2, 3, 4):
			c4.append(sector_path(pb, synth, dist, dev, rank, world, args.config4_ffis * world, 27, 2, ccd4, SEED + 200 + 10 * ccd4, chunk, args.streams, barrier,
				"synthetic 200-s cadence camera 2 ccd %d, %d cadences per GPU (bounded sample of 12,000 per CCD)" % (ccd4, args.config4_ffis)))
		tot_ms = sum(x["ms"] for x in c4)
		config4 = {"value": sum(x["total_ffis"] for x in c4) / (tot_ms * 1e-3), "unit": "FFIs/s", "scaling": "weak", "ccds": 4,
			"ffis_per_gpu_per_ccd": args.config4_ffis, "time_smooth": 27, "halo_w": 13, "ms": tot_ms,
			"halo_ms": sum(x["halo_ms"] for x in c4), "reduce_ms": sum(x["reduce_ms"] for x in c4), "per_ccd": c4}

	if rank != 0:
		if world > 1:
			dist.destroy_process_group()
		return

	# ---- CPU baseline on this box's host cores (rank 0, N = 1 only; bounded sample)
	cpu = None
	if world == 1 and not args.no_cpu:
		cores = os.cpu_count() or 1
		ns = min(max(cores, 8), 64)
		imgs = cpu_sample
		try:
			v, dt = cpu_reference_run(list(imgs), hdrs[:ns], cores)
			cpu = {"value": v, "unit": "FFIs/s", "cores": cores, "kind": "port",
				"sample": f"first {ns} FFIs of the benchmark cube, oracle restatement (numpy/scipy), spawn Pool({cores}).imap, {dt:.1f} s"}
		except Exception as err:  # the baseline must never take the bench line down
			cpu = {"value": None, "unit": "FFIs/s", "cores": cores, "kind": "port", "sample": f"failed: {err!r}"}

	line = {
		"metric": "FFIs/sec background-fitted (2048x2048)", "value": value, "unit": "FFIs/s", "n_gpus": world,
		"steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
		"scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": "synthetic single CCD 2048x2048 x 1,340 FFIs (30-min cadence sector), TESS path camera 1 ccd 2, 3 rounds",
			"ffis_per_gpu": n, "ffis_per_launch": chunk, "streams": args.streams, "l2": "inputs (22.5 GB/GPU) larger than L2", "parallelism": f"cadence shards x{world}"},
		"roofline": roofline, "cpu_baseline": cpu,
		"e2e": {"value": e2e_value, "unit": "FFIs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
			"ffis_per_step": ne, "copy_ceiling": copy_ceiling, "frac_of_ceiling": e2e_value / copy_ceiling,
			"copy_ceiling_gbs": copy_ceiling * (4 + 4 + 0.125) * H * W / 1e9, "affinity": aff,
			"note": "one e2e step = fit_stack_host over a pinned host stack, wall clock between synchronisations; results (bkg f32 + mask u8) "
				"end up in pinned host memory -- the mask crosses the link as bits and is expanded by host threads (the D2H direction is the "
				"bottleneck: 21 MB out against 16.8 MB in per FFI with a byte mask); "
				"copy_ceiling = the same pinned buffers and bytes in both directions with no kernels, at this N"},
		"gpu_launches": launches, "clocks": clocks,
		"kernel_ms": {k: round(v, 3) for k, v in prof.items()}, "prepare_path": prep,
		"shenanigans_path": shen, "stamps_path": stamps_path,
		"config3_path": config3, "config4_path": config4, "sharded_parity": parity,
	}
	line.update(reference_probe())
	print(json.dumps(line), flush=True)
	if world > 1:
		dist.destroy_process_group()


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=3)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
	ap.add_argument('--ffis', type=int, default=1340, help='FFIs per GPU per step')
	ap.add_argument('--chunk', type=int, default=64, help='FFIs per tbk_fit_batch launch')
	ap.add_argument('--streams', type=int, default=4, help='CUDA streams the chunks alternate between')
	ap.add_argument('--no-configs', dest='configs', action='store_false', help='skip the config 3 / config 4 / sharded parity legs')
	ap.add_argument('--config3-ffis', type=int, default=4000, help='cadences of the sharded 10-min sector (config 3, N > 1 only)')
	ap.add_argument('--config4-ffis', type=int, default=192, help='cadences per GPU and CCD of the bounded 200-s camera (config 4)')
	ap.add_argument('--no-affinity', action='store_true', help='do not bind the process to the CPUs local to its GPU')
	ap.add_argument('--e2e-ffis', type=int, default=512, help='pinned host stack size for the end-to-end leg')
	ap.add_argument('--e2e-chunk', type=int, default=16)
	ap.add_argument('--prepare-ffis', type=int, default=256)
	ap.add_argument('--no-prepare', dest='prepare', action='store_false')
	ap.add_argument('--no-cpu', action='store_true')
	ap.add_argument('--ref-ffis', type=int, default=None)
	args = ap.parse_args()
	if args.impl == 'reference':
		run_reference(args)
	else:
		run_b200(args)


if __name__ == '__main__':
	main()
