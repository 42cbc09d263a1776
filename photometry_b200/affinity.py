"""
Host-side placement for the end-to-end path: one process per GPU moves ~38 MB per FFI through pinned host buffers, so the
process (and, by first touch, its pinned pages) should live on the NUMA node the GPU hangs off.  ``bind_to_gpu`` restricts the
calling process to the CPUs local to the GPU (sysfs ``local_cpulist`` of the PCI device); pinned buffers allocated afterwards
are then placed on that node by the kernel's first-touch policy.  Everything here is best effort: on a box without the sysfs
entries (or a single NUMA node) it does nothing.
"""
import os


def _parse_cpulist(text):
	cpus = set()
	for part in text.strip().split(','):
		if not part:
			continue
		if '-' in part:
			a, b = part.split('-')
			cpus.update(range(int(a), int(b) + 1))
		else:
			cpus.add(int(part))
	return cpus


def gpu_local_cpus(device_index):
	"""CPUs local to CUDA device ``device_index`` (empty set when unknown)."""
	bus = None
	try:
		import torch
		pr = torch.cuda.get_device_properties(device_index)
		bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
	except Exception:
		try:
			import pynvml
			pynvml.nvmlInit()
			h = pynvml.nvmlDeviceGetHandleByIndex(_visible_to_physical(device_index))
			bus = pynvml.nvmlDeviceGetPciInfo(h).busId
			bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
			if len(bus.split(':')[0]) == 8:      # nvml style 00000000:1b:00.0 -> sysfs 0000:1b:00.0
				bus = bus[4:]
		except Exception:
			return set()
	path = f'/sys/bus/pci/devices/{bus}/local_cpulist'
	try:
		with open(path) as fid:
			return _parse_cpulist(fid.read())
	except OSError:
		return set()


def _visible_to_physical(device_index):
	vis = os.environ.get('CUDA_VISIBLE_DEVICES')
	if vis:
		try:
			return int(vis.split(',')[device_index])
		except (ValueError, IndexError):
			pass
	return device_index


def bind_to_gpu(device_index, world_on_node=1, local_rank=0):
	"""
	Restrict this process to the CPUs local to its GPU; when several ranks share those CPUs they are split evenly so the
	ranks' copy threads do not sit on each other.  Returns a dict describing what was done (for the bench line).
	"""
	info = {"bound": False, "cpus": None, "numa_cpus": 0}
	if not hasattr(os, 'sched_setaffinity'):
		return info
	cpus = sorted(gpu_local_cpus(device_index) & os.sched_getaffinity(0))
	info["numa_cpus"] = len(cpus)
	if not cpus:
		return info
	try:
		os.sched_setaffinity(0, set(cpus))
		info["bound"] = True
		info["cpus"] = f"{cpus[0]}-{cpus[-1]} ({len(cpus)})"
	except OSError:
		pass
	return info
