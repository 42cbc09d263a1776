"""
Turn the raw ncu exports brought back in gpurun_out/ into the tracked summaries under profiles/:
  python scripts/make_profiles.py ROUND LAUNCHES_CSV FIT_RAW_CSV NFFI [SHE_RAW_CSV]
  -> profiles/rNN_launches_bench.csv, rNN_launches_bench_shares.txt, rNN_ncu_fit_summary.txt / .json, rNN_ncu_shenanigans_summary.txt
"""
import csv, io, json, os, shutil, subprocess, sys, collections

rnd, launches_csv, fit_raw, nffi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
she_raw = sys.argv[5] if len(sys.argv) > 5 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = os.path.join(root, 'profiles')

# ---- launch list of the bench command
lines = open(launches_csv).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
cmd = next((l for l in lines[:start] if 'bench.py' in l), '')
tot = collections.OrderedDict()
for r in rows:
	if r.get('Metric Name') != 'gpu__time_duration.sum':
		continue
	name = r['Kernel Name'].split('(')[0].replace('void ', '').replace('(bool)', '').replace('(int)', '')
	us = float(r['Metric Value'].replace(',', ''))
	unit = r['Metric Unit']
	us *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)
	t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += us
shutil.copy(launches_csv, os.path.join(prof, f'{rnd}_launches_bench.csv'))
fit_names = [k for k in tot if not k.startswith('k_bkgshe') and not k.startswith('k_time') and not k.startswith('k_sum') and not k.startswith('k_zero') and not k.startswith('k_decode')]
total = sum(tot[k][1] for k in fit_names)
with open(os.path.join(prof, f'{rnd}_launches_bench_shares.txt'), 'w') as f:
	f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_  python bench.py --steps 2 --warmup 1 --ffis 128 --no-cpu --no-prepare --e2e-ffis 32\n")
	f.write(f"# (the bench command on a 128-FFI cube so that the serialised ncu pass stays short; full list: {rnd}_launches_bench.csv)\n")
	f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
	for k in sorted(fit_names, key=lambda k: -tot[k][1]):
		f.write(f"{k:28s} launches={tot[k][0]:5d} total_us={tot[k][1]:12.1f} share={100 * tot[k][1] / total:5.1f}%\n")

# ---- full capture of one tbk_fit_batch
txt = subprocess.run([sys.executable, os.path.join(root, 'scripts', 'ncu_summary.py'), fit_raw, '--all'], capture_output=True, text=True).stdout
open(os.path.join(prof, f'{rnd}_ncu_fit_summary.txt'), 'w').write(txt)
rows = list(csv.reader(open(fit_raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
def g(d, k):
	try: return float(d[idx[k]].replace(',', ''))
	except Exception: return float('nan')
def scale(unit, kind):
	return {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)
kern = collections.OrderedDict()
for d in data:
	name = d[idx['Kernel Name']].split('(')[0].replace('void ', '')
	t = g(d, 'gpu__time_duration.sum') * scale(units[idx['gpu__time_duration.sum']], 't')
	rd = g(d, 'dram__bytes_read.sum') * scale(units[idx['dram__bytes_read.sum']], 'b')
	wr = g(d, 'dram__bytes_write.sum') * scale(units[idx['dram__bytes_write.sum']], 'b')
	e = kern.setdefault(name, dict(launches=0, time_us=0.0, dram_read_bytes=0.0, dram_write_bytes=0.0, inst=0.0))
	e['launches'] += 1; e['time_us'] += t; e['dram_read_bytes'] += rd; e['dram_write_bytes'] += wr
	e['inst'] += g(d, 'smsp__inst_executed.sum')
total_us = sum(e['time_us'] for e in kern.values())
for e in kern.values():
	e['share'] = e['time_us'] / total_us
nl = sum(e['launches'] for e in kern.values())
out = dict(source=f"ncu --set full --clock-control none --import-source on -k regex:^k_ -s {nl} -c {nl} python scripts/prof_run.py {nffi}: the {nl} kernel launches of one tbk_fit_batch over {nffi} synthetic 2048x2048 TESS FFIs (cold-cache, serialised)",
	ffis_per_launch=nffi, kernels=kern, total_time_us=total_us,
	dram_bytes_per_ffi=sum(e['dram_read_bytes'] + e['dram_write_bytes'] for e in kern.values()) / nffi, algorithmic_bytes_per_ffi=2048 * 2048 * 9)
json.dump(out, open(os.path.join(prof, f'{rnd}_ncu_fit_summary.json'), 'w'), indent=1)
print('fit chain: %d launches, %.1f us, %.1f MB DRAM per FFI' % (sum(e['launches'] for e in kern.values()), out['total_time_us'], out['dram_bytes_per_ffi'] / 1e6))
if she_raw:
	txt = subprocess.run([sys.executable, os.path.join(root, 'scripts', 'ncu_summary.py'), she_raw], capture_output=True, text=True).stdout
	open(os.path.join(prof, f'{rnd}_ncu_shenanigans_summary.txt'), 'w').write(
		"# ncu --set full --clock-control none -k regex:k_bkgshe python scripts/dev_she.py 8: the three kernels of the background-shenanigans stage on 8 synthetic 2048x2048 frames\n" + txt)
