"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv): per kernel time, issue rate, stalls, pipes."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
def g(d, k):
	try: return float(d[idx[k]].replace(',', ''))
	except Exception: return float('nan')
seen = set()
for d in data:
	name = d[idx['Kernel Name']].split('(')[0]
	if name in seen and '--all' not in sys.argv: continue
	seen.add(name)
	print(f"== {name}  grid={d[idx['Grid Size']]} block={d[idx['Block Size']]}")
	print(f"   time {g(d,'gpu__time_duration.sum'):.4f} {units[idx['gpu__time_duration.sum']]}  regs {g(d,'launch__registers_per_thread'):.0f}  warps_active {g(d,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}%  issue_active {g(d,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}%")
	print(f"   dram rd {g(d,'dram__bytes_read.sum'):.1f} wr {g(d,'dram__bytes_write.sum'):.1f} {units[idx['dram__bytes_read.sum']]}  dram% {g(d,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f}  l1hit {g(d,'l1tex__t_sector_hit_rate.pct'):.1f}  l2hit {g(d,'lts__t_sector_hit_rate.pct'):.1f}")
	print(f"   inst {g(d,'smsp__inst_executed.sum')/1e6:.1f} M  thr/inst {g(d,'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f}  smem bank conflicts {g(d,'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')/1e6:.2f} M")
	pipes = ['alu', 'fma', 'fp64', 'xu', 'lsu', 'adu', 'cbu', 'uniform']
	print('   pipes% ' + ' '.join(f"{p}={g(d, f'sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active'):.1f}" for p in pipes))
	st = [(g(d, h), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for h in hdr if h.startswith('smsp__average_warps_issue_stalled_')]
	st.sort(reverse=True)
	print('   stalls(warps/issue) ' + ' '.join(f"{n}={v:.2f}" for v, n in st[:7]))
