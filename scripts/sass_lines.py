"""
Join an ncu source-page CSV (per-SASS-instruction counters) with nvdisasm line info and aggregate by source line.

  cuobjdump -xelf all photometry_b200/lib/libtbk.so ; nvdisasm -gi -c tbk_fit.sm_100a.cubin > fit_gi.sass
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
  python scripts/sass_lines.py fit_gi.sass MANGLED_SUBSTRING src.csv [--top 40] [--by outer|inner]
"""
import argparse, csv, io, re, collections, sys

ap = argparse.ArgumentParser()
ap.add_argument('sass'); ap.add_argument('kernel'); ap.add_argument('csv')
ap.add_argument('--top', type=int, default=40)
ap.add_argument('--depth', type=int, default=0, help='0 = outermost frame, 1 = one level below, ... -1 = innermost')
ap.add_argument('--ops', action='store_true', help='aggregate by SASS opcode instead')
ap.add_argument('--under', default=None, help='only instructions whose inline chain contains this file:line')
ap.add_argument('--launch', type=int, default=0, help='which launch of the kernel in the csv')
args = ap.parse_args()

# 1. instruction offset -> (inner file:line, outer tbk_fit.cu line)
lines = open(args.sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('\t.section\t.text.') and args.kernel in l)
loc = {}
cur_inner = cur_outer = None
chain = []
cont = False
for l in lines[start + 1:]:
	if l.startswith('\t.section'):
		break
	m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
	if m:
		a = f"{m.group(1).split('/')[-1]}:{m.group(2)}"
		b = f"{m.group(3).split('/')[-1]}:{m.group(4)}" if m.group(3) else None
		if not (chain and chain[-1] == a and cont):
			chain = [a]
		if b:
			chain.append(b)
		cont = b is not None
		continue
	m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', l)
	if m:
		loc[int(m.group(1), 16)] = (tuple(chain), m.group(2).strip())
		cont = False

if args.csv == '-':   # static mode: SASS instructions per source line (code size)
	agg = collections.Counter()
	for off, (ch, op) in loc.items():
		ch = ch or ('?',)
		if args.under and args.under not in ch:
			continue
		agg[ch[0] if args.depth < 0 else ch[max(len(ch) - 1 - args.depth, 0)]] += 1
	print(f"{sum(agg.values())} SASS instructions")
	for k, n in agg.most_common(args.top):
		print(f"{k:34s} {n:6d}")
	sys.exit(0)

# 2. csv rows of the requested launch
txt = open(args.csv).read()
blocks = re.split(r'(?m)^"Kernel Name",', txt)[1:]
blk = blocks[args.launch]
rows = list(csv.reader(io.StringIO(blk[blk.index('\n') + 1:])))
hdr = rows[0]; rows = rows[1:]
ia, ii, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = int(rows[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
for r in rows:
	if len(r) <= isamp or not r[ia].startswith('0x'):
		continue
	off = int(r[ia], 16) - base
	ch, op = loc.get(off, (('?',), ''))
	ch = ch or ('?',)
	if args.under and args.under not in ch:
		continue
	key = ch[0] if args.depth < 0 else ch[max(len(ch) - 1 - args.depth, 0)]
	if args.ops:
		key = op.split()[1] if op.startswith('@') else op.split()[0]
	n, s = int(r[ii] or 0), int(r[isamp] or 0)
	agg[key][0] += n; agg[key][1] += s
	tot[0] += n; tot[1] += s
print(f"total warp-instructions {tot[0]}  samples {tot[1]}  ({len(rows)} SASS instructions, {len(loc)} with line info)")
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
	print(f"{k:34s} inst {100.0 * n / tot[0]:6.2f} %   samples {100.0 * s / max(tot[1], 1):6.2f} %")
