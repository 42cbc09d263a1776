"""
Per-kernel count of the SASS mnemonics that show how tiles are moved (UBLKCP = cp.async.bulk, UTMALDG = tensor-map TMA,
SYNCS = mbarrier, LDGSTS = cp.async, LDG / STG = plain global access) in the shipped library:
  python scripts/sass_grep.py [libtbk.so] > profiles/rNN_sass_grep.txt
"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'photometry_b200', 'lib', 'libtbk.so')
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
pat = re.compile(r'\b(UBLKCP[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|SYNCS[.\w]*|LDGSTS[.\w]*|LDG[.\w]*|STG[.\w]*|ATOMS[.\w]*|REDG?[.\w]*|ATOMG[.\w]*)')
per = collections.OrderedDict(); cur = None
for line in txt.split('\n'):
	m = re.search(r'Function : (\S+)', line)
	if m:
		cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0].replace('void ', '')
		per[cur] = collections.Counter(); continue
	if cur and '/*' in line:
		m = pat.search(line.split('/*', 2)[1] if line.count('/*') >= 2 else line)
		m = pat.search(line)
		if m:
			key = m.group(1)
			key = re.sub(r'\.(E|64|128|32|U8|U16|S8|S16|CONSTANT|STRONG|SYS|GPU|SM|EF|EL|LTC\d+B|BYPASS|ZFILL|MIN|MAX|ADD|POPC|INC|OR|AND|CAS|CAST|SPIN|F32|F64|FTZ|RN|S32|U32)\b', '', key)
			per[cur][key] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib)} (sm_100a), mnemonic counts per kernel; arch flags: see photometry_b200/build.py")
for k, c in per.items():
	if not c: continue
	print(f"{k}")
	print("    " + "  ".join(f"{n}={v}" for n, v in sorted(c.items())))
