"""cuobjdump -sass convasr_b200/libconvasr_b200.so | python tools/sass_histogram.py > profiles/rNN_sass_opcode_histogram.md
Per-kernel SASS instruction counts with the Blackwell-specific opcode families singled out."""
import collections
import re
import subprocess
import sys

INTERESTING = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'FFMA2', 'RED', 'ATOM', 'LDGSTS', 'MUFU', 'SHFL', 'DADD']
fn, hist = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
	m = re.search(r'Function : (\S+)', line)
	if m:
		fn = m.group(1)
		continue
	m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
	if m and fn:
		hist[fn][m.group(1)] += 1
names = subprocess.run(['c++filt'], input = '\n'.join(hist), capture_output = True, text = True).stdout.splitlines()
rows = []
for (f, c), name in zip(hist.items(), names):
	fam = collections.Counter()
	for op, n in c.items():
		for key in INTERESTING:
			if op.startswith(key):
				fam[key] += n
				break
	rows.append((re.sub(r'\(.*', '', name).replace('cab::', ''), sum(c.values()), fam))
rows.sort(key = lambda r: -r[1])
print('# SASS opcode histogram of convasr_b200/libconvasr_b200.so (sm_100a)\n')
print('`cuobjdump -sass convasr_b200/libconvasr_b200.so | python tools/sass_histogram.py`.  Columns: total SASS instructions, then the telling opcode')
print('families: `UTCHMMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld (TMEM -> registers), `UTMALDG` / `UTMASTG` = TMA tensor load /')
print('store, `UBLKCP` = cp.async.bulk (1-D bulk copy), `SYNCS` = mbarrier operations, `LDGSTS` = cp.async, `FFMA2` = packed fp32 FMA, `RED` / `ATOM` =')
print('reductions / atomics, `DADD` = fp64 accumulation.  Template instantiations of one kernel are listed once (the largest).\n')
print('| kernel | instr | ' + ' | '.join(INTERESTING) + ' |')
print('|---|---|' + '---|' * len(INTERESTING))
seen = set()
for name, tot, fam in rows:
	base = re.sub(r'<.*', '', name)
	if base in seen or tot < 40:
		continue
	seen.add(base)
	print(f'| `{name[:90]}` | {tot} | ' + ' | '.join(str(fam.get(k, '')) for k in INTERESTING) + ' |')
