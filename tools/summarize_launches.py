"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the
per-step share of each kernel over the LAST complete device step (frontend .. ctc_grad_scatter for the
inference workload, frontend .. mt_update_kernel for a training step)."""
import collections
import csv
import sys


def load(path):
	with open(path) as f:
		lines = [l for l in f if not l.startswith('==')]
	rows = list(csv.DictReader(lines))
	out = []
	for r in rows:
		v = float(r['Metric Value'].replace(',', ''))
		unit = r['Metric Unit']
		us = v * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(unit, 1e-3)
		out.append((r['Kernel Name'], r['Grid Size'], r['Block Size'], us))
	return out


def short(name):
	name = name.replace('cab::', '').replace('void ', '')
	return name.split('(')[0][:48]


def main(path):
	rows = load(path)
	print(f'# {len(rows)} launches in {path}')
	# last complete step: from the last absmax launch that is followed by the step's final kernel
	end_name = sys.argv[2] if len(sys.argv) > 2 else ('mt_update_kernel' if any('mt_update_kernel' in r[0] for r in rows) else 'ctc_grad_scatter')
	starts = [i for i, r in enumerate(rows) if 'absmax' in r[0]]
	ends = [i for i, r in enumerate(rows) if end_name in r[0]]
	step = None
	for s in reversed(starts):
		e = [x for x in ends if x > s]
		nxt = [x for x in starts if x > s]
		if e and (not nxt or e[0] < nxt[0]):
			step = (s, e[0])
			break
	if step is None:
		print('no complete step found')
		return
	seg = rows[step[0]:step[1] + 1]
	tot = sum(r[3] for r in seg)
	print(f'\n## last complete device step: launches {step[0]}..{step[1]} ({len(seg)} launches, {tot:.1f} us serialised, cold cache)\n')
	agg = collections.OrderedDict()
	for name, grid, block, us in seg:
		a = agg.setdefault(short(name), [0, 0.0])
		a[0] += 1
		a[1] += us
	print('| kernel | launches | total us | share |\n|---|---|---|---|')
	for k, (n, t) in sorted(agg.items(), key = lambda kv: -kv[1][1]):
		print(f'| {k} | {n} | {t:.1f} | {t / tot * 100:.1f}% |')
	print('\n### tensor-pipe kernel launches in step order\n')
	print('| # | kernel | grid | us |\n|---|---|---|---|')
	i = 0
	for name, grid, block, us in seg:
		if 'conv1d_umma' in name or 'wgrad_umma' in name:
			print(f'| {i} | {short(name)} | {grid} | {us:.1f} |')
			i += 1


if __name__ == '__main__':
	main(sys.argv[1])
