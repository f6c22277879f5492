"""Per-launch roofline table from an `ncu --set full` raw CSV export (ncu -i X.ncu-rep --page raw --csv)."""
import csv
import json
import os
import sys


def fnum(x):
	try:
		return float(x.replace(',', ''))
	except Exception:
		return float('nan')


def to_bytes(v, unit):
	return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def to_us(v, unit):
	return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3, 'second': 1e6}.get(unit, 1)


def main(path, peaks_path = None):
	rows = list(csv.reader(open(path)))
	hdr, units, data = rows[0], rows[1], rows[2:]
	idx = {h: i for i, h in enumerate(hdr)}
	hbm = 6548.2
	if peaks_path and os.path.exists(peaks_path):
		hbm = json.load(open(peaks_path))['hbm_gbs']

	def col(r, name):
		return fnum(r[idx[name]]), units[idx[name]]

	print(f'| # | kernel | grid x block | time us | SM clk GHz | DRAM rd MB | DRAM wr MB | DRAM GB/s (% of {hbm:.0f} measured) | tensor pipe % | L2 thr % | regs |')
	print('|---|---|---|---|---|---|---|---|---|---|---|')
	for i, r in enumerate(data):
		name = r[idx['Kernel Name']].replace('cab::', '').split('(')[0][:40]
		t = to_us(*col(r, 'gpu__time_duration.sum'))
		rd = to_bytes(*col(r, 'dram__bytes_read.sum'))
		wr = to_bytes(*col(r, 'dram__bytes_write.sum'))
		cyc = fnum(r[idx['sm__cycles_elapsed.max']]) if 'sm__cycles_elapsed.max' in idx else float('nan')
		ghz = cyc / t / 1e3 if t > 0 else float('nan')
		tens = fnum(r[idx['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]) if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in idx else float('nan')
		l2 = fnum(r[idx['lts__throughput.avg.pct_of_peak_sustained_elapsed']])
		gbs = (rd + wr) / (t * 1e-6) / 1e9
		grid = r[idx['launch__grid_size']] if 'launch__grid_size' in idx else '?'
		block = r[idx['launch__block_size']] if 'launch__block_size' in idx else '?'
		regs = r[idx['launch__registers_per_thread']]
		print(f'| {i} | {name} | {grid} x {block} | {t:.1f} | {ghz:.2f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} ({gbs / hbm * 100:.1f}%) | {tens:.1f} | {l2:.1f} | {regs} |')


if __name__ == '__main__':
	main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
