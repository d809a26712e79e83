"""Summarise an .ncu-rep (raw page) into a compact table: one row per captured kernel launch."""
import csv
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'dur'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
    ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64%'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu%'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smem%'),
    ('smsp__inst_executed.sum', 'inst'),
]
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print('| kernel | ' + ' | '.join(n for _, n in KEYS) + ' |')
print('|---|' + '---|' * len(KEYS))
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')
    cells = []
    for k, _ in KEYS:
        if k in hdr:
            v, u = r[hdr.index(k)], units[hdr.index(k)]
            try:
                f = float(v.replace(',', ''))
                v = f'{f:.4g}'
            except ValueError:
                pass
            cells.append(f'{v} {u}'.strip() if u not in ('%', '') else v)
        else:
            cells.append('-')
    print(f'| {name} | ' + ' | '.join(cells) + ' |')
