"""Summarise an ncu report of the round kernel: key raw metrics + hottest SASS by executed count and stall samples."""
import csv, collections, subprocess, sys
rep = sys.argv[1]; nchunks = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
for h, u, v in zip(hdr, units, vals):
    if h in want or 'issue_stalled' in h and 'per_issue_active' in h and float(v or 0) > 0.2:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = rows[2:]
iS, iE, iP = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iP]) for r in data)
print('SASS rows', len(data), 'warp-instr', tot, 'per chunk', tot / nchunks if nchunks else '')
by = collections.Counter(); bys = collections.Counter()
for r in data:
    m = r[iS].split(); op = (m[1] if m[0].startswith('@') else m[0]).split('.')[0]
    by[op] += int(r[iE]); bys[op] += int(r[iP])
for op, c in by.most_common(16): print(f'  {op:10s} {100 * c / tot:5.1f}% instr  {100 * bys[op] / totS:5.1f}% samples')
