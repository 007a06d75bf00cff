"""usage: python scripts/debug/ncu_top.py <file.ncu-rep> [n]  -- key metrics + the SASS lines with the most stall samples"""
import csv, subprocess, sys
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); h, u = r[0], r[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
want += [k for k in h if 'issue_stalled' in k and 'per_issue_active' in k and 'not_issued' not in k]
for row in r[2:3]:
    for w in want:
        if w in h:
            v = row[h.index(w)]
            try:
                if 'issue_stalled' in w and float(v) < 0.15: continue
            except ValueError:
                pass
            print(f"{w:90s} {v} {u[h.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows, hdr, k = [], None, 0
for row in csv.reader(src.splitlines()):
    if row and row[0] == 'Kernel Name': k += 1; continue
    if row and row[0] == 'Address': hdr = row; continue
    if k == 1 and hdr: rows.append(row)
ia, ie = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
tot = sum(int(x[ia]) for x in rows if x[ia].isdigit())
print('total samples', tot, 'sass rows', len(rows))
top = sorted(range(len(rows)), key=lambda i: -int(rows[i][ia]) if rows[i][ia].isdigit() else 0)[:top_n]
for i in sorted(top):
    print(f"{i:5d} {rows[i][1].strip()[:70]:70s} {rows[i][ia]:>7s} {rows[i][ie]:>10s}")
