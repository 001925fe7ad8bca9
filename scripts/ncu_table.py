#!/usr/bin/env python
"""Per-kernel table from an `ncu --set full` report: mean duration, DRAM bytes, DRAM / L2 / issue / tensor
utilisation and the top stall reasons, aggregated over the launches of each kernel name.
usage: ncu_table.py report.ncu-rep out.json [out.txt]"""
import collections
import csv
import json
import subprocess
import sys

M = {'dur': 'gpu__time_duration.sum', 'rd': 'dram__bytes_read.sum', 'wr': 'dram__bytes_write.sum',
     'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'lts_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
     'l1_pct': 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
     'issue_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
     'warps_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
     'tensor_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
     'regs': 'launch__registers_per_thread', 'grid': 'launch__grid_size', 'block': 'launch__block_size',
     'bank_conf': 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
     'st_long_sb': 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
     'st_short_sb': 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
     'st_barrier': 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
     'st_mio': 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
     'st_math': 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
     'st_wait': 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
     'st_lg': 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
     'st_nosel': 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
     'st_sleep': 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
     'st_membar': 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio'}
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0,
         'nsecond': 1e-3, 'msecond': 1e3, 'second': 1e6}
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0]
    d = agg.setdefault(name, collections.defaultdict(list))
    for k, col in M.items():
        if col in hdr:
            i = hdr.index(col)
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            d[k].append(v * SCALE.get(units[i], 1.0))
out = []
for name, d in agg.items():
    o = {'kernel': name, 'launches': len(d['dur'])}
    for k in M:
        if d[k]:
            o[k] = sum(d[k]) / len(d[k])
    out.append(o)
json.dump(out, open(sys.argv[2], "w"), indent=1)
lines = ["%-44s %3s %8s %8s %6s %5s %5s %5s %5s %4s  top stalls" % ("kernel", "n", "us", "dramMB", "dram%", "lts%", "iss%", "wrp%", "tens%", "regs")]
for o in sorted(out, key=lambda o: -o.get('dur', 0) * o['launches']):
    st = sorted(((k[3:], v) for k, v in o.items() if k.startswith('st_')), key=lambda kv: -kv[1])[:3]
    lines.append("%-44s %3d %8.1f %8.2f %6.1f %5.1f %5.1f %5.1f %5.1f %4d  %s" % (
        o['kernel'][-44:], o['launches'], o.get('dur', 0), (o.get('rd', 0) + o.get('wr', 0)) / 1e6, o.get('dram_pct', 0),
        o.get('lts_pct', 0), o.get('issue_pct', 0), o.get('warps_pct', 0), o.get('tensor_pct', 0), int(o.get('regs', 0)),
        " ".join("%s=%.1f" % kv for kv in st)))
print("\n".join(lines))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join(lines) + "\n")
