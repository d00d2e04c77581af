"""Builds profiles/<tag>_summary.md and profiles/traffic.json from the ncu exports of one round:
    python profiles/make_summary.py r01g
expects profiles/<tag>_ncu_full_wavefront_raw.csv (ncu -i prof.ncu-rep --page raw --csv) and
profiles/<tag>_launches_bench_steps2.csv (ncu --metrics gpu__time_duration.sum ... --csv)."""
import collections
import csv
import json
import sys

import os

tag = sys.argv[1]
rows = list(csv.reader(open(f'profiles/{tag}_ncu_full_wavefront_raw.csv')))
hdr, units = rows[0], rows[1]
sep = f'profiles/{tag}_ncu_full_separate_raw.csv'          # the separate sample / pdf kernels (same columns)
if os.path.exists(sep):
    extra = list(csv.reader(open(sep)))
    assert extra[0] == hdr
    rows += extra[2:]


def val(r, k):
    i = hdr.index(k)
    return float(r[i].replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[i], 1)


def num(r, k):
    return float(r[hdr.index(k)].replace(',', ''))


out, summ = {}, []
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    short = name.split('<')[1].split('>')[0].split('<')[0].split(',')[0]
    t = num(r, 'gpu__time_duration.sum')
    tu = units[hdr.index('gpu__time_duration.sum')]
    t_ms = t / 1e3 if tu == 'us' else (t / 1e6 if tu == 'ns' else t)
    out[short] = {"dram_bytes_read": val(r, 'dram__bytes_read.sum'), "dram_bytes_write": val(r, 'dram__bytes_write.sum'), "gpu_time_ms": t_ms,
                  "source": f"profiles/{tag}_ncu_full_wavefront_raw.csv (ncu --set full, 16Mi vertices per launch)"}
    summ.append((short, t_ms, num(r, 'smsp__inst_executed.sum'), num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                 num(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'), num(r, 'lts__t_sector_hit_rate.pct'),
                 num(r, 'l1tex__t_sector_hit_rate.pct'), num(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
                 num(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
                 num(r, 'launch__registers_per_thread'), val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum'),
                 num(r, 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed') if 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed' in hdr else float('nan'),
                 num(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed') if 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed' in hdr else float('nan')))
json.dump(out, open('profiles/traffic.json', 'w'), indent=1)

lines = [l for l in open(f'profiles/{tag}_launches_bench_steps2.csv') if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[row['Kernel Name']][0] += 1
    agg[row['Kernel Name']][1] += v
    seq.append((row['Kernel Name'], v))
wf = [(n, v) for n, v in seq if 'k_wavefront' in n and 'Locate' not in n]
steps = [wf[i:i + 2] for i in range(len(wf) - 1) if 'SamplePdfLane' in wf[i][0] and 'SplatRecords' in wf[i + 1][0]]
st = steps[4] if len(steps) > 4 else steps[-1]
tot = sum(v for _, v in st)
with open(f'profiles/{tag}_summary.md', 'w') as f:
    f.write(f"# Profile summary {tag} (B200, bench.py config: 16 Mi vertices per launch, tree 8193 spatial nodes / 1.62 M quadtree nodes)\n\n")
    f.write(f"Source files: `{tag}_ncu_full_wavefront_raw.csv` (ncu --set full --clock-control none, one launch of each wavefront kernel), "
            f"`{tag}_launches_bench_steps2.csv` (gpu__time_duration.sum of every launch of `bench.py --steps 2 --warmup 3`), `{tag}_bench_n1.json` (the plain bench line).\n\n")
    f.write("| kernel | ncu time ms | warp instr | issue active % | threads/instr | L2 hit % | L1 hit % | L2 throughput % | DRAM throughput % | warps active % | regs | DRAM bytes | L1->L2 request port busy % | LSU data pipe % |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for s_ in summ:
        f.write(f"| k_wavefront<{s_[0]}> | {s_[1]:.3f} | {s_[2]:.3e} | {s_[3]:.1f} | {s_[4]:.1f} | {s_[5]:.1f} | {s_[6]:.1f} | {s_[7]:.1f} | {s_[8]:.1f} | {s_[9]:.1f} | {int(s_[10])} | {s_[11] / 1e6:.0f} MB | {s_[12]:.1f} | {s_[13]:.1f} |\n")
    f.write("\n(L1->L2 request port = `l1tex__m_l1tex2xbar_req_cycles_active`: one slot per 32 B sector a lane asks for -- the unit the per-lane descents saturate first; "
            "`bench.py`'s `roofline.request_roof` measures the same against the random-sector gather probe.)\n")
    f.write("\nTensor pipes: idle by design (nothing on this path is a dense contraction).\n\n")
    f.write("Share of one step in the serialised launch list (cold-cache, per ncu): " +
            ", ".join(f"{n.split('<')[1].split('>')[0].split('<')[0].split(',')[0]} {v:.0f} us ({v / tot * 100:.0f} %)" for n, v in st) +
            f"; the CUDA-event shares in `{tag}_bench_n1.json` agree.\n\n")
    ref = sum(v for n, (c, v) in agg.items() if 'k_wavefront' not in n and 'k_l2_read' not in n)
    cnt = sum(c for n, (c, v) in agg.items() if 'k_wavefront' not in n and 'k_l2_read' not in n)
    f.write(f"Refine + sweeps (all `k_scan_*`, `k_items<...>`, `k_single<...>` launches): {cnt} launches, {ref / 1e3:.2f} ms of kernel time over the 7 refines "
            f"of the bench run (tree-build iterations + timed ones), serialised by ncu; the wall time of one refine is `per_iteration.refine_ms` of the bench line "
            f"({json.load(open(f'profiles/{tag}_bench_n1.json'))['per_iteration']['refine_ms']:.2f} ms: a chain of dependent launches whose latencies overlap by programmatic dependent launch).\n")
print(open(f'profiles/{tag}_summary.md').read())
