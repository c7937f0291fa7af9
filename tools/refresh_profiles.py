"""Rebuild profiles/ from gpurun_out/: <round>_launches.csv (ncu launch list), a full capture
(.ncu-rep given on the command line) -> launch summary, full-capture summary, source hotspots,
traffic.json.  Usage: python tools/refresh_profiles.py gpurun_out/r02_prof.ncu-rep 1048576 [r02]"""
import collections, contextlib, csv, io, json, shutil, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'tools'))
import ncu_summary as ns

rep, n_evals = sys.argv[1], float(sys.argv[2])
RND = sys.argv[3] if len(sys.argv) > 3 else 'r02'
prof = ROOT / 'profiles'
shutil.copy(ROOT / 'gpurun_out' / f'{RND}_launches.csv', prof / f'{RND}_launches.csv')
rows = list(csv.reader(open(prof / f'{RND}_launches.csv')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; kn, mv, mn = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
        continue
    tot[r[kn]] += float(r[mv].replace(',', '')); cnt[r[kn]] += 1
T = sum(tot.values())
with open(prof / f'{RND}_launch_summary.csv', 'w') as f:
    f.write('kernel,launches,total_ms,avg_us,share_pct\n')
    for k, v in tot.most_common():
        f.write(f'"{k}",{cnt[k]},{v / 1e6:.3f},{v / cnt[k] / 1e3:.1f},{100 * v / T:.1f}\n')
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(out)))
d = dict(zip(rr[0], zip(rr[1], rr[2])))
keep = [k for k in ns.KEYS if k in d] + [k for k in ('launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed') if k in d]
with open(prof / f'{RND}_like_kernel_ncu_full_summary.csv', 'w') as f:
    f.write('metric,unit,value\n')
    f.write(f'kernel,,"{d["Kernel Name"][1]}"\n')
    for k in dict.fromkeys(keep):
        f.write(f'{k},{d[k][0]},{d[k][1]}\n')
u = d['dram__bytes_read.sum'][0]
mult = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[u]
traffic = (float(d['dram__bytes_read.sum'][1]) + float(d['dram__bytes_write.sum'][1])) * mult
json.dump({"nf_like_kernel_dram_bytes_per_launch": traffic,
           "source": f"profiles/{RND}_like_kernel_ncu_full_summary.csv (ncu --set full, one launch of {int(n_evals)} evals of {d['Kernel Name'][1]})"},
          open(prof / 'traffic.json', 'w'))
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    lines = ns.source(rep, n_evals)
rows2 = [(n / n_evals, s, src) for n, s, src in lines]
tot_s = sum(r[1] for r in rows2)
levels = collections.Counter()
for n, s, src in rows2:
    levels[round(n, 2)] += 1
with open(prof / f'{RND}_source_hotspots.txt', 'w') as f:
    f.write(f"{d['Kernel Name'][1]}, one launch of {int(n_evals)} evals (ncu --set full --import-source on, source page)\n")
    f.write(buf.getvalue())
    f.write("\nexecutions per eval x instructions at that count = warp-instructions per eval, share of stall samples\n")
    agg = collections.OrderedDict()
    for n, s, src in rows2:
        k = round(n, 2)
        a = agg.setdefault(k, [0, 0.0, 0])
        a[0] += 1; a[1] += n; a[2] += s
    for k, (ci, ti, si) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if ti >= 15:
            f.write(f"  {k:9.2f} x {ci:4d} = {ti:8.1f}   {100 * si / tot_s:5.1f}%\n")
print(open(prof / f'{RND}_launch_summary.csv').read())
print(open(prof / f'{RND}_source_hotspots.txt').read()[-900:])
print(open(prof / 'traffic.json').read())
