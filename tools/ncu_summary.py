"""Summarise an .ncu-rep: key metrics of each captured kernel (raw page) and, with --source,
per-opcode / per-line instruction counts and stall samples from the source page."""
import csv, subprocess, sys, collections, io

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.sum.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']

def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('kernel:', d.get('Kernel Name'))
        for k in KEYS:
            if k in d: print(f'  {k:90s} {d[k]}')

def source(rep, n_evals):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and any('Instructions Executed' in c for c in r))
    hdr = rows[hi]
    ci = hdr.index('Source'); ce = hdr.index('Instructions Executed')
    cs = next(i for i, c in enumerate(hdr) if c.startswith('Warp Stall Sampling (All'))
    ops = collections.Counter(); stalls = collections.Counter(); tot = 0; stot = 0
    lines = []
    for r in rows[hi + 1:]:
        if len(r) <= max(ci, ce, cs): continue
        try: n = int(r[ce]); s = int(r[cs])
        except ValueError: continue
        toks = r[ci].replace('@', ' ').split()
        op = toks[1] if toks and (toks[0].startswith('P') or toks[0].startswith('!P') or toks[0].startswith('UP') or toks[0].startswith('!UP')) and len(toks) > 1 else (toks[0] if toks else '?')
        op = op.split('.')[0]
        ops[op] += n; stalls[op] += s; tot += n; stot += s
        lines.append((n, s, r[ci]))
    print(f'total warp-instructions {tot}  per eval {tot / n_evals:.1f}   stall samples {stot}')
    for op, n in ops.most_common(28):
        print(f'  {op:10s} {n / n_evals:9.1f} {100 * n / tot:5.1f}%   stall {100 * stalls[op] / max(stot, 1):5.1f}%')
    return lines

if __name__ == '__main__':
    rep = sys.argv[1]
    raw(rep)
    if '--source' in sys.argv:
        n_evals = float(sys.argv[sys.argv.index('--source') + 1])
        source(rep, n_evals)

def dump_lines(rep, n_evals, out_path):
    lines = source(rep, n_evals)
    with open(out_path, 'w') as f:
        for n, s, src in lines:
            f.write(f'{n / n_evals:9.2f} {s:7d}  {src}\n')
