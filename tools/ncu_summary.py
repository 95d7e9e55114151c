"""Summarise an .ncu-rep (raw + source pages) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'sm__cycles_elapsed.max',
        'smsp__thread_inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_cbu.sum',
        'sm__inst_executed_pipe_adu.sum']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, 'raw')
    hdr, units = raw[0], raw[1]
    for k, row in enumerate(raw[2:]):
        print('== launch %d: %s' % (k, row[hdr.index('Kernel Name')][:70]))
        for key in KEYS:
            if key in hdr:
                print('   %-66s %s %s' % (key, row[hdr.index(key)], units[hdr.index(key)]))
        st = [(h, row[i]) for i, h in enumerate(hdr) if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
        st.sort(key=lambda x: -float(x[1] or 0))
        print('   stalls (warps per issue):', ', '.join('%s=%.2f' % (h.split('stalled_')[1].split('_per_issue')[0], float(v)) for h, v in st[:7]))
    src = page(rep, 'source')
    hdr, kern, data = None, 0, {}
    for row in src:
        if row and row[0] == 'Kernel Name':
            kern += 1
            continue
        if row and row[0] == 'Address':
            hdr = row
            continue
        if hdr and len(row) == len(hdr):
            data.setdefault(kern, []).append(row)
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    d = data.get(want, [])
    if not d:
        return
    iS, iSrc = hdr.index('# Samples'), hdr.index('Source')
    tot = sum(int(r[iS]) for r in d) or 1
    print('== source page, launch %d: %d samples over %d SASS instructions' % (want, tot, len(d)))
    region, acc, cnt = 0, collections.Counter(), collections.Counter()
    for r in d:
        acc[region] += int(r[iS]); cnt[region] += 1
        if 'BAR.SYNC' in r[iSrc]:
            region += 1
    print('   samples by region between barriers:', {k: '%.1f%% (%d instr)' % (100 * v / tot, cnt[k]) for k, v in acc.items()})
    ie = [h for h in hdr if h.endswith('Instructions Executed') and 'Thread' not in h]
    if ie:
        iE = hdr.index(ie[0])
        mix = collections.Counter()
        for r in d:
            parts = r[iSrc].split()
            op = parts[1] if parts and parts[0].startswith('@') and len(parts) > 1 else (parts[0] if parts else '?')
            mix[op.split('.')[0]] += int(r[iE] or 0)
        tote = sum(mix.values()) or 1
        print('   dynamic warp-instruction mix (%d executed):' % tote, ', '.join('%s %.1f%%' % (k, 100 * v / tote) for k, v in mix.most_common(18)))
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {s: sum(int(r[hdr.index(s)]) for r in d) for s in stalls}
    print('   stall samples:', {s: v for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
    for r in sorted(d, key=lambda r: -int(r[iS]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
        top = max(stalls, key=lambda s: int(r[hdr.index(s)]))
        print('   %5.1f%%  %-14s %s' % (100 * int(r[iS]) / tot, top, r[iSrc][:100]))


if __name__ == '__main__':
    main()
