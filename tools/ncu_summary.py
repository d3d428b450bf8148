"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/."""
import collections, csv, subprocess, sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']

def main(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')]
        print('== %s' % name)
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('   %-70s %s %s' % (w, vals[i], units[i]))
        st = [(h, vals[i]) for i, h in enumerate(hdr) if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
        st = [(h, float(v)) for h, v in st if v not in ('', 'n/a')]
        st.sort(key=lambda x: -x[1])
        for h, v in st[:7]:
            print('   stall %-62s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    seen = set()
    for b in blocks:
        if b['name'] in seen or len(b['rows']) < 2:
            continue
        seen.add(b['name'])
        h = b['rows'][0]
        ix = {x: i for i, x in enumerate(h)}
        data = b['rows'][1:]
        tot = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
        agg = collections.Counter()
        for r in data:
            s = r[ix['Source']]
            op = s.split()[1] if s.startswith('@') else s.split()[0]
            agg[op.split('.')[0]] += int(r[ix['Instructions Executed']] or 0)
        print('-- SASS mix of %s: %d instructions, %d warp-instructions executed' % (b['name'], len(data), tot))
        print('   ' + ', '.join('%s %.1f%%' % (k, 100.0 * v / max(tot, 1)) for k, v in agg.most_common(12)))

if __name__ == '__main__':
    main(sys.argv[1])
