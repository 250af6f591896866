"""Summarise an ncu report: python profiles/ncu_summary.py <file.ncu-rep> [kernel-regex] [--traffic-json "<how it was captured>"]
Prints the headline metrics per captured launch and the top stall locations (needs -lineinfo).  With --traffic-json the DRAM
bytes of the first stress_tma / particle_tma launch go to profiles/ncu_traffic.json together with the hash of the kernel sources
of this tree (bench.py reports them as roofline.traffic and says whether they still belong to the current kernels)."""
import csv, io, json, os, subprocess, sys
traffic_note = None
if '--traffic-json' in sys.argv:
    i = sys.argv.index('--traffic-json')
    traffic_note = sys.argv[i + 1]
    del sys.argv[i:i + 2]
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr:
            print('   %-62s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
    st = sorted(((float(r[i]), h) for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and r[i]), reverse=True)
    for v, h in st[:6]:
        print('   stall %-40s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
if traffic_note is not None:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out = {'source': traffic_note, 'kernel_sha16': bench.kernel_sha16()}
    for key in ('stress_tma', 'particle_tma'):
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            if key + '<' in name or name.startswith(key):
                def val(metric):        # ncu prints Mbyte / Gbyte / usecond ...: normalise to bytes and microseconds
                    v, u = float(r[hdr.index(metric)].replace(',', '')), units[hdr.index(metric)]
                    return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[u]
                rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
                out[key] = {'kernel': name, 'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_launch': rd + wr,
                            'duration_us': val('gpu__time_duration.sum')}
                break
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ncu_traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote profiles/ncu_traffic.json:', json.dumps(out)[:300])
pat = sys.argv[2] if len(sys.argv) > 2 else None
if pat:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    iS, iSrc, iA, iE = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Address'), hdr.index('Instructions Executed')
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) > iE and r[iA].startswith('0x') and r[iA] not in seen:
            seen.add(r[iA]); data.append(r)
    tot = sum(int(r[iS]) for r in data)
    texec = sum(int(r[iE]) for r in data)
    print('-- source page: %d samples, %d warp-instructions, %d SASS lines' % (tot, texec, len(data)))
    base = int(data[0][iA], 16)
    for name in ['stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_short_sb', 'stall_mio', 'stall_lg', 'stall_math', 'stall_not_selected', 'stall_selected', 'stall_branch_resolving', 'stall_no_inst', 'stall_dispatch', 'stall_sleep', 'stall_membar', 'stall_drain', 'stall_tex', 'stall_misc']:
        if name in hdr:
            i = hdr.index(name)
            print('   %-24s %6d (%.1f%%)' % (name, sum(int(r[i] or 0) for r in data), 100.0 * sum(int(r[i] or 0) for r in data) / max(tot, 1)))
    for r in sorted(data, key=lambda r: -int(r[iS]))[:28]:
        print('   %5d %4.1f%% +0x%04x exec=%8s %s' % (int(r[iS]), 100 * int(r[iS]) / tot, int(r[iA], 16) - base, r[iE], r[iSrc].strip()[:84]))
