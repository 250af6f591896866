"""Per-CUDA-source-line instruction and stall-sample totals of one kernel of an ncu report (needs -lineinfo and
--import-source on):  python profiles/ncu_lines.py <file.ncu-rep> <kernel-regex> [top]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat, '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, fname, hdr, first = [], None, None, True
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r[0] == 'Function Name':
        if not first and 'tma' not in r[1]:
            pass
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0].isdigit():
        iS, iE = hdr.index('# Samples'), hdr.index('Instructions Executed')
        lines.append((fname, int(r[0]), r[1].strip(), int(r[iS]) if r[iS].isdigit() else 0, int(r[iE]) if r[iE].isdigit() else 0))
seen = {}
for f, n, src, s, e in lines:          # the same launch may be listed once per captured instance: keep the first
    seen.setdefault((f, n), (src, s, e))
tot_s = sum(v[1] for v in seen.values()); tot_e = sum(v[2] for v in seen.values())
print('total samples %d, warp instructions %d' % (tot_s, tot_e))
for (f, n), (src, s, e) in sorted(seen.items(), key=lambda kv: -kv[1][2])[:top]:
    print('%6.2f%% inst %5.2f%% smp  %s:%d  %s' % (100.0 * e / max(tot_e, 1), 100.0 * s / max(tot_s, 1), f, n, src[:110]))
