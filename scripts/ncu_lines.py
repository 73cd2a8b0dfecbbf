"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo).
usage: python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + rx],
                     capture_output=True, text=True).stdout
cur, out, tot, ts = None, [], 0, 0
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''):
        continue
    try:
        out.append((cur, int(r[0]), r[1].strip()[:120], int(r[6]), int(r[7])))
        tot += int(r[7]); ts += int(r[6])
    except ValueError:
        pass
print(f'{rx}: {tot} warp instructions, {ts} samples')
for f, l, s, sm, ie in sorted(out, key=lambda x: -x[4])[:top]:
    print(f'{f}:{l:4d} {100 * ie / tot:5.1f}% instr {100 * sm / max(ts, 1):5.1f}% samples | {s}')
