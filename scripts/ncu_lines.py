"""Per-source-line instruction and stall shares of one kernel from an .ncu-rep captured with --import-source on.
usage: ncu_lines.py file.ncu-rep kernel_regex [min_pct]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, H, lines, seen_funcs = None, None, {}, 0
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur_file = r[1].split('/')[-1]
            continue
        if len(r) >= 2 and r[0] == 'Function Name':
            continue
        if r and r[0] == 'Line No':
            H = r
            iL, iS = 0, 1
            iI = H.index('Instructions Executed')
            iW = H.index('Warp Stall Sampling (All Samples)')
            continue
        if H is None or len(r) < len(H):
            continue
        x = len(r) - len(H)          # source text with unescaped quotes/commas splits into extra fields
        vi, vw = r[iI + x], r[iW + x]
        if r[iL].isdigit() and vi.replace('.', '').isdigit():         # a CUDA source line row (aggregated over its SASS)
            key = (cur_file, int(r[iL]))
            e = lines.setdefault(key, [','.join(r[iS:iS + 1 + x]).strip(), 0, 0])
            e[1] += int(float(vi))
            e[2] += int(float(vw)) if vw.replace('.', '').isdigit() else 0
    tot = sum(v[1] for v in lines.values()) or 1
    tw = sum(v[2] for v in lines.values()) or 1
    print(f'instructions {tot}  stall samples {tw}')
    for (f, ln), v in sorted(lines.items()):
        if v[1] > tot * min_pct / 100 or v[2] > tw * min_pct / 100:
            print(f'{f:18s}:{ln:4d} exec {v[1]/tot*100:5.1f}%  stall {v[2]/tw*100:5.1f}%  {v[0][:110]}')


if __name__ == '__main__':
    main()
