"""Hot SASS instructions of one kernel from an .ncu-rep (source page).  usage: ncu_hot.py file.ncu-rep kernel_regex [min_pct]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
    H = rows[hi]
    iS, iI, iW, iT = H.index('Source'), H.index('Instructions Executed'), H.index('Warp Stall Sampling (All Samples)'), H.index('Avg. Threads Executed')
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= iT or not r[iI].isdigit():
            if len(r) > 1 and r[0] == 'Kernel Name':
                break            # next launch of the same kernel
            continue
        data.append((r[iS].strip(), int(r[iI]), int(r[iW] or 0), r[iT]))
    tot = sum(d[1] for d in data) or 1
    tw = sum(d[2] for d in data) or 1
    print(len(data), 'SASS instructions; warp-instructions executed', tot, '; stall samples', tw)
    for i, d in enumerate(data):
        if d[1] > tot * min_pct / 100 or d[2] > tw * min_pct / 100:
            print(f"{i:5d} exec {d[1]/tot*100:5.2f}%  stall {d[2]/tw*100:5.2f}%  thr={d[3]:>5s}  {d[0][:100]}")


if __name__ == '__main__':
    main()
