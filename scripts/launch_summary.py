"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_summary.py launches.csv "command" """
import csv
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(open(path, errors="ignore")) if r and not r[0].startswith("==")]
    H = rows[0]
    iN, iM, iV, iU = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
    tot = {}
    for r in rows[1:]:
        if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
            continue
        v = float(r[iV].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)
        e = tot.setdefault(r[iN], [0.0, 0])
        e[0] += v
        e[1] += 1
    total = sum(v[0] for v in tot.values())
    print(f"ncu launch list of `{cmd}` (gpu__time_duration.sum, --clock-control none)")
    print("cold-cache / serialised times: compare SHARES.  Includes the e2e leg (pack_fixed_kernel<0,1> = direct-insert chunks,")
    print("scan kernels of the uint8 lengths) and the synthetic-read generator.\n")
    for name, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:10.3f} ms  {ms/total*100:5.1f}% n={n:4d} {name[:110]}")


if __name__ == "__main__":
    main()
