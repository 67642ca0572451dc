"""profiles/traffic.json from a full-size `ncu --set full` capture of one fused pass.
usage: make_traffic.py file.ncu-rep "description of the capture command" """
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, desc = sys.argv[1], sys.argv[2]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    res, seen = {}, set()
    for r in rows[2:]:
        name = r[H.index('Kernel Name')]
        short = 'ssq::' + name.replace('void ', '').split('(')[0].split('<')[0]
        if short in seen:
            continue                       # first captured launch of each kernel
        seen.add(short)
        rd = float(r[H.index('dram__bytes_read.sum')])
        wr = float(r[H.index('dram__bytes_write.sum')])
        unit_r, unit_w = rows[1][H.index('dram__bytes_read.sum')], rows[1][H.index('dram__bytes_write.sum')]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        res[short] = int(rd * scale[unit_r] + wr * scale[unit_w])
    res['pass'] = sum(res.values())
    res['source'] = desc
    json.dump(res, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
