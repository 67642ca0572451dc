#!/bin/bash
# Round-end check without the ncu captures: smoke, GPU tests, the bench line and the reference arm (see final_measure.sh).
# usage: scripts/final_measure_lite.sh r02z
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $o/${tag}_smoke.txt
t0=$(date +%s); timeout 900 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; echo "bench.py wall: $(( $(date +%s) - t0 )) s" >> $o/${tag}_smoke.txt
t0=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_bench_n1.err; echo "bench.py --impl reference wall: $(( $(date +%s) - t0 )) s" >> $o/${tag}_smoke.txt
cat $o/${tag}_smoke.txt | grep -v from_numpy; head -c 300 $o/${tag}_bench_n1.json; echo
