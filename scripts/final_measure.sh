#!/bin/bash
# Round-end measurements on one B200 (run from the repo root under gpurun); outputs under gpurun_out/ with the given tag.
# usage: scripts/final_measure.sh r02z
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $o/${tag}_smoke.txt
t0=$(date +%s); timeout 900 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; echo "bench.py wall: $(( $(date +%s) - t0 )) s" >> $o/${tag}_smoke.txt
t0=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_bench_n1.err; echo "bench.py --impl reference wall: $(( $(date +%s) - t0 )) s" >> $o/${tag}_smoke.txt
# launch list of the bench command (kernel shares of the step); never a bench value
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --config none --no-cpu-baseline --no-e2e --no-parity > $o/${tag}_ncu_launches.log 2>&1
# full captures: one fused C2 pass at full size; the ShortSeq192 pass; the ShortSeqVar kernels
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"pack32|region_scatter|count_regions2" --launch-skip 6 --launch-count 3 \
    -o $o/${tag}_pass_1e9 python scripts/prof_count.py 1e9 1e8 32 > $o/${tag}_ncu_pass.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"pack_fixed|count_parts192" --launch-skip 4 --launch-count 2 \
    -o $o/${tag}_192 python scripts/prof_count.py 2e8 1.25e7 75 > $o/${tag}_ncu_192.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"pack_var2" --launch-skip 3 --launch-count 1 \
    -o $o/${tag}_packvar python scripts/prof_var.py 1e7 > $o/${tag}_ncu_var.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"decode_var2" --launch-skip 3 --launch-count 1 \
    -o $o/${tag}_decvar python scripts/prof_var.py 1e7 >> $o/${tag}_ncu_var.log 2>&1
{ python scripts/percall.py; python scripts/fastq_bench.py 2e7 2e6 32; python scripts/prof_var.py 1e7; python scripts/quickbench.py 5e8 5e7 32; python scripts/quickbench.py 2e8 1.25e7 75; } > $o/${tag}_other.txt 2>&1
tail -2 $o/${tag}_smoke.txt; head -c 300 $o/${tag}_bench_n1.json; echo; ls -la $o | grep ${tag}
