#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full captures.
# usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
lscpu | head -20 >> $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?" >> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transit_logp_grad_kernel -s 3 -c 1 -f -o $OUT/${TAG}_prof_transit \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-kepler > $OUT/${TAG}_ncu_transit.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kepler_fwd_kernel -s 4 -c 1 -f -o $OUT/${TAG}_prof_kepler \
    python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_kepler.log 2>&1
tail -5 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_smoke.log | tail -3; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
