#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, tuning sweep, ncu launch list + full captures.
# usage (under gpurun): bash tools/gpu_round.sh [tag] [quick]
TAG=${1:-r01}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
if [ "$MODE" = "tune" ]; then
# short visit while tuning the logp kernel: fused parity tests, launch-width sweep, one ncu capture
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_reference_glue.py -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
: > $OUT/${TAG}_sweep.txt
for mb in 3 4; do for w in 8 12 16 24 32 48; do
  echo -n "minblocks=$mb warps_per_sm=$w " >> $OUT/${TAG}_sweep.txt
  XO_TRANSIT_MINBLOCKS=$mb XO_TRANSIT_WARPS_PER_SM=$w timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-kepler --no-extra 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])" >> $OUT/${TAG}_sweep.txt
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transit_logp_grad_kernel -s 3 -c 1 -f -o $OUT/${TAG}_prof_transit \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-kepler --no-extra > $OUT/${TAG}_ncu_transit.log 2>&1
tail -5 $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_sweep.txt
exit 0
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
for ML in 1 2; do XO_MODEL_LAUNCHES=$ML timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q > $OUT/${TAG}_pytest_launches$ML.log 2>&1; echo "pytest(XO_MODEL_LAUNCHES=$ML) exit $?" >> $OUT/${TAG}_pytest.log; done
XO_MODEL_PDL=0 timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q > $OUT/${TAG}_pytest_nopdl.log 2>&1; echo "pytest(XO_MODEL_PDL=0) exit $?" >> $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?" >> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
: > $OUT/${TAG}_sweep.txt; for ML in 1 2 3; do echo -n "XO_MODEL_LAUNCHES=$ML: " >> $OUT/${TAG}_sweep.txt; XO_MODEL_LAUNCHES=$ML timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-kepler --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_resident_data']['value'])" >> $OUT/${TAG}_sweep.txt; done
for HF in 0 1; do echo -n "XO_MODEL_HEAVY_FIRST=$HF: " >> $OUT/${TAG}_sweep.txt; XO_MODEL_HEAVY_FIRST=$HF timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-kepler --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_resident_data']['value'])" >> $OUT/${TAG}_sweep.txt; done
if [ "$MODE" = "full" ]; then
# tuning sweep of the logp kernel (register cap x launch width)
for SWEEPFLAG in "" "--full-sweep"; do for mb in 2 3 4 5; do for w in 16 24 32; do
  echo -n "flag=$SWEEPFLAG minblocks=$mb warps_per_sm=$w " >> $OUT/${TAG}_sweep.txt
  XO_TRANSIT_MINBLOCKS=$mb XO_TRANSIT_WARPS_PER_SM=$w timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-kepler --no-extra $SWEEPFLAG 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])" >> $OUT/${TAG}_sweep.txt
done; done; done
for mb in 2 3 4; do echo -n "kepler minblocks=$mb " >> $OUT/${TAG}_sweep.txt; XO_KEPLER_MINBLOCKS=$mb timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kepler'])" >> $OUT/${TAG}_sweep.txt; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"transit_|model_|kepler_fwd|fold_" -c 80 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:transit_logp_grad_kernel -s 3 -c 1 -f -o $OUT/${TAG}_prof_transit \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-kepler > $OUT/${TAG}_ncu_transit.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kepler_fwd_kernel -s 4 -c 1 -f -o $OUT/${TAG}_prof_kepler \
    python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_kepler.log 2>&1
# the other kernels of the section-8 rows: RV likelihood (configs[2]) and the celerite scans (configs[3])
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rv_logp_grad_kernel -s 4 -c 1 -f -o $OUT/${TAG}_prof_rv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-kepler > $OUT/${TAG}_ncu_rv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gp_loglike_kernel|gp_grad_final_kernel" -s 6 -c 1 -f -o $OUT/${TAG}_prof_gp \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-kepler > $OUT/${TAG}_ncu_gp.log 2>&1
# sanitizers on the small smoke problem (memory errors, shared-memory races in the queue kernels)
timeout 600 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/${TAG}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/${TAG}_racecheck.log
fi
tail -3 $OUT/${TAG}_memcheck.log $OUT/${TAG}_racecheck.log 2>/dev/null
tail -8 $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_smoke.log; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_sweep.txt
