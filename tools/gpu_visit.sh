#!/bin/bash
# One GPU-box visit (round 2).  usage (under gpurun): bash tools/gpu_visit.sh TAG MODE
#   MODE = check   : full -m gpu test suite, smoke, bench (ours + reference arm twice), launch list, transit ncu capture
#   MODE = gp      : celerite tests + bench extra configs + ncu capture of the celerite kernels
#   MODE = transit : fused tests + bench (no extras) + ncu capture of the logp kernel
#   MODE = final   : everything, captures of every kernel, sanitizers
TAG=${1:-r02}
MODE=${2:-check}
OUT=gpurun_out
mkdir -p $OUT
NCU_FULL="ncu --set full --metrics smsp__inst_executed_pipe_fp64.sum --clock-control none --import-source on"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt
(echo -n "cgroup cpu.max: "; cat /sys/fs/cgroup/cpu.max 2>/dev/null || cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>/dev/null; grep Cpus_allowed_list /proc/self/status; uptime) >> $OUT/${TAG}_gpu.txt 2>&1

run_pytest() { # name, args...
  local name=$1; shift
  timeout 2400 python -m pytest "$@" -m gpu -q > $OUT/${TAG}_pytest_$name.log 2>&1; echo "pytest($name) exit $?" | tee -a $OUT/${TAG}_pytest_$name.log
}
bench_line() { # name, args...
  local name=$1; shift
  timeout 1200 python bench.py "$@" > $OUT/${TAG}_bench_$name.json 2> $OUT/${TAG}_bench_$name.err; echo "bench($name) exit $?" >> $OUT/${TAG}_bench_$name.err
}
capture() { # name, kernel regex, skip, bench args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 $NCU_FULL -k regex:"$rx" -s $skip -c 1 -f -o $OUT/${TAG}_prof_$name python bench.py "$@" > $OUT/${TAG}_ncu_$name.log 2>&1
}

case $MODE in
check)
  run_pytest all tests
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
  bench_line ours --steps 20 --warmup 5
  bench_line ref1 --impl reference --steps 20 --warmup 5
  bench_line ref2 --impl reference --steps 20 --warmup 5
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"transit_|model_|kepler_fwd|fold_|gp_" -c 120 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1
  capture transit transit_logp_grad_kernel 3 --steps 2 --warmup 3 --no-cpu --no-kepler --no-extra
  ;;
gp)
  run_pytest gp tests/test_gpu_gp.py tests/test_gpu_configs.py
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
  bench_line ours --steps 20 --warmup 5 --no-cpu --no-kepler --extra gp
  for V in $GP_VARIANTS; do XO_B200_LIB=exoplanet_b200/libxo_b200_$V.so bench_line $V --steps 20 --warmup 5 --no-cpu --no-kepler --extra gp; done
  # B = 16 launches: the bench runs 13 launches at B = 1 first
  # launch order of bench --extra gp: 13 x forward<0> (B=1), 13 x (forward<2>, backward) (B=1), then the same at B=16
  capture gpfwd "gp_forward_kernel" 27 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  capture gpbwd "gp_backward_kernel" 14 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  ;;
transit)
  run_pytest fused tests/test_gpu_fused.py tests/test_reference_glue.py tests/test_gpu_configs.py tests/test_gpu_sampling.py
  bench_line ours --steps 20 --warmup 5 --no-cpu --no-kepler --extra c4,hmc
  for V in $TRANSIT_VARIANTS; do XO_B200_LIB=exoplanet_b200/libxo_b200_$V.so bench_line $V --steps 20 --warmup 5 --no-cpu --no-kepler --no-extra; done
  capture transit transit_logp_grad_kernel 3 --steps 2 --warmup 3 --no-cpu --no-kepler --no-extra
  ;;
both)
  run_pytest all tests/test_gpu_fused.py tests/test_reference_glue.py tests/test_gpu_configs.py tests/test_gpu_sampling.py tests/test_gpu_gp.py
  bench_line ours --steps 20 --warmup 5 --no-cpu --no-kepler
  for R in 16 32; do XO_GP_RUN=$R bench_line gprun$R --steps 20 --warmup 5 --no-cpu --no-kepler --extra gp; done
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hmc_|posterior_|transit_|model_|gp_" -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-kepler --extra gp,hmc > $OUT/${TAG}_ncu_bench.log 2>&1
  capture transit transit_logp_grad_kernel 3 --steps 2 --warmup 3 --no-cpu --no-kepler --no-extra
  capture gpagg "gp_aggregate_kernel" 27 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  capture gpll "gp_loglike_kernel" 27 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  capture gpgagg "gp_grad_aggregate_kernel" 14 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  capture gpgfin "gp_grad_final_kernel" 14 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  ;;
scale)
  # usage: gpurun --gpus N -- 'NG=N bash tools/gpu_visit.sh TAG scale'
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517"
  timeout 900 $TR bench.py --gpus $NG --steps 20 --warmup 5 > $OUT/${TAG}_bench_${NG}gpu.json 2> $OUT/${TAG}_bench_${NG}gpu.err; echo "bench($NG gpu) exit $?" >> $OUT/${TAG}_bench_${NG}gpu.err
  timeout 900 $TR bench.py --gpus $NG --steps 20 --warmup 5 --same-seed --no-kepler > $OUT/${TAG}_bench_${NG}gpu_sameseed.json 2> $OUT/${TAG}_bench_${NG}gpu_sameseed.err; echo "bench($NG gpu, same seed) exit $?" >> $OUT/${TAG}_bench_${NG}gpu_sameseed.err
  [ "$NG" = "2" ] && timeout 600 $TR bench.py --impl reference --gpus $NG --steps 20 --warmup 5 > $OUT/${TAG}_bench_${NG}gpu_ref.json 2> $OUT/${TAG}_bench_${NG}gpu_ref.err; echo "bench($NG gpu, reference arm) exit $?" >> $OUT/${TAG}_bench_${NG}gpu_ref.err
  for f in $OUT/${TAG}_bench_${NG}gpu*.json; do echo "== $f"; head -c 1500 $f; echo; done
  tail -3 $OUT/${TAG}_bench_${NG}gpu*.err
  exit 0
  ;;
sweep)
  run_pytest host tests/test_gpu_fused.py tests/test_gpu_ops.py tests/test_gpu_configs.py tests/test_reference_glue.py tests/test_pytensor_ops.py
  bench_line ours --steps 20 --warmup 5 --no-cpu --no-kepler --extra gp
  : > $OUT/${TAG}_sweep.txt
  one() { echo -n "$1: " >> $OUT/${TAG}_sweep.txt; env $1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-kepler --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],5), round(d['e2e']['value']), round(d['e2e_resident_data']['value']))" >> $OUT/${TAG}_sweep.txt; }
  one XO_NOOP=1; one XO_NOOP=2
  for ML in 1 2; do one XO_MODEL_LAUNCHES=$ML; done
  one XO_MODEL_HEAVY_FIRST=1; one XO_MODEL_PDL=0
  for W in 12 16 20 28 32 40 48; do one XO_TRANSIT_WARPS_PER_SM=$W; done
  for MB in 2 4; do one XO_TRANSIT_MINBLOCKS=$MB; done
  cat $OUT/${TAG}_sweep.txt
  ;;
ab)
  # usage: AB_ENVS="XO_A=1 XO_B=2,XO_C=3" (space separated variants; commas join several assignments of one variant)
  run_pytest fused tests/test_gpu_fused.py tests/test_reference_glue.py tests/test_gpu_configs.py tests/test_gpu_sampling.py
  bench_line ours --steps 20 --warmup 5 --no-cpu --no-kepler --extra c4,hmc
  : > $OUT/${TAG}_sweep.txt
  one() { echo -n "$1: " >> $OUT/${TAG}_sweep.txt; env ${1//,/ } timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-kepler --extra c4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],5), round(d['e2e']['value']), round(d['e2e_resident_data']['value']), [(k, v.get('value')) for k, v in d.get('other_configs', {}).items() if isinstance(v, dict)])" >> $OUT/${TAG}_sweep.txt 2>&1; }
  one XO_NOOP=1
  for E in $AB_ENVS; do one $E; done
  cat $OUT/${TAG}_sweep.txt
  [ -n "$AB_CAPTURE" ] && capture transit transit_logp_grad_kernel 3 --steps 2 --warmup 3 --no-cpu --no-kepler --no-extra
  ;;
final)
  run_pytest all tests
  for ML in 1 2; do XO_MODEL_LAUNCHES=$ML run_pytest launches$ML tests/test_gpu_fused.py; done
  XO_MODEL_PDL=0 run_pytest nopdl tests/test_gpu_fused.py
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
  bench_line ours --steps 20 --warmup 5
  bench_line ref1 --impl reference --steps 20 --warmup 5
  bench_line ref2 --impl reference --steps 20 --warmup 5
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"transit_|model_|kepler_fwd|fold_|gp_|rv_|hmc_|posterior_" -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu --extra c2,gp,hmc > $OUT/${TAG}_ncu_bench.log 2>&1
  capture transit transit_logp_grad_kernel 3 --steps 2 --warmup 3 --no-cpu --no-kepler --no-extra
  capture kepler kepler_fwd_kernel 4 --steps 1 --warmup 3 --no-cpu --no-extra
  capture rv rv_logp_grad_kernel 4 --steps 1 --warmup 3 --no-cpu --no-kepler --extra c2
  capture gpagg "gp_aggregate_kernel" 27 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  capture gpll "gp_loglike_kernel" 27 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  capture gpgagg "gp_grad_aggregate_kernel" 14 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  capture gpgfin "gp_grad_final_kernel" 14 --steps 1 --warmup 3 --no-cpu --no-kepler --extra gp
  timeout 600 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck exit $?" >> $OUT/${TAG}_memcheck.log
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck exit $?" >> $OUT/${TAG}_racecheck.log
  ;;
esac
tail -4 $OUT/${TAG}_pytest_*.log 2>/dev/null; tail -2 $OUT/${TAG}_smoke.log 2>/dev/null
for f in $OUT/${TAG}_bench_*.json; do echo "== $f"; head -c 3000 $f; echo; done
tail -3 $OUT/${TAG}_bench_*.err 2>/dev/null
exit 0
