#!/bin/bash
# e2e of the bench line, several fresh processes in a row (is the host<->device path bimodal on this box, and why?)
# usage (under gpurun): bash tools/e2e_repeat.sh TAG N
TAG=${1:-e2e}; N=${2:-6}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_e2e.txt
for i in $(seq 1 $N); do
  python bench.py --steps 20 --warmup 5 --no-cpu --no-kepler --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print(round(d['value']), round(e['value']), round(d['e2e_resident_data']['value']), e.get('warmup'), e.get('pcie_link_after'))" >> gpurun_out/${TAG}_e2e.txt 2>&1
  nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pstate,clocks.sm --format=csv,noheader >> gpurun_out/${TAG}_e2e.txt
done
cat gpurun_out/${TAG}_e2e.txt
