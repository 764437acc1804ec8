#!/bin/bash
# tests/probes/run_mma_energy.sh [seconds]: every variant of build/mma_energy next to an nvidia-smi power / clock log
# -> gpurun_out/mma_energy.txt  (median power and SM clock of the second half of each run)
cd "$(dirname "$0")/../.."
S=${1:-3}
mkdir -p gpurun_out
: > gpurun_out/mma_energy.txt
for v in n128 n256 n128s n256s cg2; do
  nvidia-smi --query-gpu=power.draw,clocks.sm --format=csv,noheader,nounits -lms 100 -i 0 > gpurun_out/mma_energy_$v.smi &
  SMI=$!
  timeout 60 build/mma_energy $v $S >> gpurun_out/mma_energy.txt 2>&1
  kill $SMI 2>/dev/null; wait $SMI 2>/dev/null
  python3 - "$v" <<'PY' >> gpurun_out/mma_energy.txt
import statistics, sys
v = sys.argv[1]
rows = [l.split(",") for l in open(f"gpurun_out/mma_energy_{v}.smi") if l.count(",") == 1]
rows = rows[len(rows) // 2:]
if rows:
    print(f'  {v}: median power {statistics.median(float(r[0]) for r in rows):.0f} W, median SM clock {statistics.median(float(r[1]) for r in rows):.0f} MHz, {len(rows)} samples')
PY
done
cat gpurun_out/mma_energy.txt
