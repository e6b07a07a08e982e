#!/bin/bash
# First GPU call of the next round: one pass over the experiments DESIGN.md section 5 lists, ~2 min of box time.
#   tools/gpurun_retry.sh gpurun_out/r2_exp.log 600 'bash tools/r2_experiments.sh'
# Each line: configuration, event-timed step (median/min us) and where the GEMM chain starts/ends.
run() {  # label, env...
  local label=$1; shift
  for v in lrt bbb; do
    printf '%-34s %s ' "$label" "$v"
    env "$@" timeout 100 python tools/timeline.py $v 2>&1 | grep "event-timed\|first kernel" | \
      sed -e 's/BBBAlexNet [a-z]* B=512: event-timed replay //' -e 's/; [0-9]* instrumented.*//' -e 's/first kernel start -> //' | tr '\n' ' '
    echo
  done
}
echo "== correctness of the experimental on-chip x^2 variant (must pass before its timing means anything)"
BBB_B200_SQ_ONCHIP=1 timeout 200 python -m pytest tests -m gpu -q -x -k "fused or smoke" 2>&1 | tail -2
echo "== timelines"
run "default"
run "UNITS=1 (4 x 1-block stages)" BBB_B200_UNITS=1
run "SQ_ONCHIP=1" BBB_B200_SQ_ONCHIP=1
run "SQ_ONCHIP=1 UNITS=1" BBB_B200_SQ_ONCHIP=1 BBB_B200_UNITS=1
run "PDL=0" BBB_B200_PDL=0
run "PREP_CHAINS=2" BBB_B200_PREP_CHAINS=2
run "PREP_CARVEOUT=0" BBB_B200_PREP_CARVEOUT=0
echo "== per-step trace of the tap-GEMMs (default, then SQ)"
timeout 100 python tools/trace_tapgemm.py lrt 2>&1 | grep "^layer\|mma_full" | cut -c1-240
BBB_B200_SQ_ONCHIP=1 timeout 100 python tools/trace_tapgemm.py lrt 2>&1 | grep "^layer\|mma_full" | cut -c1-240
echo "== forwards in flight on S streams (bench.py --streams): headline vs S-stream figure"
for S in 2 3; do
  timeout 200 python bench.py --streams $S --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('S=$S headline', round(d['ms_per_step']*1e3,1), 'us', round(d['value']), ' streams', d['streams'] and (round(d['streams']['ms_per_step']*1e3,1), round(d['streams']['value'])))"
done
