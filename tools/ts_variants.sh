#!/bin/bash
# Times every built variant of the TS rollout kernel (B200RL_TS_VARIANT) and runs the rollout parity tests on it.
out=${1:-gpurun_out/ts_variants.log}
: > $out
for v in ${VARIANTS:-0 2}; do   # add a case to the launcher switch in csrc/rollout_ts.cu to build another bit combination
  echo "== variant $v" >> $out
  B200RL_TS_VARIANT=$v python tools/time_rollout.py --modes ts --envs 65536 --launches 12 >> $out 2>&1
  B200RL_TS_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "rollout and (ts or agree)" 2>&1 | tail -1 >> $out
done
