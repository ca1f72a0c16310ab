set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_s5_pytest.log; cat gpurun_out/r02_s5_pytest.log
timeout 300 python tools/time_forward.py > gpurun_out/r02_forward_tc_timing.log 2>&1; cat gpurun_out/r02_forward_tc_timing.log
(timeout 120 python tools/time_e2e_gaps.py; timeout 120 python tools/time_e2e_gaps.py side) > gpurun_out/r02_e2e_gaps.log 2>&1; cat gpurun_out/r02_e2e_gaps.log
(for v in 2 18; do echo "== variant $v"; B200RL_TS_VARIANT=$v timeout 120 python tools/time_rollout.py --modes ts --envs 65536,8192 --launches 12; done) > gpurun_out/r02_ts_no_critic.log 2>&1; cat gpurun_out/r02_ts_no_critic.log
timeout 300 python bench.py > gpurun_out/r02_s5_bench.json 2> gpurun_out/r02_s5_bench.err; cat gpurun_out/r02_s5_bench.json; tail -3 gpurun_out/r02_s5_bench.err
