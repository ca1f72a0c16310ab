set -x
timeout 900 python -m pytest tests -m gpu -q -x -k "update or helloworld or a2c or discrete or cycle or minibatch or reference_loop or checkpoint" 2>&1 | tail -8 > gpurun_out/r02_s6_pytest.log; cat gpurun_out/r02_s6_pytest.log
B200RL_PROFILE=1 timeout 200 python tools/profile_update_phases.py > gpurun_out/r02_update_tc_phases_resident.log 2>&1; cat gpurun_out/r02_update_tc_phases_resident.log
timeout 200 python tools/time_update_schedules.py > gpurun_out/r02_update_schedules_resident.log 2>&1; cat gpurun_out/r02_update_schedules_resident.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_s6_bench.json 2> gpurun_out/r02_s6_bench.err; cat gpurun_out/r02_s6_bench.json | cut -c1-400; tail -3 gpurun_out/r02_s6_bench.err
