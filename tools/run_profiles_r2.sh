set -x
python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k c4 2>&1 | tail -5 > gpurun_out/r2_t5.log
python tools/c5_shard_engines.py > gpurun_out/r2_c5_shard_engines.log 2>&1
python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline --no-e2e > gpurun_out/r2_b5_plain.json 2> gpurun_out/r2_b5_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_launches.log 2>&1
python tools/prof_case.py 4096 0 0 0 > gpurun_out/r2_prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_replay_flights -s 1 -c 1 -o gpurun_out/prof_r2_c3 python tools/prof_case.py 4096 0 0 0 > gpurun_out/r2_ncu_c3.log 2>&1
python tools/prof_case.py 4096 0 0 1 > gpurun_out/r2_prof_plain_fan.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_replay_flights -s 1 -c 1 -o gpurun_out/prof_r2_c3_fan python tools/prof_case.py 4096 0 0 1 > gpurun_out/r2_ncu_c3_fan.log 2>&1
python tools/prof_case.py 1184 0 0 0 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ray_setup -s 1 -c 1 -o gpurun_out/prof_r2_setup python tools/prof_case.py 1184 0 0 0 > gpurun_out/r2_ncu_setup.log 2>&1
tail -3 gpurun_out/r2_t5.log; tail -4 gpurun_out/r2_c5_shard_engines.log; ls -la gpurun_out/*.ncu-rep | tail -4
