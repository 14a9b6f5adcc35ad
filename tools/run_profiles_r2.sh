# ncu captures of round 2 (run on a B200 through gpurun; every ncu command follows the same command run plain, exit 0)
set -x
python tools/prof_case.py 4096 0 0 0 > gpurun_out/r2_prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_replay_flights -s 1 -c 1 -o gpurun_out/prof_r2_c3 python tools/prof_case.py 4096 0 0 0 > gpurun_out/r2_ncu_c3.log 2>&1
python tools/prof_case.py 4096 0 0 1 > gpurun_out/r2_prof_plain_fan.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_replay_flights -s 1 -c 1 -o gpurun_out/prof_r2_c3_fan python tools/prof_case.py 4096 0 0 1 > gpurun_out/r2_ncu_c3_fan.log 2>&1
python tools/prof_case.py 1184 0 0 0 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ray_setup -s 1 -c 1 -o gpurun_out/prof_r2_setup python tools/prof_case.py 1184 0 0 0 > gpurun_out/r2_ncu_setup.log 2>&1
python tools/c5_shard_engines.py > gpurun_out/r2_c5_shard_engines_after.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4; tail -2 gpurun_out/r2_c5_shard_engines_after.log
# the sub-tile engine at full size: config 2 (time-sliced map tiles) and config 4 (value tiles)
python tools/prof_case_c4.py c2 360000 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_replay_tiles -s 1 -c 1 -o gpurun_out/prof_r2_tiles_c2 python tools/prof_case_c4.py c2 360000 > gpurun_out/r2_ncu_tiles_c2.log 2>&1
python tools/prof_case_c4.py c4 1048576 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_replay_tiles -s 1 -c 1 -o gpurun_out/prof_r2_tiles_c4 python tools/prof_case_c4.py c4 1048576 > gpurun_out/r2_ncu_tiles_c4.log 2>&1
