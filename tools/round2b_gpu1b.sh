mkdir -p gpurun_out/r2b
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool racecheck --print-limit 20 --error-exitcode 3 python tools/sanitize_run.py > gpurun_out/r2b/sanitize_racecheck.log 2>&1; echo "racecheck: exit $? ; $(grep -E 'RACECHECK SUMMARY' gpurun_out/r2b/sanitize_racecheck.log)"
timeout 600 python tools/team_time.py heijn 4000 boxer 4000 pick 8192 heijn 16000 pick 65536 raw:albert 10000 raw:omnipanda 10000 raw:panda_gripper 10000 raw:jackal 10000 raw:boxer 10000 2>&1 | tee gpurun_out/r2b/team_ab.txt
export MPPIB_K2_TEAM=1
for sc in "pick 8192" "heijn 4000" "boxer 4000"; do set -- $sc; timeout 400 ncu --set full --import-source on --clock-control none -k regex:rollout_team -s 2 -c 1 -f -o gpurun_out/r2b/team_$1_$2 python tools/prof_push.py $1 $2 2>&1 | tail -1; done
