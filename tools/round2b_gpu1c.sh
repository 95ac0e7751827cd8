# ncu --set full captures of the team kernel on the three contact configs at their shard sizes; the reports are summarised on the box
# (tools/ncu_summary.py, tools/team_regions.py) and only the text travels back (a report with sources is ~30 MB)
mkdir -p gpurun_out/r2b /tmp/reps
export MPPIB_K2_TEAM=1
for sc in "pick 8192" "heijn 4000" "boxer 4000"; do
  set -- $sc
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:rollout_team -s 2 -c 1 -f -o /tmp/reps/team_$1_$2 python tools/prof_push.py $1 $2 2>&1 | tail -1
  { python tools/ncu_summary.py /tmp/reps/team_$1_$2.ncu-rep; python tools/team_regions.py /tmp/reps/team_$1_$2.ncu-rep; } > gpurun_out/r2b/team_$1_$2.txt 2>&1
done
MPPIB_K2_TEAM=0 timeout 400 ncu --set full --clock-control none -k regex:mppib_rollout_kernel -s 2 -c 1 -f -o /tmp/reps/thread_heijn_4000 python tools/prof_push.py heijn 4000 2>&1 | tail -1
python tools/ncu_summary.py /tmp/reps/thread_heijn_4000.ncu-rep > gpurun_out/r2b/thread_heijn_4000.txt 2>&1
ls -la gpurun_out/r2b
