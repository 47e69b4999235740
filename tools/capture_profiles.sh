#!/bin/bash
# Round-2 ncu evidence in one call on the GPU box (each ncu pass only after the same command exited 0 without ncu):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/capture_profiles.sh'
# then here: RC_PROFILE_TAG=r02 python tools/profile_summary.py gpurun_out/r02g_propagate_raw.csv
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --steps 2 --warmup 1 > $O/r02g_bench_plain.json 2> $O/r02g_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02g_launches.csv \
    python bench.py --steps 2 --warmup 1 > $O/r02g_ncu1.log 2>&1
python tools/one_eval.py c3 44 > $O/r02g_plain.log 2>&1 || exit 1
if [ -z "$RC_CAPTURE_SKIP_PROPAGATE" ]; then
ncu --set full --clock-control none -k regex:rc_propagate -s 80 -c 40 -o $O/r02g_propagate -f python tools/one_eval.py c3 44 > $O/r02g_ncu2.log 2>&1
ncu -i $O/r02g_propagate.ncu-rep --page raw --csv > $O/r02g_propagate_raw.csv 2>/dev/null
rm -f $O/r02g_propagate.ncu-rep
fi
ncu --set full --clock-control none -k 'regex:rc_traj|rc_cum' -s 18 -c 9 -o $O/r02g_traj -f python tools/one_eval.py c3 44 > $O/r02g_ncu3.log 2>&1
ncu -i $O/r02g_traj.ncu-rep --page raw --csv > $O/r02g_traj_raw.csv 2>/dev/null
rm -f $O/r02g_traj.ncu-rep
ls -la $O/r02g_*
