#!/bin/bash
# One gpurun call that regenerates the round's evidence (profiles/README.md).  Every ncu command runs only after the same
# program has exited 0 without ncu.  Outputs go to gpurun_out/; fold them here with tools/launch_summary.py,
# tools/level_instr.py and tools/ncu_summary.py and copy what is to be kept into profiles/r2/.
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh'
set -x
O=gpurun_out
python bench.py > $O/bench_final.json 2> $O/bench_final.err
B="python bench.py --steps 2 --warmup 3 --no-extras --skip-peaks"
$B > $O/b3.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/launches_final.csv $B > $O/ncu_lf.log 2>&1
B1="python bench.py --steps 1 --warmup 3 --no-extras --skip-peaks"
$B1 > /dev/null 2>&1 && ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/level_instr_f64.csv $B1 > /dev/null 2>&1
$B1 --dtype f32 > /dev/null 2>&1 && ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/level_instr_f32.csv $B1 --dtype f32 > /dev/null 2>&1
python tools/run_c3r.py time > $O/c3r_f64.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3r_f64.csv python tools/run_c3r.py time > /dev/null 2>&1
python tools/run_c3r.py time f32 > $O/c3r_f32.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3r_f32.csv python tools/run_c3r.py time f32 > /dev/null 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:level_live_screen_kernel -s 21 -c 1 -o $O/live_f64_final -f $B1 > $O/ncu_a.log 2>&1
$N -k regex:level_live_screen_kernel -s 21 -c 1 -o $O/live_f32_final -f $B1 --dtype f32 > $O/ncu_b.log 2>&1
$N -k regex:level_live_screen_run_kernel -s 31 -c 1 -o $O/live_run_f64_final -f python tools/run_c3r.py time > $O/ncu_c.log 2>&1
$N -k regex:level_live_screen_run32_kernel -s 31 -c 1 -o $O/live_run_f32_final -f python tools/run_c3r.py time f32 > $O/ncu_d.log 2>&1
$N -k regex:level1_running_kernel -s 2 -c 1 -o $O/level1_run_f32_final -f python tools/run_c3r.py time f32 > $O/ncu_e.log 2>&1
# the five reports are 24 MB each and gpurun brings back at most 64 MiB: fold them here, keep the raw pages as csv and one report
mkdir -p $O/ncu_raw
for k in live_f64 live_f32 live_run_f64 live_run_f32 level1_run_f32; do
  ncu -i $O/${k}_final.ncu-rep --page raw --csv > $O/ncu_raw/${k}.csv 2>/dev/null
done
ncu -i $O/live_f64_final.ncu-rep --page source --csv > $O/ncu_raw/live_f64_source.csv 2>/dev/null
cp profiles/level_kernel_summary.json $O/level_kernel_summary.before.json
python tools/ncu_summary.py live_f64=$O/live_f64_final.ncu-rep live_f32=$O/live_f32_final.ncu-rep \
  live_run_f64=$O/live_run_f64_final.ncu-rep live_run_f32=$O/live_run_f32_final.ncu-rep \
  level1_run_f32=$O/level1_run_f32_final.ncu-rep:5000050000 > $O/ncu_summary.log 2>&1
cp profiles/level_kernel_summary.json $O/level_kernel_summary.json
rm -f $O/live_f32_final.ncu-rep $O/live_run_f64_final.ncu-rep $O/live_run_f32_final.ncu-rep $O/level1_run_f32_final.ncu-rep
cat $O/c3r_f64.log $O/c3r_f32.log $O/ncu_summary.log
