set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --steps 2 --warmup 3 --no-extras --skip-peaks > gpurun_out/b3.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras --skip-peaks > gpurun_out/ncu_lf.log 2>&1
python tools/run_c3r.py time > gpurun_out/c3r_f64.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3r_f64.csv python tools/run_c3r.py time > /dev/null 2>&1
python tools/run_c3r.py time f32 > gpurun_out/c3r_f32.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3r_f32.csv python tools/run_c3r.py time f32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:level_tile_screen_kernel -s 20 -c 1 -o gpurun_out/scr_f64_final -f python bench.py --steps 1 --warmup 3 --no-extras --skip-peaks > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:level_tile_screen_kernel -s 20 -c 1 -o gpurun_out/scr_f32_final -f python bench.py --steps 1 --warmup 3 --no-extras --skip-peaks --dtype f32 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:level_tile_screen_run_kernel -s 30 -c 1 -o gpurun_out/run_f64_final -f python tools/run_c3r.py time > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:level_tile_screen_run32_kernel -s 30 -c 1 -o gpurun_out/run_f32_final -f python tools/run_c3r.py time f32 > gpurun_out/ncu_d.log 2>&1
cat gpurun_out/c3r_f64.log gpurun_out/c3r_f32.log
