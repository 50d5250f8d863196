set -x
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python scripts/adm_profile.py > gpurun_out/adm_profile_r1o.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1o.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_r1o.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -c 8 -f -o gpurun_out/prof_r1o_conv python scripts/adm_profile.py --reps 1 > gpurun_out/ncu_conv_r1o.log 2>&1
ncu -i gpurun_out/prof_r1o_conv.ncu-rep --page raw --csv > gpurun_out/prof_r1o_conv_raw.csv 2>/dev/null
timeout 400 ncu --set full --import-source on --clock-control none -k regex:gn_apply -c 6 -f -o gpurun_out/prof_r1o_gn python scripts/adm_profile.py --reps 1 > gpurun_out/ncu_gn_r1o.log 2>&1
ncu -i gpurun_out/prof_r1o_gn.ncu-rep --page raw --csv > gpurun_out/prof_r1o_gn_raw.csv 2>/dev/null
timeout 400 ncu --set full --import-source on --clock-control none -k regex:step_ -c 3 -f -o gpurun_out/prof_r1o_step python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_step_r1o.log 2>&1
ncu -i gpurun_out/prof_r1o_step.ncu-rep --page raw --csv > gpurun_out/prof_r1o_step_raw.csv 2>/dev/null
ls -la gpurun_out/*r1o*
