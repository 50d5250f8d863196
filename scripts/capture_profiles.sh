# Evidence for profiles/: run on the GPU box (gpurun -- 'bash scripts/capture_profiles.sh <tag>'), then condense here
# with scripts/ncu_summary.py.  Numbers printed under ncu are never bench values.
tag=${1:-r1s}
set -x
mkdir -p gpurun_out
# 1. launch list of the bench command (only the timed region: --profile-range brackets it with cudaProfilerStart/Stop)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1600 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range \
  > gpurun_out/ncu_bench_$tag.log 2>&1
# 2. the dominant kernel: 3x3 convolution 256 -> 256 at 16 x 256 x 256 with the fused GroupNorm + SiLU input transform
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 1 -c 1 -f \
  -o gpurun_out/prof_${tag}_conv_fused python scripts/conv_one.py --hw 256 --ci 256 --co 256 --fused 1 > gpurun_out/ncu_conv_$tag.log 2>&1
ncu -i gpurun_out/prof_${tag}_conv_fused.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_conv_fused_raw.csv 2>/dev/null
# 3. the same layer with the fused 1x1 skip operand (K = 9 x 256 + 512)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 1 -c 1 -f \
  -o gpurun_out/prof_${tag}_conv_skip python scripts/conv_one.py --hw 256 --ci 256 --co 256 --fused 1 --skip 512 > gpurun_out/ncu_convskip_$tag.log 2>&1
ncu -i gpurun_out/prof_${tag}_conv_skip.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_conv_skip_raw.csv 2>/dev/null
# 4. remaining GroupNorm passes (resampling blocks) and the transition kernel
timeout 400 ncu --set full --import-source on --clock-control none -k regex:gn_apply -c 4 -f -o gpurun_out/prof_${tag}_gn \
  python scripts/adm_profile.py --reps 1 > gpurun_out/ncu_gn_$tag.log 2>&1
ncu -i gpurun_out/prof_${tag}_gn.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_gn_raw.csv 2>/dev/null
timeout 400 ncu --set full --import-source on --clock-control none -k regex:step_ -c 2 -f -o gpurun_out/prof_${tag}_step \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_step_$tag.log 2>&1
ncu -i gpurun_out/prof_${tag}_step.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_step_raw.csv 2>/dev/null
rm -f gpurun_out/prof_${tag}_gn.ncu-rep gpurun_out/prof_${tag}_step.ncu-rep
ls -la gpurun_out/*$tag*
