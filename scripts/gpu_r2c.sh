# Round-2 GPU pass C: new tests, per-launch tables of configs 2 / 4, ncu evidence for profiles/.
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nn_gpu.py tests/test_step_gpu.py tests/test_samplers_gpu.py tests/test_conv_rowepi_gpu.py "tests/test_adm_gpu.py::test_every_card_native_vs_oracle_and_reference" -q --maxfail=30 > gpurun_out/pytest_gpu_$tag.txt 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu_$tag.txt
for cfg in unet64 dit_b2; do
  for epi in -1 1; do
    timeout 300 python scripts/plan_detail.py --config $cfg --rowepi $epi > gpurun_out/plan_${cfg}_epi${epi}_$tag.txt 2>&1
  done
done
tail -8 gpurun_out/plan_unet64_epi-1_$tag.txt
timeout 600 python bench.py --steps 3 --warmup 2 --no-eager-gpu --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
print(d['value'], d['step_kernel']['frac'], d['step_kernel_noise'])"
# ---- ncu evidence
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1600 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-gpu --no-extras --profile-range \
  > gpurun_out/ncu_bench_$tag.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 1 -c 1 -f \
  -o gpurun_out/prof_${tag}_conv_fused python scripts/conv_one.py --hw 256 --ci 256 --co 256 --fused 1 > gpurun_out/ncu_conv_$tag.log 2>&1
ncu -i gpurun_out/prof_${tag}_conv_fused.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_conv_fused_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:step_vec4 -c 4 -f -o gpurun_out/prof_${tag}_step \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-gpu > gpurun_out/ncu_step_$tag.log 2>&1
ncu -i gpurun_out/prof_${tag}_step.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_step_raw.csv 2>/dev/null
rm -f gpurun_out/prof_${tag}_step.ncu-rep
for cfg in unet64 dit_b2; do
  timeout 400 ncu --set full --clock-control none -k regex:"conv_gemm|attention_tc|rownorm|segment" --launch-skip 120 -c 24 -f -o gpurun_out/prof_${tag}_$cfg \
    python bench.py --config $cfg --steps 1 --warmup 1 --no-cpu-baseline --no-eager-gpu --no-extras > gpurun_out/ncu_${cfg}_$tag.log 2>&1
  ncu -i gpurun_out/prof_${tag}_$cfg.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_${cfg}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_${tag}_$cfg.ncu-rep
done
ls -la gpurun_out/*$tag*
