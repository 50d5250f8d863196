# Round-2 GPU pass B: the row-domain epilogue (EPI 2) -- its own tests first, then the suite, then A/B benches.
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_rowepi_gpu.py -q --maxfail=60 > gpurun_out/pytest_rowepi_$tag.txt 2>&1
echo "rowepi rc=$?"; tail -15 gpurun_out/pytest_rowepi_$tag.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_conv_rowepi_gpu.py --deselect "tests/test_adm_gpu.py::test_full_card_64_steps_vs_oracle" > gpurun_out/pytest_gpu_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.txt; tail -5 gpurun_out/pytest_gpu_$tag.txt
timeout 600 python -m pytest "tests/test_adm_gpu.py::test_full_card_64_steps_vs_oracle" -q -s > gpurun_out/pytest_64step_$tag.txt 2>&1
echo "64step rc=$?"; grep -E "drift|local|card" gpurun_out/pytest_64step_$tag.txt
for cfg in adm unet64 dit_b2; do
  for epi in 1 0; do
    AZB_ROWEPI=$epi timeout 600 python bench.py --config $cfg --steps 3 --warmup 2 --no-eager-gpu --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}_${cfg}_epi$epi.json 2> gpurun_out/bench_${tag}_${cfg}_epi$epi.err
    echo "bench $cfg epi=$epi rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${tag}_${cfg}_epi$epi.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('forward_kernels'))" 2>&1 | tail -2
  done
done
