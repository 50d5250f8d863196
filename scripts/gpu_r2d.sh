# Round-2 GPU pass D: the TF32 mode's tests, step-kernel variants, DiT / UNet after the SiLU change, TF32 throughput.
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tf32_gpu.py -q --maxfail=40 -s > gpurun_out/pytest_tf32_$tag.txt 2>&1
echo "tf32 rc=$?"; grep -E "rel-L2|mean\|d\||passed|failed|Error|error" gpurun_out/pytest_tf32_$tag.txt | tail -40
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_samplers_gpu.py tests/test_conv_rowepi_gpu.py tests/test_nn_gpu.py tests/test_precond_gpu.py -q --maxfail=20 > gpurun_out/pytest_gpu_$tag.txt 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_$tag.txt
for cfg in adm unet64 dit_b2; do
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 2 --no-eager-gpu --no-cpu-baseline > gpurun_out/bench_${tag}_$cfg.json 2> gpurun_out/bench_${tag}_$cfg.err
  echo "bench $cfg rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_$cfg.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['step_kernel']['frac'], d['step_kernel_noise']['frac'], d.get('forward_kernels'))" 2>&1 | tail -2
done
AZB_PRECISION=tf32 timeout 900 python bench.py --steps 2 --warmup 1 --no-eager-gpu --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}_adm_tf32.json 2> gpurun_out/bench_${tag}_adm_tf32.err
echo "bench tf32 rc=$?"; tail -c 800 gpurun_out/bench_${tag}_adm_tf32.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_adm_tf32.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('forward_kernels'))"
