# Round-2 GPU pass A: full GPU test suite, then the bench line of every BASELINE configuration.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$tag.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --durations=15 > gpurun_out/pytest_gpu_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.txt
tail -5 gpurun_out/pytest_gpu_$tag.txt
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$tag.err
for cfg in unet64 dit_b2 mlp; do
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 2 > gpurun_out/bench_${tag}_$cfg.json 2> gpurun_out/bench_${tag}_$cfg.err
  echo "bench $cfg rc=$?"; tail -c 300 gpurun_out/bench_${tag}_$cfg.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
echo "ref rc=$?"
