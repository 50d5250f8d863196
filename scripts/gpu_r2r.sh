tag=${1:-r2r}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_conv_rowepi_gpu.py tests/test_nn_gpu.py tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/pytest_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$tag.txt; tail -5 gpurun_out/pytest_$tag.txt
timeout 300 python scripts/plan_detail.py --config dit_b2 > gpurun_out/plan_dit_b2_$tag.txt 2>&1; grep -E "^#" gpurun_out/plan_dit_b2_$tag.txt; sed -n 2,9p gpurun_out/plan_dit_b2_$tag.txt
timeout 300 python scripts/plan_detail.py --config unet64 > gpurun_out/plan_unet64_$tag.txt 2>&1; grep -E "^#" gpurun_out/plan_unet64_$tag.txt
timeout 300 python scripts/plan_detail.py --config unet64 --rowepi 1 > gpurun_out/plan_unet64_epi1_$tag.txt 2>&1; grep -E "^#" gpurun_out/plan_unet64_epi1_$tag.txt
timeout 600 python bench.py --config dit_b2 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_${tag}_dit_b2.json 2> gpurun_out/bench_${tag}_dit_b2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${tag}_dit_b2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_dit_b2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline_e2e']['frac'])"
