tag=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_halo_gpu.py tests/test_nn_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$tag.txt; tail -8 gpurun_out/pytest_$tag.txt
grep -q "rc=0" gpurun_out/pytest_$tag.txt || exit 1
python scripts/unet_conv_ab.py 2>&1 | grep "h-1 r-1\|halo -1 rowepi -1" | tee gpurun_out/unet_conv_ab_64_$tag.txt
python scripts/unet_conv_ab.py --c 128 --hw 32 2>&1 | grep "halo -1 rowepi -1" | tee gpurun_out/unet_conv_ab_128_$tag.txt
python scripts/unet_conv_ab.py --c 256 --hw 16 2>&1 | grep "halo -1 rowepi -1" | tee gpurun_out/unet_conv_ab_256_$tag.txt
timeout 300 python scripts/graph_trace.py --config unet64 > gpurun_out/graph_trace_unet64_$tag.txt 2>&1; tail -3 gpurun_out/graph_trace_unet64_$tag.txt
timeout 600 python bench.py --config unet64 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_${tag}_unet64.json 2> gpurun_out/bench_${tag}_unet64.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${tag}_unet64.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_unet64.json').read().strip().splitlines()[-1]); print('unet64', d['value'], d['e2e']['value'], d['roofline_e2e']['frac'])"
