tag=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_gpu.py tests/test_conv_rowepi_gpu.py tests/test_nn_gpu.py tests/test_adm_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$tag.txt; tail -8 gpurun_out/pytest_$tag.txt
grep -q "rc=0" gpurun_out/pytest_$tag.txt || exit 1
timeout 300 python scripts/graph_trace.py --config unet64 > gpurun_out/graph_trace_unet64_$tag.txt 2>&1; tail -45 gpurun_out/graph_trace_unet64_$tag.txt
for cfg in unet64 dit_b2; do
timeout 600 python bench.py --config $cfg --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_${tag}_$cfg.json 2> gpurun_out/bench_${tag}_$cfg.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${tag}_$cfg.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_$cfg.json').read().strip().splitlines()[-1]); print('$cfg', d['value'], d['e2e']['value'], d['roofline_e2e']['frac'])"
done
timeout 900 python bench.py --no-cpu-baseline --no-eager-gpu --steps 3 --warmup 3 > gpurun_out/bench_${tag}_adm.json 2> gpurun_out/bench_${tag}_adm.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${tag}_adm.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_adm.json').read().strip().splitlines()[-1]); print('adm', d['value'], d['e2e']['value'], d['roofline_e2e']['frac'], d['roofline']['frac'])"
