tag=${1:-r2o}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ops_gpu.py tests/test_nn_gpu.py tests/test_adm_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$tag.txt; tail -5 gpurun_out/pytest_$tag.txt
timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "mma.sync" | tee gpurun_out/attn_bench_$tag.txt
timeout 600 python bench.py --config dit_b2 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_${tag}_dit_b2.json 2> gpurun_out/bench_${tag}_dit_b2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${tag}_dit_b2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_dit_b2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline_e2e']['frac'])"
