# Round-2 GPU pass H: persistent short-sequence attention kernel (logits + P resident in TMEM, q/k RMS norm in place): tests, A/B, DiT plan
tag=${1:-r2h}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k attention > gpurun_out/pytest_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$tag.txt; tail -15 gpurun_out/pytest_$tag.txt
grep -q "rc=0" gpurun_out/pytest_$tag.txt || exit 1
echo "--- t256 kernel"; timeout 300 python scripts/attn_bench.py 2>&1 | tee gpurun_out/attn_bench_small_$tag.txt
echo "--- ring kernel"; AZB_ATTN_SMALL=0 timeout 300 python scripts/attn_bench.py 2>&1 | grep -v mma.sync | tee gpurun_out/attn_bench_ring_$tag.txt
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_nn_gpu.py tests/test_adm_gpu.py -m gpu -q -x > gpurun_out/pytest2_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest2_$tag.txt; tail -5 gpurun_out/pytest2_$tag.txt
timeout 300 python scripts/plan_detail.py --config dit_b2 > gpurun_out/plan_dit_b2_$tag.txt 2>&1; grep -E "^#" gpurun_out/plan_dit_b2_$tag.txt; sed -n 2,9p gpurun_out/plan_dit_b2_$tag.txt
timeout 600 python bench.py --config dit_b2 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_${tag}_dit_b2.json 2> gpurun_out/bench_${tag}_dit_b2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${tag}_dit_b2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_dit_b2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline_e2e']['frac'])"
