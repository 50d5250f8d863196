tag=${1:-r2k}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k attention --timeout 60 > gpurun_out/pytest_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$tag.txt; tail -8 gpurun_out/pytest_$tag.txt
grep -q "rc=0" gpurun_out/pytest_$tag.txt || exit 1
for turns in 1 0; do
echo "--- AZB_ATTN_TURNS=$turns"
AZB_ATTN_TURNS=$turns timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "mma.sync\|32x32, 512" | tee gpurun_out/attn_bench_turns${turns}_$tag.txt
done
timeout 120 python scripts/attn_trace.py --qknorm 2>&1 | head -14 | tee gpurun_out/attn_trace_qk_$tag.txt
