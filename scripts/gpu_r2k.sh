tag=${1:-r2k}
mkdir -p gpurun_out
for turns in 1 0; do
echo "--- AZB_ATTN_TURNS=$turns"
AZB_ATTN_TURNS=$turns timeout 300 python scripts/attn_bench.py 2>&1 | grep -v "mma.sync\|32x32, 512" | tee gpurun_out/attn_bench_turns${turns}_$tag.txt
done
AZB_ATTN_TURNS=0 timeout 120 python scripts/attn_trace.py 2>&1 | head -14 | tee gpurun_out/attn_trace_$tag.txt
