# Round-2 GPU pass E: TF32 mode after the GroupNorm / pair work.
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tf32_gpu.py -q --maxfail=40 -s > gpurun_out/pytest_tf32_$tag.txt 2>&1
echo "tf32 rc=$?"; grep -E "rel-L2|mean\|d\||passed|failed|Error|error|assert" gpurun_out/pytest_tf32_$tag.txt | tail -40
for pair in -1 0; do
AZB_PAIR=$pair AZB_PRECISION=tf32 timeout 900 python bench.py --steps 2 --warmup 1 --no-eager-gpu --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}_adm_tf32_p$pair.json 2> gpurun_out/bench_${tag}_adm_tf32_p$pair.err
echo "bench tf32 pair=$pair rc=$?"; tail -c 600 gpurun_out/bench_${tag}_adm_tf32_p$pair.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_${tag}_adm_tf32_p$pair.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('forward_kernels'), d['roofline'])"
done
