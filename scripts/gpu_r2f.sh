# Round-2 GPU pass F: what the driver runs at round end (full GPU suite, smoke, default bench, reference arm) + the
# bench lines of the other configurations for profiles/.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.txt; tail -6 gpurun_out/pytest_gpu_$tag.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.txt 2>&1; echo "smoke rc=$?"; tail -9 gpurun_out/smoke_$tag.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$tag.err
for cfg in unet64 dit_b2 mlp; do
  timeout 600 python bench.py --config $cfg > gpurun_out/bench_${tag}_$cfg.json 2> gpurun_out/bench_${tag}_$cfg.err; echo "bench $cfg rc=$?"
done
timeout 900 python bench.py --precision tf32 --steps 2 --warmup 1 --no-cpu-baseline --no-eager-gpu > gpurun_out/bench_${tag}_adm_tf32.json 2> gpurun_out/bench_${tag}_adm_tf32.err; echo "tf32 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*${tag}*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'roofline', d.get('roofline',{}).get('frac'), 'e2e_frac', d.get('roofline_e2e',{}).get('frac'), 'eager', {k:round(v['value'],3) for k,v in d.get('eager_gpu',{}).items() if isinstance(v,dict)})
    except Exception as e:
        print(f, 'ERR', e)
PY
# launch list of the bench command's timed region (ncu, cold-cache, serialised: shares, not absolute times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1600 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-gpu --profile-range \
  > gpurun_out/ncu_bench_$tag.log 2>&1; echo "ncu launch list rc=$?"
