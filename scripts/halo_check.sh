set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_r1q.txt 2>&1; tail -8 gpurun_out/pytest_gpu_r1q.txt
timeout 300 python scripts/adm_profile.py > gpurun_out/adm_profile_r1q.txt 2>&1; head -32 gpurun_out/adm_profile_r1q.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; tail -1 gpurun_out/bench_r1q.json
