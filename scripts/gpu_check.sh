# One GPU round trip that validates a build: gpurun -- 'bash scripts/gpu_check.sh [tag]'
tag=${1:-check}
set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_$tag.txt 2>&1; tail -8 gpurun_out/pytest_gpu_$tag.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.txt 2>&1; tail -3 gpurun_out/smoke_$tag.txt
timeout 300 python scripts/adm_profile.py > gpurun_out/adm_profile_$tag.txt 2>&1; head -9 gpurun_out/adm_profile_$tag.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -1 gpurun_out/bench_$tag.json
