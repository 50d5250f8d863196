set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv
( time timeout 900 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu_r1p.txt 2>&1; tail -3 gpurun_out/pytest_gpu_r1p.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1p.txt 2>&1; tail -4 gpurun_out/smoke_r1p.txt
timeout 600 python bench.py > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err; tail -1 gpurun_out/bench_r1p.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r1p.json 2> gpurun_out/bench_ref_r1p.err; tail -1 gpurun_out/bench_ref_r1p.json
timeout 300 python scripts/adm_profile.py > gpurun_out/adm_profile_r1p.txt 2>&1; tail -12 gpurun_out/adm_profile_r1p.txt
