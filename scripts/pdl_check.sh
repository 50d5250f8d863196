set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_r1r.txt 2>&1; tail -8 gpurun_out/pytest_gpu_r1r.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1r_pdl.json 2> gpurun_out/bench_r1r_pdl.err; tail -1 gpurun_out/bench_r1r_pdl.json | cut -c1-400
AZB_PDL=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1r_nopdl.json 2> gpurun_out/bench_r1r_nopdl.err; tail -1 gpurun_out/bench_r1r_nopdl.json | cut -c1-400
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r1r_pdl2.json 2> gpurun_out/bench_r1r_pdl2.err; tail -1 gpurun_out/bench_r1r_pdl2.json | cut -c1-400
tail -3 gpurun_out/bench_r1r_pdl.err
