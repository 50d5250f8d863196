set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1t.txt 2>&1; tail -4 gpurun_out/smoke_r1t.txt
timeout 900 python bench.py > gpurun_out/bench_r1t.json 2> gpurun_out/bench_r1t.err; tail -1 gpurun_out/bench_r1t.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r1t.json 2> gpurun_out/bench_ref_r1t.err; tail -1 gpurun_out/bench_ref_r1t.json
timeout 900 python scripts/nn_bench.py --help 2>&1 | head -20
