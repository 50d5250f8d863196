set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_halo_gpu.py -q 2>&1 | tail -25 ) > gpurun_out/halo_tests.txt 2>&1; tail -25 gpurun_out/halo_tests.txt
timeout 600 python scripts/halo_ab.py > gpurun_out/halo_ab.txt 2>&1; cat gpurun_out/halo_ab.txt
