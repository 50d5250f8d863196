tag=${1:-r2u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_gpu.py tests/test_conv_rowepi_gpu.py -m gpu -q -x > gpurun_out/pytest_$tag.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$tag.txt; tail -12 gpurun_out/pytest_$tag.txt
(cd scripts; for v in plain,-1,-1 +gate+res+sums,-1,-1 norm+act,-1,-1; do echo "== $v"; python conv_timeline.py --only $v | tail -5; done) 2>&1 | tee gpurun_out/conv_timeline_$tag.txt
python scripts/unet_conv_ab.py 2>&1 | grep "halo -1 rowepi -1" | tee gpurun_out/unet_conv_ab_64_$tag.txt; python scripts/unet_conv_ab.py --c 128 --hw 32 2>&1 | grep "halo -1 rowepi -1" | tee gpurun_out/unet_conv_ab_128_$tag.txt
python scripts/unet_conv_ab.py --c 256 --hw 16 2>&1 | grep "halo -1 rowepi -1" | tee gpurun_out/unet_conv_ab_256_$tag.txt
