tag=${1:-r2t}
mkdir -p gpurun_out
for v in plain,-1,-1 plain,0,-1; do
name=$(echo $v | tr ',-' '_m')
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 1 -c 1 -f \
  -o gpurun_out/prof_${tag}_$name python scripts/unet_conv_ab.py --only $v > gpurun_out/ncu_${tag}_$name.log 2>&1
ncu -i gpurun_out/prof_${tag}_$name.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${tag}_$name.ncu-rep --page source --csv > gpurun_out/prof_${tag}_${name}_source.csv 2>/dev/null
echo "=== $v"; python scripts/ncu_attn_summary.py gpurun_out/prof_${tag}_$name 28
done
