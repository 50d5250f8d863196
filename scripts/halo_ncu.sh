set -x
mkdir -p gpurun_out
for v in "skiphalo:--fused 0 --halo 1 --skip 512" "skipfused:--fused 1 --halo 1 --skip 512" "skiptap:--fused 0 --halo 0 --skip 512"; do
  name=${v%%:*}; args=${v#*:}
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm --launch-skip 1 -c 1 -f -o gpurun_out/prof_$name python scripts/conv_one.py --hw 256 --ci 256 --co 256 $args > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_${name}_sass.csv 2>/dev/null
done
