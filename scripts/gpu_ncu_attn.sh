# ncu --set full of the short-sequence attention kernel on the DiT-B/2 shape (3rd launch of the kernel in attn_bench --once)
tag=${1:-r2j}
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention_t256 --launch-skip 2 -c 1 -f \
  -o gpurun_out/prof_${tag}_attn_t256 python scripts/attn_bench.py --once > gpurun_out/ncu_attn_$tag.log 2>&1
tail -3 gpurun_out/ncu_attn_$tag.log
ncu -i gpurun_out/prof_${tag}_attn_t256.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_attn_t256_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${tag}_attn_t256.ncu-rep --page source --csv > gpurun_out/prof_${tag}_attn_t256_source.csv 2>/dev/null
ls -la gpurun_out/*${tag}*
