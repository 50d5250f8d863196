# ncu --set full captures of this session's kernels (profiles/r2_ncu_full_*): short-sequence attention (DiT shape, fused q/k norm
# is the 2nd attention_t256 launch of attn_trace --qknorm warm-up), the 64 -> 64 U-Net convolution with the fused per-pixel norm,
# the 768 -> 3072 + SiLU token GEMM with the row-domain epilogue
tag=${1:-r2ae}
mkdir -p gpurun_out
cap() {  # name, kernel regex, launch-skip, command...
  name=$1; regex=$2; skip=$3; shift 3
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$regex --launch-skip $skip -c 1 -f -o gpurun_out/prof_${tag}_$name "$@" > gpurun_out/ncu_${tag}_$name.log 2>&1
  ncu -i gpurun_out/prof_${tag}_$name.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${tag}_$name.ncu-rep --page source --csv > gpurun_out/prof_${tag}_${name}_source.csv 2>/dev/null
  echo "=== $name"; python scripts/ncu_attn_summary.py gpurun_out/prof_${tag}_$name 12
}
cap attention_t256_qknorm attention_t256 1 python scripts/attn_trace.py --qknorm
cap unet_conv64_norm conv_gemm 1 python scripts/unet_conv_ab.py --only norm+act,-1,-1
cap unet_conv64_gate_res_sums conv_gemm 1 python scripts/unet_conv_ab.py --only +gate+res+sums,-1,-1
cap dit_fc1_silu conv_gemm 0 python scripts/gemm_one.py --act silu --once
rm -f gpurun_out/prof_${tag}_*.ncu-rep
