tag=${1:-r2p}
mkdir -p gpurun_out
for re in -1 1; do for act in none silu; do
python scripts/gemm_one.py --act $act --rowepi $re
done; done 2>&1 | tee gpurun_out/gemm_one_$tag.txt
python scripts/gemm_one.py --act silu --rowepi 1 --pair 1 2>&1 | tee -a gpurun_out/gemm_one_$tag.txt
python scripts/gemm_one.py --act silu --rowepi -1 --pair 1 2>&1 | tee -a gpurun_out/gemm_one_$tag.txt
python scripts/gemm_one.py --k 3072 --n 768 --gate 1 --res 1 --rowepi -1 2>&1 | tee -a gpurun_out/gemm_one_$tag.txt
python scripts/gemm_one.py --k 3072 --n 768 --gate 1 --res 1 --rowepi 1 2>&1 | tee -a gpurun_out/gemm_one_$tag.txt
python scripts/gemm_one.py --k 768 --n 768 --res 1 --rowepi -1 2>&1 | tee -a gpurun_out/gemm_one_$tag.txt
python scripts/gemm_one.py --k 768 --n 768 --res 1 --rowepi 1 2>&1 | tee -a gpurun_out/gemm_one_$tag.txt
for re in -1 1; do python scripts/plan_detail.py --config unet64 --rowepi $re | grep "^#"; done
