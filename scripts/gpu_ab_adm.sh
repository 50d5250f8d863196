# Same-box A/B of the ADM headline between build/old_tree (an older commit: mkdir -p build/old_tree && git archive <commit> | tar -x -C build/old_tree && (cd build/old_tree && python -m azula_b200.csrc.build); MEASURED_PEAKS.json and baseline/_ref copied next to it) and the working tree
mkdir -p gpurun_out
for i in 1 2; do
  timeout 600 python bench.py --no-cpu-baseline --no-eager-gpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new-first $i', round(d['value'],3), d['roofline']['frac'], d['roofline']['us_per_launch'], d['clocks']['sm_mhz'])"
  (cd build/old_tree && timeout 600 python bench.py --no-cpu-baseline --no-eager-gpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old $i', round(d['value'],3), d['roofline']['frac'], d['roofline']['us_per_launch'], d['clocks']['sm_mhz'])")
  timeout 600 python bench.py --no-cpu-baseline --no-eager-gpu --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new $i', round(d['value'],3), d['roofline']['frac'], d['roofline']['us_per_launch'], d['clocks']['sm_mhz'])"
done 2>&1 | tee gpurun_out/ab_adm.txt
