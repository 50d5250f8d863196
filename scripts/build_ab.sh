# Same-box A/B of two builds of libazb.so: AZB_LIBRARY selects the library azula_b200/_lib.py loads.
#   gpurun -- 'bash scripts/build_ab.sh build/libazb_old.so'
old=$1
for i in 1 2 3; do
  AZB_LIBRARY=$old timeout 300 python scripts/adm_profile.py 2>&1 | sed -n 2,4p | sed "s/^/old $i: /"
  timeout 300 python scripts/adm_profile.py 2>&1 | sed -n 2,4p | sed "s/^/new $i: /"
done
