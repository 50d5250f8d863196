mkdir -p gpurun_out
(cd build/old_tree && timeout 300 python scripts/adm_profile.py --reps 5 2>/dev/null) > gpurun_out/adm_profile_old.txt
timeout 300 python scripts/adm_profile.py --reps 5 2>/dev/null > gpurun_out/adm_profile_new.txt
(cd build/old_tree && timeout 300 python scripts/adm_profile.py --reps 5 2>/dev/null) > gpurun_out/adm_profile_old2.txt
timeout 300 python scripts/adm_profile.py --reps 5 2>/dev/null > gpurun_out/adm_profile_new2.txt
head -12 gpurun_out/adm_profile_old2.txt; head -12 gpurun_out/adm_profile_new2.txt
