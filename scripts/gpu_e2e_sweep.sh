set -x
# n_maps lanes threads reps upload_group chunk tail first_group_divisor
for cfg in "256 2 16 10 64 128 0 1" "256 2 16 10 64 128 0 4" "256 2 16 10 64 128 0 8" "256 2 16 10 64 128 0 2" "256 2 16 10 32 128 0 2" "256 2 16 10 64 128 0 1" "256 2 16 10 64 128 0 4"; do
  echo "== $cfg"; python scripts/exp_heap_detect.py $cfg 2>&1 | grep "rep" | awk '{print $6}' | sort -n | head -6 | tr '\n' ' '; echo
done
CSM_HOST_TRACE=1 python scripts/exp_heap_detect.py 256 2 16 5 64 128 0 4 2>&1 | tail -26
