for v in 2 3; do
export THCM_ASM_PIPE=$v
for box in "33 5 3 1" "40 5 4 1" "40 9 6 0" "70 9 6 1"; do
timeout 120 python scripts/debug_pipe.py $box 2>&1 | grep -v "^$" | grep "prog\|stuck\|equal" | sort | head -20
done
done
