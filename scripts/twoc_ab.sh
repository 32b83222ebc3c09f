for g in 4 2 1; do for sdo in 0 1; do
echo "GRID_PER_SM=$g SI_GRAPHS=$sdo"; SI_G3_GRID_PER_SM=$g SI_GRAPHS=$sdo timeout 300 python scripts/twoc_bench.py 2>&1 | tail -2
done; done
