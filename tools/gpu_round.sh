mkdir -p gpurun_out
L=gpurun_out/r02m_cycle.log; rm -f $L
for env in "MGB_X=0" "MGB_PDL=0" "MGB_TAIL_SMEM=0" "MGB_TAIL=0" "MGB_PDL=0 MGB_TAIL=0"; do
echo "== $env" >> $L
for lv in 9 8 7 6 5 4 3; do
env $env python tools/cycle_case.py --levels $lv --cycles 20 --batch 1 >> $L 2>&1
done
done
cat $L
