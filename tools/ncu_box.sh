#!/bin/bash
# Runs ON the GPU box (under gpurun): ncu launch list of the bench command, one --set full capture of a buildgraph at the
# bench configuration, and the text summaries of that capture (the .ncu-rep itself stays on the box: it exceeds what
# gpurun_out/ may carry back).   usage: tools/ncu_box.sh TAG [N_READS] [K]
TAG=${1:-r02}; N=${2:-20000000}; K=${3:-31}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_${K}.csv \
    python bench.py --reads-per-gpu $N -k $K --steps 2 --warmup 1 --no-cpu-baseline --no-hash > $OUT/${TAG}_ncu_bench.log 2>&1
REP=/tmp/${TAG}_full.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_edge_part|k_split|k_count|k_node_part|k_node_count|k_item_part|k_row_part|k_sort_emit|k_msd|k_out_" \
    -o ${REP%.ncu-rep} -f python tests/gpu_profile.py $N $K > $OUT/${TAG}_full.log 2>&1
ls -la $REP >> $OUT/${TAG}_full.log
python tools/ncu_summary.py $REP "$TAG -- ncu --set full --clock-control none, tests/gpu_profile.py $N $K (one buildgraph at the bench configuration)" > $OUT/${TAG}_ncu_full.md
S1=$(grep -o "'n_items': [0-9]*" $OUT/${TAG}_full.log | head -1 | grep -o "[0-9]*$")
S2=$(grep -o "'n_items': [0-9]*" $OUT/${TAG}_full.log | tail -1 | grep -o "[0-9]*$")
python tools/ncu_traffic.py $REP $N 150 $K 2 ${S1:-0} ${S2:-0} > $OUT/${TAG}_traffic.json
for kern in k_edge_part k_split k_count k_node_part k_node_count k_item_part k_sort_emit; do
    python tools/ncu_lines.py $REP "$kern" 0.7 > $OUT/${TAG}_lines_${kern}.txt 2>&1
done
ls -la $OUT | tail -20
