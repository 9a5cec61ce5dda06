#!/bin/bash
# copy the evidence of a tools/gpu/run_y.sh run (gpurun_out/<tag>_*) into profiles/ under the round's names
# usage: tools/collect_profiles.sh <gpurun tag> <profiles prefix>
cd "$(dirname "$0")/.."
T=$1; P=$2
python tools/make_ncu_traffic.py gpurun_out/${T}_full.ncu-rep "ball_h0=0.02" profiles/ncu_traffic.json > /dev/null
python tools/make_ncu_traffic.py gpurun_out/${T}_eage75_full.ncu-rep "eage_shaped_grid222x700x700_hmin=75_freq=4" profiles/ncu_traffic.json > /dev/null
python profiles/summarize_ncu.py gpurun_out/${T}_full.ncu-rep profiles/${P}_ncu_full_ball_h0.02_tile_layout_summary.json > /dev/null
python profiles/summarize_ncu.py gpurun_out/${T}_eage75_full.ncu-rep profiles/${P}_ncu_full_eage_hmin75_bucket_layout_summary.json > /dev/null
cp gpurun_out/${T}_launches.csv profiles/${P}_ncu_launches_ball_h0.02.csv
grep '^{' gpurun_out/${T}_bench_default.json | tail -1 > profiles/${P}_bench_default_line.json
for w in ball_0.02 disk_0.01 eage_150 eage_75 bp2004_75 bp2004_25; do
  cp gpurun_out/${T}_kernels_$w.json profiles/${P}_kernels_$w.json
  grep '^{' gpurun_out/${T}_bench_$w.json | tail -1 > profiles/${P}_bench_$w.json
done
for tool in memcheck racecheck synccheck; do cp gpurun_out/${T}_sanitizer_${tool}_tiles.log profiles/${P}_sanitizer_${tool}_smoke_both_layouts.log; done
for f in gpurun_out/${T}_pytest*.log; do cp $f profiles/${P}_$(basename $f | sed "s/^${T}_//"); done
ls profiles | grep "^${P}_" | wc -l
