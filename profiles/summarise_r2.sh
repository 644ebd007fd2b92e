#!/bin/bash
# Reduces the raw evidence of profiles/collect_r2.sh (gpurun_out/) to the text summaries committed under profiles/r2/.
O=gpurun_out; R=profiles/r2
mkdir -p $R
cp $O/r2_bench_default.json $R/bench_default.json
cp $O/r2_bench_headline.json $R/bench_headline.json
cp $O/r2_launches_c2.csv $R/launches_bench_c2.csv
python profiles/launch_summary.py $O/r2_launches_c2.csv > $R/launches_bench_c2.txt
grep -v "^/\|_warn_once" $O/r2_timeline_c2.txt > $R/timeline_c2.txt
grep -v "^/\|_warn_once" $O/r2_timeline_headline.txt > $R/timeline_headline.txt
cp $O/r2_kbench.txt $R/kbench.txt
cp $O/r2_cloud_kbench.txt $R/cloud_kbench.txt
for w in c3 c5 c5e; do cp $O/r2_bench_$w.json $R/bench_$w.json; done
for n in r2_photo_l1_c2 r2_photo_l1_headline r2_aux_c2 r2_photo_min_c2min r2_cloud r2_velo r2_edge; do
  python profiles/ncu_summary.py $O/$n.ncu-rep > $R/$n.summary.txt
done
python profiles/ncu_hot.py $O/r2_photo_l1_c2.ncu-rep > $R/r2_photo_l1_c2.opcodes.txt
python profiles/ncu_hot.py $O/r2_photo_l1_headline.ncu-rep > $R/r2_photo_l1_headline.opcodes.txt
python profiles/ncu_lines.py $O/r2_photo_l1_c2.ncu-rep "" 40 > $R/r2_photo_l1_c2.lines.txt
python profiles/ncu_lines.py $O/r2_photo_min_c2min.ncu-rep "" 40 > $R/r2_photo_min_c2min.lines.txt
grep -v "^/\|_warn_once\|Warning" $O/r2_edge_bench.txt > $R/edge_bench.txt
grep -v "^/\|_warn_once\|Warning" $O/r2_aux_bench.txt > $R/aux_bench.txt
