#!/bin/bash
# Reduces the raw evidence of profiles/collect_r1.sh (gpurun_out/) to the text summaries committed under profiles/r1/.
O=gpurun_out; R=profiles/r1
mkdir -p $R
cp $O/r1_bench_default.json $R/bench_default.json
cp $O/r1_bench_headline.json $R/bench_headline.json
cp $O/r1_launches_c2.csv $R/launches_bench_c2.csv
python profiles/launch_summary.py $O/r1_launches_c2.csv > $R/launches_bench_c2.txt
grep -v "^/\|_warn_once" $O/r1_timeline_c2.txt > $R/timeline_c2.txt
grep -v "^/\|_warn_once" $O/r1_timeline_headline.txt > $R/timeline_headline.txt
cp $O/r1_kbench.txt $R/kbench.txt
for n in r1_photo_l1_c2 r1_photo_l1_headline r1_aux_c2 r1_photo_min_c2min r1_cloud; do
  python profiles/ncu_summary.py $O/$n.ncu-rep > $R/$n.summary.txt
done
python profiles/ncu_hot.py $O/r1_photo_l1_c2.ncu-rep > $R/r1_photo_l1_c2.opcodes.txt
python profiles/ncu_hot.py $O/r1_photo_l1_headline.ncu-rep > $R/r1_photo_l1_headline.opcodes.txt
python profiles/ncu_lines.py $O/r1_photo_l1_c2.ncu-rep "" 40 > $R/r1_photo_l1_c2.lines.txt
python profiles/ncu_lines.py $O/r1_photo_min_c2min.ncu-rep "" 40 > $R/r1_photo_min_c2min.lines.txt
grep -v "^/\|_warn_once\|Warning" $O/r1_edge_bench.txt > $R/edge_bench.txt
grep -v "^/\|_warn_once\|Warning" $O/r1_aux_bench.txt > $R/aux_bench.txt
