#!/bin/bash
# Round-2 evidence, one GPU box: plain bench first (its own exit code), then the ncu passes of the same commands.
# Raw reports land in gpurun_out/; profiles/summarise_r2.sh reduces them to the text committed under profiles/r2/.
set -x
O=gpurun_out
python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err || exit 1
python bench.py --workload headline --no-cpu --no-cloud > $O/r2_bench_headline.json 2>> $O/r2_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-graph --no-cloud > $O/r2_ncu_launches.log 2>&1
python profiles/trace_step.py c2 20 > $O/r2_timeline_c2.txt 2>&1
python profiles/trace_step.py headline 40 > $O/r2_timeline_headline.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:photo_l1 -s 2 -c 1 -o $O/r2_photo_l1_c2 python profiles/prof_photo.py c2 4 > $O/r2_ncu_a.log 2>&1
$NCU -k regex:photo_l1\|photo_finalize -s 4 -c 2 -o $O/r2_photo_l1_headline python profiles/prof_photo.py headline 4 > $O/r2_ncu_b.log 2>&1
$NCU -k regex:lowres_merge\|smooth -s 4 -c 2 -o $O/r2_aux_c2 python profiles/prof_photo.py c2 4 > $O/r2_ncu_c.log 2>&1
$NCU -k regex:photo_min -s 2 -c 1 -o $O/r2_photo_min_c2min python profiles/prof_photo.py c2min 4 > $O/r2_ncu_d.log 2>&1
$NCU -k regex:cloud_ -s 4 -c 2 -o $O/r2_cloud python profiles/prof_cloud.py > $O/r2_ncu_e.log 2>&1
$NCU -k regex:velo_ -s 4 -c 2 -o $O/r2_velo python profiles/prof_velo.py > $O/r2_ncu_f.log 2>&1
$NCU -k regex:edge_ -s 12 -c 6 -o $O/r2_edge python profiles/prof_edge.py > $O/r2_ncu_g.log 2>&1
python profiles/cloud_kbench.py velo > $O/r2_cloud_kbench.txt 2>&1
python profiles/kbench.py headline headline64 c1 c2 c3 c5 c2min > $O/r2_kbench.txt 2>&1
for w in c3 c5 c5e; do python bench.py --workload $w --no-cpu --no-cloud > $O/r2_bench_$w.json 2>> $O/r2_bench_default.err; done
grep -c Report $O/r2_ncu_?.log
python profiles/edge_bench.py > $O/r2_edge_bench.txt 2>&1
python profiles/aux_bench.py > $O/r2_aux_bench.txt 2>&1
