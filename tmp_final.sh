#!/bin/bash
timeout 600 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r02_n1.err
timeout 120 python tools/prof_extract.py 128 3 > gpurun_out/prof_extract.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python tools/prof_extract.py 128 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -c 40 -o /tmp/prof_step_r02 python tools/prof_extract.py 128 2 > gpurun_out/prof_step_r02.log 2>&1; tail -1 gpurun_out/prof_step_r02.log
ncu -i /tmp/prof_step_r02.ncu-rep --page raw --csv > gpurun_out/raw_step_r02.csv 2>/dev/null
timeout 120 python tools/prof_knn.py 2000000 2 > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:k_knn2_tc -c 1 -o /tmp/prof_knn_tc_r02 python tools/prof_knn.py 2000000 2 > gpurun_out/prof_knn_tc_r02.log 2>&1; tail -1 gpurun_out/prof_knn_tc_r02.log
ncu -i /tmp/prof_knn_tc_r02.ncu-rep --page raw --csv > gpurun_out/raw_knn_tc_r02.csv 2>/dev/null
ls -la gpurun_out/ | head -20; du -sh gpurun_out
