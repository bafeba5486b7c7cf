#!/bin/bash
# one GPU round trip: ncu launch lists + full captures of the feature kernel (scene, batch of views), then the default bench line
tag=${1:-x}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --no-cpu --no-extras --steps 2 --warmup 1 > gpurun_out/l_${tag}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_views.csv python bench.py --no-cpu --no-extras --workload views --steps 2 --warmup 1 > gpurun_out/lv_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:feature_kernel -s 1 -c 1 -o gpurun_out/prof_feat_${tag}_scene -f python bench.py --no-cpu --no-extras --steps 1 --warmup 1 > gpurun_out/p_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:feature_kernel -s 1 -c 1 -o gpurun_out/prof_feat_${tag}_views -f python bench.py --no-cpu --no-extras --workload views --steps 1 --warmup 1 > gpurun_out/pv_${tag}.log 2>&1
python bench.py > gpurun_out/bench_default_${tag}.json 2> gpurun_out/bench_default_${tag}.err
tail -c 600 gpurun_out/bench_default_${tag}.json
