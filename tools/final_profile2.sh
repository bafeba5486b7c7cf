#!/bin/bash
tag=${1:-x}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --no-cpu --no-extras --steps 2 --warmup 1 > gpurun_out/l_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:normals_knn_coop -s 1 -c 1 -o gpurun_out/prof_normals_${tag} -f python bench.py --no-cpu --no-extras --steps 1 --warmup 1 > gpurun_out/pn_${tag}.log 2>&1
python bench.py > gpurun_out/bench_default_${tag}.json 2> gpurun_out/bench_default_${tag}.err
tail -c 300 gpurun_out/bench_default_${tag}.json
