#!/bin/bash
# one GPU round trip: parity tests, then the two single-GPU bench workloads (no CPU leg)
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$tag.log
python bench.py --workload view1m --no-cpu --steps 5 > gpurun_out/bench_view1m_$tag.log 2>&1
python bench.py --no-cpu --steps 3 > gpurun_out/bench_scene10m_$tag.log 2>&1
tail -3 gpurun_out/pytest_gpu_$tag.log
