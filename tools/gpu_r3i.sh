#!/bin/bash
mkdir -p gpurun_out/r3i
python -m pytest tests/test_cuda_prologue.py -x -q 2>&1 | tail -25 | tee gpurun_out/r3i/pytest_prologue.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r3i/pytest_gpu.log
python tools/prologue_probe.py 2>&1 | tee gpurun_out/r3i/prologue_probe.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r3i/bench.json 2> gpurun_out/r3i/bench.err; tail -c 1500 gpurun_out/r3i/bench.json
