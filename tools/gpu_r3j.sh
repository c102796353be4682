#!/bin/bash
mkdir -p gpurun_out/r3j
python -m pytest tests/test_cuda_prologue.py tests/test_cuda_parity.py -x -q 2>&1 | tail -8 | tee gpurun_out/r3j/pytest.log
python tools/prologue_probe.py 2>&1 | tee gpurun_out/r3j/prologue_probe.txt
