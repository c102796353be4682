#!/bin/bash
# wide-stepper parity + C5 timing, MMA vs DFMA variant (run under gpurun)
python -m pytest tests/test_cuda_parity.py -q -m gpu -k "wide" 2>&1 | tail -3
DPS_WIDE_NO_MMA=1 python -m pytest tests/test_cuda_parity.py -q -m gpu -k "wide" 2>&1 | tail -1
python tools/bench_configs.py c5
DPS_WIDE_NO_MMA=1 python tools/bench_configs.py c5
