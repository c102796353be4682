#!/bin/bash
# First GPU pass of round 2: parity tests, smoke, bench lines (headline, 8-GPU shard size, generic variant, reference arm),
# launch list of smoke().  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python bench.py --steps 10 --warmup 3 --sims 8192 --no-cpu-baseline > gpurun_out/r2a_bench_8192.json 2>> gpurun_out/r2a_bench.err
python bench.py --variant generic --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_generic.json 2>> gpurun_out/r2a_bench.err
python bench.py --variant generic --sims 8192 --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_generic_8192.json 2>> gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.json 2>> gpurun_out/r2a_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke_ncu.log 2>&1
head -c 600 gpurun_out/r2a_bench.json; echo; head -c 300 gpurun_out/r2a_bench_8192.json; echo
