#!/bin/bash
# Final single-GPU validation of the round: the driver's sequence (GPU tests, smoke, bench both arms) + side configs.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python bench.py --variant generic --no-cpu-baseline --no-e2e > gpurun_out/final_generic.json 2>> gpurun_out/final_bench.err
python bench.py --config c4 --steps 10 > gpurun_out/final_c4.json 2>> gpurun_out/final_bench.err
python bench.py --config c5 --steps 3 > gpurun_out/final_c5.json 2>> gpurun_out/final_bench.err
python - <<'PY'
import json
for f in ('final_ref', 'final_bench', 'final_generic', 'final_c4', 'final_c5'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, 'value %.4g %s ms/step %.3f' % (d['value'], d['unit'], d['ms_per_step']), 'e2e %.4g' % d['e2e']['value'] if d.get('e2e') else '',
              'frac %.4f' % d['roofline']['frac'] if d.get('roofline') else '')
    except Exception as e:
        print(f, 'FAILED', e)
PY
