#!/bin/bash
# Runs on the GPU box (gpurun, 1 GPU): the whole single-GPU evidence set of a build -> gpurun_out/<tag>_* (copied to profiles/ afterwards).
#   GPU tests, smoke, the bench line of every config, the reference arm, the ncu launch list + one ncu --set full capture with its derived pages, sanitizers
T=${1:-r2h}
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --config ggx > $O/${T}_bench_ggx.json 2>> $O/${T}_bench.err
python bench.py --config arm > $O/${T}_bench_arm.json 2>> $O/${T}_bench.err
python bench.py --config scale --steps 3 --warmup 3 > $O/${T}_bench_scale.json 2>> $O/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
bash tools/capture_profiles.sh $T
rm -f $O/${T}_full.ncu-rep
for tool in memcheck racecheck initcheck; do timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/${T}_sanitize_$tool.log 2>&1; done
tail -2 $O/${T}_pytest_gpu.log; cat $O/${T}_smoke.log | tail -1; for f in bench bench_ggx bench_arm bench_scale bench_reference; do python -c "import json,sys; d=json.loads(open('$O/${T}_$f.json').read().strip().splitlines()[-1]); print('$f', d.get('ms_per_step'), d.get('value'), (d.get('roofline') or {}).get('frac'))"; done; for t in memcheck racecheck initcheck; do tail -n 1 $O/${T}_sanitize_$t.log; done
