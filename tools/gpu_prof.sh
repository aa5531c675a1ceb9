#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-kernel-table --from-features > $O/bench_feat.json 2> $O/bench_feat.err
python -c "import json; d=json.loads([l for l in open('$O/bench_feat.json') if l.startswith('{')][-1]); print('features', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
tail -3 $O/bench_feat.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-kernel-table --from-features --criterion --optimizer > $O/bench_full.json 2> $O/bench_full.err
python -c "import json; d=json.loads([l for l in open('$O/bench_full.json') if l.startswith('{')][-1]); print('features+criterion+optimizer', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
tail -3 $O/bench_full.err
