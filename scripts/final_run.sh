#!/bin/bash
# End-of-round measurement on a GPU box: full -m gpu suite, reference-host frame times, single-call times, both bench arms.
tag=${1:-rXX}
out=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_gputest.txt
python scripts/run_ref_host.py 376 1241 24 2>/dev/null | python -c "
import sys, json
recs=[json.loads(l) for l in sys.stdin if l.startswith('{') and 'frame' in l]
ms=[r.get('ms') for r in recs if r.get('ms') is not None]
print('ref host ms per frame:', [round(m,2) for m in ms])
" | tee $out/${tag}_ref_host.txt
python scripts/time_single_calls.py 2>&1 | tail -11 | tee $out/${tag}_single_calls.txt
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench.err; tail -2 $out/${tag}_bench.err
python bench.py --impl reference --steps 20 > $out/${tag}_bench_reference.json 2>/dev/null
python - <<PY
import json
d=[json.loads(l) for l in open("$out/${tag}_bench_n1.json") if l.startswith("{")][0]
print("value", d["value"], "e2e", d["e2e"]["value"], "rec", d["e2e"]["records_only"]["value"], "unp", d["unpipelined_value"], d["e2e"]["unpipelined_value"])
print("roof", d["roofline"]["frac"], d["roofline"]["issue_frac"], "lk", d["roofline_lk"]["issue_frac"], d["roofline_lk"]["launch_ms"], "knn", d["roofline_matching"]["frac"], d["roofline_matching"]["stage_ms"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["latency_mode"]["value"], "track", d["tracking_frame"]["ms_per_step"], "single", d["single_stream"]["ms_per_frame_e2e"], d["single_stream"]["ms_per_frame_e2e_plain_launches"])
print("steps", d["steps"], d["timed_steps"], d["ms_per_step"], d["clocks"])
r=[json.loads(l) for l in open("$out/${tag}_bench_reference.json") if l.startswith("{")][0]
print("ref", r["value"], r["cpu_baseline"]["latency_mode"]["value"])
PY
