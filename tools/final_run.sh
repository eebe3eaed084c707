#!/bin/bash
# One gpurun call that produces every artefact copied into profiles/ at the end of a round.
# usage (on the GPU box, from the repo root): bash tools/final_run.sh <tag>
T=${1:-fin}
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; tail -1 $O/${T}_smoke.log
python bench.py --impl reference > $O/${T}_bench_c2_reference.json 2> $O/${T}_bench_c2_reference.err
for c in c2 c3 c4 f1; do
  python bench.py --config $c > $O/${T}_bench_$c.json 2> $O/${T}_bench_$c.err
done
python tools/bench_line.py $O/${T}_bench_c2.json $O/${T}_bench_c3.json $O/${T}_bench_c4.json $O/${T}_bench_f1.json
# launch list of the default bench command (per-launch device times are cold-cache and serialised: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spz_ -s 16 -c 40 --csv \
    --log-file $O/${T}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/${T}_launch_bench.log 2>&1
timeout 300 python tools/dims_sweep.py > $O/${T}_dims_sweep.jsonl 2> $O/${T}_dims_sweep.err; wc -l $O/${T}_dims_sweep.jsonl
