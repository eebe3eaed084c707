#!/bin/bash
# One gpurun call that produces the artefacts copied into profiles/ at the end of a round (1 GPU).
# usage (on the GPU box, from the repo root): bash tools/final_run.sh <tag>
T=${1:-fin}
O=gpurun_out
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -2 $O/${T}_bench.err
python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
# ncu --set full of the dominant codec kernels of every config (traffic, pipes, stalls): one encode + one decode launch each
for c in c3 c2 c4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spz_(en|de)code_(lane|wide)" -s 2 -c 2 -f \
      -o $O/${T}_ncu_$c python bench.py --config $c --no-sub --no-cpu-baseline --e2e-steps 1 --steps 1 --warmup 3 > $O/${T}_ncu_$c.log 2>&1
done
# launch list of the default bench command (per-launch device times are cold-cache and serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spz_ -c 120 --csv \
    --log-file $O/${T}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-sub > $O/${T}_launch_bench.log 2>&1
ls -la $O | grep ${T}_
timeout 600 python tools/dims_sweep.py --blocks 8 > $O/${T}_dims_sweep_block8.jsonl 2> $O/${T}_dims_sweep.err; wc -l $O/${T}_dims_sweep_block8.jsonl
