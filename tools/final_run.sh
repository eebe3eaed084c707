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
# condense on the box: gpurun brings back at most 64 MiB, four reports are more
python tools/ncu_profile_csv.py $O/${T}_ncu_summary.csv $O/${T}_traffic.json c3=$O/${T}_ncu_c3.ncu-rep c2=$O/${T}_ncu_c2.ncu-rep c4=$O/${T}_ncu_c4.ncu-rep > /dev/null 2>&1
rm -f $O/${T}_ncu_c2.ncu-rep $O/${T}_ncu_c4.ncu-rep
# launch list of the default bench command (per-launch device times are cold-cache and serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spz_ -c 120 --csv \
    --log-file $O/${T}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-sub > $O/${T}_launch_bench.log 2>&1
ls -la $O | grep ${T}_
# BASELINE config 5, block 8: 1 GiB per point, 512-row chunks (bench geometry: enough chunks for every kernel family)
timeout 900 python tools/dims_sweep.py --blocks 8 --mb 1024 --chunk-rows 512 --dims 1,2,3,4,5,6,7,8,12,16,20,24,32,48,64,80,100,128,256 \
    > $O/${T}_dims_sweep_block8.jsonl 2> $O/${T}_dims_sweep.err; wc -l $O/${T}_dims_sweep_block8.jsonl
# the wide-row kernels of every (element width, predictor) at 80 dims
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spz_.*code_wide -f -o $O/${T}_ncu_wide80 \
    python tools/wide_probe.py > $O/${T}_ncu_wide80.log 2>&1; tail -2 $O/${T}_ncu_wide80.log
python tools/ncu_profile_csv.py $O/${T}_ncu_wide80_summary.csv $O/${T}_wide80_traffic.json wide80=$O/${T}_ncu_wide80.ncu-rep > /dev/null 2>&1
rm -f $O/${T}_ncu_wide80.ncu-rep; du -sh $O
