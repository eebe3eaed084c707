python bench.py --config c4 --no-sub --no-cpu-baseline --e2e-steps 1 > gpurun_out/w13_c4.json 2>gpurun_out/w13.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/w13_c4.json").read().strip().splitlines()[-1])
print(d["encode_gbs"],d["decode_gbs"],d["roofline_encode"]["frac"],d["roofline_decode"]["frac"],d["roofline_encode"]["ms_per_launch"],d["roofline_decode"]["ms_per_launch"],d["clocks"])
PY
