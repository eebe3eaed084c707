"""Print the headline numbers of bench.py JSON lines. usage: bench_line.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        rd, re = d.get("roofline_decode", {}), d.get("roofline_encode", {})
        print(f"{f}: enc {d['encode_gbs']:.0f} dec {d['decode_gbs']:.0f} GB/s  frac enc {re.get('frac', 0):.3f} dec {rd.get('frac', 0):.3f}"
              f"  ratio {d['config'].get('ratio', 0):.4f} chunk_rows {d['config'].get('chunk_rows')} e2e {d.get('e2e', {}).get('value')}")
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
