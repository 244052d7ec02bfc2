import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{"metric"')][-1])
print(round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, [(r["kernel"], round(r["ms"],2)) for r in d["roofline_all"]], d["lm"]["final_error"])
