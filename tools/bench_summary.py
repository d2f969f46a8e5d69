import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), round(d["roofline"]["frac"],3)); print(d["kernel_ms_per_step"])
