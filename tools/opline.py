import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["metric"], round(d["ms_per_step"],4), "%.4g" % d["value"], d["roofline"].get("frac"))
