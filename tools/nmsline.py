import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],4), round(d["value"]/1e6,2), {k:round(v,4) for k,v in d["roofline"].items() if k.startswith("ms_")})
