import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print(d["config"]["config_name"], d.get("math"), "value %.4g e2e %.4g"%(d["value"], d["e2e"]["value"]), d["roofline"]["per_kernel_ms"], "frac %.3f fc %s"%(d["roofline"]["frac"], d["roofline"].get("frac_compulsory")))
