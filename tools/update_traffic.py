"""profiles/ncu_traffic.json from the committed ncu summaries of the kernels bench.py times.
Usage: python tools/update_traffic.py r2f      (reads profiles/<prefix>_tree_config{2,3,4,5}_ncu_summary.txt)"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1] if len(sys.argv) > 1 else "r2f"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def api_name(ncu_name):
    """ncu prints 'void tree_kernel_s4<4, 1, 4, 0>(TreeParams)'; bppgpu_batch_kernel_name says
    'tree_kernel_s4<4,true,4,all_paths>'."""
    m = re.search(r"(tree_kernel_\w+)<([^>]*)>", ncu_name)
    name, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
    tf = lambda a: "true" if a in ("1", "true") else "false"
    if name == "tree_kernel_s4":
        return "%s<%s,%s,%s,%s>" % (name, args[0], tf(args[1]), args[2], "scaled_only" if args[3] in ("1", "true") else "all_paths")
    if name == "tree_kernel_s20t":
        return "%s<%s,%s>" % (name, args[0], tf(args[1]))
    return "%s<%s>" % (name, ",".join(args))


out = {"_comment": "dram__bytes_read.sum / dram__bytes_write.sum per launch of the dominant (tree) kernel, from the ncu "
                   "--set full capture named in source, at the config's full size, scaling off. bench.py reports it as "
                   "roofline.traffic only while the kernel instantiation it times carries the same name. "
                   "Written by tools/update_traffic.py."}
for cfg in ("config2", "config3", "config4", "config5"):
    path = os.path.join("profiles", "%s_tree_%s_ncu_summary.txt" % (prefix, cfg))
    txt = open(os.path.join(ROOT, path)).read()
    kern = re.search(r"kernels: \['([^']*)'", txt).group(1)
    vals = {}
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(re.escape(key) + r"\s+(\w+)\s+\['([0-9.]+)'\]", txt)
        vals[key] = float(m.group(2)) * UNIT[m.group(1)]
    out[cfg] = {"kernel": api_name(kern), "dram_bytes_read": vals["dram__bytes_read.sum"],
                "dram_bytes_write": vals["dram__bytes_write.sum"], "source": path}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=2)
print(json.dumps(out, indent=2))
