"""One-line digest of a bench.py JSON line (file argument)."""
import json, sys
for line in open(sys.argv[1]):
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    ex = d.get("extras", {})
    out = {"n_gpus": d["n_gpus"], "ms_per_step": d["ms_per_step"], "value": d["value"], "launches": d.get("gpu_launches"),
           "frac": d.get("roofline", {}).get("frac"), "long": d.get("long_run", {}).get("ms_per_step")}
    for k, f in (("burgers_lbfgs", "ms_per_iteration"), ("burgers_cfg1_10k", "adam_ms_per_step"), ("burgers_cfg1_10k", "lbfgs_ms_per_iteration"),
                 ("burgers_identification", "ms_per_step"), ("burgers_identification", "lbfgs_ms_per_iteration"), ("schrodinger", "ms_per_step"),
                 ("burgers_8x40_generic", "ms_per_step"), ("burgers_discrete_time", "ms_per_step")):
        if k in ex and f in ex[k]:
            out[k + "." + f] = ex[k][f]
    if "parity_check" in d:
        out["parity_ok"] = d["parity_check"].get("ok")
    if "cfg5" in d:
        out["cfg5"] = [d["cfg5"].get("adam_ms_per_step"), d["cfg5"].get("lbfgs_ms_per_iteration")]
    print(json.dumps(out))
