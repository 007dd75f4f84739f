"""profiles/k2_issue.json from an ncu launch list of tools/sweep_blocks.py (block coder kernels only):
    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:"k2_model|k2_code|k2_coder|k2_block" --csv --log-file L.csv python tools/sweep_blocks.py 1 auto 1
usage: k2_issue.py L.csv n_symbols [out.json]      (whole encode + decode passes are found from the largest decode launch)
Warp instructions per coded symbol and the time-weighted issue-slot utilisation of the coder's launches (all generations),
per direction; bench.py applies the per-symbol figure to the run it times (roofline_kernels, bound "issue")."""
import csv, json, sys
from collections import defaultdict

path, n_sym = sys.argv[1], float(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/k2_issue.json"
d = defaultdict(dict)
for r in csv.reader(open(path)):
    if len(r) > 10 and r[0].isdigit():
        i = int(r[0]); d[i]["k"] = r[4]; d[i]["grid"] = int(r[8].strip("()").split(",")[0]); d[i][r[-3]] = float(r[-1].replace(",", ""))
def is_dec(k):
    """k2_coder_kernel<MODE, LEGACY> and k2_block_kernel<MODE>: MODE 1 is the decoder. The model and the interval kernel are
    the two halves of the encoder whatever their template arguments say (k2_model_kernel<BY_CONTEXT>)."""
    if "k2_model_kernel" in k or "k2_code_kernel" in k:
        return False
    return "<(int)1" in k or "<1" in k
big = max(v["grid"] for v in d.values() if is_dec(v["k"]))
ends = [i for i in sorted(d) if is_dec(d[i]["k"]) and d[i]["grid"] == big]      # a pass ends with the decoder's last generation
iters = float(len(ends))
d = {i: v for i, v in d.items() if i <= ends[-1]}
agg = {"encode": defaultdict(float), "decode": defaultdict(float)}
per_kernel = defaultdict(lambda: defaultdict(float))
for i in sorted(d):
    k = d[i]["k"]
    dec = is_dec(k)
    side = "decode" if dec else "encode"
    t = d[i].get("gpu__time_duration.sum", 0.0); ins = d[i].get("smsp__inst_executed.sum", 0.0)
    ia = d[i].get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
    a = agg[side]; a["ns"] += t; a["inst"] += ins; a["ia_ns"] += ia * t; a["launches"] += 1
    pk = per_kernel[side + ": " + k.split("(")[0].replace("void ", "")]; pk["ns"] += t / iters; pk["inst"] += ins / iters; pk["launches"] += 1 / iters
res = {}
for side, a in agg.items():
    if not a["ns"]:
        continue
    res[side] = {"warp_inst_per_symbol": round(a["inst"] / iters / n_sym, 2), "issue_active_pct": round(a["ia_ns"] / a["ns"], 1),
                 "launches_per_pass": a["launches"] / iters, "ncu_ms_per_pass": round(a["ns"] / iters / 1e6, 3),
                 "source": f"{path}: ncu launch list of tools/sweep_blocks.py, config 2, {int(n_sym)} symbols per pass"}
res["kernels"] = {k: {"ms": round(v["ns"] / 1e6, 3), "warp_inst_M": round(v["inst"] / 1e6, 1), "launches": v["launches"]} for k, v in per_kernel.items()}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
