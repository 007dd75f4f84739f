"""Turn the outputs of tools/profile_round.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/:
bench lines per named config, the reference arm's line, the ncu launch list with per-kernel shares, ncu --set full
summaries of the hot kernels, DRAM traffic per launch (traffic.json, read by bench.py) and the block coder's
instructions per symbol (k2_issue.json, read by bench.py).   usage: make_profiles.py [tag=r02f] [round=r02]"""
import csv, io, json, os, shutil, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02f"
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"

def first_json(path):
    for line in open(path):
        if line.startswith("{"):
            return json.loads(line)

for c in (1, 2, 3, 5):
    src = os.path.join(G, f"{tag}_bench_config{c}.json")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{rnd}_bench_config{c}.json"))
shutil.copy(os.path.join(G, f"{tag}_bench_reference.json"), os.path.join(P, f"{rnd}_bench_reference_arm.json"))
shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{rnd}_launches.csv"))
shutil.copy(os.path.join(G, f"{tag}_k2_launches.csv"), os.path.join(P, f"{rnd}_k2_launches.csv"))
bench = first_json(os.path.join(P, f"{rnd}_bench_config2.json"))
n_sym = bench["symbols_per_s_encode"] * bench["stage_ms"]["k2e"] * 1e-3
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "k2_issue.py"), os.path.join("profiles", f"{rnd}_k2_launches.csv"), str(round(n_sym)),
                os.path.join("profiles", "k2_issue.json")], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)

out = [f"# Round 2: ncu summaries of the final build (B200, config 2: 3 014 484 reads x 150 bp, automatic block size, default layout)", "",
       "Commands profiled: `python bench.py --config 2 --steps 1 --warmup 1 --no-cpu` (launch list) and `python tools/sweep_blocks.py 1 auto 1`",
       "(one resident encode + decode per pass, one stream per block in every generation; per-kernel figures and the full captures); every",
       "ncu pass ran only after the same command had exited 0 without ncu (tools/profile_round.sh).", "",
       f"## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, profiles/{rnd}_launches.csv; cold-cache, serialised)", ""]
rows = [r for r in csv.reader(open(os.path.join(P, f"{rnd}_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r[4].split("(")[0].replace("void ", ""); agg[k][0] += 1; agg[k][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
out += ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
out += [f"| {k} | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |" for k, v in sorted(agg.items(), key=lambda x: -x[1][1])]
sm = bench["stage_ms"]
coder = sum(v[1] for k, v in agg.items() if "k2_" in k or "merge" in k or "snapshot" in k)
out += ["", f"bench.py's CUDA-event stage times for the same build (profiles/{rnd}_bench_config2.json): K2 encode {sm['k2e']:.2f} ms, K2 decode {sm['k2d']:.2f} ms, "
        f"K1 {sm['k1']:.2f} ms, K3 {sm['k3']:.2f} ms of a {bench.get('single_batch', bench)['ms_per_step']:.1f} ms round trip of one batch with the device to itself: the block coder (with its generation merges) is "
        f"{100 * (sm['k2e'] + sm['k2d']) / bench.get('single_batch', bench)['ms_per_step']:.0f} % of the step there and {100 * coder / tot:.0f} % of the ncu launch list "
        "(which also holds the K1-alone encodes and the pipelined host-buffer steps of the bench command).", ""]
ki = json.load(open(os.path.join(P, "k2_issue.json")))
out += [f"## Block coder per pass (profiles/{rnd}_k2_launches.csv -> profiles/k2_issue.json)", "",
        "| direction | warp instructions per symbol | issue-slot utilisation (time-weighted) | launches | ncu ms |", "|---|---|---|---|---|"]
out += [f"| {d} | {ki[d]['warp_inst_per_symbol']} | {ki[d]['issue_active_pct']} % | {ki[d]['launches_per_pass']:.0f} | {ki[d]['ncu_ms_per_pass']} |" for d in ("encode", "decode")]
out += ["", "| kernel (per pass) | ms | M warp instructions |", "|---|---|---|"]
out += [f"| {k} | {v['ms']} | {v['warp_inst_M']} |" for k, v in ki["kernels"].items()]
out += [""]
traffic = {}
def val(r, idx, units, m):
    return float(r[idx[m]]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[m]]]
for f, names in ((f"{tag}_k1k3", {"k1_extract": "k1_extract_kernel", "k3_reconstruct": "k3_reconstruct_kernel"}),
                 (f"{tag}_k2enc", {"k2_model": "k2_model_kernel", "k2_code": "k2_code_kernel"}), (f"{tag}_k2dec", {"k2_coder": "k2_coder_kernel<decode>"})):
    rep = os.path.join(G, f + ".ncu-rep")
    out += [f"## `ncu --set full --clock-control none --import-source on`: gpurun_out/{f}.ncu-rep", "```"]
    out += subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout.rstrip().split("\n")
    out += ["```", ""]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw))); hdr, units = rr[0], rr[1]; idx = {h: i for i, h in enumerate(hdr)}
    for r in rr[2:]:
        for pat, key in names.items():
            if pat in r[idx["Kernel Name"]] and key not in traffic:
                traffic[key] = int(val(r, idx, units, "dram__bytes_read.sum") + val(r, idx, units, "dram__bytes_write.sum"))
traffic["k2 block coder (encode)"] = traffic.get("k2_model_kernel", 0) + traffic.get("k2_code_kernel", 0)
traffic["k2 block coder (decode)"] = traffic.get("k2_coder_kernel<decode>", 0)
traffic["_reads_per_launch"] = 3014484
traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, config 2; K1: the launch over the tail of the batch "
                    "(2 789 k of the 3 014 k reads); the block coder entries are the last generation's launches (2 958 blocks, 92 % of the reads): "
                    "model kernel + interval kernel for encode")
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
open(os.path.join(P, f"{rnd}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print(json.dumps(traffic, indent=1))
