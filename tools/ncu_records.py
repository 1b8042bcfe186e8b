"""Turn the raw ncu outputs of tools/run_profiles.sh (gpurun_out/) into the tracked records under profiles/:
  r02_ncu_launches.csv / r02_ncu_launches_summary.txt  - launch list of the last forward (+ SI-SNRi), per-kernel shares
  ncu_traffic.json                                      - DRAM bytes per launch of the main kernels (bench.py reads it)
Run here (no GPU needed) after tools/ncu_summary.py has written the full-capture summaries."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "r02_launches.csv")
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
kn, val = hdr.index("Kernel Name"), hdr.index("Metric Value")
ours = [(re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("vatss::", "").replace("<unnamed>::", ""),
         float(r[val].replace(",", "")) / 1e3) for r in rows if "vatss" in r[kn] or re.search(r"\bk_[a-z]", r[kn])]
# the last forward: from the last k_visual_compress on
start = max(i for i, (k, _) in enumerate(ours) if k.startswith("k_visual_compress"))
last = ours[start:]
agg = collections.OrderedDict()
for k, us in last:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
shutil.copy(src, os.path.join(ROOT, "profiles", "r02_ncu_launches.csv"))
lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none, python tools/profile_forward.py (cfg-2: batch 32 x 4 s), last forward + SI-SNRi",
         "# (cold-cache, serialised: compare SHARES with bench.py's stage_ms, not absolutes)"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k:60s} {n:3d} launches {us / 1e3:8.3f} ms {100 * us / tot:5.1f}%  avg {us / n:8.1f} us")
lines.append(f"total {tot / 1e3:.3f} ms in {sum(v[0] for v in agg.values())} launches")
open(os.path.join(ROOT, "profiles", "r02_ncu_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

summ = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_full_forward_summary.json")))
def to_bytes(m):
    u = m["unit"].lower()
    return m["value"] * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
fam = {"attention": "k_tc_attn3", "lstm_recurrent": "k_tc_lstm_pp", "qkv": "k_tc_gemm<384, 128", "outproj_ln1": "k_tc_gemm<128, 128",
       "ffn_ln2": "k_tc_gemm<128, 256"}
kern = {}
for name, pat in fam.items():
    ls = [l for l in summ["launches"] if pat in l["kernel"]]
    if ls:
        kern[name] = {"kernel": pat, "dram_bytes_per_launch": int(sum(to_bytes(l["dram_read"]) + to_bytes(l["dram_write"]) for l in ls) / len(ls)),
                      "launches_captured": len(ls)}
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
json.dump({"source": "profiles/r02_ncu_full_forward_summary.json (ncu --set full --clock-control none, tools/run_profiles.sh: python tools/profile_forward.py, cfg-2 batch 32 x 4 s)",
           "commit": head + " (kernels as captured)",
           "metric": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches of the kernel",
           "kernels": kern}, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(kern, indent=1))
