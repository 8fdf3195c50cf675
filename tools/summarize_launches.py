"""Turns an ncu launch list (CSV with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch)
into profiles/<tag>_launch_summary.txt and profiles/<round>_kernel_traffic.json (<round> = the tag up to its first "_", e.g.
r2_step -> r2).   python tools/summarize_launches.py <csv> <tag>"""
import collections, csv, json, os, re, sys
src, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = [l for l in open(src) if not l.startswith("==")]
launch = collections.OrderedDict()
MUL = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in csv.DictReader(lines):
    d = launch.setdefault(row["ID"], dict(name=row["Kernel Name"], grid=row["Grid Size"], us=0.0, rd=0.0, wr=0.0))
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]; m = row["Metric Name"]
    if m == "gpu__time_duration.sum": d["us"] = v / 1e3 if u == "ns" else (v if u == "us" else v * 1e3)
    elif m.startswith("dram__bytes_read"): d["rd"] = v * MUL[u]
    elif m.startswith("dram__bytes_write"): d["wr"] = v * MUL[u]
def short(n): return re.sub(r"void |pu::|\(.*", "", n)[:64]
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(short(d["name"]), dict(n=0, us=0.0, bytes=0.0)); a["n"] += 1; a["us"] += d["us"]; a["bytes"] += d["rd"] + d["wr"]
tot = sum(d["us"] for d in launch.values())
out = [f"# ncu launch list of ONE training step (batch 4 x 180k BraTS-shaped clouds), {len(launch)} launches, {tot/1e3:.2f} ms of kernel time",
       "# (cold-cache, serialised durations: compare SHARES, not absolutes)", f"{'share':>6} {'ms':>8} {'n':>5} {'GB/s(dram)':>10}  kernel"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    out.append(f"{100*a['us']/tot:6.2f} {a['us']/1e3:8.3f} {a['n']:5d} {a['bytes']/max(a['us'],1e-9)/1e3:10.0f}  {k}")
out.append("\n# 25 longest individual launches")
for d in sorted(launch.values(), key=lambda d: -d["us"])[:25]:
    out.append(f"{d['us']:9.1f} us  {(d['rd']+d['wr'])/d['us']/1e3:7.0f} GB/s  rd {d['rd']/1e6:8.1f} MB  wr {d['wr']/1e6:8.1f} MB  {short(d['name'])} {d['grid']}")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{tag}_launch_summary.txt"), "w").write("\n".join(out) + "\n")
# per C-ABI entry traffic (bytes per launch), mapping kernel names to the entry points bench.py times
ENTRY = [(r"tc::tc_persist_kernel<\d+, 0,", "pu_tc_linear_fwd"), (r"tc::tc_linear_kernel", "pu_tc_linear_fwd"),
         (r"tc::tc_persist_kernel<\d+, 1,", "pu_tc_att_pooling_fwd"), (r"tc::tc_persist_kernel<\d+, 2,", "pu_tc_att_pooling_bwd"),
         (r"tc::tc_persist_kernel<\d+, 3,", "pu_tc_att_pooling_bwd_fused"),
         (r"tc::tc_wgrad_kernel|tc::tc_wgrad_ts_kernel", "pu_tc_wgrad"), (r"mlp::bn_act_fwd", "pu_bn_act_fwd"),
         (r"mlp::bn_bwd_reduce", "pu_bn_bwd_reduce"), (r"mlp::bn_bwd_apply", "pu_bn_bwd_apply"),
         (r"lfa::maxpool_bwd", "pu_random_sample_bwd"), (r"lfa::maxpool_fwd", "pu_random_sample_fwd"), (r"lfa::gather_rows", "pu_gather_rows_fwd"), (r"lfa::segment_sum", "pu_segment_sum"),
         (r"mlp::wgrad", "pu_wgrad"), (r"mlp::linear_narrow|mlp::gemm_kernel<\d+, \d+, \d+, \d+, 0>", "pu_linear_fwd"),
         (r"knn::knn_search_kernel|knn::knn_query_warp_kernel", "pu_knn_batch"), (r"locse::locse_mlp_fwd", "pu_locse_mlp_fwd"),
         (r"locse::locse_mlp_bwd_kernel|locse::locse_mlp_bwd_direct", "pu_locse_mlp_bwd"), (r"locse::locse_moment", "pu_locse_moments"), (r"att16::att16_fwd_kernel", "pu_att16_fwd"),
         (r"att16::att16_bwd_kernel", "pu_att16_bwd")]
tr = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k, a in agg.items():
    for pat, entry in ENTRY:
        if re.match(pat, k):
            tr[entry][0] += a["n"]; tr[entry][1] += a["bytes"]; tr[entry][2] += a["us"]
            break
json.dump({e: dict(dram_bytes_per_launch=v[1] / v[0], launches_per_step=v[0], ncu_ms_per_step=v[2] / 1e3,
                   dram_gbs_under_ncu=v[1] / max(v[2], 1e-9) / 1e3) for e, v in tr.items() if v[0]},
          open(os.path.join(ROOT, "profiles", tag.split("_")[0] + "_kernel_traffic.json"), "w"), indent=1)
print("\n".join(out[:30]))
