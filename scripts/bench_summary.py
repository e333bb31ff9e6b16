"""Print the headline numbers of a bench.py JSON line (gpurun_out/*.json)."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
print("value %.1f f/s  ms/step %.3f  regions %s" % (d["value"], d["ms_per_step"], [round(x, 2) for x in d.get("timed_regions", {}).get("ms", [])]))
if d.get("e2e"): print("e2e %.1f f/s  h2d GB/s/rank %.1f" % (d["e2e"]["value"], d["e2e"].get("h2d_gbs_per_rank", 0)))
c = d["config"]; print("single clip ms", c.get("single_clip_in_flight_ms_per_step"), "launches", d.get("gpu_launches"), "kept", c.get("kept_slots"))
r = d.get("roofline") or {}
print("attention_contraction", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in (r.get("attention_contraction") or {}).items()})
hs = r.get("hbm_stages", {})
for k in ("mask_logits", "panoptic_fusion", "level_fusion"):
    if k in hs: print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in hs[k].items() if a not in ("kernels",)})
ks = r.get("kernels", {})
tot = sum(v["ms_per_step"] for v in ks.values())
print("serialised kernel ms/step %.3f" % tot)
for k, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms_per_step"])[:16]:
    print("  %-24s %.4f ms  x%-5.1f %-8s frac=%s grid=%s" % (k, v["ms_per_step"], v["launches_per_step"], v["bound"], None if v.get("frac") is None else round(v["frac"], 3), v.get("grid_ctas")))
