"""Short view of a bench.py line: python tests/tools/show_bench.py <file>"""
import json
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.strip().startswith("{")]
if not lines:
    print("no JSON line")
    sys.exit(0)
j = json.loads(lines[-1])
r = j.get("roofline", {})
print({"ms": round(j["ms_per_step"], 4), "value": round(j["value"]), "e2e_ms": round(j["e2e"]["ms_per_step"], 4),
       "serial_ms": round(j["e2e"]["serial_ms_per_step"], 4), "launches": j["gpu_launches"],
       "host_enqueue_ms": round(j.get("host_enqueue_ms_per_step", 0), 3), "single": j.get("single_sweep_forward"),
       "clk": (j.get("clocks") or {}).get("sm_mhz")})
print("parity:", json.dumps(j.get("parity")))
print("roofline:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ("achieved", "frac", "ms_per_launch", "share_of_step")},
      "step:", {k: round(v, 4) for k, v in r.get("whole_step", {}).items() if isinstance(v, float) and v < 1e6})
for name, k in sorted(r.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"   {name:28s} x{k['launches_per_step']:5.1f}  {k['ms_per_launch'] * 1e3:9.1f} us  {k['ms_per_step']:8.4f} ms/step  {100 * k['share_of_profiled_step']:5.1f} %")
if "nvlink" in r:
    print("nvlink:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r["nvlink"].items()})
print("eager:", j.get("gpu_eager_baseline"))
print("extras:", json.dumps(j.get("extras")))
print("cpu:", json.dumps(j.get("cpu_baseline")))
