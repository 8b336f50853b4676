"""Print the one-line bench JSON as a short table (developer convenience)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    pc = r.get("per_class", {})
    print("value %.0f Msamples/s  %.2f ms/step  e2e %.0f  launches %s  clocks %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), d.get("clocks", {}).get("sm_mhz")))
    for k, v in pc.items():
        print("   %-10s %7.2f ms  %6.0f launches  %6.2f GB  %7.0f GB/s" % (
            k, v["ms_per_step"], v["launches_per_step"], v["algorithmic_GB_per_step"], v["GBps"] or 0))
    ws = r.get("whole_step", {})
    print("   whole step: %.1f GB, %.1f B/segment, %.2f segments/sample, frac of peak %.3f" % (
        ws.get("algorithmic_GB", 0), ws.get("bytes_per_segment", 0), ws.get("segments_per_sample", 0),
        ws.get("frac_of_peak_per_gpu", 0)))
