import sys, json
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        r = d.get("roofline", {})
        print(tag, "value=%.0f ms=%.4f e2e=%.0f hop_us=%.2f frac=%.3f in_graph_us=%s graph_bracket=%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("avg_launch_us", 0), r.get("frac", 0), r.get("in_graph_us"), r.get("in_graph_bracketed_us")))
