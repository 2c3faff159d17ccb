import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
lines = [l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
r = d.get("roofline") or {}
e = d.get("e2e") or {}
print(tag, "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms %s" % r.get("kernel_ms"),
      "frac %s" % r.get("frac"), "e2e %s" % e.get("value"), r.get("kernel"))
