import json, sys
d = json.load(open(sys.argv[1]))
print("it/s %.2f  ms/step %.2f  kkt_ms %.2f  iter_frac %.3f  launches/it %.1f" % (
    d["value"], d["ms_per_step"], d["kkt_solve_ms"], d["iter_roofline_frac"], d["gpu_launches"] / d["steps"]))
for k, v in d["roofline"]["kernels"].items():
    print("  %-14s %8.3f ms/it  launches/it %.1f" % (k, v["ms"] / d["steps"], v["launches"] / d["steps"]))
