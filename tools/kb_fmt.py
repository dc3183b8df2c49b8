import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        print(l.rstrip()); continue
    print(d["kernel"].ljust(40), "ms %.3f best %.3f GB/s %.0f frac %.3f" % (d["ms_mean"], d["ms_best"], d["achieved_gbs"], d.get("frac", 0)))
