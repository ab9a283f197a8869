"""Side-by-side per-op device times (ms per launch) of several bench.py --profile-out files."""
import json, sys
files = sys.argv[1:]
P = [json.load(open(f)) for f in files]
print("%-34s" % "op", " ".join("%9s" % f.split("/")[-1].replace("dbg_", "").replace("prof_", "").replace(".json", "")[:9] for f in files))
for w in ("detector", "recogniser"):
    for i, o in enumerate(P[0]["ops"][w]):
        name = "%s%-2d %s %dx%d %d->%d k%d s%d" % (w[0], o["index"], "conv" if o["kind"] == 0 else "pool", o["H"], o["W"], o["Cin"], o["Cout"], o["KH"], o["stride"])
        print("%-34s" % name[:34], " ".join("%9.3f" % (p["ops"][w][i]["ms"] / max(p["ops"][w][i]["launches"], 1)) for p in P))
for i, o in enumerate(P[0]["ops"].get("stages", [])):
    print("%-34s" % ("stage " + o["name"]), " ".join(("%9.3f" % (p["ops"]["stages"][i]["ms"] / max(p["ops"]["stages"][i]["launches"], 1))) if "stages" in p["ops"] else "%9s" % "-" for p in P))
print("%-34s" % "step total", " ".join("%9.3f" % (p["ms_total"] / p["steps"]) for p in P))
