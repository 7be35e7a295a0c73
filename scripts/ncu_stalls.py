"""Per-segment warp-stall breakdown of an ncu report's source page (segments end at branches/barriers).
Usage: python scripts/ncu_stalls.py report.ncu-rep"""
import csv, subprocess, sys

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    tables, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            tables.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for t in tables:
        hdr, data = t["hdr"], t["data"]
        ix = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        total = sum(int(r[ix["# Samples"]]) for r in data) or 1
        print("##", t["name"][:110])
        seg, c = [], {"n": 0, "ex": 0, "st": {s: 0 for s in stalls}, "start": 0}
        for k, r in enumerate(data):
            src = r[ix["Source"]].strip()
            c["n"] += int(r[ix["# Samples"]]); c["ex"] += int(r[ix["Instructions Executed"]])
            for s in stalls: c["st"][s] += int(r[ix[s]])
            if any(x in src for x in ("BAR.SYNC", "BRA", "EXIT", "WARPSYNC")):
                c["end"] = k; c["mark"] = src[:44]; seg.append(c)
                c = {"n": 0, "ex": 0, "st": {s: 0 for s in stalls}, "start": k + 1}
        seg.append(c)
        tot = {s: sum(x["st"][s] for x in seg) for s in stalls}
        print("  samples", total, {k[6:]: round(v / total, 3) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]})
        for x in seg:
            if x["n"] < total * 0.004: continue
            top = sorted(x["st"].items(), key=lambda y: -y[1])[:5]
            print(f"  instr {x['start']:5d}-{x.get('end', 0):5d} samples {x['n']:6d} ({x['n']/total:.3f}) exec {x['ex']/1e6:8.1f}M"
                  f" s/Mexec {x['n']/max(x['ex'],1)*1e6:6.1f} | " + ", ".join(f"{a[6:]} {b/max(x['n'],1):.2f}" for a, b in top), "|", x.get("mark", ""))

if __name__ == "__main__":
    for p in sys.argv[1:]: main(p)
