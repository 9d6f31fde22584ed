"""Per-instruction view of an .ncu-rep (source page, SASS): top shared-memory bank-conflict sites, stall sample totals by reason,
instruction mix.  usage: python scripts/ncu_source_hot.py rep.ncu-rep [kernel_index]"""
import csv, subprocess, sys, io, collections

def main(path, which=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    # split per kernel: blocks start with a "Kernel Name" line
    blocks, cur = [], []
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            if cur: blocks.append(cur)
            cur = []
        else:
            cur.append(line)
    if cur: blocks.append(cur)
    rows = list(csv.reader(blocks[which]))
    hdr, data = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    f = lambda r, h: float(r[ix[h]]) if r[ix[h]] not in ("", "-") else 0.0
    tot_samples = sum(f(r, "# Samples") for r in data)
    print(f"kernel {which}: {len(data)} SASS instructions, {tot_samples:.0f} samples")
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(f(r, h) for r in data) for h in reasons}
    print("stall samples by reason:", ", ".join(f"{h[6:]}={v / tot_samples:.3f}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
    mix = collections.Counter()
    for r in data:
        op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
        if op.startswith("@"): op = r[ix["Source"]].split()[1]
        mix[op.split(".")[0]] += f(r, "Instructions Executed")
    tot_inst = sum(mix.values())
    print("instruction mix:", ", ".join(f"{k}={v / tot_inst:.3f}" for k, v in mix.most_common(14)))
    print("shared-memory wavefronts: total", sum(f(r, "L1 Wavefronts Shared") for r in data), "ideal", sum(f(r, "L1 Wavefronts Shared Ideal") for r in data))
    sites = sorted(data, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:14]
    for r in sites:
        if f(r, "L1 Wavefronts Shared Excessive") == 0: break
        print(f"  excess {f(r, 'L1 Wavefronts Shared Excessive'):12.0f} of {f(r, 'L1 Wavefronts Shared'):12.0f}  execs {f(r, 'Instructions Executed'):10.0f}  {r[ix['Source']].strip()[:70]}")
    hot = sorted(data, key=lambda r: -f(r, "# Samples"))[:12]
    print("hottest instructions by samples:")
    for r in hot:
        print(f"  {f(r, '# Samples'):8.0f}  {r[ix['Source']].strip()[:90]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
