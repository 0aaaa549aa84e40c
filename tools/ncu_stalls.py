"""Per-kernel digest of an ncu --set full capture's SOURCE page (no GPU needed): where the warps' time goes by stall reason,
the instruction mix, and the copy-engine / mbarrier instructions with their execution counts.
   python tools/ncu_stalls.py <prof.ncu-rep> <kernel regex> [<frames per launch>] [<gathered slots per frame>]"""
import collections, csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 100
slots = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
name = rows[0][1] if rows and len(rows[0]) > 1 else kre
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
# (the page is repeated once per captured launch: keep the first)
first = data[0][hdr.index("Address")]
for k in range(1, len(data)):
    if data[k][hdr.index("Address")] == first:
        data = data[:k]
        break
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
tot = sum(num(r[isamp]) for r in data)
inst = sum(num(r[iex]) for r in data)
print(f"## `{name[:110]}`\n")
print(f"warp-level instructions executed: {inst:,} per launch of {frames} frames" + (f" = {inst / frames / slots:.1f} per gathered slot" if slots else ""))
print(f"code: {len(data)} SASS instructions = {len(data) * 16 / 1024:.0f} KB\n")
agg = collections.Counter()
for r in data:
    for i in stall: agg[hdr[i][6:]] += num(r[i])
print("| warp state (sampled, all warps) | share |\n|---|---|")
for k, v in agg.most_common(10): print(f"| {k} | {100 * v / tot:.1f} % |")
op, opx = collections.Counter(), collections.Counter()
for r in data:
    t = [x for x in r[isrc].split() if not x.startswith("@")]
    o = t[0] if t else "?"
    op[o] += num(r[isamp]); opx[o] += num(r[iex])
print("\n| opcode | executed | per slot | samples |\n|---|---|---|---|")
for o, v in opx.most_common(22):
    print(f"| {o} | {v:,} | " + (f"{v / frames / slots:.2f}" if slots else "") + f" | {100 * op[o] / tot:.1f} % |")
print("\ncopy-engine / mbarrier instructions (address, executed, samples):\n```")
for r in data:
    if any(x in r[isrc] for x in ("UTMALDG", "UBLKCP", "SYNCS")):
        print(f"{r[hdr.index('Address')][-5:]}  {num(r[iex]):>10,}  {num(r[isamp]):>6}  {r[isrc].strip()[:90]}")
print("```")
