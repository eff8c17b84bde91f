"""Debug aid: distance (in instructions and in MUFU issues) between every MUFU.EX2 of a kernel and the first instruction that reads
its result.  ptxas tracks MUFU (variable latency) with a few scoreboards, so the consumer often sits 1-2 MUFUs behind the producer and
the warp stalls for the MUFU latency at every element; this prints the histogram.
    python scripts/sass_mufu_lag.py <object-or-cubin> <mangled-kernel-name-substring>"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = [m.group(1).strip() for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", b)]
    hist_i, hist_m = collections.Counter(), collections.Counter()
    n = 0
    for i, t in enumerate(ins):
        m = re.match(r"(@!?U?P\d+\s+)?MUFU\.EX2\s+(R\d+),", t)
        if not m:
            continue
        n += 1
        dst = m.group(2)
        mufus = 0
        for j in range(i + 1, min(i + 400, len(ins))):
            u = ins[j]
            ops = u.split(None, 1)[1] if " " in u else ""
            srcs = ops.split(",", 1)[1] if "," in ops else ""
            if re.search(rf"\b{dst}\b", srcs):
                hist_i[min(j - i, 40)] += 1
                hist_m[mufus] += 1
                break
            if "MUFU" in u:
                mufus += 1
    print(name[:80], "MUFU.EX2:", n)
    print(" consumer distance in MUFU issues:", dict(sorted(hist_m.items())))
    print(" consumer distance in instructions (capped 40):", dict(sorted(hist_i.items())))
