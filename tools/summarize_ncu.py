"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) into the markdown summary that is
committed under profiles/.  Usage: python tools/summarize_ncu.py launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        rows.append((r["Kernel Name"], r["Grid Size"], r["Block Size"], v))
    tot = sum(v for *_, v in rows)
    print("# ncu launch list summary: one training step (configs[1], bs 16, 256x256, classic, bf16, eager launches)\n")
    print("`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` around "
          "`tools/profile_step.py` (cold-cache, serialised launches: compare shares, not absolutes).\n")
    print("launches: %d, summed duration: %.1f ms\n" % (len(rows), tot / 1e3))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, b, v in rows:
        m = re.search(r"(conv_\w+_kernel<[^>]*>|in_\w+_kernel<[^>]*>|pack_kernel<[^>]*>|\w+_kernel)", n)
        ours = "sscg::" in n or (m and not n.startswith("void at::"))
        key = m.group(1) if m else re.sub(r"\(.*", "", n)[:60]
        agg[("ours" if "sscg" in n or "conv_" in key or key.startswith(("in_", "pack", "unpack", "wprep", "wgrad", "bias"))
             else "torch", key)][0] += 1
        agg[("ours" if "sscg" in n or "conv_" in key or key.startswith(("in_", "pack", "unpack", "wprep", "wgrad", "bias"))
             else "torch", key)][1] += v
    print("| share | total ms | launches | avg us | origin | kernel |\n|---:|---:|---:|---:|---|---|")
    for (org, k), (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("| %.1f%% | %.2f | %d | %.1f | %s | `%s` |" % (100 * v / tot, v / 1e3, c, v / c, org, k))
    ours = sum(v for (o, _), (_, v) in agg.items() if o == "ours")
    print("\nthis repo's kernels: %.1f%% of the summed launch time; torch glue (losses, softmax, one-hot, Adam, "
          "fills): %.1f%%" % (100 * ours / tot, 100 * (tot - ours) / tot))


if __name__ == "__main__":
    main(sys.argv[1])
