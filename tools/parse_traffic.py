#!/usr/bin/env python
"""ncu CSV (tools/gram_traffic.py) -> profiles/gram_traffic.json: {"D<D>_N<N>": bytes_read + bytes_written per launch}."""
import csv
import json
import sys

from gram_traffic import SHAPES


def main(path, out):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    per_launch = {}
    for r in rows[hdr + 1:]:
        if len(r) <= vi or "gram_tma" not in r[ki]:
            continue
        per_launch.setdefault(int(r[0]), 0.0)
        per_launch[int(r[0])] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    vals = [per_launch[k] for k in sorted(per_launch)]
    assert len(vals) == len(SHAPES), (len(vals), len(SHAPES))
    res = {f"D{D}_N{N}": v for (D, N), v in zip(SHAPES, vals)}
    res["_source"] = f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of gram_tma_kernel, one launch per shape ({path})"
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
