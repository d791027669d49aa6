#!/usr/bin/env python
"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py gpurun_out/launches.csv [n_hessian_syrk_name_substring]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    name = re.sub(r"\(.*$", "", r[ik]).replace("cvxb::<", "").strip()
    tot[name] += v
    cnt[name] += 1
probe = sum(v for k, v in tot.items() if "dmma_peak" in k)
total = sum(tot.values()) - probe
syrk = [k for k in cnt if "gemm_dmma_kernel<128, 128, 2, 4, 1, 1>" in k or "gemm_dmma_streamk_kernel<1, 1>" in k]
steps = sum(cnt[k] for k in syrk)
print("total %.3f ms excluding the DMMA peak probe; %d Hessian SYRK launches = Newton steps in the capture -> %.3f ms per step"
      % (total / 1e3, steps, total / 1e3 / max(steps, 1)))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if "dmma_peak" in k:
        continue
    print("%-62s launches=%5d total=%8.3f ms avg=%8.2f us share=%5.1f%%" % (k[:62], cnt[k], v / 1e3, v / cnt[k], 100 * v / total))
