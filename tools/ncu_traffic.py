"""Summarise one `ncu --set full` capture of the resident Arnoldi kernel for bench.py's `roofline.traffic`:
    python tools/ncu_traffic.py <raw.csv> <k> <passes> <n> <out.json> [source note]
raw.csv = `ncu -i <rep> --page raw --csv`; k = basis size of the captured launch (reconstructed from the launch index, see
profiles/README.md); algorithmic bytes of that launch = (passes * k + 3) * 8 n."""
import csv
import json
import sys

raw, k, passes, n, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tot = 0.0
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    i = hdr.index(name)
    tot += float(vals[i]) * scale[units[i]]
dur_i = hdr.index("gpu__time_duration.sum")
dur = float(vals[dur_i]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[dur_i]]
alg = (passes * k + 3) * 8.0 * n
d = {"kernel": vals[hdr.index("Kernel Name")], "k": k, "passes": passes, "n": n, "dram_bytes": tot, "algorithmic_bytes": alg, "ratio": tot / alg,
     "duration_s": dur, "gbs_under_ncu": alg / dur / 1e9, "source": sys.argv[6] if len(sys.argv) > 6 else raw}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d))
