"""profiles/traffic.json from an ncu full-set summary (scripts/ncu_launch_summary.py full ...): DRAM bytes read +
written per launch, averaged over the captured launches of each kernel, keyed by bench.py's kernel tags.
usage: python scripts/ncu_traffic.py profiles/rXX_ncu_full_summary.csv profiles/traffic.json"""
import csv
import json
import re
import sys

TAGS = [  # (regex on the ncu kernel name, bench tag)
    (r"ks_c2c_pipe<", "ks_c2c.y"),
    (r"kz_c2r(_pipe)?<\w+, \d+, 2>", "kz_c2r.rz"),
    (r"k_cg_update<", "k_cg_update"),
    (r"kz_deriv2_pipe<", "kz_deriv2"),
    (r"ks_deriv2_pipe<\w+, \d+, 1,", "ks_deriv2.y"),
    (r"ks_deriv2_pipe<\w+, \d+, 3,", "ks_deriv2.x.matvec"),
    (r"ks_deriv2_pipe<\w+, \d+, 4,", "ks_deriv2.x.rhs"),
    (r"kz_r2c(_pipe)?<\w+, \d+, 1>", "kz_r2c.axpy"),
    (r"kz_r2c(_pipe)?<\w+, \d+, 0>", "kz_r2c"),
    (r"kz_c2r(_pipe)?<\w+, \d+, 1>", "kz_c2r.norm"),
    (r"ks_pc_pipe<", "ks_pc"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(inp, out):
    rows = [r for r in csv.reader(l for l in open(inp) if not l.startswith("#"))]
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    rd, wr = ix["dram__bytes_read.sum"], ix["dram__bytes_write.sum"]
    acc = {}
    for r in rows[2:]:
        for rx, tag in TAGS:
            if re.search(rx, r[0]):
                b = float(r[rd]) * SCALE[units[rd]] + float(r[wr]) * SCALE[units[wr]]
                acc.setdefault(tag, []).append(b)
                break
    res = {t: sum(v) / len(v) for t, v in acc.items()}
    res["_source"] = (f"{inp}: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, 256^3 f32), "
                      "mean over the captured launches of each kernel")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
