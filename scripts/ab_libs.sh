#!/bin/bash
# A/B of library builds on one B200 (how every profiles/r2[p-z]*_ab.txt of round 2 was taken; variants are built with
# scripts/build_variant.sh or by copying libglia_rd.so aside before a rebuild):  gpurun --timeout 900 -- 'bash scripts/ab_libs.sh <tag> <name> [<name> ...]'
# <name> -> glia_b200/lib/libglia_rd_<name>.so ("default" -> libglia_rd.so).  256^3 twice per library (interleaved),
# 512^3 once; a parity subset runs first on the LAST library named.  Output: gpurun_out/<tag>_*.json, <tag>_ab.txt
set -u
cd "$(dirname "$0")/.."
tag=$1; shift
B="python bench.py --no-extras --no-cpu-baseline"
L=$PWD/glia_b200/lib
mkdir -p gpurun_out
lib() { if [ "$1" = default ]; then echo $L/libglia_rd.so; else echo $L/libglia_rd_$1.so; fi; }
last=${@: -1}
GLIA_RD_LIB=$(lib $last) python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "K1 or forward_adjoint or config1 or apply_D" 2>&1 | tail -2 > gpurun_out/${tag}_ab.txt
for rep in 1 2; do for v in "$@"; do
  GLIA_RD_LIB=$(lib $v) $B --steps 3 --warmup 3 > gpurun_out/${tag}_256_${v}_$rep.json 2>> gpurun_out/${tag}.err
done; done
for v in "$@"; do
  GLIA_RD_LIB=$(lib $v) $B --workload rd512 --steps 1 --warmup 1 > gpurun_out/${tag}_512_$v.json 2>> gpurun_out/${tag}.err
done
python - $tag <<'PY' | tee -a gpurun_out/${tag}_ab.txt
import glob, json, sys
tag = sys.argv[1]
for f in sorted(glob.glob(f"gpurun_out/{tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    k = d["kernels"]
    print(f.split(tag + "_")[1][:-5].ljust(18), "%.2f steps/s" % d["value"], "frac %.3f" % d["roofline"]["whole_step"]["frac"],
          "sm %.0f MHz" % d["clocks"]["sm_mhz"],
          " ".join("%s=%.1f" % (t, k[t]["avg_us"]) for t in ("kz_deriv2", "ks_deriv2.y", "ks_deriv2.x.matvec", "ks_pc", "ks_c2c.y", "kz_c2r.rz") if t in k))
PY
