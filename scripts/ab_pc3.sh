#!/bin/bash
# A/B of the in-stage exchange form of the pure transforms (ks_pc_pipe, inverse ks_c2c_pipe: two tiles of smem,
# three CTAs per SM) against the committed three-tile form:
#   gpurun --timeout 600 -- 'bash scripts/ab_pc3.sh'
# base = glia_b200/lib/libglia_rd_base.so, pc3 = glia_b200/lib/libglia_rd_pc3.so (scripts/build_variant.sh)
set -u
cd "$(dirname "$0")/.."
B="python bench.py --no-extras --no-cpu-baseline"
L=$PWD/glia_b200/lib
mkdir -p gpurun_out
for rep in 1 2; do
for v in base pc3; do
  GLIA_RD_LIB=$L/libglia_rd_$v.so $B --steps 3 --warmup 3 > gpurun_out/r2p_256_${v}_$rep.json 2>> gpurun_out/r2p.err
done; done
for v in base pc3; do
  GLIA_RD_LIB=$L/libglia_rd_$v.so $B --workload rd512 --steps 1 --warmup 1 > gpurun_out/r2p_512_$v.json 2>> gpurun_out/r2p.err
done
python - <<'PY' | tee gpurun_out/r2p_pc3_ab.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2p_*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    k = d["kernels"]
    print(f.split("r2p_")[1][:-5].ljust(14), "%.2f steps/s" % d["value"], "frac %.3f" % d["roofline"]["whole_step"]["frac"],
          "sm %.0f MHz" % d["clocks"]["sm_mhz"],
          " ".join("%s=%.1f" % (t, k[t]["avg_us"]) for t in ("ks_pc", "ks_c2c.y", "ks_deriv2.y", "ks_deriv2.x.matvec", "kz_c2r.rz") if t in k))
PY
