#!/bin/bash
# One gpurun call that measures the two unmeasured 512-point candidates of round 1 (lever 1):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/round2_ab.sh'
# Before the call, build the probe libraries here (CPU box):
#   scripts/build_variant.sh tw1 -DGLIA_TW_LDG=1 ; scripts/build_variant.sh tw2 -DGLIA_TW_LDG=2
# Output: gpurun_out/r2ab_*.json + a summary on stdout.
set -u
cd "$(dirname "$0")/.."
B="python bench.py --no-extras --no-cpu-baseline"
L=$PWD/glia_b200/lib
run() { tag=$1; shift
  env "$@" $B --steps 3 --warmup 3 > gpurun_out/r2ab_256_$tag.json 2>> gpurun_out/r2ab.err
  env "$@" $B --workload rd512 --steps 1 --warmup 1 > gpurun_out/r2ab_512_$tag.json 2>> gpurun_out/r2ab.err; }
run base GLIA_RD_ZPIPE=0
run zpipe GLIA_RD_ZPIPE=1
for v in tw1 tw2; do
  [ -f $L/libglia_rd_$v.so ] && run $v GLIA_RD_LIB=$L/libglia_rd_$v.so && run ${v}_zpipe GLIA_RD_LIB=$L/libglia_rd_$v.so GLIA_RD_ZPIPE=1
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2ab_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    k = d["kernels"]
    print(f.split("r2ab_")[1][:-5].ljust(16), "%.2f steps/s" % d["value"], "frac %.3f" % d["roofline"]["whole_step"]["frac"],
          " ".join("%s=%.1f" % (t, k[t]["avg_us"]) for t in ("kz_deriv2", "ks_deriv2.y", "ks_deriv2.x.matvec", "kz_c2r.rz") if t in k))
PY
