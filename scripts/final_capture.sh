#!/bin/bash
# One gpurun call that re-validates and re-profiles the final code of a round on one B200:
#   gpurun --timeout 1500 -- 'bash scripts/final_capture.sh r2w'
# smoke(), pytest -m gpu, the default bench line (timed by `time`), the 512^3 line, the ncu launch list of the bench
# command and one ncu --set full capture (+ source pages of the four sweep kernels).  Outputs: gpurun_out/<tag>_*.
set -u
cd "$(dirname "$0")/.."
t=$1
o=gpurun_out
mkdir -p $o
python -c "import __graft_entry__ as g; g.smoke()" > $o/${t}_smoke.txt 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $o/${t}_pytest_gpu_tail.txt
( time python bench.py > $o/${t}_bench_256.json 2> $o/${t}_bench_256.err ) 2> $o/${t}_bench_256_time.txt
python bench.py --workload rd512 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $o/${t}_bench_512.json 2> $o/${t}_bench_512.err
N1="ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1500"
$N1 --csv --log-file $o/${t}_launches_raw.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > $o/${t}_ncu_bench.log 2>&1
python scripts/ncu_launch_summary.py launches $o/${t}_launches_raw.csv $o/${t}_launches_summary.csv "$N1 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline (256^3 f32; launches 4000..5500 = inside the PCG loops of the warm-up step)"
N2="ncu --set full --clock-control none --import-source on -s 40 -c 22"
$N2 -f -o $o/${t}_full python scripts/prof_driver.py > $o/${t}_ncu_full.log 2>&1
ncu -i $o/${t}_full.ncu-rep --page raw --csv > $o/${t}_full_raw.csv 2>> $o/${t}_ncu_full.log
python scripts/ncu_launch_summary.py full $o/${t}_full_raw.csv $o/${t}_ncu_full_summary.csv "$N2 python scripts/prof_driver.py (256^3 f32, one diffusion solve; final kernels of the round)"
for k in ks_deriv2_pipe ks_pc_pipe kz_c2r kz_deriv2_pipe; do
  ncu -i $o/${t}_full.ncu-rep --page source --csv -k regex:$k -c 1 > $o/${t}_src_$k.csv 2>> $o/${t}_ncu_full.log
done
rm -f $o/${t}_full.ncu-rep
tail -3 $o/${t}_pytest_gpu_tail.txt; cat $o/${t}_smoke.txt | tail -1; cat $o/${t}_bench_256_time.txt | tail -3
python scripts/show_bench.py $o/${t}_bench_256.json | head -14
python scripts/show_bench.py $o/${t}_bench_512.json | head -10
