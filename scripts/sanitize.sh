#!/bin/bash
# compute-sanitizer pass over the kernels of the hot path (memcheck, racecheck, synccheck) through the GPU parity tests:
#   gpurun --timeout 900 -- 'bash scripts/sanitize.sh'
# Output: gpurun_out/r2q_sanitize_<tool>.txt (tail of each run) + a one-line verdict per tool on stdout.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SMALL="tests/test_gpu_parity.py::test_K1_reference_unit_test tests/test_gpu_parity.py::test_apply_D tests/test_gpu_parity.py::test_ensemble_batch_handle tests/test_gpu_parity.py::test_adjoint_without_store tests/test_gpu_parity.py::test_apply_D_512_f32"
K='K1 or (apply_D and 64-False) or (apply_D and 128-False) or (apply_D and 256-False-float32) or (ensemble_batch_handle and 64-3-float32) or adjoint_without_store or apply_D_512'
python -c "import torch; print(torch.cuda.get_device_name(0))"
for tool in memcheck racecheck synccheck; do
  extra=""
  [ $tool = racecheck ] && extra="--racecheck-report all"
  timeout 420 $CS --tool $tool $extra --target-processes all --print-limit 30 --error-exitcode 9 \
    python -m pytest $SMALL -m gpu -q -x -k "$K" > gpurun_out/r2q_sanitize_$tool.log 2>&1
  rc=$?
  tail -40 gpurun_out/r2q_sanitize_$tool.log > gpurun_out/r2q_sanitize_$tool.txt
  grep -c "=========.*\(Error\|Race\|Hazard\|hazard\)" gpurun_out/r2q_sanitize_$tool.log > /tmp/cnt.txt
  echo "$tool: rc=$rc flagged_lines=$(cat /tmp/cnt.txt) $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r2q_sanitize_$tool.log | tr '\n' ' ')"
  grep -E "=========.*(Error|Race|Hazard|hazard|at |in )" gpurun_out/r2q_sanitize_$tool.log | head -40 > gpurun_out/r2q_sanitize_${tool}_first.txt
  rm -f gpurun_out/r2q_sanitize_$tool.log.big
done
