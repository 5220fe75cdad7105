#!/bin/bash
# compute-sanitizer targets over small inputs of the hot path (run on the GPU box: `gpurun -- tools/sanitize.sh`).
#   memcheck  : out-of-bounds / misaligned accesses of every kernel family (radius graph, exact fp32, fused tcgen05)
#   racecheck : shared-memory hazards of the fused kernels (barrier / mbarrier protocol of the pipelines)
# The inputs are the smoke() sizes: the tools replay every launch, a full workload would take hours.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=${SAN:-compute-sanitizer}
run() {   # tool, log
  echo "== $SAN --tool $1" | tee "gpurun_out/sanitize_$1.log"
  timeout "${SAN_TIMEOUT:-900}" $SAN --tool "$1" --error-exitcode 3 --launch-timeout 0 \
      python tools/sanitize_inputs.py >> "gpurun_out/sanitize_$1.log" 2>&1
  echo "exit code $?" | tee -a "gpurun_out/sanitize_$1.log"
  tail -4 "gpurun_out/sanitize_$1.log"
}
for tool in ${SAN_TOOLS:-memcheck racecheck}; do run "$tool"; done
