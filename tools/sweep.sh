#!/bin/bash
# Workload sweep on one B200 (BASELINE.json configs[3], configs[4] per-GPU shard at both cutoffs); lean bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for a in "--workload cfg2_lipo_train" "--workload cfg4_bace_cls" "--workload cfg5_cov2_stress --molecules 128" \
         "--workload cfg5_cov2_stress --molecules 128 --cutoff 5"; do
  timeout 300 python bench.py --lean --steps 10 $a 2>gpurun_out/sweep.err | tail -1 | tee -a gpurun_out/sweep.jsonl | python -c '
import sys, json
d = json.loads(sys.stdin.read()); c = d["config"]; r = d["roofline"]
print(c["workload"], "cutoff", c["cutoff"], "atoms", c["atoms"], "edges", c["edges"], "conf/s", round(d["value"]),
      "ms", round(d["ms_per_step"], 3), "frac", round(r["frac"], 4), "us", round(r["avg_launch_us"], 1))' || tail -5 gpurun_out/sweep.err
done
