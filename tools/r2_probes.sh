#!/bin/bash
# Round-2 opening probes (one gpurun call, ~1 min): what paces a k step of the chain kernels' MMA issuer - 462 clocks measured,
# 384 of tensor-pipe time?  (DESIGN.md §5.1 "Where the remaining time is".)  Prints, for the sdf-only chain, the per-k-step stamps
# of the MMA warp (wait a_ready | wait full | issue | commit) under four ablations:
#   baseline | one MMA per k step (1/3 of the operand reads) | no weight stream (no ring writes) | both
# Reading: if "mma issue" + "commit" stay ~the same and the step shortens only with fewer MMAs  -> tensor pipe / operand reads;
#          if the step shortens with NOSTREAM                                                   -> shared-memory write traffic / L2 stream;
#          if nothing changes                                                                   -> the issue loop itself (R2UR / TRYWAIT chain).
# Usage:  gpurun --timeout 300 -- 'bash tools/r2_probes.sh > gpurun_out/r2_probes.txt 2>&1'
cd "$(dirname "$0")/.."
for v in "X=0" "I2SDF_DEBUG_MMAS=1" "I2SDF_DEBUG_NOSTREAM=1" "I2SDF_DEBUG_MMAS=1 I2SDF_DEBUG_NOSTREAM=1"; do
    echo "=================== $v"
    env $v timeout 120 python tools/timeline.py 2>/dev/null | tail -28
done
echo "=================== kernel times (tools/quick_bench.py): 8-column items (default) and 16-column items"
timeout 120 python tools/quick_bench.py 2>/dev/null | head -3
I2SDF_SDF16=1 timeout 120 python tools/quick_bench.py 2>/dev/null | head -3
echo "=================== stand-alone k-step probe (tools/probe_kstep.cu): MMA pace / epilogue pace per traffic source"
bash tools/build_probe.sh >/dev/null 2>&1 && timeout 120 ./tools/probe_kstep
