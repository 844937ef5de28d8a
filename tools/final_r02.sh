#!/bin/bash
# Last GPU call of round 2 (one B200): the whole GPU suite incl. tests/test_zz_direct_gpu.py with every test's printed measurements,
# then the bench lines of the three training variants.  Every step writes its own file, so a clamped call keeps what finished.
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > gpurun_out/r02k_gputest_1gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02k_gputest_1gpu.log
tail -3 gpurun_out/r02k_gputest_1gpu.log
timeout 120 python bench.py --config synthetic_light_mask --cpu-rays 128 --steps 30 --warmup 5 > gpurun_out/r02k_bench_train_light_1gpu.json 2> gpurun_out/r02k_bench_light.err
timeout 120 python bench.py --bubble 1024 --cpu-rays 128 --steps 30 --warmup 5 > gpurun_out/r02k_bench_train_bubble_1gpu.json 2> gpurun_out/r02k_bench_bubble.err
timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/r02k_bench_train_1gpu.json 2> gpurun_out/r02k_bench.err
wc -c gpurun_out/r02k_*
