#!/bin/bash
# final single-GPU records of the round (run under gpurun): default bench line + the other BASELINE.json configurations
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_default.log 2>&1; tail -1 gpurun_out/r02_bench_default.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.log 2>&1; tail -1 gpurun_out/r02_bench_reference_arm.log | cut -c1-300
for m in dptn_wav dptn_mask dprnn; do
  python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_$m.log 2>&1; tail -1 gpurun_out/r02_bench_$m.log | cut -c1-200
done
python bench.py --model dptn_wav --batch 4 --steps 5 --warmup 3 --no-eager-baseline > gpurun_out/r02_bench_cfg1.log 2>&1; tail -1 gpurun_out/r02_bench_cfg1.log | cut -c1-200
python bench.py --seconds 10 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_10s.log 2>&1; tail -1 gpurun_out/r02_bench_10s.log | cut -c1-200
python bench.py --engine tensor-f16res --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_f16res.log 2>&1; tail -1 gpurun_out/r02_bench_f16res.log | cut -c1-200
python tools/attn_bench.py --shapes all --versions 1,3 > gpurun_out/r02_attn_bench_final.log 2>&1; cat gpurun_out/r02_attn_bench_final.log
# lipreader front end: parity of both engines + device time, ablation of the tcgen05 convolution kernel
python tools/lipreader_bench.py 32 100 5 > gpurun_out/r02_lipreader_bench.log 2>&1; tail -1 gpurun_out/r02_lipreader_bench.log
python tools/lipreader_ablate.py 32 100 > gpurun_out/r02_lipreader_ablation.log 2>&1; head -4 gpurun_out/r02_lipreader_ablation.log
