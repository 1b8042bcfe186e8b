#!/bin/bash
# ncu evidence of the round (run under gpurun, one GPU): launch list of the whole forward and one --set full capture of
# the first intra + inter sub-block (QKV, attention, out-projection, LSTM, FFN) plus the encoder and the tail kernel of
# the SECOND forward of tools/profile_forward.py (cfg-2: batch 32 x 4 s).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python tools/profile_forward.py > gpurun_out/r02_prof_plain.log 2>&1 || { cat gpurun_out/r02_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python tools/profile_forward.py > gpurun_out/r02_ncu_launches.log 2>&1
# per forward the regex matches: encoder (1) + 12 sub-blocks x 5 + tail (1) = 62 launches
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_tc_|k_tail_staged|k_encoder_vec" \
    --launch-skip 62 --launch-count 11 -f -o gpurun_out/r02_full_a python tools/profile_forward.py > gpurun_out/r02_ncu_full_a.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_tail_staged|k_visual" \
    --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02_full_b python tools/profile_forward.py > gpurun_out/r02_ncu_full_b.log 2>&1
tail -n 2 gpurun_out/r02_ncu_full_a.log; tail -n 2 gpurun_out/r02_ncu_full_b.log
