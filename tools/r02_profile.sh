#!/bin/bash
# Round-2 ncu evidence (GPU box): launch list of the bench command, full captures of the encoder (helper-lane
# schedule), the mono decoder, and the WAV-order stereo decoder of the one-stream path.  Each capture runs only
# after the same command has exited 0 without ncu.  Outputs under gpurun_out/.
set -x
B="python bench.py --warmup 3 --no-e2e --no-cpu --no-long --no-sweeps"
$B --steps 2 > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B --steps 2 > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aad_encode_roles -s 3 -c 1 -f -o gpurun_out/prof_r2_enc $B --steps 1 > gpurun_out/r02_ncu_enc.log 2>&1
$B --steps 1 --clips 1500 --trials 0 > gpurun_out/r02_plain_dec.json 2> gpurun_out/r02_plain_dec.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:aad_decode_fast -s 3 -c 1 -f -o gpurun_out/prof_r2_dec $B --steps 1 --clips 1500 --trials 0 > gpurun_out/r02_ncu_dec.log 2>&1
L="python bench.py --steps 1 --warmup 1 --clips 64 --no-sweeps --no-e2e --no-cpu"
$L > gpurun_out/r02_plain_long.json 2> gpurun_out/r02_plain_long.err || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:aad_decode_fast<4, 2, 1' -s 4 -c 1 -f -o gpurun_out/prof_r2_dec_il $L > gpurun_out/r02_ncu_dec_il.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:aad_decode_wide<3, 1' -s 4 -c 1 -f -o gpurun_out/prof_r2_wide_il $L > gpurun_out/r02_ncu_wide_il.log 2>&1
ls -la gpurun_out/prof_r2_*
