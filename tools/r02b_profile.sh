#!/bin/bash
# Round-2 (second session) ncu evidence (GPU box): launch list of the bench command and a full capture of the encoder
# (helper-lane schedule, shift-free quantiser).  Each capture runs only after the same command has exited 0 without ncu.
set -x
B="python bench.py --warmup 3 --no-e2e --no-cpu --no-long --no-sweeps"
$B --steps 2 > gpurun_out/r02b_plain.json 2> gpurun_out/r02b_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv $B --steps 2 > gpurun_out/r02b_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aad_encode_roles -s 3 -c 1 -f -o gpurun_out/prof_r2d_enc $B --steps 1 > gpurun_out/r02b_ncu_enc.log 2>&1
ls -la gpurun_out/prof_r2d_* gpurun_out/r02b_*
