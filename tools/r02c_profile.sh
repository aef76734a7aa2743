#!/bin/bash
# Round-2 (third session) ncu evidence (GPU box): full captures of the tensor-map staged decoder (aad_decode_tma, kernel
# path 7) and of the default decoder on the same aligned 1,500-clip batch.  Runs only after the same command has exited 0
# without ncu.
set -x
export AB_CLIPS=1500
B="python tools/dec_tma_ab.py c1b4"
$B > gpurun_out/r02c_tma_plain.json 2> gpurun_out/r02c_tma_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:aad_decode_tma -s 3 -c 1 -f -o gpurun_out/prof_r2e_tma $B > gpurun_out/r02c_ncu_tma.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:aad_decode_fast -s 3 -c 1 -f -o gpurun_out/prof_r2e_fast $B > gpurun_out/r02c_ncu_fast.log 2>&1
ls -la gpurun_out/prof_r2e_* gpurun_out/r02c_*
