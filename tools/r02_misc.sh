set -x
for s in 8 16 24 48; do AAD_B200_SLICES=$s python bench.py --steps 3 --warmup 3 --no-cpu --no-long --no-sweeps > gpurun_out/r2_slices_$s.json 2> gpurun_out/r2_slices_$s.err; done
L="python bench.py --steps 1 --warmup 1 --clips 64 --no-sweeps --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:aad_decode_fast<\(int\)4, \(int\)2, \(int\)1' -s 4 -c 1 -f -o gpurun_out/prof_r2_dec_il $L > gpurun_out/r02_ncu_dec_il.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:aad_decode_wide<\(int\)3, \(int\)1' -s 4 -c 1 -f -o gpurun_out/prof_r2_wide_il $L > gpurun_out/r02_ncu_wide_il.log 2>&1
ls -la gpurun_out/prof_r2_*
