#!/bin/bash
# Round-2 evidence run (one B200, under gpurun): launch list, ncu --set full of the on-path kernels, compute-sanitizer.
set -x
mkdir -p gpurun_out
python tools/profile_aux.py --reps 3 > gpurun_out/r02_aux_timing.json 2> gpurun_out/r02_aux.err
python tools/profile_step.py --cost L1 --iters 30 --time-kernels > gpurun_out/r02_step_L1.json 2>> gpurun_out/r02_aux.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_l1.csv \
    python tools/profile_step.py --cost L1 --iters 30 --no-profile > /dev/null 2>> gpurun_out/r02_aux.err
ncu --set full --clock-control none --import-source on -k regex:k_pcg_smem -s 22 -c 1 -o gpurun_out/r02_prof_pcg_smem -f \
    python tools/profile_step.py --cost L1 --iters 30 --no-profile > /dev/null 2>> gpurun_out/r02_aux.err
ncu --set full --clock-control none --import-source on -k regex:"k_residual|k_sell_rhs|k_weights|k_pair_best|k_attach_best" -s 40 -c 5 -o gpurun_out/r02_prof_edge -f \
    python tools/profile_step.py --cost L1 --iters 12 --no-profile > /dev/null 2>> gpurun_out/r02_aux.err
ncu --set full --clock-control none --import-source on -k regex:"k_mst" -c 3 -o gpurun_out/r02_prof_mst -f \
    python tools/profile_aux.py --reps 1 > /dev/null 2>> gpurun_out/r02_aux.err
for tool in memcheck racecheck synccheck; do
  SAN_N=12000 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r02_sanitizer_$tool.log
done
cat gpurun_out/r02_aux_timing.json
