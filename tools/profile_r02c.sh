#!/bin/bash
# Round-2 evidence, part 3 (one B200), after the block-Jacobi thresholds moved to (0.5, 0.001): per-kernel step profile,
# ncu launch list of the same step, ncu --set full of the on-path PCG kernel summarised ON the box.
set -x
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --cost L1 --iters 30 --time-kernels > gpurun_out/r02_step_L1_v2.json 2> gpurun_out/r02c.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_l1_v2.csv \
    python tools/profile_step.py --cost L1 --iters 30 --no-profile > /dev/null 2>> gpurun_out/r02c.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg_persistent_reg -s 22 -c 1 -o /tmp/r02_prof_pcg_reg -f \
    python tools/profile_step.py --cost L1 --iters 30 --no-profile > /dev/null 2>> gpurun_out/r02c.err
python tools/ncu_summary.py /tmp/r02_prof_pcg_reg.ncu-rep > gpurun_out/r02_ncu_prof_pcg_reg.json
ncu -i /tmp/r02_prof_pcg_reg.ncu-rep --page raw --csv | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; d=rows[2]
keep={k:d[i] for i,k in enumerate(h) if any(t in k for t in ('dram__bytes','lts__t_bytes','lts__t_sectors_srcunit_tex','l1tex__t_bytes','smsp__cycles','shared','launch__','gpu__time'))}
json.dump(keep,open('gpurun_out/r02_ncu_prof_pcg_reg_raw_subset.json','w'),indent=1)"
tail -3 gpurun_out/r02c.err
head -c 600 gpurun_out/r02_ncu_prof_pcg_reg.json
