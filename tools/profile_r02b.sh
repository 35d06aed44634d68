#!/bin/bash
# Round-2 evidence, part 2 (one B200): init_mst after the chain fast path, ncu --set full of the persistent PCG kernel
# summarised ON the box (the .ncu-rep is too large to bring back with the others).
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_init_mst.py tests/test_ref_pin.py -m gpu -x -q 2>&1 | tail -3
python tools/profile_aux.py --reps 3 > gpurun_out/r02_aux_timing_v2.json
ncu --set full --clock-control none --import-source on -k regex:k_pcg_smem -s 22 -c 1 -o /tmp/r02_prof_pcg_smem -f \
    python tools/profile_step.py --cost L1 --iters 30 --no-profile > /dev/null
python tools/ncu_summary.py /tmp/r02_prof_pcg_smem.ncu-rep > gpurun_out/r02_ncu_prof_pcg_smem.json
ncu -i /tmp/r02_prof_pcg_smem.ncu-rep --page raw --csv | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; d=rows[2]
keep={k:d[i] for i,k in enumerate(h) if any(t in k for t in ('dram__bytes','lts__t_bytes','lts__t_sectors_srcunit_tex','l1tex__t_bytes','smsp__cycles','shared','launch__'))}
json.dump(keep,open('gpurun_out/r02_ncu_prof_pcg_smem_raw_subset.json','w'),indent=1)"
ncu --set full --clock-control none -k regex:"k_mst" -c 3 -o /tmp/r02_prof_mst2 -f python tools/profile_aux.py --reps 1 > /dev/null
python tools/ncu_summary.py /tmp/r02_prof_mst2.ncu-rep > gpurun_out/r02_ncu_prof_mst_v2.json
cat gpurun_out/r02_aux_timing_v2.json
