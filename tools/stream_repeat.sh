# replays a stream a few times per PCG variant; with IRA_DEBUG every global call prints its iteration counts (2: every Newton solve)
for sv in ${SOLVERS:-0 256}; do for i in $(seq 1 ${RUNS:-3}); do IRA_DEBUG=${DBG:-1} IRA_SOLVER=$sv timeout 200 python tools/bench_stream.py --frames ${FRAMES:-4002} --cpu-frames 0 2>gpurun_out/stream_dbg_${sv}_$i.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('solver',$sv,d['result_sha1'],repr(d['geodesic_rms_vs_ground_truth_rad']),[round(c['ms'],1) for c in d['global_calls']])"; done; done
