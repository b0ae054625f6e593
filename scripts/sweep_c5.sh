#!/bin/bash
# BASELINE config 5: collocation-count sweep, N_f in {1e5, 1e6, 1e7, 1e8} points IN TOTAL on W GPUs (one bench.py line each)
# usage: bash scripts/sweep_c5.sh W [workload]   -> gpurun_out/r2_sweep_W<W>_<workload>.jsonl
W=${1:-1}; WL=${2:-ev}
OUT=gpurun_out/r2_sweep_W${W}_${WL}.jsonl
: > $OUT
for TOT in 100000 1000000 10000000 100000000; do
  N=$((TOT / W))
  ARGS="bench.py --gpus $W --steps 5 --warmup 3 --workload $WL --n-f $N --no-cpu-baseline --no-e2e --no-small"
  if [ "$W" = "1" ]; then python $ARGS 2>/dev/null | grep '^{' >> $OUT
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29541 $ARGS 2>/dev/null | grep '^{' >> $OUT; fi
done
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    print(f"W={d['n_gpus']} n_f/GPU={d['config']['n_f_per_gpu']:>9d} total={d['config']['n_f_per_gpu']*d['n_gpus']:>9d}  {d['value']:.4g} pts/s  {d['ms_per_step']:.3f} ms/step  kernel share {d['roofline']['kernel_share_of_step']:.3f}" + (f"  dp_identity grad {d['dp_identity']['grad_rel']:.1e}" if 'dp_identity' in d else ""))
PY
