#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus 8): aggregate pinned-memory copy bandwidth with 1, 2,
# 4 and 8 ranks copying at once, then the bench at 8 GPUs.
mkdir -p gpurun_out
python scripts/pcie_bw.py > gpurun_out/r02_pcie_1.json 2>/dev/null
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      scripts/pcie_bw.py > gpurun_out/r02_pcie_$n.json 2>/dev/null
done
cat gpurun_out/r02_pcie_*.json
for n in 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r02_bench_gpus$n.log 2>gpurun_out/r02_bench_gpus$n.err
  tail -c 400 gpurun_out/r02_bench_gpus$n.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_gpus$n.log").read().strip().splitlines()[-1])
print("gpus", d["n_gpus"], "value %.3e"%d["value"], "ms", round(d["ms_per_step"],2), "e2e %.3e"%d["e2e"]["value"], "e2e ms", round(d["e2e"]["ms_per_step"],2))
PY
done
