#!/bin/bash
# Multi-GPU evidence of a round, one gpurun call per N:   gpurun --gpus N --timeout 1500 -- 'bash tools/scale_round.sh N'
# bench.py under torchrun (weak-scaled C3, strong-scaled replicas, row-sharded C5), the row-sharded solve checked against the
# single-GPU result on every rank (count data and real-valued data: exact-sum and running-sum screens), the C++ facade's
# ReplicaSolver over the N visible devices, the two-process shard tests.  Outputs: gpurun_out/scale_N*.
set -x
N=${1:-2}
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $N --steps 5 --warmup 3 > $O/scale_n${N}_bench.json 2> $O/scale_n${N}_bench.err
$T tools/shard_check.py --size 1000000 --parts 32 --check > $O/scale_n${N}_shard_counts.log 2>&1
$T tools/shard_check.py --size 1000000 --parts 32 --data real --check > $O/scale_n${N}_shard_real.log 2>&1
LD_LIBRARY_PATH=partitionsolvers_b200/_lib:$LD_LIBRARY_PATH partitionsolvers_b200/_lib/test_facade > $O/scale_n${N}_facade.log 2>&1
python -m pytest tests/test_gpu_shard.py -x -q > $O/scale_n${N}_shard_tests.log 2>&1
tail -n 3 $O/scale_n${N}_shard_counts.log $O/scale_n${N}_shard_real.log $O/scale_n${N}_facade.log $O/scale_n${N}_shard_tests.log
python - <<EOF
import json
d = json.loads(open("$O/scale_n${N}_bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, d.get("mc_replicas", {}).get("value"), json.dumps(d.get("sharded"))[:500])
EOF
