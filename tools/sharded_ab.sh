#!/bin/bash
# sharded-state check on N GPUs, fused peer-memory swap vs NCCL all-to-all:  tools/sharded_ab.sh N QUBITS [LAYERS]
N=$1; Q=$2; L=${3:-2}
for path in p2p nccl; do
  echo "== QB_SWAP=$path"
  QB_SWAP=$path python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/sharded_check.py --qubits $Q --layers $L 2>&1 | grep -E "^\{|Error|error" | tail -3
done
