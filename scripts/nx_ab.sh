#!/bin/bash
# A/B of the gradient exchange on N GPUs of one box: NCCL (two communicators) vs the fused LL exchange over NVLink peer memory.
# Usage (under gpurun --gpus N): bash scripts/nx_ab.sh N [tags...]   tags: nccl ll2 ll1
N=${1:-2}; shift
tags=${@:-"nccl ll2"}
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/n${N}_$tag.json 2>gpurun_out/n${N}_$tag.err; python -c "
import json
d=json.loads(open('gpurun_out/n${N}_$tag.json').read().strip().splitlines()[-1]); print('$tag', $N, round(d['value']/1e6,1), round(d['ms_per_step'],3), d['phases_ms'], {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if k in ('fused_minibatch','reduce_partials','adam')})" || tail -5 gpurun_out/n${N}_$tag.err; }
for t in $tags; do
  case $t in
    nccl) run nccl CRUX_PEER_FLOATS=0 ;;
    ll2) run ll2 CRUX_PEER_FLOATS=8192 ;;
    default) run default X=1 ;;
    ll1) run ll1 CRUX_PEER_FLOATS=8192 CRUX_LL_ADAM_V1=1 ;;
  esac
done
