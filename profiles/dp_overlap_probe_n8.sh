#!/bin/bash
# N = 8: granularity of the gradient collectives (one per encoder layer / one per encoder) x SM budget of the backward GEMMs.
# usage: bash profiles/dp_overlap_probe_n8.sh
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --no-configs 2>gpurun_out/n8.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']/1e6,3), 'M samples/s')"; }
echo "blocks=encoder COMM_SMS=148"; MMN_DP_BLOCKS=encoder MMN_WIDE_COMM_SMS=148 run
echo "blocks=encoder COMM_SMS=112"; MMN_DP_BLOCKS=encoder MMN_WIDE_COMM_SMS=112 run
