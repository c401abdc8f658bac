#!/bin/bash
# What the gradient all-reduce costs a data-parallel C4 step as a function of the SMs the backward GEMMs use while collectives
# are in flight (MMN_WIDE_COMM_SMS; 148 = all of them).  usage (N GPUs): bash profiles/dp_overlap_probe.sh N [budgets...]
N=${1:-2}; shift
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']/1e6,3), 'M samples/s')"; }
for s in ${@:-148 136 128 120}; do echo "N=$N MMN_WIDE_COMM_SMS=$s"; MMN_WIDE_COMM_SMS=$s run; done
