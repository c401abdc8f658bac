#!/bin/bash
# N = 8: one configuration of (NVLS channels, SM budget of the backward GEMMs) per call; the variants tried are recorded in
# profiles/r2_dp_overlap_probe.txt.  usage: bash profiles/dp_overlap_probe_n8.sh <nvls_channels> <comm_sms>
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --no-configs 2>gpurun_out/n8.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']/1e6,3), 'M samples/s')"; }
echo "NCCL_NVLS_NCHANNELS=$1 MMN_WIDE_COMM_SMS=$2"; NCCL_NVLS_NCHANNELS=$1 MMN_WIDE_COMM_SMS=$2 run
