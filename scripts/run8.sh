export NGPU=8
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/probe/a2a_probe.py 1024 2>&1 | grep "NCCL all" | tee gpurun_out/a2a_probe.txt
export GBX_PART_PROFILE=1
GBX_EXCHANGE_PIECES=2 STEPS=30 TAG=_8g_p2 timeout 150 bash scripts/gpu_cycle.sh partrun
STEPS=30 TAG=_8g_direct BARGS="--part-exchange 0" timeout 150 bash scripts/gpu_cycle.sh partrun
