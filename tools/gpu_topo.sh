#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nvidia-smi nvlink -s >> gpurun_out/topo.txt 2>&1 | head -40
cat gpurun_out/topo.txt | head -30
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/nccl_probe.py > gpurun_out/nccl_probe.log 2>&1; echo rc=$?
grep -i "NVLS\|P2P\|SHM\|via\|Connected\|probe" gpurun_out/nccl_probe.log | head -30
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 tools/two_gpu_check.py 2>&1 | tail -3
