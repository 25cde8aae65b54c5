#!/bin/bash
# usage: multi.sh N TAG  -- the bench on N GPUs of one box (torchrun, one rank per GPU), as the driver launches it
N=$1; TAG=$2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
else
  LDSR_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
fi
echo "rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench.json; grep "ldsr_em_batch: device" gpurun_out/${TAG}_bench.err | tail -8
if [ "$N" != "1" ]; then python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_parity.py -m gpu -q -k "device" > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log; fi
