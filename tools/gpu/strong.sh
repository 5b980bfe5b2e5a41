#!/bin/bash
# strong scaling of ONE 18 000-frame clip on N GPUs: tools/gpu/strong.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --scaling strong --steps 3 > gpurun_out/strong_${N}gpu.json 2> gpurun_out/strong_${N}gpu.err
echo "exit $?"; tail -2 gpurun_out/strong_${N}gpu.err; cut -c1-260 gpurun_out/strong_${N}gpu.json
