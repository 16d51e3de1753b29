#!/bin/bash
# e2e leg only, 8 ranks, with and without NUMA placement of the pinned staging buffers
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
cat /sys/devices/system/node/online >> gpurun_out/topo.txt
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ] && [ "$(cat $d/class | cut -c1-6)" = "0x0302" ]; then echo "$d $(cat $d/numa_node) $(cat $d/local_cpulist)"; fi; done >> gpurun_out/topo.txt
nproc >> gpurun_out/topo.txt
N=${1:-8}
for mode in 0 1; do
  B200Q_NO_NUMA=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$mode bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 4 --awq-layers 0 --moe-layers 0 --moe-awq-experts 0 --no-cpu-baseline > gpurun_out/numa_$mode.json 2> gpurun_out/numa_$mode.err
done
tail -n 3 gpurun_out/numa_1.err | cut -c1-200
