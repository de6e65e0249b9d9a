#!/bin/bash
# Records the GPU box's host topology (NUMA nodes, allowed CPUs/memory nodes, GPU affinity) for the e2e-scaling work.
out=gpurun_out/box_topology.txt
{
  echo "== nproc"; nproc
  echo "== allowed"; grep -i "allowed" /proc/self/status
  echo "== numa nodes"; for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  echo "== lscpu"; lscpu | grep -iE "model name|socket|numa|^cpu\(s\)|thread"
  echo "== topo"; nvidia-smi topo -m
  echo "== gpus"; nvidia-smi --query-gpu=index,name,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv
  echo "== gpu numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) class=$(cat $d/class)"; fi; done
  echo "== libnuma"; ls /usr/lib/x86_64-linux-gnu | grep -i numa
  echo "== mem"; free -g
} > $out 2>&1
