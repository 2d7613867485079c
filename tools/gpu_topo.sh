#!/bin/bash
nvidia-smi topo -m 2>&1 | head -30
echo "--- nodes"; ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
for n in /sys/devices/system/node/node*; do echo "$n cpulist $(cat $n/cpulist 2>/dev/null) mem $(grep MemTotal $n/meminfo 2>/dev/null | awk '{print $4,$5}')"; done
echo "--- allowed"; grep -i "allowed_list" /proc/self/status
echo "--- nproc $(nproc)"; 
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then echo "$(basename $d) numa_node $(cat $d/numa_node) local_cpulist $(cat $d/local_cpulist)"; fi; done
which numactl; python -c "import ctypes.util; print(ctypes.util.find_library('numa'))"
