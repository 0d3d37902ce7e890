#!/bin/bash
# Host / topology facts of the GPU box (for the scaling analysis).
nproc; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" ; nvidia-smi topo -m 2>/dev/null | head -14
python - <<'PY'
import torch, os
for r in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(r)
    bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    try:
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip(); cl = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
    except OSError as e:
        node, cl = "?", str(e)
    print(r, bdf, "numa", node, "cpus", cl)
print("affinity", len(os.sched_getaffinity(0)))
PY
