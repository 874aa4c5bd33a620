"""Host-side placement for the host-resident mode: keep a rank's pinned buffers (and the thread which fills and
submits them) on the NUMA node its GPU hangs off, so that concurrent H2D streams of the GPUs of one box do not
all pull from one socket's memory.  Pure host plumbing; nothing here touches the decode path."""
import os


def numa_node_of_gpu(device_index):
    """The NUMA node of a CUDA device according to sysfs, or None when it cannot be determined."""
    try:
        import torch
        properties = torch.cuda.get_device_properties(device_index)
        address = "%04x:%02x:%02x.0" % (properties.pci_domain_id, properties.pci_bus_id, properties.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % address) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def cpus_of_numa_node(node):
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            text = f.read().strip()
    except Exception:
        return None
    cpus = set()
    for part in text.split(","):
        if "-" in part:
            low, high = part.split("-")
            cpus.update(range(int(low), int(high) + 1))
        elif part:
            cpus.add(int(part))
    return cpus or None


def bind_to_numa_node_of_gpu(device_index):
    """Restrict this process to the CPUs of the GPU's NUMA node (memory then follows by first touch).
    Returns the node, or None if nothing was changed."""
    node = numa_node_of_gpu(device_index)
    if node is None:
        return None
    cpus = cpus_of_numa_node(node)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
