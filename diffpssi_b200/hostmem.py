"""Host-side placement of the buffers the recorder is streamed into (``Pss.run(host_out=...)``).

On a two-socket host every GPU hangs off one socket; a pinned buffer whose pages live on the other socket makes every
device->host byte cross the inter-socket link, and with one process per GPU all of them allocate wherever the launcher
happened to start them.  ``bind_to_gpu_node`` pins the calling process to the CPUs of the GPU's NUMA node (read from
sysfs: /sys/bus/pci/devices/<bus id>/numa_node and /sys/devices/system/node/node<N>/cpulist), after which first touch --
and therefore ``tensor.pin_memory()`` -- places pages on that node.  Pure host bookkeeping; no effect on results.
"""
import os

import torch


def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus += list(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_info(index):
    """{'bus_id', 'numa_node', 'cpus'} of CUDA device ``index``; numa_node is None when the platform does not tell."""
    info = {'bus_id': None, 'numa_node': None, 'cpus': None}
    try:
        prop = torch.cuda.get_device_properties(index)
        bus = '%04x:%02x:%02x.0' % (getattr(prop, 'pci_domain_id', 0), prop.pci_bus_id, prop.pci_device_id)
        info['bus_id'] = bus
        with open('/sys/bus/pci/devices/%s/numa_node' % bus) as f:
            node = int(f.read().strip())
        if node >= 0:
            info['numa_node'] = node
            with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
                info['cpus'] = _parse_cpulist(f.read())
    except Exception as exc:      # no sysfs entry, containerised view, ...: placement stays with the OS
        info['error'] = repr(exc)
    return info


def bind_to_gpu_node(index):
    """Restrict the calling process to the CPUs of GPU ``index``'s NUMA node (intersection with the CPUs it may already
    use).  Returns the info dict; a no-op when the node is unknown."""
    info = gpu_numa_info(index)
    if info.get('cpus'):
        try:
            allowed = os.sched_getaffinity(0)
            want = set(info['cpus']) & allowed
            if want:
                os.sched_setaffinity(0, want)
                info['bound'] = len(want)
        except Exception as exc:
            info['error'] = repr(exc)
    return info
