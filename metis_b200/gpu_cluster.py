"""Cluster description (mirrors gpu_cluster.py:8-58 of the reference)."""
from __future__ import annotations

from typing import List

from .utils import DeviceType, GPUNode, parse_hostfile, parse_nodefile


class GPUCluster:
    def __init__(self, hostfile_path: str, clusterfile_path: str, strict: bool = False):
        """``strict=True`` (opt-in): multi-digit ``slots=`` counts, and a ValueError that names the host for a
        node missing from the clusterfile, an unknown instance type or a missing field; the default reproduces
        the reference's parsing (quirk Q10) and its bare KeyError / ValueError."""
        self.host_entries = parse_hostfile(hostfile_path, strict=strict)
        self.nodes_info = parse_nodefile(clusterfile_path)
        if strict:
            for host in self.host_entries.values():
                info = self.nodes_info.get(host['ip'])
                if info is None:
                    raise ValueError(f"host {host['ip']!r} of {hostfile_path} is not in {clusterfile_path}")
                for key in ('instance_type', 'memory', 'intra_bandwidth', 'inter_bandwidth'):
                    if key not in info:
                        raise ValueError(f"{clusterfile_path}: node {host['ip']!r} has no {key!r}")
                if str(info['instance_type']).upper() not in DeviceType.__members__:
                    raise ValueError(f"{clusterfile_path}: node {host['ip']!r} has unknown instance_type "
                                     f"{info['instance_type']!r} (known: {', '.join(DeviceType.__members__)})")
        self.nodes = {
            node_id: GPUNode(device_type=DeviceType.from_string(self.nodes_info[host['ip']]['instance_type']),
                             num_devices=host['num_device'])
            for node_id, host in self.host_entries.items()
        }

    def get_num_nodes(self) -> int:
        return len(self.nodes)

    def get_num_nodes_by_device_type(self, device_type: str) -> int:
        """gpu_cluster.py:22-23: despite the name this is the number of DEVICES of the type."""
        return sum(n.num_devices for n in self.nodes.values() if n.device_type.name == device_type)

    def get_num_devices_per_node(self) -> int:
        return self.nodes[0].num_devices                      # node 0 for all nodes (quirk Q10)

    def get_total_num_devices(self) -> int:
        return sum(n.num_devices for n in self.nodes.values())

    def get_device_types(self) -> List[DeviceType]:
        return [n.device_type for n in self.nodes.values()]

    def get_str_device_types(self) -> str:
        return '_'.join(t.name for t in set(self.get_device_types()))

    def get_device_memory(self, node_id: int):
        return self.nodes_info[self.host_entries[node_id]['ip']]['memory'] * 1024

    def get_device_memory_for_device_type(self, device_type: str):
        for info in self.nodes_info.values():                 # gpu_cluster.py:47-50 (first match, raw string)
            if info['instance_type'] == device_type:
                return info['memory'] * 1024
        return None

    def get_intra_bandwidth(self, node_id: int):
        return self.nodes_info[self.host_entries[node_id]['ip']]['intra_bandwidth']

    def get_inter_bandwidth(self, node_id: int):
        # gpu_cluster.py:56-58 returns the *intra* value (quirk Q2); kept for parity
        return self.nodes_info[self.host_entries[node_id]['ip']]['intra_bandwidth']
