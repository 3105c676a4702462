"""Input records of the search (mirrors the reference's utils.py:8-85 interface).

Same names and meaning as the reference so callers can switch imports:
``DeviceType``, ``ModelConfig``, ``GPUNode``, ``parse_hostfile``, ``parse_nodefile``.
``DeviceType`` additionally knows H100 and B200 (the reference enum stops at T4,
utils.py:46-50, so BASELINE configs 2/4/5 cannot be expressed with it).
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from enum import Enum
from typing import Dict, Union


class DeviceType(Enum):
    A100 = "a100"
    V100 = "v100"
    P100 = "p100"
    T4 = "t4"
    H100 = "h100"
    B200 = "b200"

    @staticmethod
    def from_string(s: str) -> 'DeviceType':
        """utils.py:52-57: case-insensitive lookup, ValueError for unknown names."""
        member = DeviceType.__members__.get(s.upper())
        if member is None:
            raise ValueError
        return member


@dataclass
class ModelConfig:
    """utils.py:71-79 (keyword construction as in cost_het_cluster.py:63-65)."""
    model_name: str
    num_layers: int
    sequence_length: int
    vocab_size: int
    hidden_size: int
    attention_head_size: int


@dataclass
class GPUNode:
    device_type: DeviceType
    num_devices: int


def parse_hostfile(file_path: str, strict: bool = False) -> Dict[int, Dict[str, Union[str, int]]]:
    """utils.py:8-24.  The device count is the single character at offset 6 of the
    second token (quirk Q10: ``IP4 8888888`` and ``host slots=8`` both give 8, ``slots=16`` gives 1).
    ``strict=True`` (opt-in) parses ``<host> slots=<N>`` with a multi-digit N, skips blank lines and
    rejects anything else with a ValueError naming the line."""
    entries: Dict[int, Dict[str, Union[str, int]]] = {}
    with open(file_path, 'rt') as fh:
        if not strict:
            for node_id, line in enumerate(iter(fh.readline, '')):
                fields = line.split(' ')
                entries[node_id] = {'ip': fields[0], 'num_device': int(fields[1][6:7])}
            return entries
        for lineno, line in enumerate(fh, 1):
            fields = line.split()
            if not fields:
                continue
            if len(fields) != 2 or not fields[1].startswith('slots=') or not fields[1][6:].isdigit() \
                    or int(fields[1][6:]) < 1:
                raise ValueError(f'{file_path}:{lineno}: expected "<host> slots=<N>", got {line.strip()!r}')
            entries[len(entries)] = {'ip': fields[0], 'num_device': int(fields[1][6:])}
    return entries


def parse_nodefile(file_path: str) -> Dict[str, Dict[str, Union[str, int, float]]]:
    """utils.py:27-31: clusterfile JSON keyed by host ip."""
    with open(file_path, 'r') as fh:
        return json.loads(fh.read())
