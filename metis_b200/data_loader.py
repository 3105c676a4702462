"""Profile JSON loader (mirrors data_loader.py:10-61 of the reference; same dict schema)."""
from __future__ import annotations

import json
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

_TYPE_RE = re.compile(r"DeviceType\.(\w+?)_")
_TP_RE = re.compile(r"tp(\d+)")
_BS_RE = re.compile(r"bs(\d+)")


class ProfileSchemaError(ValueError):
    """A profile file does not have the schema the search reads (opt-in check, ``validate=True``)."""


_NUMERIC_LISTS = (('execution_time', 'layer_compute_total_ms'), ('execution_memory', 'layer_memory_total_mb'),
                  ('model', 'parameters', 'parameters_per_layer_bytes'))
_NUMERIC_SCALARS = (('execution_time', 'forward_backward_time_ms'), ('execution_time', 'optimizer_time_ms'),
                    ('execution_time', 'batch_generator_time_ms'))


def _dig(raw: Dict, path: Tuple[str, ...], name: str):
    node = raw
    for key in path:
        if not isinstance(node, dict) or key not in node:
            raise ProfileSchemaError(f"{name}: missing key {'/'.join(path)}")
        node = node[key]
    return node


def validate_profile(raw: Dict, name: str, num_layers: Optional[int] = None) -> int:
    """Schema of one profile JSON as consumed by data_loader.py:20-36 of the reference: three per-layer
    numeric lists of one common length and three scalar times.  Returns the layer count."""
    if not (_TYPE_RE.search(name) and _TP_RE.search(name) and _BS_RE.search(name)):
        raise ProfileSchemaError(f"{name}: file name must look like DeviceType.<TYPE>_tp<N>_bs<M>.json")
    if int(_TP_RE.search(name).group(1)) < 1 or int(_BS_RE.search(name).group(1)) < 1:
        raise ProfileSchemaError(f"{name}: tp and bs must be positive")
    lengths = set()
    for path in _NUMERIC_LISTS:
        values = _dig(raw, path, name)
        if not isinstance(values, list) or not values:
            raise ProfileSchemaError(f"{name}: {'/'.join(path)} must be a non-empty list")
        if not all(isinstance(v, (int, float)) and not isinstance(v, bool) for v in values):
            raise ProfileSchemaError(f"{name}: {'/'.join(path)} holds a non-numeric entry")
        lengths.add(len(values))
    if len(lengths) != 1:
        raise ProfileSchemaError(f"{name}: per-layer lists differ in length {sorted(lengths)}")
    for path in _NUMERIC_SCALARS:
        value = _dig(raw, path, name)
        if not isinstance(value, (int, float)) or isinstance(value, bool):
            raise ProfileSchemaError(f"{name}: {'/'.join(path)} must be a number")
    layers = lengths.pop()
    if num_layers is not None and layers != num_layers:
        raise ProfileSchemaError(f"{name}: {layers} layers, other profiles have {num_layers}")
    return layers


class ProfileDataLoader:
    def __init__(self, profile_dir: str, file_order: Optional[Sequence[str]] = None, sort_files: bool = False,
                 validate: bool = False):
        """Default behaviour is the reference's (``os.listdir`` order, no checks).  Opt-in extensions:
        ``file_order`` pins the listing order (the first listed file supplies profile_data['model'] and the
        first listed type the layer weights, quirk Q3); ``sort_files`` sorts the listing instead, so the result
        no longer depends on the file system; ``validate`` checks every file against the schema the search reads
        and raises ProfileSchemaError naming the file and the offending key."""
        self.profile_dir = profile_dir
        self.validate = validate
        listed = [f for f in os.listdir(profile_dir) if f.endswith('.json')]
        if file_order is not None:
            if sorted(file_order) != sorted(listed):
                raise ValueError('file_order does not match the .json files in the profile directory')
            listed = list(file_order)
        elif sort_files:
            listed = sorted(listed)
        self.profile_data_list = listed

    @staticmethod
    def _model_section(raw: Dict) -> Dict:
        times = raw['execution_time']
        return {'optimizer_time': times['optimizer_time_ms'] * 2,
                'num_layers': len(times['layer_compute_total_ms']),
                'batch_generator': times['batch_generator_time_ms'],
                'parameters': raw['model']['parameters']['parameters_per_layer_bytes']}

    @staticmethod
    def _device_section(raw: Dict) -> Dict:
        layer_times = list(raw['execution_time']['layer_compute_total_ms'])
        return {'time': {'layer-computes': layer_times,
                         'fb_sync': raw['execution_time']['forward_backward_time_ms'] - sum(layer_times)},
                'memory': raw['execution_memory']['layer_memory_total_mb']}

    def load_profile_data_all(self) -> Tuple[Dict, List[str]]:
        profile_data: Dict = {}
        device_types: List[str] = []
        layers: Optional[int] = None
        for name in self.profile_data_list:
            if self.validate and not _TYPE_RE.search(name):
                raise ProfileSchemaError(f"{name}: file name must look like DeviceType.<TYPE>_tp<N>_bs<M>.json")
            dev = _TYPE_RE.search(name).group(1)
            tp = _TP_RE.search(name).group(1)
            bs = _BS_RE.search(name).group(1)
            if f'DeviceType.{dev}' not in profile_data:
                profile_data[f'DeviceType.{dev}'] = {}
                device_types.append(dev)
            with open(f'{self.profile_dir}/{name}', 'r') as fh:
                raw = json.loads(fh.read())
            if self.validate:
                layers = validate_profile(raw, name, layers)
            if 'model' not in profile_data:
                profile_data['model'] = self._model_section(raw)
            profile_data[f'DeviceType.{dev}'][f'tp{tp}_bs{bs}'] = self._device_section(raw)
        return profile_data, device_types
