"""Profile JSON loader (mirrors data_loader.py:10-61 of the reference; same dict schema)."""
from __future__ import annotations

import json
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

_TYPE_RE = re.compile(r"DeviceType\.(\w+?)_")
_TP_RE = re.compile(r"tp(\d+)")
_BS_RE = re.compile(r"bs(\d+)")


class ProfileDataLoader:
    def __init__(self, profile_dir: str, file_order: Optional[Sequence[str]] = None):
        """``file_order`` (extension) pins the directory listing order: the first listed file
        supplies profile_data['model'] and the first listed type the layer weights (quirk Q3)."""
        self.profile_dir = profile_dir
        listed = [f for f in os.listdir(profile_dir) if f.endswith('.json')]
        if file_order is not None:
            if sorted(file_order) != sorted(listed):
                raise ValueError('file_order does not match the .json files in the profile directory')
            listed = list(file_order)
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
        for name in self.profile_data_list:
            dev = _TYPE_RE.search(name).group(1)
            tp = _TP_RE.search(name).group(1)
            bs = _BS_RE.search(name).group(1)
            if f'DeviceType.{dev}' not in profile_data:
                profile_data[f'DeviceType.{dev}'] = {}
                device_types.append(dev)
            with open(f'{self.profile_dir}/{name}', 'r') as fh:
                raw = json.loads(fh.read())
            if 'model' not in profile_data:
                profile_data['model'] = self._model_section(raw)
            profile_data[f'DeviceType.{dev}'][f'tp{tp}_bs{bs}'] = self._device_section(raw)
        return profile_data, device_types
