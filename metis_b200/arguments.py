"""Command-line flags of the two search CLIs (same 17 names and types as arguments.py:16-49)."""
from __future__ import annotations

import argparse
from typing import Optional, Sequence

_FLAGS = (
    # model (arguments.py:16-21)
    ('model_name', str), ('model_size', str), ('num_layers', int), ('gbs', int),
    # gpt model (:23-28)
    ('hidden_size', int), ('sequence_length', int), ('vocab_size', int), ('attention_head_size', int),
    # cluster (:30-33)
    ('hostfile_path', None), ('clusterfile_path', None),
    # search (:42-49)
    ('profile_data_path', None), ('max_profiled_tp_degree', int), ('max_profiled_batch_size', int),
    ('min_group_scale_variance', int), ('max_permute_len', int),
    # environment (:36-39)
    ('log_path', None), ('home_dir', None),
)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser()
    for name, kind in _FLAGS:
        if kind is None:
            parser.add_argument(f'--{name}')
        else:
            parser.add_argument(f'--{name}', type=kind)
    return parser


def parse_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    return build_parser().parse_args(argv)
