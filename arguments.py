"""Same module name as the reference's arguments.py; the flag table lives in metis_b200.arguments."""
from metis_b200.arguments import build_parser, parse_args  # noqa: F401
