"""Synthetic inputs live in agrl/pytorch_b200/synthetic.py (bench.py needs them without importing oracle/)."""
from agrl.pytorch_b200.synthetic import *  # noqa: F401,F403
