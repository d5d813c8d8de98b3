"""oracle -- CPU checkers for the sample-based hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (sample-based-gnn_b200/, include/) never does.
"""
from .oracle import *  # noqa: F401,F403
