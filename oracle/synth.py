"""Seeded synthetic inputs (moved to autoposeestimation_b200/synthetic.py so that bench.py and the
product never import oracle/); re-exported here for the tests and the oracle tools."""
from autoposeestimation_b200.synthetic import *  # noqa: F401,F403
from autoposeestimation_b200.synthetic import INTR, DEPTH_SCALE  # noqa: F401
