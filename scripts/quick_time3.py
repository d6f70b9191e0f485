"""Scratch: the three timing shapes that matter for A/B runs of the tensor-core kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_time import run
run(16, 862, 8, True)
run(64, 862, 28, False)
run(32, 5168, 8, False)
