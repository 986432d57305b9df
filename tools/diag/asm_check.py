"""diagnostic: normal equations of the tile gather against the atomic-scatter kernels on the same point (LVI_ASM_CHECK=1 prints the differences)"""
import os, sys
os.environ["LVI_ASM_CHECK"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lvi_exc_b200.backend import CudaBackend, CudaProblem
from tests import problems
b = CudaBackend(0)
for stage in sys.argv[1:] or ["so3", "surfel", "lvi", "lvi_locked"]:
    pd = problems.make_lvi_problem(stage)
    prob = CudaProblem(b, pd)
    print("stage", stage, flush=True)
    prob.evaluate(gradient=True)
