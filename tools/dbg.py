import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
eng = B.Engine()
b = w.config1(4)
rec, cig = eng.align(b)
print(rec)
