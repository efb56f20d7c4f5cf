import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autometa_b200 import synth
from autometa_b200.binning import recursive_dbscan as R
table, markers = synth.make_sweep_input()
R.recursive_dbscan(table.copy(), markers, 20.0, 95.0, 25.0, 5.0)
tb = table.copy()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
cl, un = R.recursive_dbscan(tb, markers, 20.0, 95.0, 25.0, 5.0)
dt = time.perf_counter() - t0
pr.disable()
print("api seconds", dt, cl.shape, un.shape)
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
