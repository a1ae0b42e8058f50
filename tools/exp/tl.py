import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
which = sys.argv[1] if len(sys.argv) > 1 else "d1"
sys.argv = [sys.argv[0]]
import importlib.util
spec = importlib.util.spec_from_file_location("bc", os.path.join(os.path.dirname(__file__), "..", "bench_conv.py"))
bc = importlib.util.module_from_spec(spec); spec.loader.exec_module(bc)
bc.timeit = lambda fn, reps=1: (fn(), torch.cuda.synchronize(), 0.001)[2]
B = 64
if which == "d1": bc.run_dgrad(B, 128, 128, 128, 224, 256, 1, 2)
if which == "f1": bc.run(B, 128, 128, 224, 256, 128, 1, 1, 1)
if which == "f1p": bc.run(B, 128, 128, 128, 256, 128, 1, 0, 0)
if which == "f3": bc.run(B, 128, 128, 128, 128, 32, 3, 1, 1)
if which == "d3": bc.run_dgrad(B, 128, 128, 32, 128, 128, 3, 0)
if which == "w1": bc.run_wgrad(128, 32, 32, 768, 1024, 128, 1, 1)
if which == "w1b": bc.run_wgrad(128, 128, 128, 224, 256, 128, 1, 1)
