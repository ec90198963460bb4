"""Developer tool: run bench.py's k1_micro leg alone, optionally with the trainer's global state switched on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import vlpet_b200 as V
import vlpet_b200.functional as F_
if "--direct" in sys.argv:
    F_.set_direct_grad_accumulation(True)
for rep in range(3):
    print(bench.k1_micro(V, F_, 6527.1), flush=True)
