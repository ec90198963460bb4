"""Where does a training step go?  torch.profiler over one task cycle: CUDA kernel time vs wall time, top kernels."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlpet_b200.host as H
from torch.profiler import profile, ProfilerActivity

BS = int(sys.argv[1]) if len(sys.argv) > 1 else 300     # per-rank --batch-size (38 ~ one of 8 ranks of the bs=300 bench)
torch.manual_seed(0)
cfg = H.bart_base_vlpet_large(assume_no_padding=True)
model = H.VLBart(cfg).train()
tr = H.PetTrainer(model, cfg, "cuda", total_steps=20000)
tr.step_idx = 2000
cyc = [{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()} for b in H.multitask_cycle(BS, H.TASKS)]
for i in range(8):
    tr.train_step(cyc[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(8):
    tr.train_step(cyc[i % 4])
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 8
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4):
        tr.train_step(cyc[i % 4])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev) / 4 / 1e3
by = {}
for e in ev:
    by[e.name] = by.get(e.name, 0) + e.device_time
gtr = H.GraphedPetTrainer(H.VLBart(cfg).train(), cfg, torch.device("cuda", 0), lr=1e-3, total_steps=20000)
gtr.set_step(2000)
for i in range(12):
    gtr.train_step(cyc[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(16):
    gtr.train_step(cyc[i % 4])
torch.cuda.synchronize()
print(f"batch-size {BS}: graphed wall per step {(time.perf_counter() - t0) / 16 * 1e3:.2f} ms")
print(f"wall per step {wall*1e3:.2f} ms; sum of CUDA kernel time per step {tot:.2f} ms; kernels per step {len(ev)/4:.0f}")
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:28]:
    print(f"  {v/4/1e3:8.3f} ms/step  {k[:110]}")
