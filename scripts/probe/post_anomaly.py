"""Per-step stage times of the detector pipeline (is a stage time an average of outliers?)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from gencomm_b200 import pipeline

dev = "cuda"
F, N, P = 8, 4, 100000
pipe = pipeline.DetectorPipeline(F, N, P, shape="opv2v_h", fusion="att", device=dev)
sets = []
for s in range(3):
    p, pw = pipeline.synthetic_detector_inputs(50 + s, F, N, P, pipe.lidar_range)
    sets.append((torch.from_numpy(p).to(dev), torch.from_numpy(pw).to(dev)))
pipe.enable_stage_timing()
for i in range(4):
    pipe.step(*sets[i % 3])
torch.cuda.synchronize()
pipe._marks.clear()
rows = []
for i in range(12):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.step(*sets[i % 3])
    e1.record()
    torch.cuda.synchronize()
    st = pipe.stage_ms(1)
    rows.append((e0.elapsed_time(e1), st))
for ms, st in rows:
    print(f"{ms:7.2f}  " + "  ".join(f"{k[:5]}={v:5.2f}" for k, v in st.items()))
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_reserved() / 2**30, "GiB reserved")
import time
pipe._marks = None
torch.cuda.synchronize()
t = []
for i in range(12):
    t0 = time.perf_counter()
    pipe.step(*sets[i % 3])
    t.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
print("host enqueue ms per step (no sync, GPU running behind):", " ".join(f"{x:.2f}" for x in t))
