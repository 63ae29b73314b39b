import os, sys, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from xva_trainer_b200 import ops
B, T = 32, 880
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
x = r(B, T, 384); w1 = r(3, 1536, 384) * 0.03; b1 = r(1536)
out = torch.empty(B, T, 1536, device="cuda")
def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        rows.append(p.stdout.readline().strip())
    p.terminate()
fn = lambda: ops.conv_fwd(x, w1, (-1, 0, 1), bias=b1, relu=True, out=out)
for _ in range(5): fn()
torch.cuda.synchronize()
stop, rows = threading.Event(), []
th = threading.Thread(target=sample, args=(stop, rows)); th.start()
time.sleep(0.3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5000
e0.record()
for _ in range(n): fn()
e1.record(); torch.cuda.synchronize()
stop.set(); th.join()
ms = e0.elapsed_time(e1) / n
print(f"{os.environ.get('TAG')}: {ms*1e3:.1f} us {2*B*T*1536*1152/ms/1e9:.1f} TF/s; clocks/power: {rows[4:-1:3]}")
