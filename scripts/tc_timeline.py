"""Print the tcgen05 kernel's per-step timeline (cycles) of CTA 0.  Run with NMR_TC_TIMELINE=1 on the GPU box."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfmatch_b200 import _lib as L
from nerfmatch_b200.engine import Engine, EngineConfig
from nerfmatch_b200.synthetic import make_state_dict

embed = sys.argv[1] if len(sys.argv) > 1 else "normal"
from nerfmatch_b200.config import build_renderer
from nerfmatch_b200.synthetic import K_7SCENES, make_pose

sd = make_state_dict(0, embed_type=embed, sigma_gain=120.0, sigma_bias=(8.0, 8.0))
case = dict(embed=embed, nc=64, nf=128, app=False, white_bg=False, stop_layer=3, feat_comb="lin")
ren = build_renderer(case, sd, precision="bf16", device="cuda:0")
ren.rng = "philox"
eng = ren.engine(torch.device("cuda:0"))
rays, _, _ = eng.raygen(K_7SCENES, make_pose(1), 480, 640, 4)     # 19200 rays: one chunk, fine pass = last launch
for _ in range(2):
    ren.predict(rays, 1, 1)
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 9 * 4 + 64))()
FINE = 8 * 9 * 4 + 16
n = L.load().nmr_debug_timeline(buf, len(buf))
t = np.array(buf[:8 * 9 * 4], dtype=np.int64).reshape(8, 9, 4)
t0 = t[1:6]
print("per step (cycles), iterations 1..5 averaged")
print("step  mma_issue  mma_end->epi_start  epilogue  epi_end->next_mma_start   total")
for s in range(9):
    issue = (t0[:, s, 1] - t0[:, s, 0]).mean()
    lat1 = (t0[:, s, 2] - t0[:, s, 1]).mean()
    epi = (t0[:, s, 3] - t0[:, s, 2]).mean()
    nxt = t[1:6, s + 1, 0] if s < 8 else t[2:7, 0, 0]
    lat2 = (nxt - t0[:, s, 3]).mean()
    print(f"{s:4d} {issue:10.0f} {lat1:18.0f} {epi:9.0f} {lat2:22.0f} {issue + lat1 + epi + lat2:8.0f}")
en = np.array(buf[8 * 9 * 4:8 * 9 * 4 + 16], dtype=np.int64).reshape(4, 4)
print("encoder warps, iterations 2..5 (cycles): encode next tile, wait for the views partial sums, combine + local compositing")
for i in range(4):
    print(f"{i + 2:4d} {en[i, 1] - en[i, 0]:8d} {en[i, 2] - en[i, 1]:8d} {en[i, 3] - en[i, 2]:8d}")
print("iteration period:", np.diff(t[1:7, 0, 0]).mean())

fine = np.array(buf[FINE:FINE + 36], dtype=np.int64).reshape(9, 4)
if not fine[1:8].all():
    sys.exit(0)          # this build carries no fine-grained stamps
print("iteration 3, hand-over of layer s -> s+1 (cycles): epilogue wake -> first slice in registers -> chunk 0 handed over -> MMA warp has it;"
      " chunk 0 -> chunk 7 seen by the MMA warp; last MMA issue of layer s -> epilogue wake")
for s_ in range(1, 8):
    wake = t[3, s_, 2]
    print(f"{s_:4d}  ld {fine[s_, 0] - wake:6d}  cvt+st+arrive {fine[s_, 1] - fine[s_, 0]:6d}  seen {fine[s_ + 1, 2] - fine[s_, 1]:6d}"
          f"  | chunks 0->7 {fine[s_ + 1, 3] - fine[s_ + 1, 2]:6d}  | last issue -> wake {wake - t[3, s_, 1]:6d}  | first issue(s+1) - wake {t[3, s_ + 1, 0] - wake:6d}")
