"""Two runs of N training steps from the same seed on the same batch: prints both loss trajectories side by side.
Split-K reductions use fp32 atomics, so the runs differ at the 1e-7 level on step 0 and drift apart from there;
anything larger on the first steps points at a race."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vibravox_b200
from vibravox_b200 import data as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
body, air = O.synthetic_pairs(B, 48000, seed=42)
batch = {"audio_body_conducted": body.cuda(), "audio_airborne": air.cuda()}
runs = []
for r in range(2):
    lm = vibravox_b200.build_model(seed=42, device="cuda")
    tr = []
    for it in range(N):
        lm.training_step(batch)
        tr.append((float(lm.logged["train/generator/backprop_loss"]), float(lm.logged["train/discriminator/backprop_loss"])))
    runs.append(tr)
for it in range(N):
    a, b = runs[0][it], runs[1][it]
    print(f"{it:3d}  G {a[0]:.7f} {b[0]:.7f} (rel {abs(a[0]-b[0])/abs(a[0]):.1e})   D {a[1]:.7f} {b[1]:.7f} (rel {abs(a[1]-b[1])/abs(a[1]):.1e})")
