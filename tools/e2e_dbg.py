import os, sys
sys.path.insert(0, '/root/repo')
import torch, vibravox_b200
from bench import synthetic_pairs
dev = torch.device("cuda", 0)
B, S = 32, 48000
lm = vibravox_b200.build_model(seed=42, device=dev)
body_h, air_h = synthetic_pairs(B, S, 42)
body_h, air_h = body_h.pin_memory(), air_h.pin_memory()
body_d, air_d = body_h.to(dev), air_h.to(dev)
batch = {"audio_body_conducted": body_d, "audio_airborne": air_d}
mode = sys.argv[1]
stage_body, stage_air = torch.empty_like(body_d), torch.empty_like(air_d)
loss_h = torch.empty(2, dtype=torch.float32).pin_memory()
for it in range(30):
    if mode == "e2e":
        stage_body.copy_(body_h, non_blocking=True)
        stage_air.copy_(air_h, non_blocking=True)
        lm.training_step({"audio_body_conducted": stage_body, "audio_airborne": stage_air})
    else:
        lm.training_step(batch)
    loss_h[0:1].copy_(lm.logged["train/generator/backprop_loss"].view(1), non_blocking=True)
    loss_h[1:2].copy_(lm.logged["train/discriminator/backprop_loss"].view(1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    print(it, mode, float(loss_h[0]), float(loss_h[1]), float(lm.logged["train/generator/reconstructive_loss_freq"]), flush=True)
