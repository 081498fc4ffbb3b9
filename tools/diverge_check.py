"""Run N training steps on the same synthetic batch with the CUDA path and (optionally) the CPU oracle and
print the logged losses, to tell genuine GAN divergence on noise inputs from a kernel bug."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vibravox_b200
from oracle import eben_oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
with_oracle = len(sys.argv) > 3 and sys.argv[3] == "oracle"
body, air = O.synthetic_pairs(B, 48000, seed=42)
lm = vibravox_b200.build_model(seed=42, device="cuda")
batch = {"audio_body_conducted": body.cuda(), "audio_airborne": air.cuda()}
orc = O.OracleEBENStep(seed=42) if with_oracle else None
torch.set_num_threads(os.cpu_count())
for it in range(N):
    lm.training_step(batch)
    logs = {k.split("/", 1)[1]: float(v) for k, v in lm.logged.items()}
    line = f"{it:3d} gpu: " + " ".join(f"{k.split('/')[-1][:10]}={v:.4g}" for k, v in logs.items())
    line += " | norms " + " ".join(f"{float(x):.3g}" for x in lm.last_norms) + " lam " + " ".join(f"{float(x):.3g}" for x in lm.last_lambdas)
    if orc is not None:
        ol = orc.step(body, air)
        line += "\n    cpu: " + " ".join(f"{k.split('/')[-1][:10]}={v:.4g}" for k, v in ol.items())
        line += " | norms " + " ".join(f"{x:.3g}" for x in orc.last["norms"])
    print(line, flush=True)
    if any(v != v for v in logs.values()):
        print("NaN at step", it)
        break
