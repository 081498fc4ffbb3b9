"""Repeat the graphed-vs-eager trajectory comparison on the golden config; prints the loss at each step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vibravox_b200
from vibravox_b200 import data as O
gold = torch.load("tests/golden/train_step.pt")
body, air = O.synthetic_pairs(gold["B"], gold["S"], seed=gold["data_seed"])
batch = {"audio_body_conducted": body.cuda(), "audio_airborne": air.cuda()}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
def run(graphed):
    lm = vibravox_b200.build_model(seed=gold["model_seed"], device="cuda")
    tr = []
    for it in range(7):
        (lm.training_step_graphed if graphed else lm.training_step)(batch)
        tr.append(float(lm.logged["train/generator/backprop_loss"]))
    return tr
for r in range(reps):
    print("eager  ", " ".join(f"{v:9.3f}" for v in run(False)))
for r in range(reps):
    print("graphed", " ".join(f"{v:9.3f}" for v in run(True)))
