import os, sys, traceback, types, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vibravox_b200
from vibravox_b200 import ops, functional
from oracle import eben_oracle as O

import ctypes, glob
_rt = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so.12"))[0])
state = {"bad": False, "n": 0, "last_ok": None}
def status():
    s = torch.cuda.current_stream()
    out = ctypes.c_int(0)
    e = _rt.cudaStreamIsCapturing(ctypes.c_void_p(s.cuda_stream), ctypes.byref(out))
    return (e, out.value)
def wrap(name, fn):
    def w(*a, **k):
        out = fn(*a, **k)
        if not state["bad"]:
            st = status()
            state["n"] += 1
            code = int(st[1]) if isinstance(st, tuple) else -1
            if isinstance(st, tuple) and (int(st[0]) != 0 or code == 2):
                state["bad"] = True
                print("INVALIDATED after op", name, "n", state["n"], "status", st, "thread", threading.current_thread().name,
                      "last ok", state["last_ok"], flush=True)
                traceback.print_stack(limit=12)
            else:
                state["last_ok"] = (name, code, threading.current_thread().name)
        return out
    return w
for n in dir(ops):
    f = getattr(ops, n)
    if isinstance(f, types.FunctionType) and not n.startswith("_") and n not in ("require_cuda", "use_tc", "get_pack"):
        setattr(ops, n, wrap(n, f))
print("status sample outside capture:", status())
B, S = int(sys.argv[1]), int(sys.argv[2])
body, air = O.synthetic_pairs(B, S, seed=1)
batch = {"audio_body_conducted": body.cuda(), "audio_airborne": air.cuda()}
lm = vibravox_b200.build_model(seed=42, device="cuda")
try:
    for it in range(5):
        lm.training_step_graphed(batch)
    torch.cuda.synchronize()
    print("OK", B, S, lm.graph_launches(), float(lm.logged["train/generator/backprop_loss"]))
except Exception as e:
    print("FAIL", B, S, str(e).splitlines()[0])
