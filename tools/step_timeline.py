"""Kernel timeline of one training step (torch.profiler / CUPTI): per-stream busy time, overlap, and the longest kernels
with start offsets - to see which chain is the critical path.  usage: python tools/step_timeline.py [graph]"""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vibravox_b200
from vibravox_b200 import data as D
from torch.profiler import profile, ProfilerActivity

graphed = len(sys.argv) > 1 and sys.argv[1] == "graph"
body, air = D.synthetic_pairs(32, 48000, seed=1)
lm = vibravox_b200.build_model(seed=42, device="cuda")
batch = {"audio_body_conducted": body.cuda(), "audio_airborne": air.cuda()}
step = lm.training_step_graphed if graphed else lm.training_step
for _ in range(5 if graphed else 3):
    step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(batch)
    torch.cuda.synchronize()
path = "gpurun_out/step_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
t1 = max(e["ts"] + e["dur"] for e in ev)
print(f"kernels {len(ev)}  span {(t1 - t0) / 1e3:.2f} ms  sum of durations {sum(e['dur'] for e in ev) / 1e3:.2f} ms")
streams = collections.defaultdict(list)
for e in ev:
    streams[e["args"].get("stream")].append(e)
for s, es in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print(f"stream {s}: {len(es)} kernels, busy {sum(e['dur'] for e in es) / 1e3:.2f} ms, first {(es[0]['ts'] - t0) / 1e3:.2f} last {(es[-1]['ts'] + es[-1]['dur'] - t0) / 1e3:.2f}")
# concurrency profile: time with k kernels running
pts = []
for e in ev:
    pts.append((e["ts"], 1)); pts.append((e["ts"] + e["dur"], -1))
pts.sort()
cur, last, hist = 0, t0, collections.Counter()
for t, d in pts:
    hist[cur] += t - last
    cur += d; last = t
print("time (ms) with k kernels in flight:", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})
# phases: coarse 1-ms buckets listing the dominant kernel per stream
print("longest 40 kernels: start(ms) dur(us) stream name")
for e in sorted(ev, key=lambda e: -e["dur"])[:40]:
    print(f"  {(e['ts'] - t0) / 1e3:7.2f} {e['dur']:8.1f} {e['args'].get('stream')} {e['name'][:70]} grid={e['args'].get('grid')}")
# time each kernel spends as the ONLY kernel in flight, grouped by (name, grid): the exposed part of the critical path
evs = sorted(ev, key=lambda e: e["ts"])
bounds = sorted(set([e["ts"] for e in evs] + [e["ts"] + e["dur"] for e in evs]))
alone = collections.Counter()
total = collections.Counter()
count = collections.Counter()
import bisect
starts = [e["ts"] for e in evs]
active = []
j = 0
for a, b in zip(bounds[:-1], bounds[1:]):
    while j < len(evs) and evs[j]["ts"] <= a:
        active.append(evs[j]); j += 1
    active = [e for e in active if e["ts"] + e["dur"] > a]
    if len(active) == 1:
        e = active[0]
        alone[(e["name"][:60], str(e["args"].get("grid")))] += b - a
for e in evs:
    k = (e["name"][:60], str(e["args"].get("grid")))
    total[k] += e["dur"]; count[k] += 1
print("exposed (alone in flight) time by kernel and grid: alone_ms total_ms n name grid")
for k, v in sorted(alone.items(), key=lambda kv: -kv[1])[:45]:
    print(f"  {v / 1e3:7.3f} {total[k] / 1e3:7.3f} {count[k]:4d} {k[0]} {k[1]}")
# gaps on the whole device (no kernel running)
print("idle gaps > 20 us:", [(round((a - t0) / 1e3, 2), round(b - a)) for a, b in []])
