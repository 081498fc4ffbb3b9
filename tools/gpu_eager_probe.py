"""One-off probe: the eager cuDNN baseline (tf32 / ieee) and the CPU port at the bench shape."""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
a = types.SimpleNamespace(batch=int(sys.argv[1]) if len(sys.argv) > 1 else 32, seconds=3.0, workload="bwe", cpu_batch=32)
print(json.dumps(bench.gpu_eager_baseline(a, "cuda:0")), flush=True)
if "--cpu" in sys.argv:
    print(json.dumps(bench.cpu_baseline(a, steps=2, warmup=1)), flush=True)
