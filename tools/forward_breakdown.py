"""Where the forward's time goes when its two streams overlap: times the whole forward (CUDA-graph replay, bench shape) with parts
REMOVED through the profiling build's PMCE_SKIP mask (results are garbage, only the time is meaningful):
    bit 0 no recurrent GRU steps, bit 1 no image-feature stream, bit 2 no lifter, bit 3 no decoder.
Needs the profiling library: PMCE_B200_PROFILING=1 python -m pmce_b200.build. Each mask runs in its own process (knobs are read once)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(B):
    import torch
    sys.path.insert(0, REPO)
    import bench
    from pmce_b200 import synth
    from tools.stage_times import timed
    dev = torch.device("cuda")
    model, _ = bench.build_model(dev)
    model.engine().use_graph = False
    p2d, feat = [t.to(dev) for t in synth.make_inputs(B, bench.T, bench.J, seed=3)]
    print(json.dumps({"us": round(timed(lambda: model(p2d, feat)), 1)}))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    if os.environ.get("_PMCE_BREAKDOWN_CHILD") == "1":
        return child(B)
    names = {0: "whole forward", 1: "without the recurrent GRU steps", 2: "without the image-feature stream", 4: "without the lifter",
             6: "decoder alone", 10: "lifter alone", 12: "image-feature stream alone", 8: "without the decoder"}
    for mask, name in names.items():
        env = dict(os.environ, PMCE_B200_PROFILING="1", PMCE_SKIP=str(mask), _PMCE_BREAKDOWN_CHILD="1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), str(B)], env=env, capture_output=True, text=True, timeout=300)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
        print(f"PMCE_SKIP={mask:2d}  {name:40s} {line}")


if __name__ == "__main__":
    main()
