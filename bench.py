#!/usr/bin/env python
"""Benchmark of the PMCE per-clip forward hot path on B200 (driver contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = `PMCE.forward` over one batch of B synthetic clips per GPU (T=16 frames each). Rank 0 prints ONE JSON
line. `--impl reference` times the reference algorithm's CPU implementation (the oracle port, torch CPU, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

# headline workload: BASELINE.json configs[1] / north_star target ("B=64, T=16 ... C=512")
B_PER_GPU, T, J, C, DEPTH = 64, 16, 17, 512, 3
V = 6890
METRIC = "clips/sec (B x T=16 frames) through PMCE.forward"
WORKLOAD = f"PMCE.forward (lifter+GRU+CoEvoDecoder+upsample) B={B_PER_GPU}/GPU T={T} J={J} V={V} C={C}"


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock / throttle-reason sampling through NVML (nvidia_ml_py) from a thread, every few milliseconds, started before
    the warm-up and stopped after the last timed region; the summary covers the samples taken inside the timed windows
    (`window()`), so even a 50 ms region holds >= 5 samples. Falls back to `nvidia-smi -lms` (B200_PROFILING.md recipe) when
    NVML cannot be loaded. Works under torch.distributed.run: the device is looked up by the UUID of this rank's GPU."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
               (0x80, "hw_power_brake_slowdown"))

    def __init__(self, torch_device_index, period_s=0.004):
        self.rows, self.windows, self.period = [], [], period_s
        self.idx, self.h, self.nv, self.th, self.stop_flag, self.proc = torch_device_index, None, None, None, False, None

    def start(self):
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            try:
                self.h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.nv = nv
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.nv = None
            self._start_smi()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(self.period)

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def _start_smi(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mx = None

            def rd():
                for line in self.proc.stdout:
                    r = [c.strip() for c in line.split(",")]
                    try:
                        bits = sum(b for (b, _), v in zip(((0x8, 0), (0x40, 0), (0x20, 0), (0x4, 0)), r[5:9]) if v.lower().startswith("active"))
                        self.rows.append((time.perf_counter(), float(r[1]), bits, float(r[3])))
                        self.mx = float(r[2])
                    except Exception:
                        continue
            self.th = threading.Thread(target=rd, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def window(self, t0, t1):
        """Register a timed region [t0, t1] (time.perf_counter values taken right around it on the host)."""
        self.windows.append((t0, t1))

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        if self.th:
            self.th.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        inside = [r for r in self.rows if any(a <= r[0] <= b for a, b in self.windows)] if self.windows else list(self.rows)
        used = inside if len(inside) >= 3 else list(self.rows)
        sm = sorted(r[1] for r in used)
        bits = 0
        for r in used:
            bits |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.mx, "power_w_max": max(r[3] for r in used),
                "samples": len(used), "samples_total": len(self.rows), "in_timed_regions": used is inside,
                "source": "nvml" if self.nv else "nvidia-smi", "reasons": [n for b, n in self.REASONS if bits & b]}


def build_model(device):
    import numpy as np
    import torch
    from pmce_b200 import synth
    from pmce_b200 import build as _b
    _b.build()
    root = tempfile.mkdtemp(prefix="pmce_bench_")
    synth.prepare_data_root(root, os.path.join(REPO, "tests", "golden", "J_regressors_sparse.npz"))
    os.environ["PMCE_DATA_ROOT"] = root
    os.environ["PMCE_B200_STANDALONE_CFG"] = "1"
    from pmce_b200 import models
    from pmce_b200.config import cfg
    cfg.DATASET.seqlen = T
    m = models.PMCE.get_model(J, C, DEPTH)
    sd = synth.make_state_dict(0, init_vertices=m.state_dict()["pose_mesh_coevo.init_vertices"].numpy(), lifter_out_scale=300.0,
                               num_joint=J, embed_dim=C, depth=DEPTH, seqlen=T)
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval(), sd


def _cpu_forward_fn(sd, vj_port):
    """The reference's CPU implementation of the path: the UNMODIFIED reference module when it is importable here
    (`/root/reference` in the build container, `oracle/_ref` - materialised by oracle/build_ref.py - on the GPU box), else the
    oracle port. -> (callable(p2d, feat), kind)"""
    import torch
    from oracle import ref_harness as rh
    if rh.which() is not None:
        model = rh.build_pmce(J, C, DEPTH, T)            # models.PMCE.get_model of the reference, eval mode, CPU
        model.load_state_dict(sd, strict=True)
        return (lambda p2d, feat: model(p2d, feat)), "reference"
    from oracle import pmce_oracle as po
    return (lambda p2d, feat: po.pmce_forward(sd, p2d, feat, vj_port, depth=DEPTH)), "port"


def best_cpu_threads(fwd):
    """The box may expose more logical CPUs than the container can use (oversubscription makes torch CPU slower, not
    faster): give the reference its best shot by calibrating the thread count on a small batch."""
    import torch
    from pmce_b200 import synth
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cands = sorted({max(1, avail), max(1, avail // 2), max(1, avail // 4), min(avail, 32), min(avail, 16), min(avail, 8)})
    p2d, feat = synth.make_inputs(8, T, J, seed=2)
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for n in cands:
            torch.set_num_threads(n)
            fwd(p2d, feat)
            t0 = time.perf_counter()
            fwd(p2d, feat)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = n, dt
    torch.set_num_threads(best)
    return best, avail


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of PMCE.forward on the host cores (rank 0 only), on the same
    workload (one B=64 batch per step), bounded to a few steps."""
    if rank != 0:
        return
    os.environ["CUDA_VISIBLE_DEVICES"] = ""        # this arm is the CPU implementation: the reference's hard-coded .cuda() calls stay on the host
    import numpy as np
    import torch
    from pmce_b200 import synth
    g = np.load(os.path.join(REPO, "tests", "golden", f"pmce_J{J}_C{C}_T{T}_B2.npz"))
    sd = synth.make_state_dict(0, init_vertices=g["init_vertices"], lifter_out_scale=300.0, num_joint=J, embed_dim=C, depth=DEPTH, seqlen=T)
    fwd, kind = _cpu_forward_fn(sd, g["vj_relation"])
    cores, avail = best_cpu_threads(fwd)
    p2d, feat = synth.make_inputs(B_PER_GPU, T, J, seed=1)
    steps = max(1, min(args.steps, 6))
    warm = max(1, min(args.warmup, 2))
    with torch.no_grad():
        for _ in range(warm):
            out = fwd(p2d, feat)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fwd(p2d, feat)
        dt = time.perf_counter() - t0
    assert tuple(out[0].shape) == (B_PER_GPU, V, 3)
    val = B_PER_GPU * steps / dt
    what = ("the unmodified reference module (lib/models/PMCE.py:15-26 via oracle/ref_harness.py)" if kind == "reference"
            else "the oracle port (reference module not importable here)")
    sample = (f"{steps} steps of the same B={B_PER_GPU} batch through {what} (bounded from --steps {args.steps}); torch CPU fp32, "
              f"{cores} threads (best of a calibration sweep; {avail} logical CPUs visible)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": bench_config(args.gpus),      # the own arm's config, key for key
        "note": "reference arm: one B=64 batch per step through the reference's CPU implementation on the host cores (rank 0 only)",
        "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def cpu_baseline_subprocess(steps=6, warmup=2, timeout=240):
    """cpu_baseline of the own arm = the reference arm run in a CHILD process (no CUDA context, its own thread settings)."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", str(warmup)],
                       env=env, capture_output=True, text=True, timeout=timeout)
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)["cpu_baseline"]
    raise RuntimeError("reference arm produced no JSON line: " + r.stderr[-500:])


def bench_config(world, note=None):
    """`config` of the JSON line; the reference arm prints the same keys as the own arm."""
    cfg = {"workload": WORKLOAD, "global_batch": world * B_PER_GPU,
           "parallelism": f"batch-shard x{world}" + (" + 1 all-gather" if world > 1 else ""),
           "weights": "seeded random init (no checkpoints are published)",
           "l2": "per-step working set (weights 0.46 GB + activations) exceeds the 126 MB L2; inputs rotate over 4 resident sets",
           "cuda_graph": True,
           "in_flight": "2 forwards (models.PMCE.forward_iter: two buffer slots, each with its own workspace, graph and stream)"}
    if note:
        cfg["note"] = note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="pmce_b200", choices=["pmce_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pmce_b200 import synth, _lib
    from pmce_b200 import dist as pdist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    W = max(args.warmup, 3)
    K = args.steps
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    model, sd = build_model(dev)
    eng = model.engine()
    lib = _lib.load()
    B = B_PER_GPU
    NSETS = 4
    sets = [tuple(t.to(dev) for t in synth.make_inputs(B, T, J, seed=100 + rank * NSETS + i)) for i in range(NSETS)]
    host_sets = [tuple(t.pin_memory() for t in synth.make_inputs(B, T, J, seed=100 + rank * NSETS + i)) for i in range(NSETS)]

    # N > 1: the forward writes straight into this rank's row of an all-gather buffer; the single collective of the path (an
    # in-place all-gather of the per-rank outputs) runs on a communication stream under the next step's forward
    sf = pdist.ShardedForward(model, B, J, dev) if world > 1 else None

    def step(i):
        p2d, feat = sets[i % NSETS]
        if sf is not None:
            return sf.step(p2d, feat)
        return model(p2d, feat)[0]

    host_out = (torch.empty(B, V, 3).pin_memory(), torch.empty(B, J, 3).pin_memory(), torch.empty(B, J, 3).pin_memory())

    def step_e2e(i):
        hp, hf = host_sets[i % NSETS]
        return model.forward_host(hp, hf, host_out)      # H2D of this step's inputs, forward, D2H of all three outputs, sync

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # launches per forward (counted inside the library on one eager pass; graph replays execute the same kernel nodes)
    eng.use_graph = False
    n0 = lib.pmce_launch_count()
    model(*sets[0])
    launches_per_step = int(lib.pmce_launch_count() - n0)
    eng.use_graph = True

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                  # sampling runs from the warm-up to the end of the last timed region
    # multi-GPU hooks of the pipelined loops: the forwards write into the rank's rows of the two all-gather buffers, every step's
    # all-gather is issued from the slot's forward stream onto the communication stream
    hooks = {}
    if sf is not None:
        hooks = dict(out_slots=[sf.views(sf.bufs[k], rank) for k in range(2)], before_forward=sf.wait_slot_free, after_forward=sf.gather_async)

    def run_value(k):
        """k steps through the device-resident public loop (models.PMCE.forward_iter): inputs already in HBM, two forwards in flight"""
        n = 0
        for _ in model.forward_iter((sets[i % NSETS] for i in range(k)), **hooks):
            n += 1
        if sf is not None:
            for k2 in range(2):
                sf.result(k2)                                           # the last gathers are part of the timed region
        assert n == k
        return n

    def timed_block(run, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(k)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    run_value(W)
    t_a = time.perf_counter()
    ms = timed_block(run_value, K)
    sampler.window(t_a, time.perf_counter())
    value = world * B * K / (ms * 1e-3)
    # the same K steps as K separate module calls on one stream (one forward at a time), for comparison
    for i in range(W):
        step(i)
    ms_single = timed(step, K)

    # e2e: the public host-buffer API, pipelined over the K batches (H2D of batch i+1 / D2H of batch i-1 overlap forward i);
    # every step copies its own inputs from pinned host memory and its three outputs back, and the consumer reads them
    # e2e at N > 1 includes the all-gather of every step (same hooks as the device-resident loop)

    def run_e2e(k):
        acc = 0.0
        for mesh_h, pose_h, p3_h in model.forward_host_iter((host_sets[i % NSETS] for i in range(k)), **hooks):
            acc += float(mesh_h[0, 0, 0]) + float(pose_h[0, 0, 0])      # the host really reads each step's result
        if sf is not None:
            for k2 in range(2):
                sf.result(k2)                                           # the last gathers are part of the timed region
        return acc

    run_e2e(W)
    t_a = time.perf_counter()
    ms_e2e = timed_block(run_e2e, K)
    sampler.window(t_a, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    for i in range(W):
        step_e2e(i)
    ms_e2e_sync = timed(step_e2e, K)
    e2e_value = world * B * K / (ms_e2e * 1e-3)
    h2d = B * (T * J * 2 + T * 2048) * 4
    d2h = B * (V * 3 + 2 * J * 3) * 4

    # BASELINE.json configs[3]: B=1024 sharded over 8 GPUs (128 clips per GPU) + the all-gather, beside the weak-scaling line
    b1024 = None
    if world == 8:
        Bs = 1024 // world
        sets_s = [tuple(t.to(dev) for t in synth.make_inputs(Bs, T, J, seed=500 + rank * NSETS + i)) for i in range(NSETS)]
        sfs = pdist.ShardedForward(model, Bs, J, dev)
        fn = lambda i: sfs.step(*sets_s[i % NSETS])
        for i in range(W):
            fn(i)
        ms_s = timed(fn, K)
        b1024 = {"workload": f"B=1024 T={T} C={C} batch-sharded over 8 GPUs ({Bs} clips per GPU) + 1 all-gather", "value": 1024 * K / (ms_s * 1e-3),
                 "unit": "clips/s", "ms_per_step": ms_s / K}
        del sfs, sets_s

    # roofline of the dominant kernel, timed alone with CUDA events on the launch stream
    roof = dominant_kernel_roofline(lib, dev, peaks, B)
    roof_ca = cross_attn_roofline(lib, eng, dev, peaks, B)
    big = cross_attn_roofline(lib, eng, dev, peaks, 256, nsets=4)      # BASELINE.json configs[2] batch: 3.5 items per CTA
    roof_ca["at_B256"] = {k: big[k] for k in ("achieved", "frac", "us_per_launch", "achieved_kernel_io", "frac_kernel_io", "bytes_per_launch", "embed_mode",
                                              "kernel_io_bytes_per_launch", "clips_per_launch")}

    roof_lbs = lbs_roofline(lib, dev, peaks) if rank == 0 else None
    spin = spin_leg(lib, dev, peaks, cpu=(world == 1 and not args.no_cpu_baseline)) if rank == 0 else None

    out = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(world),
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K,
                "path": "models.PMCE.forward_host_iter: pinned host inputs -> H2D -> forward -> D2H of the 3 outputs, every step, copies of "
                        "neighbouring steps overlapped with the forward (the reference loop lib/core/base.py:218-238, pipelined)",
                "unpipelined": {"value": world * B * K / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync / K,
                                "path": "models.PMCE.forward_host: H2D -> forward -> D2H -> sync per step"}},
        "one_forward_at_a_time": {"value": world * B * K / (ms_single * 1e-3), "ms_per_step": ms_single / K,
                                  "path": "K separate models.PMCE.forward calls on one stream (CUDA-graph replay each)"},
        "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
        "config3_B1024_x8": b1024, "clocks": clocks, "roofline": roof, "roofline_cross_attn": roof_ca, "roofline_lbs": roof_lbs, "spin_feature_extractor": spin,
        "peaks": peaks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_subprocess()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(key, B):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed ncu-derived table
    `profiles/ncu_traffic.json` (written by profiles/collect.sh from one `ncu --set full` capture per kernel; never a literal
    in this file). None when the table has no entry for (kernel, batch)."""
    try:
        tab = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json")))
        e = tab.get(f"{key}@B{B}")
        return float(e["dram_read_bytes"] + e["dram_write_bytes"]) if e else None
    except Exception:
        return None


def dominant_kernel_roofline(lib, dev, peaks, B):
    """The kernel with the largest share of the step: the tcgen05 split-bf16 GEMM, on the lifter fc1 shape
    (tokens x 2C x C, bias + GELU fused). Timed alone with CUDA events on the launch stream, operands pre-split and
    larger than... resident in L2/HBM as in the forward. achieved = ALGORITHMIC flops 2*M*N*K / time; the kernel issues
    3 bf16 MMAs per product (bf16x3 split precision), so 1/3 of the bf16 peak is its ceiling."""
    import ctypes as Ct
    import torch
    M, N, K = B * T * J, 2 * C, C
    x = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    oh = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    xs = [torch.empty(M, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    wsp = [torch.empty(N, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: Ct.c_void_p(t.data_ptr())
    assert lib.pmce_split_bf16(P(x), M, K, P(xs[0]), P(xs[1]), st) == 0
    assert lib.pmce_split_bf16(P(w), N, K, P(wsp[0]), P(wsp[1]), st) == 0

    def call():
        Z = Ct.c_void_p(0)   # exactly the forward's fc1 call: bias + GELU fused, split-bf16 output by TMA store
        rc = lib.pmce_linear_tc_presplit(P(xs[0]), P(xs[1]), P(wsp[0]), P(wsp[1]), P(b), M, N, K, 1, Z, P(oh[0]), P(oh[1]), Z, st)
        assert rc == 0, lib.pmce_last_error()
    for _ in range(5):
        call()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    torch.cuda.synchronize(dev)
    sec = e0.elapsed_time(e1) * 1e-3 / n
    flops = 2.0 * M * N * K
    ach = flops / sec / 1e12
    return {"kernel": "linear_tc_kernel<256, TC_SPLIT_GELU, pair> (tcgen05 cta_group::2 / TMA / TMEM split-bf16 GEMM; lifter fc1 + bias + GELU)", "bound": "tensor", "achieved": ach,
            "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
            "traffic": ncu_traffic("linear_tc_kernel_fc1", B), "traffic_unit": "B", "frac_of_split_ceiling": 3.0 * ach / peaks["bf16_tflops"],
            "flops_per_launch": flops, "mma_flops_per_launch": 3 * flops, "us_per_launch": sec * 1e6, "shape_MNK": [M, N, K],
            "peak_source": peaks["source"], "note": "3 bf16 MMAs per product (bf16x3): frac ceiling is 1/3"}


def _graph_time(fns, dev, rounds=5):
    """seconds per call of the zero-argument callables `fns` (one per rotating buffer set), timed as CUDA-graph replays."""
    import torch
    for f in fns:
        f()
    torch.cuda.synchronize(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for f in fns:
                f()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) * 1e-3 / (rounds * len(fns))


def lbs_roofline(lib, dev, peaks, B=256):
    """north_star's LBS-bandwidth roofline: `SMPL_Layer.forward` (smpl_layer.py:65-158) = smpl_pose_kernel + blend-shape GEMM +
    the sparse skinning kernel, timed as one call (`smpl_lbs_forward_sparse`) over rotating output sets larger than L2.
    achieved = SURVEY §8(d)'s algorithmic bytes (340 B in + 82,968 B out per sample) / time; `kernel_io` adds the v_posed
    intermediate the GEMM writes and the skinning kernel reads (82,688 B each way per sample)."""
    import ctypes as Ct
    import torch
    from pmce_b200 import synth
    from pmce_b200.smpl_layer import SMPL_Layer
    layer = SMPL_Layer.from_buffers(synth.make_smpl_buffers(11)).to(dev)
    p = layer._pack()
    pose, betas, trans = [t.to(dev) for t in synth.make_smpl_inputs(B, seed=13)]
    nsets = 8
    verts = [torch.empty(B, 6890, 3, device=dev) for _ in range(nsets)]
    joints = torch.empty(B, 24, 3, device=dev)
    ws = torch.empty(lib.smpl_workspace_bytes(B), dtype=torch.uint8, device=dev)
    P = lambda t: Ct.c_void_p(t.data_ptr()) if t is not None else Ct.c_void_p(0)

    def call(i):
        rc = lib.smpl_lbs_forward_sparse(P(p["blend"]), P(p["blend_hi"]), P(p["blend_lo"]), P(p["vt_pad"]), P(p["jt"]), P(p["js"]), P(p["weights"]),
                                         P(p["idx4"]), P(p["w4"]), P(p["parents"]), P(pose), P(betas), P(trans), B, 1.0, P(verts[i]), P(joints),
                                         P(ws), ws.numel(), Ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.pmce_last_error()
    sec = _graph_time([lambda i=i: call(i) for i in range(nsets)], dev)
    alg = B * (340 + 82968)
    io = alg + B * 2 * 82688
    return {"kernel": "smpl_lbs_forward_sparse (smpl_pose_kernel + blend-shape GEMM on tcgen05 + smpl_skin4_kernel: <= 4 joints per vertex)",
            "bound": "hbm", "achieved": alg / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg / sec / 1e9 / peaks["hbm_gbs"],
            "traffic": ncu_traffic("smpl_skin_kernel", B), "traffic_unit": "B", "bytes_per_launch": alg, "us_per_launch": sec * 1e6,
            "achieved_kernel_io": io / sec / 1e9, "frac_kernel_io": io / sec / 1e9 / peaks["hbm_gbs"], "samples_per_launch": B,
            "peak_source": peaks["source"]}


def spin_leg(lib, dev, peaks, cpu=False, frames=16):
    """(f)2, the step before the path: the SPIN ResNet-50 feature extractor (lib/models/spin.py:129-143) on one clip's 16 frame
    crops; algorithmic FLOPs 2 x 4.09 GMAC per frame. Also the 3x3 convolution of layer1 as the GEMM it runs as (explicit
    im2col rows [frames*3136, 576] x [64, 576]^T) on the tcgen05 GEMM, timed alone."""
    import ctypes as Ct
    import torch
    from pmce_b200 import synth
    from pmce_b200.spin import HMR
    m = HMR()
    sd = synth.make_spin_state_dict(17)
    m.load_state_dict(sd)
    m = m.to(dev)
    x = synth.make_frames(frames, 3).to(dev)
    for _ in range(3):
        m.feature_extractor(x)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        m.feature_extractor(x)
    e1.record()
    torch.cuda.synchronize(dev)
    sec = e0.elapsed_time(e1) * 1e-3 / n
    flops = frames * 2 * 4.09e9
    out = {"frames_per_s": frames / sec, "ms_per_clip_of_16_frames": sec * 1e3 * 16 / frames, "tflops_algorithmic": flops / sec / 1e12,
           "frac_of_bf16_peak": flops / sec / 1e12 / peaks["bf16_tflops"], "note": "split-bf16 (3 MMAs per product) GEMMs, explicit im2col"}
    # the 3x3 convolution of layer1 as a GEMM
    M, N, K = frames * 3136, 64, 576
    P = lambda t: Ct.c_void_p(t.data_ptr())
    a = [torch.empty(M, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    w = [torch.empty(N, K, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.pmce_split_bf16(P(torch.randn(M, K, device=dev)), M, K, P(a[0]), P(a[1]), st) == 0
    assert lib.pmce_split_bf16(P(torch.randn(N, K, device=dev) * 0.05), N, K, P(w[0]), P(w[1]), st) == 0
    b = torch.randn(N, device=dev)
    o = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(2)]
    Z = Ct.c_void_p(0)

    def conv():
        assert lib.pmce_linear_tc_presplit(P(a[0]), P(a[1]), P(w[0]), P(w[1]), P(b), M, N, K, 2, Z, P(o[0]), P(o[1]), Z,
                                           Ct.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    sec_c = _graph_time([conv] * 4, dev)
    fl = 2.0 * M * N * K
    out["conv3x3_layer1"] = {"shape_MNK": [M, N, K], "us_per_launch": sec_c * 1e6, "tflops_algorithmic": fl / sec_c / 1e12,
                             "frac": fl / sec_c / 1e12 / peaks["bf16_tflops"], "bound": "hbm (N = 64: 231 MB of im2col rows per launch)",
                             "achieved_gbs": (M * K * 4 + M * N * 4) / sec_c / 1e9, "frac_hbm": (M * K * 4 + M * N * 4) / sec_c / 1e9 / peaks["hbm_gbs"]}
    if cpu:
        from oracle import spin_oracle as so
        xc = synth.make_frames(4, 3)
        with torch.no_grad():
            so.feature_extractor(sd, xc)
            t0 = time.perf_counter()
            so.feature_extractor(sd, xc)
            dt = time.perf_counter() - t0
        out["cpu_frames_per_s"] = 4 / dt
        out["cpu_sample"] = f"4 frames through the oracle restatement (torch CPU fp32, {torch.get_num_threads()} threads)"
    return out


def cross_attn_roofline(lib, eng, dev, peaks, B, nsets=16, rounds=6):
    """north_star's graded kernel: the fused vertex<-joint cross-attention (ca_vertex_fused_kernel, csrc/ca_fused.cuh), HBM-bound.
    Timed alone with CUDA events over `nsets` rotating buffer sets whose total size exceeds the 126 MB L2, so every launch
    streams its query rows from HBM (`pmce_ca_vertex_fused`, the mode the forward runs); `embed_mode` = the opt-in variant that
    builds the stream from the [B,431,3] coordinates (`pmce_ca_vertex_fused_embed`). achieved = ALGORITHMIC bytes / time with SURVEY.md
    §8(d)'s per-clip figure for the fused block (231,424 B: q/k/v streams in + q stream out + gamma/beta);
    `achieved_kernel_io` counts what the launch really moves per clip through HBM."""
    import ctypes as Ct
    import torch
    Vd, D = 431, 64
    P = lambda t: Ct.c_void_p(t.data_ptr())
    gen = torch.Generator(device=dev).manual_seed(5)
    xq = [torch.randn(B, Vd, D, device=dev, generator=gen) for _ in range(nsets)]
    coords = [torch.randn(B, Vd, 3, device=dev, generator=gen) * 0.3 for _ in range(nsets)]
    Kt = torch.randn(B, J, D, device=dev, generator=gen)
    Vt = torch.randn(B, J, D, device=dev, generator=gen)
    gb = torch.randn(B, lib.pmce_adaln_slots(), 2, D, device=dev, generator=gen)
    fold_ws = torch.empty(lib.pmce_ca_fold_bytes(B), dtype=torch.uint8, device=dev)
    table_ws = torch.empty(Vd * D, device=dev)

    def call_inplace(i, fold=0):
        rc = lib.pmce_ca_vertex_fused(eng._dp, P(eng.weights), 1, P(xq[i]), P(Kt), P(Vt), P(gb), B, Ct.c_void_p(0), Ct.c_void_p(0), P(fold_ws), fold,
                                      Ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.pmce_last_error()

    def call_embed(i, fold=0):
        rc = lib.pmce_ca_vertex_fused_embed(eng._dp, P(eng.weights), 1, P(coords[i]), P(xq[i]), P(Kt), P(Vt), P(gb), B, P(fold_ws), fold, P(table_ws),
                                            Ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.pmce_last_error()

    def timed(call):
        call(0, fold=1)          # per-clip folded operands (+ the table): separate small kernels in the forward, made once here
        for i in range(nsets):
            call(i)
        torch.cuda.synchronize(dev)
        # one CUDA graph holding a full rotation over the buffer sets: the timed region is device time only, as in the forward
        # (which replays a captured graph); eager launches from Python would be bounded by the host at this kernel size
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for i in range(nsets):
                    call(i)      # in-place mode: xq[i] keeps being updated (values stay finite: every pass re-normalises the row)
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(rounds):
            graph.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) * 1e-3 / (rounds * nsets)

    sec = timed(call_inplace)
    sec_e = timed(call_embed)
    survey_bytes = ((J + Vd + Vd + J) * D * 4 + 2048) * B
    folded = 4 * (48 * 64 * 2) + 48 * 4
    io_bytes = (2 * Vd * D * 4 + folded) * B
    io_e = (Vd * 3 * 4 + Vd * D * 4 + folded) * B                 # coordinates in, stream out, folded operands in (the table stays in L2)
    flops = (4 * 2 * Vd * J * 32 + 2 * 2 * Vd * D * D) * B       # attention core + Wq + Wp
    ach = survey_bytes / sec / 1e9
    return {"kernel": "ca_vertex_fused_kernel (AdaLN_q + Wq + scores + softmax + P.V + Wp + bias + residual in one pass over the query stream; "
                      "4 independent 4-warp groups = 4 items in flight per CTA, in-place buffers, residual held in the TMEM accumulator)",
            "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
            "traffic": ncu_traffic("ca_vertex_fused_kernel", B), "traffic_unit": "B", "bytes_per_launch": survey_bytes, "us_per_launch": sec * 1e6,
            "achieved_kernel_io": io_bytes / sec / 1e9, "kernel_io_bytes_per_launch": io_bytes,
            "frac_kernel_io": io_bytes / sec / 1e9 / peaks["hbm_gbs"], "flops_per_launch": flops, "clips_per_launch": B,
            "embed_mode": {"us_per_launch": sec_e * 1e6, "frac": survey_bytes / sec_e / 1e9 / peaks["hbm_gbs"], "kernel_io_bytes_per_launch": io_e,
                           "note": "PMCE_CA_EMBED=1 (off by default): the kernel also embeds the coordinates and only WRITES the stream - half "
                                   "the HBM bytes, slower: the kernel is latency / instruction bound, not byte bound"},
            "l2": f"{nsets} rotating buffer sets ({nsets * io_bytes / 1e6:.0f} MB) > 126 MB L2", "peak_source": peaks["source"]}


if __name__ == "__main__":
    main()
