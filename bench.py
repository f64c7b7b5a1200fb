#!/usr/bin/env python
"""Benchmark of the rasterizer hot path: fwd+bwd views/s on the BASELINE.json C2 scene
(two-hand, 60k Gaussians, 512x334, precomputed colours), camera-sharded over N GPUs (config 3).

    python bench.py [--gpus N --steps K --warmup W] [--views B] [--impl reference] [--config c2|c1|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch: B distinct views (default 8 = the per-GPU share
of BASELINE config 3, a 64-view fitting step on 8 GPUs) rendered forward + backward by ONE batched
call chain on every rank, gradients summed over views in the packed buffer, then (N>1) one all-reduce
of that buffer -- libghr's own NVLink peer-memory kernel (--allreduce peer, default) or NCCL.  Weak
scaling: B views per rank per step.
  value   : views/s over all ranks, inputs resident in HBM, CUDA events per step, max over ranks, L2
            flushed between steps.  `step_ms` carries the per-step distribution of every rank.
  e2e     : same metric through the public pipelined API (guassianhand_b200.dist.PipelinedFitLoop.run over
            GraphedFitStep instances) with host buffers: H2D of Gaussian attributes + cameras and D2H of
            gradients + loss inside the timed region (wall clock, max over ranks).
  gpu_baseline (N = 1): an upstream-STRUCTURED GPU stand-in of the binning + blend stages (baseline_standin/,
            kind "restatement": not the reference's package) on the same views.
  roofline: SURVEY.md §8(d): the dominant kernel against ITS roof (FP32 for the blend kernels, HBM for
            the streaming ones), a per-stage table, and the max-sum model of the whole step
            t_roof = bytes / hbm_peak + 94 I / fp32_peak against the measured step.
  cpu_baseline / --impl reference: the CPU oracle (oracle/gs_oracle.c, OpenMP, all host threads,
            pinned) on a bounded sample of the same workload.  The reference's own implementation of this
            path (pip diff-gaussian-rasterization) is absent from /root/reference and from the box,
            so the oracle port is the reference arm (kind "port").
  --config c1 | c4 | c5: the other BASELINE.json configurations (CPU plumbing case, 1M-Gaussian SH3
            stress case, 1080p forward-only drive render) as single lines of the same shape.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

# The CPU arm's OpenMP threads stay where they start (set before any OpenMP runtime is loaded) -- in a
# single-process run only.  Under torchrun every rank would bind its main thread to the FIRST core: two ranks
# whose hosts spin on an event then share one core and take turns in ~3 ms scheduler slices (seen on 2 x B200:
# e2e 7.7k views/s with alternating 3.2 ms stalls; the late host makes the other rank wait in the all-reduce).
if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
elif hasattr(os, "sched_setaffinity"):
    try:
        os.sched_setaffinity(0, range(os.cpu_count()))      # undo a binding inherited from the launcher
    except OSError:
        pass

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(P=60000, H=512, W=334, pool_views=64, seed=0)
METRIC = "fwd+bwd views/s @512x334 two-hand Gaussians"
UNIT = "views/s"
FP32_NOMINAL_TFLOPS = 74.4      # 148 SMs x 128 lanes x 2 x 1.965 GHz
CPU_VIEWS_PER_STEP = 4          # the CPU arm's step: a bounded sample of the GPU arm's step


def _workload(B):
    return (f"BASELINE config 2 scene (two-hand, {C2['P']} Gaussians, {C2['H']}x{C2['W']}, colors_precomp, "
            f"scale/rotation covariance); step = {B} distinct views fwd+bwd per rank in one batched call "
            f"(config 3 shard: 64-view pool), grads summed over views")


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples SM clock / throttle reasons with NVML every `period` s while running."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.005):
        self.samples, self.reasons, self.period, self.on = [], 0, period, False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def _run(self):
        while self.on:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv:
            self.on = True
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        if self.nv:
            self.on = False
            self.t.join()

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        names = [n for b, n in self.REASONS.items() if self.reasons & b and n != "gpu_idle"]
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": names,
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(local_index):
    """Multi-rank runs: keep this rank's host threads -- and with them the pinned staging buffers it is about to
    allocate (first touch) -- on the NUMA node its GPU hangs off.  Seen at 8 x B200 without it: the uploads /
    downloads of four of the eight ranks ran at 17 GB/s instead of 31 GB/s (remote host memory) and their replays
    waited for them.  The rank keeps ALL cores of the node (never a single core: see the note at the top).  Returns
    a short description for the bench line, or None when the topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"
        node = int(open(path + "/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= set(range(os.cpu_count()))
        if len(cpus) < 2:
            return None
        os.sched_setaffinity(0, cpus)
        return f"host threads and pinned buffers of rank-local GPU {local_index} on NUMA node {node} ({len(cpus)} cpus)"
    except Exception:
        return None


# ----------------------------------------------------------------------------- CPU arm

def cpu_views_per_s(n_views, threads=None, budget_s=None, P=None, H=None, W=None, hands=2, sh_degree=None):
    """fwd+bwd of `n_views` views through the CPU oracle (C2 scene unless P/H/W given); returns
    (views/s, views done, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from guassianhand_b200 import scenes
    from oracle import oracle_lib as ol
    import util
    L = ol.lib()
    if threads:
        L.gso_set_num_threads(int(threads))
    nthr = L.gso_num_threads()
    P, H, W = P or C2["P"], H or C2["H"], W or C2["W"]
    sc = scenes.two_hand_scene(P, seed=C2["seed"], hands=hands, sh_degree=sh_degree)
    cams = scenes.fibonacci_cameras(C2["pool_views"], H, W, seed=C2["seed"])
    rng = np.random.default_rng(1)
    dL = (rng.normal(size=(3, H, W)) / (H * W)).astype(np.float32)
    bg = np.zeros(3, np.float32)
    done, t0 = 0, time.perf_counter()
    for v in range(n_views):
        osc = util.oracle_scene(sc, cams[v % len(cams)], bg)
        ol.forward_backward(osc, dL)
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, nthr


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = os.cpu_count()
    # warm-up (also builds/loads the oracle), then K steps of CPU_VIEWS_PER_STEP views each
    for _ in range(max(args.warmup, 1)):
        cpu_views_per_s(1, threads=ncpu)
    t0 = time.perf_counter()
    vps, done, nthr = cpu_views_per_s(args.steps * CPU_VIEWS_PER_STEP, threads=ncpu, budget_s=240.0)
    dt = time.perf_counter() - t0
    steps_done = max(done // CPU_VIEWS_PER_STEP, 1)
    sample = (f"{done} single views of the C2 scene ({CPU_VIEWS_PER_STEP} per step: a bounded sample of the GPU arm's "
              f"{args.views}-view step), oracle/gs_oracle.c with OpenMP on {nthr} pinned threads (upstream "
              f"diff-gaussian-rasterization is not installable here)")
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / steps_done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": _workload(args.views), "views_per_rank_per_step": args.views, "gaussians": C2["P"],
                   "image": [C2["H"], C2["W"]], "sample": sample},
        "cpu_baseline": {"value": vps, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample},
        "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- roofline

def algorithmic_bytes(stage, P, V, N, T, R, M=0):
    """SURVEY.md §8(d) per-stage algorithmic bytes, for V views with R instances in total."""
    sh_f = (12 * M + 12) * P * V if M else 0
    sh_b = (24 * M + 3) * P * V if M else 0
    return {
        "preprocess": 80 * P * V + sh_f,
        "tile_scan": (8 * P + 8 * T) * V,              # K2 (scan; here over tiles) + range table of K5
        "duplicate": 20 * P * V + 12 * R,              # K3
        "sort_gather": 24 * R + 8 * R,                 # K4 (one compulsory read+write of 12-B pairs) + K5
        "blend_forward": 40 * R + 20 * N * V,
        "blend_backward": 40 * R + (20 * N + 36 * P) * V,
        "preprocess_backward": 120 * P * V + sh_b,
    }[stage]


BLEND_FLOPS_PER_PAIR = {"blend_forward": 24.0, "blend_backward": 70.0}    # SURVEY.md §8(d); pair = upstream loop trip


def roofline_report(stage_ms, step_ms, P, V, N, T, R, I, hbm_peak, fp32_peak, peak_src, traffic_json, M=0):
    """Per-stage table, the dominant kernel against its own roof, and the max-sum model of the step."""
    table = {}
    for k, ms in stage_ms.items():
        b = algorithmic_bytes(k, P, V, N, T, R, M)
        row = {"launch_ms": ms, "algorithmic_bytes": b, "hbm_GBps": b / (ms * 1e-3) / 1e9,
               "hbm_frac": b / (ms * 1e-3) / 1e9 / hbm_peak}
        if k in BLEND_FLOPS_PER_PAIR:
            f = BLEND_FLOPS_PER_PAIR[k] * I
            row.update(bound="fp32", algorithmic_flops=f, fp32_TFLOPs=f / (ms * 1e-3) / 1e12,
                       fp32_frac=f / (ms * 1e-3) / 1e12 / fp32_peak)
        else:
            row["bound"] = "hbm"
        table[k] = row
    dom = max(stage_ms, key=stage_ms.get)
    d = table[dom]
    if d["bound"] == "fp32":
        main = {"bound": "fp32", "kernel": dom, "achieved": d["fp32_TFLOPs"], "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": d["fp32_frac"], "algorithmic_flops_per_launch": d["algorithmic_flops"],
                "flops_per_pair": BLEND_FLOPS_PER_PAIR[dom], "pairs_per_launch": I,
                "peak_source": f"nominal FP32 FMA ceiling {FP32_NOMINAL_TFLOPS} TFLOP/s (148 SMs x 128 lanes x 2 x 1.965 GHz); "
                               f"the library's register-operand FFMA / FFMA2 probes measure it live (roofline_fp32)"}
    else:
        main = {"bound": "hbm", "kernel": dom, "achieved": d["hbm_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": d["hbm_frac"], "peak_source": peak_src}
    main.update(traffic=(traffic_json or {}).get(dom), algorithmic_bytes_per_launch=d["algorithmic_bytes"],
                launch_ms=stage_ms[dom])
    step_bytes = sum(r["algorithmic_bytes"] for r in table.values())
    t_roof_ms = (step_bytes / (hbm_peak * 1e9) + 94.0 * I / (fp32_peak * 1e12)) * 1e3
    main["step"] = {"model": "max-sum (SURVEY.md §8d): t_roof = bytes / hbm_peak + 94 I / fp32_peak; stages are sequential",
                    "algorithmic_bytes_per_step": step_bytes, "algorithmic_flops_per_step": 94.0 * I,
                    "t_roof_ms": t_roof_ms, "t_measured_ms": step_ms, "achieved": t_roof_ms / step_ms,
                    "hbm_peak_GBps": hbm_peak, "fp32_peak_TFLOPs": fp32_peak}
    main["stages"] = table
    return main


def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    return hbm, src, traffic


def fp32_probes(NV, torch, dev, stream):
    """Register-operand dependent-FMA probes of the library: scalar FFMA and packed FFMA2 (TFLOP/s)."""
    import ctypes as C
    sink = torch.zeros(1, device=dev)
    probe_in = torch.tensor([0.999, 1e-3], device=dev)
    flops = C.c_double()

    def probe(packed):
        best = 0.0
        for _ in range(3):
            ps, pe = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ps.record()
            NV.check(NV.lib().ghr_fp32_probe(1 << 11, packed, probe_in.data_ptr(), sink.data_ptr(), C.byref(flops),
                                             stream.cuda_stream), "probe")
            pe.record()
            torch.cuda.synchronize()
            best = max(best, flops.value / (ps.elapsed_time(pe) * 1e-3) / 1e12)
        return best
    return probe(0), probe(1)


def cull_stats(views):
    """Evaluated / contributing (pixel, instance) pairs of the blend kernels, from the counting build variant
    (a separate process: the variant library must not be the one this process measures)."""
    try:
        from guassianhand_b200 import build
        if not os.path.exists(build.variant_path("count")):
            return None
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cull_stats.py"), "--views", str(views)],
                           capture_output=True, text=True, timeout=120)
        lines = r.stdout.strip().splitlines()
        if r.returncode != 0 or not lines:
            return {"error": f"tools/cull_stats.py rc {r.returncode}: {r.stderr.strip()[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def deal_views(costs, block, g, n_blocks, world):
    """The g-th deal of one block of the view pool to the ranks, balanced on `costs` (every rank computes the
    same deal).  When one step uses the whole pool (n_blocks == 1) the deals differ by a rotation of the
    tie-breaking order.  Returns one list of pool indices per rank."""
    from guassianhand_b200.dist import balanced_shards
    rot = block[g:] + block[:g] if n_blocks == 1 else block
    return [[rot[i] for i in sh] for sh in balanced_shards([costs[i] for i in rot], world)]


# ----------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from guassianhand_b200 import _native as NV, api, scenes
    from guassianhand_b200.dist import GraphedFitStep, PackedGrads, fit_step_grads
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note = None
    if world > 1:
        numa_note = bind_to_gpu_numa_node(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    NV.lib()
    peer = world > 1 and args.allreduce == "peer"
    B, K, Wm = args.views, args.steps, max(args.warmup, 3)
    P, H, W = C2["P"], C2["H"], C2["W"]
    N, T = H * W, ((W + 15) // 16) * ((H + 15) // 16)

    sc = scenes.two_hand_scene(P, seed=C2["seed"])
    cams = scenes.fibonacci_cameras(C2["pool_views"], H, W, seed=C2["seed"])
    bg = np.zeros(3, np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
                 colors_precomp=t(sc.colors))
    fits = B * world <= len(cams)
    n_blocks = max(1, len(cams) // (B * world)) if fits else 1

    # ---- which rank renders which view of a step.  The step ends at the slowest rank (the all-reduce), and a
    # view's time follows its blend pairs (sum of n_contrib), not its instance count: the step's world*B
    # views are dealt to the ranks so that the per-rank pair sums are even (deal_views; every rank computes
    # the same deal from the same costs, measured once here).  When one step uses the whole pool (8 GPUs x
    # 8 views), several different balanced deals rotate so that no rank keeps the same 8 views.
    deals, costs = None, None
    if world > 1 and fits:
        costs = []
        for c in cams:
            v1 = util.gpu_views([c], bg, dev)
            r1 = api.forward_raw(v1.cams(), gauss["means3D"], gauss["opacities"], gauss["scales"], gauss["rotations"],
                                 None, None, gauss["colors_precomp"], 0, 1.0)
            lay1 = NV.layout(P, 1, H, W, 0, 0, r1.R_cap)
            nc = r1.state[lay1.off_ncontrib: lay1.off_ncontrib + N * 4].view(torch.int32)
            costs.append(float(nc.sum(dtype=torch.int64).item()))
        n_deals = n_blocks if n_blocks > 1 else 4
        deals = []
        for g in range(n_deals):
            b0 = (g % n_blocks) * B * world
            deals.append(deal_views(costs, list(range(b0, b0 + B * world)), g, n_blocks, world))
    n_groups = len(deals) if deals is not None else n_blocks

    def group_cams(g, r=rank):
        if deals is not None:
            return [cams[i] for i in deals[g % n_groups][r]]
        base = (g % n_groups) * B * world + r * B
        return [cams[(base + j) % len(cams)] for j in range(B)]

    view_groups = [util.gpu_views(group_cams(g), bg, dev) for g in range(n_groups)]

    def dL_of(r):
        return t((np.random.default_rng(1 + r).normal(size=(B, 3, H, W)) / N).astype(np.float32))
    dL = dL_of(rank)
    grads = PackedGrads(P, 0, device=dev, peer=peer)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    # ---- calibration: instance counts and blend pairs per view group (poll mode, exact) ----
    Rs, pairs = [], []
    for g in range(n_groups):
        res = fit_step_grads(gauss, view_groups[g], dL, grads, group=False)
        Rs.append(res.R)
        lay = NV.layout(P, B, H, W, 0, 0, res.R_cap)
        nc = res.state[lay.off_ncontrib: lay.off_ncontrib + B * N * 4].view(torch.int32)
        pairs.append(int(nc.sum(dtype=torch.int64).item()))
    R_cap = int(max(Rs) * 1.25) + (1 << 14)
    # view groups of a step on concurrent streams (dist.fit_step_grads overlap): per-group capacities
    G = max(1, min(int(args.overlap), B))
    caps = R_cap
    if G > 1:
        per_group = [[] for _ in range(G)]
        for g in range(n_groups):
            r = fit_step_grads(gauss, view_groups[g], dL, grads, overlap=G, group=False)
            for j, x in enumerate(r.results):
                per_group[j].append(x.R)
        caps = [int(max(v) * 1.25) + (1 << 14) for v in per_group]
    # libghr kernels per group: preprocess, tile scan, duplicate, sort_big + sort_chunks, merge_gather, blend forward |
    # blend backward, preprocess backward (9; the ncu launch list of this command shows the same count); + the
    # partial-sum adds; + the all-reduce kernel
    launches_per_step = (1 + 1 + 1 + 2 + 1 + 1 + 2) * G + (G - 1) + (1 if world > 1 else 0)

    def step(i, fe=None, be=None):
        return fit_step_grads(gauss, view_groups[i % n_groups], dL, grads, R_cap=R_cap, check="none",
                              fwd_events=fe, bwd_events=be)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(Wm):
        step(i)
    sync_all()

    # ---- timed region: K steps.  Default: the step is replayed from a CUDA graph (one launch per
    # step); the cameras of the step's view group are copied into the graph's static buffers inside
    # the timed region.  --no-graph launches the same kernels eagerly. ----
    ev_s = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev_e = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stream = torch.cuda.current_stream()
    status_pin = torch.zeros(K, G, 4, dtype=torch.int64).pin_memory()
    if not args.no_graph:
        # the graph's static camera inputs are blocks of ONE flat buffer, so a step's cameras arrive with a
        # single device-to-device copy
        cnames = ["viewmatrix", "projmatrix", "campos", "tanfov"]
        vg_flat = [torch.cat([getattr(vg, k).reshape(-1) for k in cnames]) for vg in view_groups]
        static_flat = vg_flat[0].clone()
        parts, o = {}, 0
        for k in cnames:
            ref = getattr(view_groups[0], k)
            parts[k] = static_flat[o:o + ref.numel()].view(ref.shape)
            o += ref.numel()
        static_views = view_groups[0]._replace(**parts)
        gstep = GraphedFitStep(gauss, static_views, dL, grads, R_cap=caps, overlap=G)

        def run_step(i):
            static_flat.copy_(vg_flat[i % n_groups])
            return gstep.replay()
    else:
        def run_step(i):
            return step(i)
    # multi-GPU: the first replays of a graph that holds a collective pay one-time costs on some rank (and
    # every rank waits for it at the all-reduce): at least 20 untimed replays
    for i in range(max(Wm, 20) if world > 1 else Wm):
        run_step(i)
    # clocks are sampled on rank 0 only (it prints the line): NVML calls take the driver's lock, and with one
    # sampler per rank a delayed launch on ANY rank stalls every rank at the step's all-reduce.  The sampler is
    # created and started BEFORE the barrier (nvmlInit takes milliseconds: started inside the timed loop it made
    # rank 0 late for its first step, and the other ranks' first interval absorbed the wait in the all-reduce).
    sampler = ClockSampler(local, period=0.005 if world == 1 else 0.02) if rank == 0 else contextlib.nullcontext()
    n_states = 1
    with sampler as clk:
        sync_all()
        # two untimed steps after the host-side barrier: ranks leave a device synchronisation up to milliseconds
        # apart (host wake-up latency), and the all-reduce of a step is what re-aligns the GPUs
        for i in range(2):
            flush.zero_()
            run_step(i)
        for i in range(K):
            flush.zero_()                                   # L2 flush, outside the timed events
            ev_s[i].record()
            res = run_step(i)
            ev_e[i].record()
            states = [x.state for x in res.results] if hasattr(res, "results") else [res.state]
            n_states = len(states)
            for j, st in enumerate(states):
                NV.check(NV.lib().ghr_read_status_async(st.data_ptr(), status_pin[i, j].data_ptr(),
                                                        stream.cuda_stream), "status")
        sync_all()
    got_last = grads.flat.double().clone()                  # the last timed step's (all-reduced) gradients
    per_step = np.array([s.elapsed_time(e) for s, e in zip(ev_s, ev_e)], np.float64)
    if int((status_pin[:K, :n_states, 1] & 0xFFFFFFFF).sum()) != 0:
        raise RuntimeError("bench: instance capacity overflow inside the timed region; result invalid")
    if peer and grads.comm.status()[1]:
        raise RuntimeError("bench: the peer all-reduce timed out waiting for a rank")
    all_steps = torch.tensor(per_step, device=dev, dtype=torch.float64)
    if world > 1:
        gathered = [torch.zeros_like(all_steps) for _ in range(world)]
        dist.all_gather(gathered, all_steps)
        all_steps = torch.stack(gathered)
    else:
        all_steps = all_steps[None]
    all_steps = all_steps.cpu().numpy()                       # [world, K] ms
    total_ms_max = float(all_steps.sum(axis=1).max())
    value = world * B * K / (total_ms_max / 1000.0)
    med = np.median(all_steps, axis=1)
    step_stats = {
        "per_rank": [{"median": float(np.median(r)), "p10": float(np.percentile(r, 10)), "p90": float(np.percentile(r, 90)),
                      "max": float(r.max()), "mean": float(r.mean())} for r in all_steps],
        "value_from_median": world * B / (float(med.max()) / 1000.0),
        "steps": [[round(float(x), 4) for x in r] for r in all_steps] if world > 1 else None,
        "note": "ms per step, CUDA events around every step on every rank (the all-reduce ends a step at the slowest rank)"}

    # ---- N>1: the all-reduced gradients of the LAST timed step against rank 0 rendering all its views ----
    ar_check = None
    if world > 1 and deals is not None:
        g_last = (K - 1) % n_groups
        if rank == 0:
            ref = PackedGrads(P, 0, device=dev)
            tot = torch.zeros_like(ref.flat, dtype=torch.float64)
            for r_ in range(world):
                vr = util.gpu_views(group_cams(g_last, r_), bg, dev)
                fit_step_grads(gauss, vr, dL_of(r_), ref, group=False)
                tot += ref.flat.double()
            err = float((got_last - tot).abs().max() / tot.abs().max())
            ar_check = {"max_rel_err": err, "ok": bool(err <= 1e-5),
                        "what": f"packed gradients after the {world}-rank step vs rank 0 rendering all "
                                f"{world * B} views itself (max |a-b| / max |b|)"}
        torch.cuda.synchronize()
        dist.barrier()

    # ---- per-stage durations: the same K steps again, launched eagerly with the library's
    # per-stage CUDA events (events cannot be timed from inside a replayed graph) ----
    fwd_ev = [NV.StageEvents(NV.GHR_NSTAGES_FWD) for _ in range(K)]
    bwd_ev = [NV.StageEvents(NV.GHR_NSTAGES_BWD) for _ in range(K)]
    for i in range(K):
        flush.zero_()
        step(i, fwd_ev[i], bwd_ev[i])
    sync_all()
    stage_ms = {}
    for si, name in enumerate(NV.FWD_STAGES):
        stage_ms[name] = float(np.mean([fwd_ev[i].elapsed_ms(si) for i in range(K)]))
    for si, name in enumerate(NV.BWD_STAGES):
        stage_ms[name] = float(np.mean([bwd_ev[i].elapsed_ms(si) for i in range(K)]))
    for e in fwd_ev + bwd_ev:
        e.close()
    R_mean = float(np.mean(status_pin[:K, :n_states, 0].numpy().sum(axis=1)))
    I_mean = float(np.mean(pairs))

    line_extra = {}
    if world == 1 and not args.quick:
        line_extra.update(single_gpu_extras(args, torch, api, scenes, util, gauss, cams, bg, dL, grads, dev, t,
                                            fit_step_grads, GraphedFitStep))
    fp32_scalar, fp32_packed = fp32_probes(NV, torch, dev, stream)

    # ---- e2e through the public pipelined API (dist.PipelinedFitLoop) with host buffers: the headline e2e ----
    # Every step uploads ITS Gaussian attributes + cameras from pinned host memory into a slot's static input
    # buffers (two H2D copies), replays the captured forward+backward(+all-reduce) and downloads the packed
    # gradients + the loss (two D2H copies).  Three graph instances with their own buffers rotate: inputs go up
    # two steps ahead, the host reads a step's results two launches after it (PipelinedFitLoop.run).
    from guassianhand_b200.dist import PipelinedFitLoop
    NG = max(1, int(os.environ.get('GHR_BENCH_E2E_SLOTS', '3')))      # (diagnostics: other pipeline depths)
    bg_dev = t(bg)
    v0 = view_groups[0]
    loop_views = api.ViewBatch(image_height=H, image_width=W, viewmatrix=v0.viewmatrix, projmatrix=v0.projmatrix,
                               campos=v0.campos, tanfov=v0.tanfov, bg=bg_dev)
    loop = PipelinedFitLoop(gauss, loop_views, dL, caps, overlap=G, slots=NG, peer=peer, trace=True)
    host_flat = loop.pack_attributes(gauss)
    host_cams = [loop.pack_cameras(vg) for vg in view_groups]
    g_h2d, g_d2h = loop.h2d_bytes_per_step, loop.d2h_bytes_per_step

    def g_run(n):
        loop.reset()
        base = torch.cuda.Event(enable_timing=True)
        base.record(torch.cuda.current_stream())
        marks, last = [], None
        for res in loop.run((host_flat, host_cams[i % n_groups]) for i in range(n)):
            last = res.loss
            marks.append(time.perf_counter())
        torch.cuda.synchronize()
        return last, marks, base

    g_run(8 if world > 1 else 4)
    sync_all()
    t0 = time.perf_counter()
    g_loss, g_marks, g_base = g_run(K)
    g_s = time.perf_counter() - t0
    g_steps = np.diff(np.array([t0] + g_marks)) * 1e3          # wall ms between consecutive results on this rank
    g_gpu, g_copy, g_host = loop.events["replay"], loop.events, loop.host_ms
    g_host = {"upload": g_host["upload"], "render": g_host["launch"], "collect": g_host["result"]}
    g_gpu_ms = [a_.elapsed_time(b_) for a_, b_ in g_gpu]       # device ms of every replay on this rank
    def rel(ev):
        return round(g_base.elapsed_time(ev), 3)
    g_mine = {"timeline_ms": [{"h2d": [rel(g_copy["h2d"][i][0]), rel(g_copy["h2d"][i][1])] if i < len(g_copy["h2d"]) else None,
                               "replay": [rel(g_gpu[i][0]), rel(g_gpu[i][1])],
                               "d2h": [rel(g_copy["d2h"][i][0]), rel(g_copy["d2h"][i][1])]} for i in range(len(g_gpu))],
              "replay": [round(float(x), 3) for x in g_gpu_ms],
              "h2d": [round(a_.elapsed_time(b_), 3) for a_, b_ in g_copy["h2d"]],
              "d2h": [round(a_.elapsed_time(b_), 3) for a_, b_ in g_copy["d2h"]]}
    g_all = [None] * world
    if world > 1:
        dist.all_gather_object(g_all, g_mine)
    else:
        g_all = [g_mine]
    tg = torch.tensor([g_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    g_value = world * B * K / float(tg.item())

    if rank == 0:
        hbm_peak, peak_src, traffic = load_peaks()
        step_ms = total_ms_max / K
        roof = roofline_report(stage_ms, step_ms, P, B, N, T, R_mean, I_mean, hbm_peak, FP32_NOMINAL_TFLOPS, peak_src,
                               traffic)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": _workload(B), "views_per_rank_per_step": B, "gaussians": P, "image": [H, W],
                       "instances_per_step": R_mean, "blend_pairs_per_step": I_mean, "R_cap": R_cap,
                       "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": f"camera-sharded dp{world}" + (" (views dealt to ranks by blend pairs)" if deals else ""),
                       "launch": "eager, one stream" if args.no_graph else
                       f"CUDA graph replay (1 launch/step) + 1 camera copy; the step's views run as {G} "
                       f"independent forward->backward chains on {G} streams inside the graph (stage_ms: the same "
                       f"kernels launched eagerly on one stream)",
                       "collective": "none (N=1)" if world == 1 else
                       ("libghr NVLink peer-memory all-reduce kernel (two-shot, in place) of the packed grads (56 B x P), "
                        "inside the graph" if peer else "NCCL all-reduce of packed grads (56 B x P)")},
            "roofline": roof,
            "roofline_fp32": {"peak_tflops_nominal": FP32_NOMINAL_TFLOPS, "probe_ffma_tflops": fp32_scalar,
                              "probe_ffma2_tflops": fp32_packed,
                              "note": "dependent-FMA chains with register operands, 128 FMA instructions per loop trip"},
            "stage_ms": stage_ms,
            "step_ms": step_stats,
            "e2e": {"value": g_value, "unit": UNIT, "h2d_bytes_per_step": g_h2d, "d2h_bytes_per_step": g_d2h,
                    "loss": g_loss,
                    "step_wall_ms": {"median": float(np.median(g_steps)), "p90": float(np.percentile(g_steps, 90)),
                                     "max": float(g_steps.max()), "note": "rank 0, wall clock between consecutive results",
                                     "host_call_ms_max": {k_: float(max(v_)) if v_ else None for k_, v_ in g_host.items()},
                                     "host_call_ms_median": {k_: float(np.median(v_)) if v_ else None for k_, v_ in g_host.items()},
                                     "steps": [round(float(x), 3) for x in g_steps],
                                     "replay_device_ms_per_rank": g_all},
                    "api": "guassianhand_b200.dist.PipelinedFitLoop.run() over GraphedFitStep instances (captured "
                           "ghr_forward + ghr_backward [+ all-reduce] of the step's views); per step: H2D of the "
                           "Gaussian attributes + cameras from pinned host memory into a slot's static inputs, D2H of "
                           "the packed gradients + the loss; three graph instances rotate so uploads run two steps "
                           "ahead and results are read two launches later, wall clock",
                    "host_placement": numa_note},
            "gpu_launches": launches_per_step * K,
            "clocks": clk.summary(),
        }
        if ar_check is not None:
            line["allreduce_check"] = ar_check
        line.update(line_extra)
        if world == 1 and not args.quick:
            line["cull_efficiency"] = cull_stats(B)
        if world == 1 and not args.no_cpu:
            vps, done, nthr = cpu_views_per_s(256, threads=os.cpu_count(), budget_s=12.0)
            line["cpu_baseline"] = {"value": vps, "unit": UNIT, "cores": nthr, "kind": "port",
                                    "sample": f"{done} single views of the same C2 scene, fwd+bwd, "
                                              f"oracle/gs_oracle.c OpenMP on {nthr} pinned threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroy_process_group() with collectives captured in live
        # CUDA graphs can block forever (seen at N=2), and the line above is already out.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def gpu_baseline_leg(torch, api, util, gauss, cams, bg, dL, dev, s1, e1, n_views=8, reps=10):
    from baseline_standin.standin import Standin
    from guassianhand_b200 import _native as NV
    P = gauss["means3D"].shape[0]
    H, W = cams[0].H, cams[0].W
    bgt = torch.from_numpy(bg).float().to(dev)
    sis, keep = [], []
    for v in range(n_views):
        v1 = util.gpu_views([cams[v]], bg, dev)
        r = api.forward_raw(v1.cams(), gauss["means3D"], gauss["opacities"], gauss["scales"], gauss["rotations"],
                            None, None, gauss["colors_precomp"], 0, 1.0)
        lay = NV.layout(P, 1, H, W, 0, 0, r.R_cap)
        geom = r.state[lay.off_geom: lay.off_geom + P * 64].view(torch.float32)
        keep.append(r)
        sis.append(Standin(geom, H, W, bgt))
    dLs = [dL[v].contiguous() for v in range(n_views)]

    def run(fwd, bwd):
        for v in range(n_views):
            if fwd:
                sis[v].forward()
            if bwd:
                sis[v].backward(dLs[v])
    for _ in range(3):
        run(True, True)
    torch.cuda.synchronize()
    res = {}
    for name, fwd, bwd in (("forward_K2_K6", True, False), ("backward_K7", False, True), ("total", True, True)):
        s1.record()
        for _ in range(reps):
            run(fwd, bwd)
        e1.record()
        torch.cuda.synchronize()
        res[name] = s1.elapsed_time(e1) / reps
    R = sum(si.R for si in sis)
    for si in sis:
        si.close()
    return {"kind": "restatement", "ms_per_%d_views" % n_views: res, "instances": R,
            "views_per_s_K2_K7": n_views / (res["total"] * 1e-3),
            "note": "upstream-STRUCTURED stand-in of binning + blend (baseline_standin/standin.cu, from SURVEY.md "
                    "Appendix A; not the reference package): CUB InclusiveSum + blocking D2H of num_rendered, "
                    "duplicateWithKeys, CUB DeviceRadixSort on 64-bit keys, identifyTileRanges, one 16x16 CTA per "
                    "tile forward and backward with 9 per-thread atomicAdd per pair; one view per call, L2 warm, "
                    "CUDA events around the loop (host sync gaps included); preprocess (K1) and its backward "
                    "(K8/K9) are NOT in it -- compare with stage_ms minus preprocess and preprocess_backward"}


def single_gpu_extras(args, torch, api, scenes, util, gauss, cams, bg, dL, grads, dev, t, fit_step_grads,
                      GraphedFitStep):
    """N=1 only: single-view latency and the reference-as-shipped call pattern through the drop-in API."""
    out = {}
    s1, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ---- single-view latency (the shape the reference itself runs: 1 view per call) ----
    v1 = util.gpu_views([cams[0]], bg, dev)
    dL1 = dL[:1].contiguous()
    r1 = fit_step_grads(gauss, v1, dL1, grads)
    cap1 = int(r1.R * 1.25) + (1 << 14)
    for _ in range(5):
        fit_step_grads(gauss, v1, dL1, grads, R_cap=cap1, check="none")
    torch.cuda.synchronize()
    n1 = 50
    s1.record()
    for _ in range(n1):
        fit_step_grads(gauss, v1, dL1, grads, R_cap=cap1, check="none")
    e1.record()
    torch.cuda.synchronize()
    single_ms = s1.elapsed_time(e1) / n1
    single_graph_ms = None
    if not args.no_graph:
        g1 = GraphedFitStep(gauss, v1, dL1, grads, R_cap=cap1)
        for _ in range(5):
            g1.replay()
        torch.cuda.synchronize()
        s1.record()
        for _ in range(n1):
            g1.replay()
        e1.record()
        torch.cuda.synchronize()
        single_graph_ms = s1.elapsed_time(e1) / n1
        if g1.status()[1]:
            raise RuntimeError("bench: overflow in the single-view graph")
    out["single_view"] = {"ms_per_view_eager": single_ms, "views_per_s_eager": 1000.0 / single_ms,
                          "ms_per_view_graph": single_graph_ms,
                          "views_per_s_graph": (1000.0 / single_graph_ms) if single_graph_ms else None,
                          "note": "1 view per call (the shape the reference runs), L2 warm"}

    # ---- GPU-class denominator: an upstream-STRUCTURED stand-in of stages K2-K7 (baseline_standin/: CUB scan +
    # blocking D2H of the instance count, global 64-bit CUB radix sort, one CTA per tile, per-thread atomicAdd)
    # on the same 8 views, one view per call as the reference's loop does (renderer_one_shot.py:494-503),
    # against libghr's same stages.  kind = "restatement": NOT the reference's package, earns no "x upstream" claim.
    try:
        out["gpu_baseline"] = gpu_baseline_leg(torch, api, util, gauss, cams, bg, dL, dev, s1, e1)
    except Exception as ex:                       # bench-only leg: never takes the line down
        out["gpu_baseline"] = {"kind": "restatement", "unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    # ---- the reference-as-shipped shape: 98,562 Gaussians, 256x256, one view per
    # call through the drop-in GaussianRasterizer, an RGB render and an all-ones mask render of the same
    # geometry per view (renderer_one_shot.py:338-346, :372-379), both differentiated ----
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    sc_s = scenes.two_hand_scene(98562, seed=0)
    cam_s = scenes.fibonacci_cameras(4, 256, 256, seed=0)
    leafs = [t(x).requires_grad_(True) for x in (sc_s.means3D, sc_s.opacities, sc_s.scales, sc_s.rotations, sc_s.colors)]
    ones_s = torch.ones_like(leafs[0])
    w_s = t((np.random.default_rng(5).normal(size=(3, 256, 256)) / 65536).astype(np.float32))

    def settings_of(c, bgv):
        return GaussianRasterizationSettings(
            image_height=c.H, image_width=c.W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bgv, scale_modifier=1.0,
            viewmatrix=t(c.viewmatrix), projmatrix=t(c.projmatrix), sh_degree=0, campos=t(c.campos),
            prefiltered=False, debug=False)
    rs_s = [settings_of(c, torch.zeros(3, device=dev)) for c in cam_s]

    def shipped_pair(i, fused):
        xyz, op, scl, rot, colr = leafs
        m2d = torch.zeros_like(xyz, requires_grad=True)
        r = GaussianRasterizer(raster_settings=rs_s[i % len(rs_s)])
        if fused:
            img, _, msk = r.forward_with_mask(means3D=xyz, means2D=m2d, opacities=op, colors_precomp=colr, scales=scl,
                                              rotations=rot)
            loss = (img * w_s).sum() + (msk * w_s[0]).sum()
        else:
            img, _ = r(means3D=xyz, means2D=m2d, shs=None, colors_precomp=colr, opacities=op, scales=scl,
                       rotations=rot, cov3D_precomp=None)
            msk, _ = r(means3D=xyz, means2D=m2d, shs=None, colors_precomp=ones_s, opacities=op, scales=scl,
                       rotations=rot, cov3D_precomp=None)
            loss = (img * w_s).sum() + (msk[0] * w_s[0]).sum()
        loss.backward()

    shipped = {}
    for name, fused in (("two_calls", False), ("fused_mask", True)):
        for i in range(5):
            shipped_pair(i, fused)
        torch.cuda.synchronize()
        s1.record()
        for i in range(40):
            shipped_pair(i, fused)
        e1.record()
        torch.cuda.synchronize()
        shipped[name + "_pairs_per_s"] = 40 / (s1.elapsed_time(e1) * 1e-3)
    out["as_shipped"] = {**shipped, "note": "reference-as-shipped shape: 98,562 Gaussians, 256x256, one view per "
                         "call through the drop-in GaussianRasterizer + autograd (eager), RGB + all-ones mask "
                         "render pair fwd+bwd; two_calls = the reference's call pattern unchanged, fused_mask = "
                         "forward_with_mask (coverage from the same pass)"}
    return out


# ----------------------------------------------------------------------------- other BASELINE configs

def run_config(args):
    """--config c1 | c4 | c5: one line per BASELINE.json configuration other than the headline one."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.config
    base = {"steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "n_gpus": world}
    if cfg == "c1":
        # BASELINE config 1: single right hand, 30k Gaussians, SH degree 0, one 256x256 camera, fwd+bwd through the
        # pure-PyTorch CPU re-expression (oracle/torch_ref.py, autograd), all host threads
        if rank != 0:
            return
        from guassianhand_b200 import scenes
        from oracle import torch_ref
        nthr = os.cpu_count()
        torch.set_num_threads(nthr)
        sc = scenes.two_hand_scene(30000, seed=0, hands=1, sh_degree=0)
        cam = scenes.fibonacci_cameras(4, 256, 256, seed=0)[0]
        dL = (np.random.default_rng(1).normal(size=(3, 256, 256)) / 65536).astype(np.float32)
        bg = np.zeros(3, np.float32)
        for _ in range(max(1, min(args.warmup, 2))):
            torch_ref.forward_backward(sc, cam, bg, dL, threads=nthr)
        n = max(1, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n):
            torch_ref.forward_backward(sc, cam, bg, dL, threads=nthr)
        dt = time.perf_counter() - t0
        vps = n / dt
        cvps, cdone, cthr = cpu_views_per_s(64, threads=nthr, budget_s=5.0, P=30000, H=256, W=256, hands=1, sh_degree=0)
        line = {**base, "impl": "reference", "metric": "fwd+bwd views/s @256x256 one-hand 30k Gaussians SH0 (CPU)",
                "value": vps, "unit": UNIT, "steps": n, "ms_per_step": 1000 * dt / n,
                "config": {"workload": "BASELINE config 1: one hand, 30000 Gaussians, SH degree 0, 256x256, 1 view fwd+bwd, "
                                       "oracle/torch_ref.py (pure PyTorch, autograd) on the host CPU"},
                "cpu_baseline": {"value": vps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{n} views, torch {torch.__version__} CPU, {torch.get_num_threads()} threads"},
                "cpu_c_oracle": {"value": cvps, "unit": UNIT, "cores": cthr, "views": cdone,
                                 "note": "same config through oracle/gs_oracle.c (OpenMP)"},
                "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    # GPU configs
    import torch.distributed as dist
    from guassianhand_b200 import _native as NV, api, scenes
    from guassianhand_b200.dist import PackedGrads, fit_step_grads
    import util
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    bg = np.zeros(3, np.float32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K, Wm = args.steps, max(args.warmup, 3)
    hbm_peak, peak_src, _ = load_peaks()
    if cfg == "c4":
        # BASELINE config 4: 1M Gaussians (hand geometry tiled 4x4), SH degree 3, 1024x1024, 16 views fwd+bwd;
        # views are sharded over the ranks, a step = V views per rank in calls of Bc views
        P, H, W, M, V = 1000000, 1024, 1024, 16, max(1, 16 // world)
        Bc = min(4, V)
        sc = scenes.two_hand_scene(P, seed=0, sh_degree=3, tile=4)
        cams = scenes.fibonacci_cameras(16, H, W, seed=0)
        mine = cams[rank * V:(rank + 1) * V]
        gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
                     shs=t(sc.shs))
        N, T = H * W, ((W + 15) // 16) * ((H + 15) // 16)
        groups = [util.gpu_views(mine[i:i + Bc], bg, dev, sh_degree=3) for i in range(0, V, Bc)]
        dL = t((np.random.default_rng(1 + rank).normal(size=(Bc, 3, H, W)) / N).astype(np.float32))
        grads = PackedGrads(P, M, device=dev, peer=world > 1)
        caps, Rs, pairs = [], [], []
        for vg in groups:
            nv = vg.viewmatrix.shape[0]
            r = fit_step_grads(gauss, vg, dL[:nv], grads, sh_degree=3, group=False)
            caps.append(int(r.R * 1.1) + (1 << 16))
            Rs.append(r.R)
            lay = NV.layout(P, nv, H, W, M, 3, r.R_cap)
            pairs.append(int(r.state[lay.off_ncontrib: lay.off_ncontrib + nv * N * 4].view(torch.int32)
                             .sum(dtype=torch.int64).item()))
            del r

        def one_step(fe=None, be=None):
            # (gradients of the calls of a step would be accumulated by the caller; the timing does not depend on it)
            for gi, (vg, cap) in enumerate(zip(groups, caps)):
                nv = vg.viewmatrix.shape[0]
                fit_step_grads(gauss, vg, dL[:nv], grads, sh_degree=3, R_cap=cap, check="none", group=False,
                               fwd_events=fe[gi] if fe else None, bwd_events=be[gi] if be else None)
            grads.all_reduce_()
        for _ in range(Wm):
            one_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.zero_()
            evs[i][0].record()
            one_step()
            evs[i][1].record()
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in evs])
        tm = torch.tensor([ms.sum()], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total = float(tm.item())
        value = world * V * K / (total / 1000)
        fe = [NV.StageEvents(NV.GHR_NSTAGES_FWD) for _ in groups]
        be = [NV.StageEvents(NV.GHR_NSTAGES_BWD) for _ in groups]
        flush.zero_()
        one_step(fe, be)
        torch.cuda.synchronize()
        stage_ms = {}
        for si, name in enumerate(NV.FWD_STAGES):
            stage_ms[name] = float(sum(e.elapsed_ms(si) for e in fe))
        for si, name in enumerate(NV.BWD_STAGES):
            stage_ms[name] = float(sum(e.elapsed_ms(si) for e in be))
        if rank == 0:
            roof = roofline_report(stage_ms, total / K, P, V, N, T, float(sum(Rs)), float(sum(pairs)), hbm_peak,
                                   FP32_NOMINAL_TFLOPS, peak_src, None, M=M)
            line = {**base, "metric": "fwd+bwd views/s @1024x1024 1M Gaussians SH3", "value": value, "unit": UNIT,
                    "steps": K, "warmup": Wm, "ms_per_step": total / K,
                    "config": {"workload": f"BASELINE config 4: 1M Gaussians (two-hand geometry tiled 4x4), SH degree 3, "
                                           f"1024x1024, {V} views per rank per step in calls of {Bc}, fwd+bwd, eager launches",
                               "instances_per_step": float(sum(Rs)), "blend_pairs_per_step": float(sum(pairs)),
                               "l2": "flushed between timed steps"},
                    "roofline": roof, "stage_ms": stage_ms,
                    "e2e": {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                            "note": "not measured for this configuration (see the headline config)"},
                    "gpu_launches": (13 * len(groups) + (1 if world > 1 else 0)) * K}
            print(json.dumps(line), flush=True)
    elif cfg == "c5":
        # BASELINE config 5: novel-pose drive render, forward only, 1920x1080, 256 poses sharded over 8 ranks
        # (32 poses per rank; per-pose rigid + noise perturbation of the means), calls of 4 views
        P, H, W, n_pose, Bc = 60000, 1080, 1920, 32, 4
        sc = scenes.two_hand_scene(P, seed=0)
        cams = scenes.fibonacci_cameras(64, H, W, seed=2)
        rng = np.random.default_rng(10 + rank)
        base_means = t(sc.means3D)
        gauss = dict(opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations), colors=t(sc.colors))
        poses = [base_means + t((rng.normal(size=(1, 3)) * 0.01).astype(np.float32)) +
                 t((rng.normal(size=sc.means3D.shape) * 2e-4).astype(np.float32)) for _ in range(8)]
        vgs = [util.gpu_views([cams[(rank * n_pose + i + j) % 64] for j in range(Bc)], bg, dev)
               for i in range(0, n_pose, Bc)]
        r0 = api.forward_raw(vgs[0].cams(), poses[0], gauss["opacities"], gauss["scales"], gauss["rotations"], None, None,
                             gauss["colors"], 0, 1.0)
        cap = int(r0.R * 1.5) + (1 << 16)
        del r0

        def one_step():
            for i, vg in enumerate(vgs):
                api.forward_raw(vg.cams(), poses[i % len(poses)], gauss["opacities"], gauss["scales"], gauss["rotations"],
                                None, None, gauss["colors"], 0, 1.0, check="none", R_cap=cap)
        for _ in range(Wm):
            one_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.zero_()
            evs[i][0].record()
            one_step()
            evs[i][1].record()
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in evs])
        tm = torch.tensor([ms.sum()], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total = float(tm.item())
        value = world * n_pose * K / (total / 1000)
        if rank == 0:
            line = {**base, "metric": "forward poses/s @1920x1080 two-hand Gaussians", "value": value, "unit": "poses/s",
                    "steps": K, "warmup": Wm, "ms_per_step": total / K,
                    "config": {"workload": f"BASELINE config 5: forward only, 1920x1080, 60000 Gaussians, {n_pose} poses per "
                                           f"rank per step in calls of {Bc} views, pose-sharded (no collective), eager launches",
                               "l2": "flushed between timed steps"},
                    "e2e": {"value": None, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                            "note": "not measured for this configuration"},
                    "gpu_launches": 9 * len(vgs) * K}
            print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--views", type=int, default=8, help="views per rank per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c1", "c4", "c5"],
                    help="c2/c3 = the headline workload (c3 = the same under torchrun); c1, c4, c5 = the other BASELINE configs")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"],
                    help="N>1: libghr's NVLink peer-memory all-reduce kernel (default) or NCCL")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="skip the single-view / as-shipped / cull-efficiency extras")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--overlap", type=int, default=2,
                    help="view groups of a step run as concurrent chains on this many streams (graph mode)")
    args = ap.parse_args()
    if args.config in ("c1", "c4", "c5"):
        run_config(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
