"""bench.py -- headline benchmark: spectral path samples/s and ms/frame at 1080p.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): `config - Apollo 11.txt` full-disk view, 1920x1080, 1024 spp,
seeded synthetic 8192x4096 textures (the NASA maps are not available offline).  One STEP = one full
frame = W*H*spp path samples through the product (wavefront) integrator.

  value   : path samples/s, inputs resident in HBM, device time of K steps (CUDA events, max over ranks)
  e2e     : the same metric through the reference-facing API with host buffers: per step the scene
            parameters come from the host config (H2D), the frame is rendered, resolved/tonemapped and
            copied back into pinned host memory (D2H), like one Renderer.accumulate()*spp + fetch_image()
  roofline: FP32/SFU issue roofline of the render kernel (this path is neither HBM- nor
            tensor-bound): algorithmic FLOP/path from the event counters x paths / kernel time against
            148 SM x 128 lanes x 2 x sm_max_clock; `traffic` = DRAM bytes/launch from ncu if captured
  cpu_baseline : the CPU oracle (a port of the reference's Taichi code, oracle/de_oracle.c) on all host
            cores on a bounded sample (the same view at 480x272, 8 spp)
N > 1: the frame's samples are sliced across ranks (rank r renders sample indices [r*spp/N, (r+1)*spp/N)
of every pixel: perfectly balanced, the union is the 1-GPU sample set) and the float accumulation
buffers are sum-reduced to rank 0 with NCCL inside the timed region; total work is fixed -> "strong".
--partition tile / tile+spp splits by interleaved 16x8 film tiles (x sample slices) instead; after the timed
region rank 0 re-renders a 64x32 crop alone and the line carries the N-GPU == 1-GPU difference ("identity").
--impl reference: the reference itself needs Taichi (absent); its CPU implementation is therefore the
oracle port, timed on all host threads on the bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CFG_DIR = os.path.join(ROOT, "digital-earth_b200", "assets", "configs")
SCENES = {"apollo": "Apollo 11", "florida": "florida", "sunset": "sunset hurricane"}
METRIC, UNIT = "spectral_path_samples_per_sec_1080p", "path samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="apollo", choices=list(SCENES))
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--tex", default="8192x4096")
    ap.add_argument("--mode", default="wavefront", choices=["wavefront", "megakernel", "parity", "preview"])
    ap.add_argument("--cpu-res", default="480x272")
    ap.add_argument("--cpu-spp", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "fused"],
                    help="N>1: ncclReduce of the accumulation buffers, or rank 0's resolve kernel reading the peers' buffers over NVLink P2P")
    ap.add_argument("--partition", default="spp", choices=["spp", "tile", "tile+spp"],
                    help="N>1: sample slices, interleaved 16x8 film tiles, or tile groups x sample slices (SURVEY 8e)")
    ap.add_argument("--tile-groups", type=int, default=0, help="tile+spp: number of tile groups (default 2)")
    ap.add_argument("--no-identity", action="store_true", help="N>1: skip the N-GPU == 1-GPU check of a 64x32 crop after the timed region")
    return ap.parse_args()


def tile_groups_of(a, world):
    if a.partition == "spp" or world == 1:
        return 1
    if a.partition == "tile":
        return world
    g = a.tile_groups or 2
    if world % g:
        raise SystemExit("--tile-groups %d does not divide %d ranks" % (g, world))
    return g


def ncu_metrics(a):
    """issue / pipe / lane figures of the render kernel are hardware counters: they come from the committed ncu capture of this
    view (profiles/r2_ncu_metrics.json, written by tools/ncu_metrics_json.py from the .ncu-rep), never from a run under a profiler."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_metrics.json"))).get(a.scene)
    except Exception:
        return None


def flop_per_path(c):
    """60*N_rmo + 45*N_cloud + 45*N_sdf + 400*N_seg + 200 (minimal necessary work)."""
    p = max(c["paths"], 1)
    return 60.0 * c["rmo_steps"] / p + 45.0 * c["cloud_steps"] / p + 45.0 * c["sdf_evals"] / p + 400.0 * c["segments"] / p + 200.0


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def scene_cfg(a):
    import digital_earth_b200 as de
    return de.load_config(os.path.join(CFG_DIR, "config - %s.txt" % SCENES[a.scene]))


def textures_for(a, tw, th):
    import digital_earth_b200 as de
    return de.textures.synthetic(tw, th, cloud_cover=0.8 if a.scene == "sunset" else 0.5, hurricane=a.scene == "sunset", seed=0)


def oracle_events(a):
    """N_* of SURVEY 8(d)'s work model: the ORACLE's event counters for this view (the reference algorithm's own step counts;
    the product kernel removes null collisions and misses, which must not shrink the numerator).  Read from the committed
    fixture profiles/oracle_events.json (tools/oracle_events.py); the cpu_baseline leg refreshes them live when it runs."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "oracle_events.json"))).get("%s_%s" % (a.scene, a.tex))
    except Exception:
        return None


def cpu_baseline(a, steps=1, textures=None):
    """Oracle port on all host cores, bounded sample of the same view; returns (paths/s, description, cores)."""
    from oracle import oracle as orc
    cfg = scene_cfg(a)
    cw, ch = map(int, a.cpu_res.split("x"))
    tw, th = map(int, a.tex.split("x"))
    tex = textures if textures is not None else textures_for(a, tw, th)
    s = orc.Scene(tex, cw, ch, cam_pos=cfg["cam_pos"], look_at=cfg["look_at"], up=cfg["up"], fov=cfg["fov"], aspect_scale=cfg["aspect_scale"],
                  sun_angle=cfg["sun_angle"], sun_path_rot=cfg["sun_path_rot"])
    cores = os.cpu_count() or 1
    times = []
    for k in range(steps):
        t0 = time.perf_counter()
        _, cnt = orc.render(s, a.cpu_spp, first_sample=k * a.cpu_spp, nthreads=cores)
        times.append(time.perf_counter() - t0)
        cpu_baseline.events = cnt  # the reference algorithm's event counts of this view (SURVEY 8d work model)
    paths = cw * ch * a.cpu_spp
    sample = "%s view at %dx%d, %d spp (%d paths/step), %dx%d synthetic textures" % (SCENES[a.scene], cw, ch, a.cpu_spp, paths, tw, th)
    return paths, times, sample, cores, tex


def taichi_reference(a, tex):
    """The unmodified reference on real Taichi (ti.cpu), when both are on the machine (BASELINE.md baseline B): same bounded sample as the
    oracle port.  Returns (times, cores) or None -- neither Taichi nor the reference exist on this project's build container / GPU box."""
    try:
        from oracle import taichi_harness as th
        if not th.available()[0]:
            return None
        cw, ch = map(int, a.cpu_res.split("x"))
        ref = th.TaichiReference(tex, (cw, ch), archs=("cpu",))
        try:
            cfg = scene_cfg(a)
            for _ in range(max(a.warmup, 0)):
                ref.render(cfg, a.cpu_spp)
            return [ref.render(cfg, a.cpu_spp)[1] for _ in range(a.steps)], ref.cores
        finally:
            ref.close()
    except Exception as e:  # e.g. ti.Texture unsupported on the CPU backend of the installed Taichi
        print("taichi reference unavailable (%s: %s); timing the oracle port" % (type(e).__name__, e), file=sys.stderr)
        return None


def run_reference(a):
    """--impl reference: CPU implementation of the path -- the reference itself on Taichi's CPU backend when that is installed
    (kind "taichi"), else the oracle port (kind "port"; Taichi is not installable in this project's containers)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    paths, _, sample, cores, tex = cpu_baseline(a, steps=0)
    kind = "port"
    tr = taichi_reference(a, tex)
    if tr:
        times, cores, kind = tr[0], tr[1], "taichi"
    else:
        _, wt, _, _, _ = cpu_baseline(a, steps=max(a.warmup, 0), textures=tex) if a.warmup else (0, [], 0, 0, 0)
        _, times, _, _, _ = cpu_baseline(a, steps=a.steps, textures=tex)
    total = sum(times)
    v = paths * a.steps / total
    W, H = map(int, a.res.split("x"))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s (config - %s.txt), %dx%d, %d spp, synthetic %s textures" % (a.scene, SCENES[a.scene], W, H, a.spp, a.tex), "scene": a.scene,
                   "note": "CPU arm renders a bounded sample of this workload per step: " + sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    import digital_earth_b200 as de

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H = map(int, a.res.split("x"))
    tw, th = map(int, a.tex.split("x"))
    from digital_earth_b200.distributed import partition, reduce_accumulation, render_partition
    groups = tile_groups_of(a, world)
    part = partition(a.spp, rank, world, groups)
    first, spp_local = part["first_sample"], part["n_spp"]
    tiles = (part["tile_stride"], part["tile_offset"]) if groups > 1 else None

    tex = textures_for(a, tw, th)
    cfg = scene_cfg(a)
    r = de.Renderer((W, H), (0, 1, 0), textures=tex, device=local, mode=a.mode)
    r.apply_config(cfg)
    r.copy_textures()
    host_img = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    fused = a.exchange == "fused" and world > 1
    peers, token = [], torch.zeros(1, device=dev)
    if fused:  # map the other ranks' accumulation buffers once (CUDA IPC over NVLink peer memory)
        handles = [None] * world
        dist.all_gather_object(handles, r.export_accum_handle())
        if rank == 0:
            peers = [r.open_peer(handles[k]) for k in range(1, world)]
    peer_offsets = [k % groups for k in range(1, world)]

    def step(e2e):
        flush.fill_(1)  # evict L2 between timed iterations (textures alone are 416 MB > L2 as well)
        if e2e:
            r.apply_config(cfg)   # host -> device: scene parameters for this frame
        render_partition(r, part)  # reset + this rank's (tile group, sample slice)
        if fused:
            # exchange fused into the resolve: a 4-byte all-reduce is the stream-ordered "all ranks have rendered" barrier,
            # rank 0's resolve kernel then sums the peers' buffers in place, a second token releases the buffers
            dist.all_reduce(token)
            img = r.fetch_image_peers(peers, a.spp, tile_stride=groups, own_offset=0, peer_offsets=peer_offsets) if rank == 0 else None
            dist.all_reduce(token)
        else:
            reduce_accumulation(r.color_buffer, dst=0)  # ncclReduce(sum) over NVLink: the path's one exchange step (no-op at N=1)
            img = r.fetch_image(spp=a.spp) if e2e and rank == 0 else None
        if e2e and rank == 0:
            host_img.copy_(img.permute(1, 0, 2), non_blocking=True)  # device -> pinned host ([H][W][3] storage order)

    def timed(e2e, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step(e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # event counters -> algorithmic FLOP/path (outside the timed region, reduced spp)
    r.set_counting(True)
    r.reset_framebuffer(); r.accumulate(4, first_sample=first, tiles=tiles)
    counters = r.counters()
    r.set_counting(False)
    tex_peak = r.tex_gather_peak() if rank == 0 else None   # measured rate of the fetch instruction the kernel uses (outside the timed region)

    for _ in range(max(a.warmup, 0)):
        step(False)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(False, a.steps)
    sampler.stop_flag = True
    # kernel-only duration of the dominant kernel (render), measured live on its stream
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r.reset_framebuffer()
    k0.record(); r.accumulate(spp_local, first_sample=first, tiles=tiles); k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1)
    step(True)
    ms_e2e = timed(True, a.steps)

    # the reference's only shipped mode is ONE sample per displayed frame (earth_viewer.py:186,241-243): accumulate() + fetch_image(),
    # measured as such on this GPU (median of 9, warm L2 -- consecutive frames of a progressive render)
    one = []
    r.reset_framebuffer()
    for _ in range(12):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(); r.accumulate(1); r.fetch_image(); f1.record()
        torch.cuda.synchronize()
        one.append(f0.elapsed_time(f1))
    ms_1spp = sorted(one[3:])[len(one[3:]) // 2]

    # N-GPU == 1-GPU (SURVEY 8e determinism check): the reduced frame of one more partitioned render against rank 0 rendering a
    # 64x32 crop of it alone with the same (pixel, sample) keys; equal up to float summation order
    identity = None
    if world > 1 and not a.no_identity:
        render_partition(r, part)
        reduce_accumulation(r.color_buffer, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            cx, cy = (W // 2 - 32) // 16 * 16, (H // 2 - 16) // 8 * 8
            multi = r.color_buffer[cy:cy + 32, cx:cx + 64].clone()
            r.reset_framebuffer(); r.accumulate(a.spp, window=(cx, cy, 64, 32), first_sample=0)
            single = r.color_buffer[cy:cy + 32, cx:cx + 64]
            den = float(single.abs().max())
            identity = {"crop": [cx, cy, 64, 32], "max_abs_diff_over_max": float((multi - single).abs().max()) / max(den, 1e-30),
                        "rel_diff_of_mean": abs(float(multi.mean()) - float(single.mean())) / max(abs(float(single.mean())), 1e-30)}
        dist.barrier()

    paths_step = W * H * a.spp
    value = paths_step * a.steps / (ms * 1e-3)
    e2e_value = paths_step * a.steps / (ms_e2e * 1e-3)
    clocks = sampler.summary()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        cpu = None
        if world == 1 and not a.no_cpu_baseline:  # the one leg that executes oracle/: CPU baseline + live event counts of the view
            paths, times, sample, cores, _ = cpu_baseline(a, steps=1, textures=tex)
            cpu = {"value": paths / times[0], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        ocnt = getattr(cpu_baseline, "events", None) or oracle_events(a)
        events_src = "oracle counters of this run's cpu_baseline leg" if getattr(cpu_baseline, "events", None) else "profiles/oracle_events.json (oracle counters)"
        if not ocnt:  # unknown view: the kernel's own counters under-count the reference's work; say so
            ocnt, events_src = counters, "KERNEL counters (no oracle fixture for this view): lower bound of the reference's work"
        fpp = flop_per_path(ocnt)
        peak_tflops = 148 * 128 * 2 * sm_max * 1e6 / 1e12  # FP32 FMA issue peak
        paths_launch = W * H * spp_local / groups                   # paths of this rank's launch (tile groups render 1/groups of the film)
        achieved = fpp * paths_launch / (kernel_ms * 1e-3) / 1e12
        executed = flop_per_path(counters) * paths_launch / (kernel_ms * 1e-3) / 1e12
        kpp = {k: counters[k] / max(counters["paths"], 1) for k in ("segments", "rmo_steps", "cloud_steps", "sdf_evals", "tex_fetches", "surface_hits")}
        fetch_rate = kpp["tex_fetches"] * paths_launch / (kernel_ms * 1e-3)   # 2x2 footprints per second (one gather each for r8 maps)
        hw = ncu_metrics(a) or {}
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = prof.get("%s_%s_%d" % (a.scene, a.res, a.spp))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s (config - %s.txt), %dx%d, %d spp, synthetic %dx%d textures" % (a.scene, SCENES[a.scene], W, H, a.spp, tw, th),
                       "scene": a.scene, "integrator": a.mode, "ms_per_frame": ms / a.steps, "ms_per_spp": ms / a.steps / a.spp,
                       "ms_1spp": ms_1spp,
                       "ms_1spp_note": "measured: one accumulate(1) + fetch_image() per frame (the reference's interactive mode), median of 9; ms_per_spp is the amortised figure",
                       "partition": ("%s + %s" % ("spp-slice x%d" % world if groups == 1 else ("interleaved 16x8 film tiles x%d" % world if groups == world else
                                                  "%d tile groups x %d spp slices" % (groups, world // groups)),
                                                  "resolve kernel reading the peers' f32 accumulation buffers over NVLink P2P (CUDA IPC)" if fused
                                                  else "ncclReduce(sum) of the f32 accumulation buffer")) if world > 1 else "single GPU",
                       "l2": "256 MiB flush between steps; textures (416 MB) exceed the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps, "h2d_bytes_per_step": 96, "d2h_bytes_per_step": W * H * 3 * 4},
            "gpu_launches": a.steps * (3 if fused else 2),  # k_render_wavefront + k_space_tiles (+ k_resolve_peers with --exchange fused) per step; memsets and NCCL are not ours
            "roofline": {"bound": "fp32_issue", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         "traffic": traffic, "kernel": "k_render_wavefront", "kernel_ms": kernel_ms, "flop_per_path": fpp,
                         "peak_source": "148 SM x 128 FP32 lanes x 2 x %.0f MHz (max SM clock of MEASURED_PEAKS.json); HBM/tensor peaks do not bound this path" % sm_max,
                         "flop_model": "60*N_rmo+45*N_cloud+45*N_sdf+400*N_seg+200 (SURVEY 8d); N from " + events_src,
                         "oracle_events_per_path": {k: ocnt[k] / max(ocnt["paths"], 1) for k in ("segments", "rmo_steps", "cloud_steps", "sdf_evals", "tex_fetches", "surface_hits")},
                         "kernel_events_per_path": kpp,
                         "note": "achieved / frac = the REFERENCE algorithm's work (oracle event counts, SURVEY 8d) per second: a reference-equivalent rate. "
                                 "executed_* = the same FLOP model on the kernel's own counters (local majorants and miss tests remove 2-4x of the steps): hardware utilisation",
                         "executed_tflops": executed, "executed_frac": executed / peak_tflops,
                         "tex_fetches_per_s": fetch_rate, "tex_peak_fetches_per_s": tex_peak, "tex_frac": (fetch_rate / tex_peak) if tex_peak else None,
                         "tex_peak_source": "de_bench_tex_gather: tex2Dgather r8, L1-resident footprints, 2048 threads/SM, measured in this run",
                         "issue_active_pct": hw.get("issue_active_pct"), "lanes_per_inst": hw.get("lanes_per_inst"), "xu_pct": hw.get("xu_pct"),
                         "alu_pct": hw.get("alu_pct"), "fma_pct": hw.get("fma_pct"), "l1tex_hit_pct": hw.get("l1tex_hit_pct"), "l2_hit_pct": hw.get("l2_hit_pct"),
                         "icache_hit_pct": hw.get("icc_hit_pct"), "gpc_icache_requests_pct_of_peak": hw.get("gcc_inst_requests_pct_of_peak"),
                         "stall_no_instruction_per_issue": hw.get("stall_no_instruction"),
                         "hw_counter_source": hw.get("source")},
            "clocks": clocks,
        }
        if identity:
            line["identity"] = identity
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if fused:
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            r.close_peers()
        dist.barrier()
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
