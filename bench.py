#!/usr/bin/env python
"""bench.py — pose samples/sec of the per-image sampling hot path on N B200s (BASELINE.json metric).

One step = one pass of the hot path over one batch of synthetic input on every rank:
    (B,18,256,256) fp32 image -> ResNet-50 encoder -> heads -> 23-joint SO(3) flow (N draws / image + point
    estimate) -> SMPL LBS of the B*N sampled bodies (6890 vertices + 90 joints each).
Workload = BASELINE.json configs[2] at the evaluation batch of configs[1]/[3]: B=32, N=100 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm (CUDA, C-ABI library)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # CPU restatement of the reference path

N > 1 is launched by torchrun (one rank per GPU); ranks shard the image axis (weak scaling: B images per GPU),
there is no data-path collective; one all_gather collects the per-image metric rows at the end of every step.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'pose_samples_per_sec'
UNIT = 'samples/s'
LBS_BYTES_PER_SAMPLE = 84664            # 6890*3*4 + 90*3*4 + 24*9*4 + 10*4
LBS_FLOP_PER_SAMPLE = 17.28e6           # dense, as the reference computes it
FLOW_FLOP_PER_SAMPLE = 1.684e6
FLOW_BYTES_PER_SAMPLE = 1148
ENC_FLOP_PER_IMAGE = {50: 12.22e9, 18: 6.28e9}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def build_cpu_problem(B, N, layers, seed=0):
    import humaniflow_b200 as hb
    from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_proxy_input, synthetic_smpl_data
    torch.manual_seed(seed)
    cfg = hb.get_model_cfg_defaults()
    cfg.NUM_RESNET_LAYERS = layers
    m = hb.HumaniflowModel('cpu', cfg, SMPL_PARENTS).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    data = synthetic_smpl_data(seed=0, skinning='body_parts')
    x = synthetic_proxy_input(B, 18, 256, seed=1)
    g = torch.Generator().manual_seed(2)
    z = torch.randn(B, N, 23, 3, generator=g) * 0.6
    se = torch.randn(B, N, 10, generator=g)
    return cfg, sd, data, x, z, se, SMPL_PARENTS


def cpu_step(cfg, sd, data, x, z, se, parents):
    """The reference path restated on the CPU (oracle/): encoder -> heads -> flow samples -> SMPL LBS."""
    from oracle import model as om
    from oracle import smpl as osmpl
    B, N = z.shape[:2]
    with torch.no_grad():
        out = om.forward(sd, cfg, parents, input=x, num_samples=N, shape_eps=se, base_noise=z)
        R = out['pose_rotmats_samples'].reshape(B * N, 23, 3, 3)
        glob = out['glob_rotmat'][:, None].expand(-1, N, -1, -1).reshape(B * N, 1, 3, 3)
        v, j = osmpl.smpl_forward(data, out['shape_samples'].reshape(B * N, 10), R, glob, pose2rot=False)
    return v, j


def time_cpu(B, N, layers, steps, warmup):
    prob = build_cpu_problem(B, N, layers)
    for _ in range(warmup):
        cpu_step(*prob)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_step(*prob)
        ts.append(time.perf_counter() - t0)
    return B * N * len(ts) / sum(ts), sum(ts) / len(ts)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    B_ref = min(args.B, args.ref_images)
    val, sec = time_cpu(B_ref, args.N, args.layers, args.steps, args.warmup)
    line = {
        'metric': METRIC, 'value': val, 'unit': UNIT, 'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic (random-init weights; SMPL-shaped synthetic body model, skinning weights grouped by body part like the real SMPL)',
        'config': workload_config(args, images_per_step=B_ref),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': '%d images x %d samples per step (of the %d-image batch), oracle/ = CPU restatement of the '
                                   'reference path (pyro/smplx not installable: SURVEY F3)' % (B_ref, args.N, args.B)},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(args, images_per_step=None):
    return {'workload': 'predict_B%d_N%d_resnet%d_18ch_256px_smpl6890' % (args.B, args.N, args.layers),
            'images_per_gpu_per_step': images_per_step if images_per_step is not None else args.B,
            'samples_per_image': args.N, 'encoder': 'resnet%d' % args.layers, 'joints': 23, 'vertices': 6890,
            'l2_policy': 'inputs+outputs (151 MB in, 268 MB out per step) exceed the 126 MB L2; no explicit flush',
            'sharding': 'image axis across ranks, all_gather of per-image metric rows'}


# ------------------------------------------------------------------------------------------------ our arm
_FULL_AFFINITY = None


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank to the CPUs NVML reports as local to its GPU before any pinned host buffer is allocated, so
    first-touch places the staging buffers on the GPU's own NUMA node (matters for the H2D/D2H legs at N > 1)."""
    global _FULL_AFFINITY
    try:
        _FULL_AFFINITY = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def run_ours(args):
    import humaniflow_b200 as hb
    from humaniflow_b200 import _lib
    from humaniflow_b200.sharding import gather_rows, sample_diversity_rows
    from humaniflow_b200.synthetic import SMPL_PARENTS, synthetic_proxy_input, synthetic_smpl_data
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    B, N = args.B, args.N
    torch.manual_seed(0)
    cfg = hb.get_model_cfg_defaults()
    cfg.NUM_RESNET_LAYERS = args.layers
    model = hb.HumaniflowModel(dev, cfg, SMPL_PARENTS).eval().to(dev)
    smpl = hb.SMPL.from_arrays(synthetic_smpl_data(seed=0, skinning='body_parts'), create_transl=False).to(dev)
    x_host = synthetic_proxy_input(B, 18, 256, seed=1 + rank).pin_memory()
    x_dev = x_host.to(dev)
    g = torch.Generator().manual_seed(2 + rank)
    z = (torch.randn(B, N, 23, 3, generator=g) * 0.6).to(dev)
    se = torch.randn(B, N, 10, generator=g).to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    pending, last_metric = [None], [None]

    def step(x, marks=None, outs=None):
        out = model(x, num_samples=N, base_noise=z, shape_eps=se)
        if marks is not None:
            marks[0].record()
        R = out['pose_rotmats_samples'].view(B * N, 23, 3, 3)
        glob = out['glob_rotmat'][:, None].expand(-1, N, -1, -1).reshape(B * N, 1, 3, 3)
        so = smpl(betas=out['shape_samples'].view(B * N, 10), body_pose=R, global_orient=glob, pose2rot=False,
                  out_vertices=None if outs is None else outs[0], out_joints=None if outs is None else outs[1])
        if marks is not None:
            marks[1].record()
        # per-image metric row (sample diversity: mean over joints of the std over samples), gathered across ranks;
        # the collective is asynchronous: its result is read when the NEXT step issues its own gather, so the cross-rank
        # rendezvous never stalls the following step's kernels
        if pending[0] is not None:
            last_metric[0] = pending[0].result()
        pending[0] = gather_rows(sample_diversity_rows(so.joints, B, N), num_images=world * B, async_op=True)
        return so, pending[0]

    def sync_all():
        if pending[0] is not None:
            last_metric[0] = pending[0].result()
            pending[0] = None
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = ev(), ev()
    enc_marks = [(ev(), ev(), ev()) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        enc_marks[i][2].record()
        step(x_dev, marks=enc_marks[i])
    e1.record()
    sync_all()
    launches = _lib.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B * N / (ms_step * 1e-3)
    ms_model = statistics.mean(m[2].elapsed_time(m[0]) for m in enc_marks)
    ms_lbs = statistics.mean(m[0].elapsed_time(m[1]) for m in enc_marks)

    # ---------------- per-stage kernel times, measured live with CUDA events on the launching stream
    def time_call(fn, iters=10):
        fn(); torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    with torch.no_grad():
        feats = model.image_encoder(x_dev)
        ms_enc = time_call(lambda: model.image_encoder(x_dev))
        ms_flow = time_call(lambda: model(None, input_feats=feats, num_samples=N, base_noise=z, shape_eps=se))
        out = model(None, input_feats=feats, num_samples=N, base_noise=z, shape_eps=se)
        R = out['pose_rotmats_samples'].view(B * N, 23, 3, 3)
        full = torch.cat([out['glob_rotmat'][:, None].expand(-1, N, -1, -1).reshape(B * N, 1, 3, 3), R], 1).contiguous()
        betas = out['shape_samples'].view(B * N, 10).contiguous()
        ms_lbs_k = time_call(lambda: smpl.lbs(betas, full))
    pk = peaks()
    stages = {
        'encoder': {'ms': ms_enc, 'bound': 'tensor', 'achieved': ENC_FLOP_PER_IMAGE[args.layers] * B / (ms_enc * 1e-3) / 1e12,
                    'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s'},
        'flow': {'ms': ms_flow, 'bound': 'fp32-fma', 'achieved_tflops': FLOW_FLOP_PER_SAMPLE * B * N / (ms_flow * 1e-3) / 1e12,
                 'achieved_gbs': FLOW_BYTES_PER_SAMPLE * B * N / (ms_flow * 1e-3) / 1e9},
        'lbs': {'ms': ms_lbs_k, 'bound': 'hbm', 'achieved': LBS_BYTES_PER_SAMPLE * B * N / (ms_lbs_k * 1e-3) / 1e9,
                'peak': pk['hbm_gbs'], 'unit': 'GB/s'},
    }
    for s in ('encoder', 'lbs'):
        stages[s]['frac'] = stages[s]['achieved'] / stages[s]['peak']
    dom = 'encoder' if ms_enc >= ms_lbs_k else 'lbs'
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    if dom == 'encoder' and args.layers == 50 and B == 32 and os.path.exists(tpath):
        traffic = json.load(open(tpath))['dram_bytes_per_step']     # ncu dram bytes of the 49 conv launches of one forward
    roofline = {'kernel': 'conv_tcgen05_kernel (ResNet-%d trunk, 49 launches per step)' % args.layers if dom == 'encoder' else 'lbs_skin_tc2_kernel (+ pose / extra-joint kernels)',
                'bound': stages[dom]['bound'], 'achieved': stages[dom]['achieved'], 'peak': stages[dom]['peak'],
                'unit': stages[dom]['unit'], 'frac': stages[dom]['frac'], 'traffic': traffic,
                'peak_source': pk['source'] + (' (sustained bf16)' if dom == 'encoder' else ' (copy bandwidth)')}

    # ---------------- end to end through the public API with HOST buffers (H2D of the images, D2H of the meshes)
    # Every step copies its own input from pinned host memory and its own result back to pinned host memory, all
    # inside the timed region.  Steps are software-pipelined over three streams (H2D | compute | D2H) with
    # double-buffered staging, so the copy of step i+1 / i-1 overlaps the kernels of step i (PCIe is full duplex).
    V = smpl.v_template.shape[0]
    v_host = [torch.empty(B * N, V, 3).pin_memory() for _ in range(2)]
    j_host = [torch.empty(B * N, smpl.num_joints_out, 3).pin_memory() for _ in range(2)]
    x_stage = [torch.empty_like(x_dev) for _ in range(2)]
    out_dev = [(torch.empty(B * N, V, 3, device=dev), torch.empty(B * N, smpl.num_joints_out, 3, device=dev)) for _ in range(2)]
    ev_copied = [torch.cuda.Event() for _ in range(2)]  # outputs of buffer b have been copied to the host
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    s_main = torch.cuda.current_stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]      # input buffer b filled
    ev_used = [torch.cuda.Event() for _ in range(2)]    # input buffer b consumed by the encoder
    ev_done = [torch.cuda.Event() for _ in range(2)]    # outputs of the step using buffer b are ready
    keep = [None, None]

    def e2e_steps(n):
        for b in range(2):
            ev_used[b].record(s_main)
            ev_copied[b].record(s_out)
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_used[b])
                x_stage[b].copy_(x_host, non_blocking=True)
                ev_in[b].record(s_in)
            s_main.wait_event(ev_in[b])
            s_main.wait_event(ev_copied[b])           # device output buffer b is free again
            so, metric = step(x_stage[b], outs=out_dev[b])
            ev_used[b].record(s_main)
            ev_done[b].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                v_host[b].copy_(so.vertices, non_blocking=True)
                j_host[b].copy_(so.joints, non_blocking=True)
                ev_copied[b].record(s_out)
            keep[b] = so
        s_main.wait_stream(s_out)
        s_main.wait_stream(s_in)

    # evaluate-style end to end (BASELINE configs[3]): the per-sample PVE / PVE-SC / PVE-PA errors against a
    # target mesh per image are reduced on the device (hf_pointset_errors) and only the per-image rows go back to the host
    from humaniflow_b200.metrics import pointset_errors, samples_min
    tgt = smpl.tpose(torch.zeros(B, 10, device=dev)).vertices.clone()                 # synthetic ground-truth meshes
    rows_host = [torch.empty(B, 4).pin_memory() for _ in range(2)]
    ev_rows = [torch.cuda.Event() for _ in range(2)]

    def eval_steps(n):
        for b in range(2):
            ev_used[b].record(s_main)
            ev_rows[b].record(s_out)
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_used[b])
                x_stage[b].copy_(x_host, non_blocking=True)
                ev_in[b].record(s_in)
            s_main.wait_event(ev_in[b])
            s_main.wait_event(ev_rows[b])
            so, metric = step(x_stage[b], outs=out_dev[b])
            err = pointset_errors(so.vertices.view(B, N, V, 3), tgt)
            rows = torch.stack([samples_min(err['plain']), samples_min(err['sc']), samples_min(err['pa']), err['plain'].mean(1)], 1)
            ev_used[b].record(s_main)
            ev_done[b].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                rows_host[b].copy_(rows, non_blocking=True)
                ev_rows[b].record(s_out)
            keep[b] = (so, rows)
        s_main.wait_stream(s_out)
        s_main.wait_stream(s_in)

    # image-to-metrics end to end (SURVEY 8f N4 + N1 around the hot path): the host sends the RGB crops and 2-D joints, the
    # proxy representation is built on the device (hf_proxy_rep), and only the per-image metric rows come back
    from humaniflow_b200.proxy_rep import build_proxy_representation
    rgb_host = torch.rand(B, 3, 256, 256, generator=torch.Generator().manual_seed(7 + rank)).pin_memory()
    j2d_host = (torch.rand(B, 17, 2, generator=torch.Generator().manual_seed(8 + rank)) * 256).pin_memory()
    rgb_stage = [torch.empty(B, 3, 256, 256, device=dev) for _ in range(2)]
    j2d_stage = [torch.empty(B, 17, 2, device=dev) for _ in range(2)]

    def rgb_steps(n):
        for b in range(2):
            ev_used[b].record(s_main)
            ev_rows[b].record(s_out)
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_used[b])
                rgb_stage[b].copy_(rgb_host, non_blocking=True)
                j2d_stage[b].copy_(j2d_host, non_blocking=True)
                ev_in[b].record(s_in)
            s_main.wait_event(ev_in[b])
            s_main.wait_event(ev_rows[b])
            so, metric = step(build_proxy_representation(rgb_stage[b], j2d_stage[b]), outs=out_dev[b])
            err = pointset_errors(so.vertices.view(B, N, V, 3), tgt)
            rows = torch.stack([samples_min(err['plain']), samples_min(err['sc']), samples_min(err['pa']), err['plain'].mean(1)], 1)
            ev_used[b].record(s_main)
            ev_done[b].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                rows_host[b].copy_(rows, non_blocking=True)
                ev_rows[b].record(s_out)
            keep[b] = (so, rows)
        s_main.wait_stream(s_out)
        s_main.wait_stream(s_in)

    # PCIe sanity numbers for the e2e line (plain pinned copies of the same buffers, not part of any timed region)
    def copy_gbs(dst, src, iters=3):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        c0, c1 = ev(), ev()
        c0.record()
        for _ in range(iters):
            dst.copy_(src, non_blocking=True)
        c1.record(); torch.cuda.synchronize()
        return src.numel() * src.element_size() * iters / (c0.elapsed_time(c1) * 1e-3) / 1e9
    v_dev_tmp = torch.empty(B * N, V, 3, device=dev)
    pcie = {'h2d_gbs': copy_gbs(x_stage[0], x_host), 'd2h_gbs': copy_gbs(v_host[0], v_dev_tmp)}
    del v_dev_tmp
    e2e_steps(5)
    sync_all()
    a, b_ev = ev(), ev()
    a.record()
    e2e_steps(args.steps)
    b_ev.record()
    sync_all()
    t = torch.tensor([a.elapsed_time(b_ev)], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    eval_steps(5)
    sync_all()
    a2, b2 = ev(), ev()
    a2.record()
    eval_steps(args.steps)
    b2.record()
    sync_all()
    t = torch.tensor([a2.elapsed_time(b2)], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eval_ms = t.item() / args.steps
    rgb_steps(5)
    sync_all()
    a3, b3 = ev(), ev()
    a3.record()
    rgb_steps(args.steps)
    b3.record()
    sync_all()
    t = torch.tensor([a3.elapsed_time(b3)], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rgb_ms = t.item() / args.steps
    clocks = sampler.stop() if rank == 0 else None

    line = None
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16 (encoder) / f32 (flow, LBS; f64 exp/log maps)',
            'data': 'synthetic (random-init weights; SMPL-shaped synthetic body model, skinning weights grouped by body part like the real SMPL)', 'config': workload_config(args),
            'e2e': {'value': world * B * N / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': x_host.numel() * 4, 'd2h_bytes_per_step': (v_host[0].numel() + j_host[0].numel()) * 4,
                    'pipelining': 'H2D | kernels | D2H on three streams, double-buffered', 'pcie_measured': pcie},
            'e2e_evaluate': {'value': world * B * N / (eval_ms * 1e-3), 'unit': UNIT, 'ms_per_step': eval_ms,
                             'h2d_bytes_per_step': x_host.numel() * 4, 'd2h_bytes_per_step': rows_host[0].numel() * 4,
                             'what': 'same step + per-sample PVE / PVE-SC / PVE-PA against one target mesh per image reduced on the '
                                     'device (hf_pointset_errors); only the (B,4) per-image rows are copied back'},
            'e2e_from_rgb': {'value': world * B * N / (rgb_ms * 1e-3), 'unit': UNIT, 'ms_per_step': rgb_ms,
                             'h2d_bytes_per_step': (rgb_host.numel() + j2d_host.numel()) * 4, 'd2h_bytes_per_step': rows_host[0].numel() * 4,
                             'what': 'RGB crops + 2-D joints from the host -> proxy representation on the device (hf_proxy_rep) -> the '
                                     'step -> on-device PVE / PVE-SC / PVE-PA rows back to the host'},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'stages': stages,
            'step_breakdown_ms': {'model_forward': ms_model, 'lbs': ms_lbs},
        }
        if world == 1 and not args.no_cpu_baseline:
            if _FULL_AFFINITY:
                os.sched_setaffinity(0, _FULL_AFFINITY)      # the CPU baseline may use every host core again
            cores = len(os.sched_getaffinity(0))
            torch.set_num_threads(cores)
            Bc = min(B, args.ref_images)
            val, sec = time_cpu(Bc, N, args.layers, steps=2, warmup=1)
            line['cpu_baseline'] = {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                                    'sample': '%d images x %d samples, 2 timed passes after 1 warm-up (%.1f s per pass); oracle/ = CPU '
                                              'restatement of the reference path' % (Bc, N, sec)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--B', type=int, default=32, help='images per GPU per step')
    ap.add_argument('--N', type=int, default=100, help='pose samples per image')
    ap.add_argument('--layers', type=int, default=50)
    ap.add_argument('--ref-images', type=int, default=8, help='images per step of the CPU arm (bounded sample)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
