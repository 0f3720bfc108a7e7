#!/usr/bin/env python
"""Benchmark of the reverse-diffusion sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete FullDPM.sample() (T=100 reverse steps) over one synthetic batch per GPU.
Metric: sampled CDR residues / s = (#GPUs x B x n_gen x K) / time, whole job.
  ours      : ab_opt_b200.FullDPM.sample -> libabopt_b200 (sm_100a kernels), inputs resident in HBM (`value`);
              `e2e` = the C-ABI host entry point abopt_design_host (atoms in pinned HOST buffers -> featurisation + loop ->
              101-frame trajectory in host memory), H2D + D2H inside the timed region;
              `gpu_eager_baseline` = the reference's eager-PyTorch evaluation (its port) on the same GPU, `cpu_baseline` on
              the host cores (N=1 only).
  reference : the reference's CPU path (oracle port, evaluation order of the reference: broadcast-multiply-sum
              with full temporaries) on the host cores, on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
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

CONFIGS = {
    # SURVEY.md section 8d.  c2 is the configuration BASELINE.json's metric is quoted on.
    'c2': dict(name='C2 AbDesign CDR-H3 co-design', B=64, L=256, gen=((120, 136),), flavour='abdesign',
               sample_structure=True, sample_sequence=True, obj='pred_noise'),
    'c3': dict(name='C3 AbDock pose diffusion', B=64, L=256, gen=((120, 136),), flavour='abdock',
               sample_structure=True, sample_sequence=False, obj='pred_x0'),
    'c4': dict(name='C4 AbDesign all-6-CDR co-design (per-GPU shard of B=256)', B=32, L=320,
               gen=((20, 30), (50, 60), (95, 105), (160, 170), (190, 200), (230, 240)), flavour='abdesign',
               sample_structure=True, sample_sequence=True, obj='pred_noise'),
}
# c5: SURVEY.md 8d config 5: one training iteration (forward + backward) on the CUDA path
TRAIN_CFG = dict(name='C5 AbDesign train.py FullDPM.forward + backward', B=128, L=256, gen=((120, 136),),
                 flavour='abdesign', obj='pred_noise')
NUM_LAYERS, T_STEPS = 6, 100


def algorithmic_bytes_per_complex_layer(L):
    """stream z once per layer + node state (x in/out, R, t, mask)."""
    return L * L * 64 * 4 + L * (2 * 128 * 4 + 36 + 12 + 1)


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'MEASURED_PEAKS.json'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def synthetic_batch(cfg, seed, device, B=None):
    """Seeded synthetic inputs of SURVEY.md 8d on `device`."""
    B = B or cfg['B']
    L = cfg['L']
    g = torch.Generator(device=device).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device=device)
    q = torch.nn.functional.normalize(rn(B, L, 4), dim=-1)            # uniform rotations -> so(3) vectors
    ang = 2 * torch.acos(q[..., :1].clamp(-1, 1))
    ang = torch.where(ang > torch.pi, ang - 2 * torch.pi, ang)
    v = torch.nn.functional.normalize(q[..., 1:], dim=-1) * ang
    mask_generate = torch.zeros(B, L, dtype=torch.bool, device=device)
    for a, b in cfg['gen']:
        mask_generate[:, a:b] = True
    return dict(v=v.contiguous(), p=rn(B, L, 3) * 10.0, s=torch.randint(0, 20, (B, L), generator=g, device=device),
                res_feat=rn(B, L, 128), pair_feat=rn(B, L, L, 64), mask_generate=mask_generate,
                mask_res=torch.ones(B, L, dtype=torch.bool, device=device))


def build_model(cfg, device):
    import ab_opt_b200
    torch.manual_seed(1234)
    if cfg['flavour'] == 'abdock':
        m = ab_opt_b200.FullDPM(128, 64, T_STEPS, eps_net_opt=dict(num_layers=NUM_LAYERS), obj=cfg['obj'], num_bins=40)
        for lin in (m.eps_net.prmsd_predictor.linear_1, m.eps_net.prmsd_predictor.linear_2, m.eps_net.prmsd_predictor.linear_3):
            torch.nn.init.normal_(lin.weight, std=0.05)
    else:
        m = ab_opt_b200.FullDPMAbDesign(128, 64, T_STEPS, eps_net_opt=dict(num_layers=NUM_LAYERS))
    return m.to(device).eval()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
            # wait for the first sample: nvidia-smi's start-up (NVML initialisation over all GPUs of the box) takes seconds and was
            # seen to stall kernel submission for 0.1-0.2 s when it fell into the timed region; the periodic queries do not
            t0 = time.time()
            while not self.lines and time.time() - t0 < 15.0 and self.proc.poll() is None:
                time.sleep(0.05)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx.append(int(f[1]))
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {'sm_mhz': statistics.median(busy) if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_run(cfg, steps, warmup, n_reverse_steps=2, B_cpu=1):
    """Times the oracle port of the reference's CPU path: `n_reverse_steps` reverse steps of a B_cpu-complex batch at
    the config's L / n_gen per bench step, extrapolated linearly in B and T (BASELINE.md section 2 shows linearity)."""
    from oracle import weights as ow, sampler as osamp, transitions as OT
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W = ow.make_state_dict(seed=1234, num_layers=NUM_LAYERS, flavour=cfg['flavour'])
    inp = synthetic_batch(cfg, 1234, 'cpu', B=B_cpu)
    n_gen = int(inp['mask_generate'][0].sum())
    gen = torch.Generator().manual_seed(0)
    L = cfg['L']

    def one():
        v, p, s = inp['v'], inp['p'] / 10.0, inp['s']
        for k in range(n_reverse_steps):
            nz = OT.draw_step_noise(B_cpu, L, gen)
            st = osamp.reverse_step(W, T_STEPS - k, v, p, s, inp['res_feat'], inp['pair_feat'], inp['mask_generate'],
                                    inp['mask_res'], nz, obj=cfg['obj'], materialize=True)
            v, p, s = st['v_next'], st['p_next'], st['s_next']
    with torch.no_grad():
        for _ in range(warmup):
            one()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dt = (time.perf_counter() - t0) / steps
    sec_per_reverse_step_per_complex = dt / (n_reverse_steps * B_cpu)
    value = n_gen / (sec_per_reverse_step_per_complex * T_STEPS)          # residues/s for a full T=100 sample
    sample = (f'{n_reverse_steps} reverse steps x B={B_cpu} complexes at L={L}, n_gen={n_gen}, {NUM_LAYERS} IPA layers, fp32, '
              f'{cores} threads; {dt:.2f} s per sample; scaled linearly to T={T_STEPS}')
    return value, dt, cores, sample, n_gen


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    value, dt, cores, sample, n_gen = cpu_reference_run(cfg, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': 'sampled CDR residues/sec', 'value': value, 'unit': 'residues/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(cfg), 'B': cfg['B'], 'L': cfg['L'], 'n_gen': n_gen, 'T': T_STEPS, 'layers': NUM_LAYERS},
        'cpu_baseline': {'value': value, 'unit': 'residues/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'residues/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(cfg):
    return (f"{cfg['name']}: FullDPM.sample, B={cfg['B']} complexes/GPU, L={cfg['L']}, n_gen={sum(b - a for a, b in cfg['gen'])}, "
            f"{NUM_LAYERS} IPA layers, T={T_STEPS}")


# ------------------------------------------------------------------------------------------ baseline legs on the GPU box
def gpu_eager_baseline(cfg, dev, n_reverse_steps=3):
    """The bar SURVEY.md finding 1 / BASELINE.md 4.2 name: the reference's eager-PyTorch evaluation ON THE SAME B200.  The
    reference tree does not travel to the GPU box, so this leg runs its line-by-line port (oracle/, materialize=True = the
    reference's broadcast-multiply-sum order with the full (B,L,L,H,.) temporaries, cuBLAS fp32 linears, ATen softmax) on cuda
    at the FULL batch for a few reverse steps and scales linearly to T.  A baseline leg, never the product path."""
    from oracle import weights as ow, sampler as osamp, transitions as OT
    W = {k: v.to(dev) for k, v in ow.make_state_dict(seed=1234, num_layers=NUM_LAYERS, flavour=cfg['flavour']).items()}
    inp = synthetic_batch(cfg, 1234, dev)
    B, L = cfg['B'], cfg['L']
    n_gen = int(inp['mask_generate'][0].sum())
    gen = torch.Generator(device=dev).manual_seed(0)

    def steps(k):
        v, p, s = inp['v'], inp['p'] / 10.0, inp['s']
        for i in range(k):
            nz = OT.draw_step_noise(B, L, gen, device=dev)
            st = osamp.reverse_step(W, T_STEPS - i, v, p, s, inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'],
                                    nz, obj=cfg['obj'], materialize=True)
            v, p, s = st['v_next'], st['p_next'], st['s_next']
        return v
    with torch.no_grad():
        steps(1)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps(n_reverse_steps)
        e1.record()
        torch.cuda.synchronize(dev)
    ms_step = e0.elapsed_time(e1) / n_reverse_steps
    peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    del W, inp
    torch.cuda.empty_cache()
    return {'value': B * n_gen / (ms_step * T_STEPS / 1e3), 'unit': 'residues/s', 'ms_per_sample': ms_step * T_STEPS,
            'ms_per_reverse_step': ms_step, 'kind': 'port of the reference evaluated eagerly with PyTorch on the same GPU',
            'sample': f'{n_reverse_steps} reverse steps at the full batch B={B}, L={L}, {NUM_LAYERS} layers, fp32 (allow_tf32 off as in the '
                      f'reference runners), scaled linearly to T={T_STEPS}; peak memory {peak_gb:.1f} GB'}


def synthetic_atoms(cfg, seed, B=None):
    """Seeded protein-like batch fields of the reference's data loader (host tensors): a noisy CA walk (~3.8 A steps), heavy atoms
    around it, two chains, consecutive numbering; generate_flag = the config's CDR segments."""
    B = B or cfg['B']
    L, A = cfg['L'], 15
    g = torch.Generator().manual_seed(seed)
    ca = torch.cumsum(torch.randn(B, L, 3, generator=g) * 2.2, dim=1)
    ca = ca - ca.mean(1, keepdim=True)
    pos = ca[:, :, None, :] + 1.6 * torch.randn(B, L, A, 3, generator=g)
    pos[:, :, 1] = ca
    mask_atoms = torch.rand(B, L, A, generator=g) < 0.8
    mask_atoms[:, :, :4] = True
    chain_nb = torch.zeros(B, L, dtype=torch.long)
    chain_nb[:, L // 2:] = 1
    res_nb = torch.arange(1, L + 1).expand(B, L).contiguous()
    gen = torch.zeros(B, L, dtype=torch.bool)
    for a, b in cfg['gen']:
        gen[:, a:b] = True
    ft = torch.ones(B, L, dtype=torch.long)
    ft[:, L // 2:] = 2
    return dict(aa=torch.randint(0, 20, (B, L), generator=g), res_nb=res_nb, chain_nb=chain_nb, pos_heavyatom=pos.contiguous(),
                mask_heavyatom=mask_atoms, fragment_type=ft, generate_flag=gen, mask=torch.ones(B, L, dtype=torch.bool))


def timed_config(cfg, dev, rank, world, steps, warmup, init_calls=2):
    """`value` leg of one configuration: W warm-up + K timed FullDPM.sample calls on this rank's complexes through
    ab_opt_b200.sharding.sample_shard (N > 1: the loop on the shard + the one packed gather).  -> (ms per step max over ranks,
    launches, last trajectory, model, inputs)."""
    import torch.distributed as dist
    import ab_opt_b200
    from ab_opt_b200 import sharding
    model = build_model(cfg, dev)
    inp = synthetic_batch(cfg, 1000 + rank, dev)          # each rank owns its own complexes (weak scaling, no exchange)
    B = cfg['B']
    kw = dict(sample_structure=cfg['sample_structure'], sample_sequence=cfg['sample_sequence'])
    a = (inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'])

    def step():
        if world > 1:
            return sharding.sample_shard(model, inp, rank * B, world * B, **kw)[1]
        return model.sample(*a, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    torch.manual_seed(rank)
    # initialisation, not warm-up: the first calls pay one-off costs (3 GB workspace cudaMalloc + memset, TMA descriptor
    # encodes, first-touch page faults of the host-side trajectory buffers) that measured 150-190 ms each on B200
    for _ in range(init_calls):
        step()
    # the warm-up keeps the previous trajectory alive while the next one is produced, exactly like the timed loop below: the
    # second set of pinned host blocks (52 MB, cudaHostAlloc) is then allocated here and not inside the second timed step
    # (measured: +30..230 ms on that step, every run)
    traj = None
    for _ in range(warmup):
        traj = step()
    barrier()
    l0 = ab_opt_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    e0.record()
    for k in range(steps):
        traj = step()
        marks[k].record()
    e1.record()
    barrier()
    launches = ab_opt_b200.launch_count() - l0
    ms = e0.elapsed_time(e1)
    timed_config.each_ms = [round(a_.elapsed_time(b_), 2) for a_, b_ in zip([e0] + marks[:-1], marks)]
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    assert torch.isfinite(traj[0][1]).all()
    return ms / steps, launches, traj, model, inp, a, kw, barrier


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, cfg, rank, world, local_rank):
    import torch.distributed as dist
    import ab_opt_b200
    from ab_opt_b200 import _capi
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, L = cfg['B'], cfg['L']
    clocks = ClockSampler(local_rank) if rank == 0 else None
    ms_per_step, launches, traj, model, inp, a, kw, barrier = timed_config(cfg, dev, rank, world, args.steps, args.warmup)
    clk = clocks.stop() if clocks else None
    n_gen = int(inp['mask_generate'][0].sum())
    value = world * B * n_gen / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with HOST buffers, the job the reference's runner does per batch (models/diffab.py:115-141):
    #      atoms in (pinned host memory) -> ResidueEmbedding + PairEmbedding + frames + 100 reverse steps on the device -> the whole
    #      101-frame trajectory back in host memory.  H2D and D2H are inside the timed region.
    nm = model.native()
    torch.manual_seed(4321)
    pe = ab_opt_b200.PairEmbedding(64, 15).to(dev).eval()
    re_ = ab_opt_b200.ResidueEmbedding(128, 15).to(dev).eval()
    atoms = synthetic_atoms(cfg, 2000 + rank)
    host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v) for k, v in atoms.items()}
    T0 = T_STEPS
    abdock = cfg['flavour'] == 'abdock'
    otv = torch.empty(T0 + 1, B, L, 3, pin_memory=True); otp = torch.empty(T0 + 1, B, L, 3, pin_memory=True)
    ots = torch.empty(T0 + 1, B, L, dtype=torch.int64, pin_memory=True)
    opr = torch.empty(T0 + 1, B, pin_memory=True) if abdock else None
    opl = torch.empty(T0 + 1, B, pin_memory=True) if abdock else None
    flags = (_capi.SAMPLE_STRUCTURE if cfg['sample_structure'] else 0) | (_capi.SAMPLE_SEQUENCE if cfg['sample_sequence'] else 0) | \
        _capi.KEEP_TRAJECTORY
    peh, reh = pe.native().handle, re_.native().handle
    _capi.check(_capi.lib().abopt_model_set_batch_offset(nm.handle, rank * B))

    def e2e_step(seed):
        _capi.check(_capi.lib().abopt_design_host(
            nm.handle, peh, reh, B, L, 15, _capi.ptr(host['aa']), _capi.ptr(host['res_nb']), _capi.ptr(host['chain_nb']),
            _capi.ptr(host['pos_heavyatom']), _capi.ptr(host['mask_heavyatom']), _capi.ptr(host['fragment_type']),
            _capi.ptr(host['generate_flag']), _capi.ptr(host['mask']), flags, 0, seed, _capi.ptr(otv), _capi.ptr(otp), _capi.ptr(ots),
            _capi.ptr(opr), _capi.ptr(opl)))
    e2e_step(1)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(2 + i)
    dt = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    assert torch.isfinite(otp[0]).all() and torch.isfinite(otp[T0 // 2]).all()
    e2e_value = world * B * n_gen / float(dt.item())
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = otv.numel() * 4 + otp.numel() * 4 + ots.numel() * 8 + (2 * (T0 + 1) * B * 4 if abdock else 0)

    # ---- per-kernel timing pass (CUDA events around every launch, outside the timed region) -> roofline
    roof, breakdown, feat_line = None, None, None
    if rank == 0:
        _capi.profile_enable(True)
        torch.manual_seed(0)
        model.sample(*a, **kw)
        torch.cuda.synchronize()
        prof = _capi.profile_collect()
        _capi.profile_enable(False)
        full_ms, full_n = prof['pair']                    # full-stream launches of pair_stream_kernel
        part_ms, part_n = prof['pair_part']               # launches that visit the generated query rows only
        pair_ms, pair_n = full_ms + part_ms, full_n + part_n
        total_ms = sum(v[0] for v in prof.values())
        peak, peak_src = measured_peak_gbs()
        # pair_stream_kernel: one launch per layer and step.  Inside the loop up to two of the six layers stream z for the GENERATED
        # query rows only: the last block of a model without the pRMSD head (focus mode) and the first block
        # (context cache: the context queries take the context part of their aggregate from a per-run cache and read only the
        # generated keys' z rows, in ctx_delta_kernel, timed as its own kind).  `frac` is bytes-weighted over ALL launches of the
        # kernel (total algorithmic bytes / total time); `full_stream` is the same for the full-stream launches alone;
        # whole_step_* keep SURVEY 8d's fixed denominator (z once per layer).
        alg_full = B * algorithmic_bytes_per_complex_layer(L)
        alg_part = B * (n_gen * L * 64 * 4 + L * (2 * 128 * 4 + 36 + 12 + 1))
        alg_bytes = (full_n * alg_full + part_n * alg_part) / max(pair_n, 1)      # mean per launch
        achieved = alg_bytes / (pair_ms / max(pair_n, 1) * 1e-3) / 1e9
        full_achieved = alg_full / (full_ms / max(full_n, 1) * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'pair_kernel_traffic.json')
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get('L') == L:
                traffic = tj['dram_bytes_per_complex'] * B * alg_bytes / alg_full
        whole = B * NUM_LAYERS * T_STEPS * algorithmic_bytes_per_complex_layer(L) / (ms_per_step * 1e-3) / 1e9
        roof = {'bound': 'hbm', 'kernel': 'pair_stream_kernel (streams pair_feat once per IPA layer: softmax-weighted pair aggregation)',
                'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': alg_bytes, 'avg_launch_ms': pair_ms / max(pair_n, 1), 'launches_per_sample': pair_n,
                'full_stream': {'launches': full_n, 'avg_launch_ms': full_ms / max(full_n, 1), 'algorithmic_bytes_per_launch': alg_full,
                                'achieved': full_achieved, 'frac': full_achieved / peak},
                'generated_rows_only': {'launches': part_n, 'avg_launch_ms': part_ms / max(part_n, 1),
                                        'algorithmic_bytes_per_launch': alg_part,
                                        'note': 'last block (focus mode) and first block (context cache; ctx_delta_kernel beside it: '
                                                '%d launches, %.1f us each)' % (prof['ctx'][1], 1e3 * prof['ctx'][0] / max(prof['ctx'][1], 1))}
                if part_n else None,
                'share_of_gpu_time': pair_ms / total_ms,
                'whole_step_achieved': whole, 'whole_step_frac': whole / peak}
        breakdown = {k: {'ms': round(v[0], 3), 'launches': v[1]} for k, v in prof.items() if v[1]}
        # the step before the loop, part of the e2e job: PairEmbedding / ResidueEmbedding at this shape
        dat = {k: v.to(dev) for k, v in atoms.items()}
        ctx = dat['mask_heavyatom'][:, :, 1] & ~dat['generate_flag']
        pa = (dat['aa'], dat['res_nb'], dat['chain_nb'], dat['pos_heavyatom'], dat['mask_heavyatom'])
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        tp_, tr_ = 0.0, 0.0
        z = x = None
        for i in range(5):
            z = x = None                   # the 1 GB output block goes back to the caching allocator BEFORE the next call asks for one:
            flush.zero_()                  # a cudaMalloc between the two event records would be timed as GPU work
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); z = pe(*pa, ctx, ctx); e[1].record(); x = re_(*pa, dat['fragment_type'], ctx, ctx); e[2].record()
            torch.cuda.synchronize()
            if i >= 2:
                tp_ += e[0].elapsed_time(e[1]) / 3; tr_ += e[1].elapsed_time(e[2]) / 3
        feat_line = {'pair_embed_ms': tp_, 'res_embed_ms': tr_, 'pairs_per_s': B * L * L / (tp_ / 1e3),
                     'note': 'PairEmbedding / ResidueEmbedding.forward at this shape, 15 atoms, L2 flushed between iterations'}
        del dat, z, x, flush

    if rank == 0:
        cpu, eager = None, None
        if world == 1 and not args.no_cpu_baseline:
            cv, cdt, cores, sample, _ = cpu_reference_run(cfg, steps=2, warmup=1, n_reverse_steps=4, B_cpu=4)
            cpu = {'value': cv, 'unit': 'residues/s', 'cores': cores, 'kind': 'port', 'sample': sample}
        if world == 1 and not args.no_gpu_eager:
            del inp, a, traj, nm
            model.invalidate_native()
            torch.cuda.empty_cache()
            try:
                eager = gpu_eager_baseline(cfg, dev)
            except Exception as ex:         # e.g. out of memory for the eager temporaries: report, do not fail the bench line
                eager = {'unavailable': str(ex)[:200]}
        line = {
            'metric': 'sampled CDR residues/sec', 'value': value, 'unit': 'residues/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(cfg), 'B_per_gpu': B, 'L': L, 'n_gen': n_gen, 'T': T_STEPS, 'layers': NUM_LAYERS,
                       'flavour': cfg['flavour'], 'rng': 'philox (in-kernel)', 'weights': 'seeded random init',
                       'l2': 'inputs larger than L2 (pair_feat %.2f GB per GPU)' % (B * L * L * 64 * 4 / 1e9),
                       'value_includes': 'init noise, 100 reverse steps, 101-frame trajectory D2H, final packed gather (N>1)',
                       'init_calls_before_warmup': 2,
                       'e2e_job': 'abopt_design_host: atoms (pinned host) -> residue + pair featurisation + frames + 100 reverse steps '
                                  '-> whole 101-frame trajectory in host memory (models/diffab.py:115-141)'},
            'clocks': clk, 'ms_each_step': getattr(timed_config, 'each_ms', None),
            'e2e': {'value': e2e_value, 'unit': 'residues/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': float(dt.item()) * 1e3},
            'gpu_launches': int(launches),
            'roofline': roof, 'cpu_baseline': cpu, 'gpu_eager_baseline': eager, 'featurisation': feat_line,
            'kernel_breakdown_ms_per_sample': breakdown,
        }
    if world == 8 and args.config == 'c2' and not args.no_c4:
        # BASELINE config 4 as configured: B=256, L=320 sharded 32 per GPU over the 8 GPUs, six generated segments, one gather
        c4 = CONFIGS['c4']
        del model
        torch.cuda.empty_cache()
        ms4, _, _, _, inp4, _, _, _ = timed_config(c4, dev, rank, world, steps=2, warmup=1, init_calls=1)
        if rank == 0:
            ng4 = int(inp4['mask_generate'][0].sum())
            peak, _ = measured_peak_gbs()
            whole4 = c4['B'] * NUM_LAYERS * T_STEPS * algorithmic_bytes_per_complex_layer(c4['L']) / (ms4 * 1e-3) / 1e9
            line['c4'] = {'workload': workload_name(c4) + f', {world} GPUs (B={world * c4["B"]} in total)', 'ms_per_step': ms4,
                          'value': world * c4['B'] * ng4 / (ms4 / 1e3), 'unit': 'residues/s', 'steps': 2, 'warmup': 1,
                          'whole_step_achieved_per_gpu': whole4, 'whole_step_frac': whole4 / peak}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train_forward(args):
    """--config c5: BASELINE config 5, `AbDesign train.py forward+backward EpsilonNet step, batch=128, N=256`: one training
    iteration up to the optimiser -- FullDPM.forward and loss.backward() (abopt_loss_backward: forward, hand-written backward,
    gradients of all parameters and of res_feat / pair_feat) -- and the forward alone (validate()).  residues/s = B * L / time.
    One GPU, not a bench line of the headline metric."""
    cfg = TRAIN_CFG
    dev = torch.device('cuda', 0)
    model = build_model(cfg, dev)
    inp = synthetic_batch(cfg, 1000, dev)
    B, L = cfg['B'], cfg['L']
    t = torch.randint(1, T_STEPS, (B,), device=dev)
    a = (inp['v'], inp['p'], inp['s'], inp['res_feat'], inp['pair_feat'], inp['mask_generate'], inp['mask_res'], True, True)
    steps = max(args.steps, 5)

    def timed(fn):
        for _ in range(2 + args.warmup):
            out = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, out
    clocks = ClockSampler(0)
    with torch.no_grad():
        ms_fwd, loss = timed(lambda: model(*a, t=t))
    ms_step, (loss_b, grads, d_res, d_pair) = timed(lambda: model.loss_and_grads(*a, t=t))
    ck = clocks.stop()
    assert all(torch.isfinite(v) for v in loss_b.values()) and torch.isfinite(d_pair).all()
    assert all(torch.isfinite(g).all() for g in grads.values())
    from ab_opt_b200 import _capi
    _capi.profile_enable(True)
    model.loss_and_grads(*a, t=t)
    torch.cuda.synchronize()
    prof = _capi.profile_collect()
    _capi.profile_enable(False)
    peak, peak_src = measured_peak_gbs()
    alg_f = B * NUM_LAYERS * algorithmic_bytes_per_complex_layer(L)
    # backward: z is read once more per layer by the recompute and once by the pair backward, d z is written (first layer) or
    # read + written (the others): (2 + 2 - 1/6) z-passes on top of the forward's one
    zl = B * L * L * 64 * 4
    alg_b = alg_f + NUM_LAYERS * 3 * zl + (NUM_LAYERS - 1) * zl
    print(json.dumps({'metric': 'training step residues/sec (FullDPM.forward + backward)', 'value': B * L / (ms_step / 1e3),
                      'unit': 'residues/s', 'n_gpus': 1, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
                      'higher_is_better': True, 'dtype': 'f32', 'data': 'synthetic',
                      'config': {'workload': f"{cfg['name']}: B={B}, L={L}, n_gen=16, {NUM_LAYERS} IPA layers; forward + backward, "
                                             'gradients of all 63 parameters + d res_feat + d pair_feat', 'l2': 'inputs larger than L2'},
                      'forward_only_ms': ms_fwd, 'losses': {k: float(v) for k, v in loss_b.items()}, 'clocks': ck,
                      'grad_norm': float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))),
                      'kernel_breakdown_ms': {k: {'ms': round(v[0], 3), 'launches': v[1]} for k, v in prof.items() if v[1]},
                      'roofline': {'bound': 'hbm', 'achieved': alg_b / (ms_step * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                                   'frac': alg_b / (ms_step * 1e-3) / 1e9 / peak, 'peak_source': peak_src,
                                   'algorithmic_bytes': alg_b, 'forward_frac': alg_f / (ms_fwd * 1e-3) / 1e9 / peak,
                                   'note': 'whole step against streaming z: forward once per layer; backward: recompute (1), pair '
                                           'backward read (1), d z write (1) + read-modify (5/6) per layer'}}), flush=True)


def run_pair_embed(args):
    """--config f1: PairEmbedding.forward (the O(L^2) featurisation that produces pair_feat before the
    loop) at the C2 shapes: B=64, L=256, 15 atoms per residue.  pairs/s = B * L^2 / time.  One GPU; not the headline metric."""
    import ab_opt_b200
    from oracle import pair_embed as PE
    B, L, A = 64, 256, 15
    dev = torch.device('cuda', 0)
    W = PE.make_state_dict(3, A)
    mod = ab_opt_b200.PairEmbedding(64, A)
    mod.load_state_dict(W, strict=True)
    mod = mod.to(dev).eval()
    host = PE.synthetic_complex(77, B, L)
    inp = {k: v.to(dev) for k, v in host.items()}
    a = (inp['aa'], inp['res_nb'], inp['chain_nb'], inp['pos_atoms'], inp['mask_atoms'], inp['context_mask'], inp['context_mask'])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(args.warmup, 3)):
        z = mod(*a)
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    steps = max(args.steps, 5)
    tot = 0.0
    for _ in range(steps):
        flush.zero_()                                      # L2 flush between timed iterations (256 MiB > 126 MB L2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        z = mod(*a)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / steps
    ck = clocks.stop()
    assert torch.isfinite(z).all()
    # CPU: the oracle (= the reference's evaluation order, bit-equal to it) on one complex, all host threads
    cpu = None
    if not args.no_cpu_baseline:
        Wc = W
        one = {k: v[:1] for k, v in host.items()}
        ac = (one['aa'], one['res_nb'], one['chain_nb'], one['pos_atoms'], one['mask_atoms'], one['context_mask'], one['context_mask'])
        PE.pair_embedding(Wc, *ac)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            PE.pair_embedding(Wc, *ac)
        dt = (time.perf_counter() - t0) / reps
        cpu = {'value': L * L / dt, 'unit': 'pairs/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': f'one complex (L={L}, A={A}) x {reps} calls, {dt:.2f} s per call'}
    peak, peak_src = measured_peak_gbs()
    pairs = B * L * L
    alg = pairs * 64 * 4 + B * L * (A * 13 + 3 * 8 + 2)                         # pair_feat written once + the per-residue inputs
    flops = pairs * 2 * 64 * (A * A + 64 + 64 + 26 + 64 + 64)                  # the five dense layers as executed (tables pre-multiplied)
    print(json.dumps({'metric': 'pair featurisation pairs/sec (PairEmbedding.forward)', 'value': pairs / (ms / 1e3), 'unit': 'pairs/s',
                      'n_gpus': 1, 'steps': steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True,
                      'dtype': 'f32', 'data': 'synthetic',
                      'config': {'workload': f'F1 PairEmbedding.forward: B={B}, L={L}, {A} atoms per residue, structure + sequence masks',
                                 'l2': 'flushed between timed iterations (256 MiB memset)'},
                      'clocks': ck, 'gpu_launches': steps,
                      'roofline': {'bound': 'hbm', 'achieved': alg / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                                   'frac': alg / (ms * 1e-3) / 1e9 / peak, 'peak_source': peak_src, 'traffic': None,
                                   'traffic_note': 'ncu: 1.02 GB written, 5 MB read per launch (profiles/r02_final_ncu_pair_embed_summary.json)',
                                   'note': 'not HBM bound: 65 kflop per pair (x3 as 3xTF32 on tcgen05, activations in tensor memory) and one exp per atom pair against 256 B written',
                                   'fp32_equivalent_tflops': flops / (ms * 1e-3) / 1e12},
                      'cpu_baseline': cpu}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS) + ['c5', 'f1'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-eager', action='store_true', help='skip the eager-PyTorch-on-GPU baseline leg')
    ap.add_argument('--no-c4', action='store_true', help='at 8 GPUs: skip the additional C4 record')
    args = ap.parse_args()
    if args.config == 'c5':
        if not torch.cuda.is_available():
            raise SystemExit('bench.py --config c5 needs a CUDA device (there is no CPU fallback)')
        run_train_forward(args)
        return
    if args.config == 'f1':
        if not torch.cuda.is_available():
            raise SystemExit('bench.py --config f1 needs a CUDA device (there is no CPU fallback)')
        run_pair_embed(args)
        return
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, cfg, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (there is no CPU fallback)')
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}')
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == '__main__':
    main()
