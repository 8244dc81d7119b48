#!/usr/bin/env python
"""bench.py -- MLP samples/sec on the synthetic VolSDF 480x270x128 render (BASELINE.json configs[1]).

One "step" = one full-frame render of 129 600 rays through the hot path (hierarchical error-bound sampler with 512
SDF-only network evaluations per ray, 192 full evaluations (sdf + d sdf/dx + radiance) per ray, sdf->sigma, front-to-back
compositing).  metric value = n_rays * 192 / t  (SURVEY.md 8d "full-MLP samples/s").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|tc]
Multi-GPU: launched by torchrun, one rank per GPU; the rays of ONE frame are block-partitioned over the ranks (strong
scaling) and the rendered RGB tiles are all-gathered over NCCL at the end of every step (north star).
`--impl reference` times the UNMODIFIED reference's own `volume_render` (PyTorch, CPU, all host threads) on a bounded strided ray
sample of the same frame; the reference tree travels to the GPU box as oracle/_ref/reference (oracle/build_ref.sh).  Where no
staged copy exists the numpy oracle port (oracle/nerfart_oracle.py) is timed instead and the line says kind "port".
The main line additionally carries `reference_gpu`: the same unmodified reference code on the same B200 (plain PyTorch CUDA),
on every 8th ray of the frame -- the "before" number BASELINE.md asks for -- with the rgb L-inf between the two renders.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

H, W = 480, 270
N_SAMPLES, N_IMPORTANCE = 128, 64
P = N_SAMPLES + N_IMPORTANCE
# matmul MACs x 2.  SDF-only evaluations skip the unused 256-wide feature head (918 016 instead of 1 049 088).
F_SDF = 2 * (39 * 256 + 256 * 256 * 2 + 256 * 217 + 256 * 256 * 4 + 256)
F_FULL = 2 * (524544 + 459008 + 265216)


def traffic_capture():
    """DRAM bytes per sample (SDF-only, full) of the MLP kernel from the dated `ncu --set full` capture committed under profiles/
    (dram__bytes_read.sum + dram__bytes_write.sum per launch / samples per launch); None when no capture file is present."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'traffic_capture.json')))
    except Exception:
        return None

DTYPES = {'fp32': 'f32',
          'tc': 'f16x2-split operands (22-bit), f32 accumulate',
          'tc2acc': 'f16x2-split operands (22-bit), f32 accumulate (two accumulators)',
          'tc_mixed': 'SDF forward (sampler + final): f16x2-split operands (22-bit); feature head, reverse sweep, radiance net: f16 operands (11-bit); f32 accumulate'}
RENDER_KW = dict(batched=True, near=0.0, far=6.0, obj_bounding_radius=3.0, perturb=False, white_bkgd=False,
                 max_upsample_steps=6, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, epsilon=0.1, max_bisection_steps=10,
                 require_nablas=True, calc_normal=True, detailed_output=False)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md clocks line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def make_inputs():
    import fixtures as fx
    c2w, K = fx.closed_form_camera(H, W)
    return c2w, K


REF_KW = dict(RENDER_KW, rayschunk=2048)                 # render.py:488,614: rayschunk default 2048


def reference_baseline(n_rays_sample, device='cpu', stride_sel=None):
    """The UNMODIFIED reference's volume_render (oracle/ref_runner.py) on a strided ray sample of the same frame, on `device`.
    Returns (samples/s, seconds, rgb [n,3] numpy, selected ray indices) or None when no reference tree is staged."""
    import ref_runner
    import fixtures as fx
    from helpers import make_volsdf
    if ref_runner.reference_root() is None:
        return None
    torch.set_num_threads(os.cpu_count())
    model = ref_runner.build_model('volsdf', make_volsdf(0.1, 0.0).state_dict(), fx.volsdf_kwargs(0.1), device=device)
    c2w, K = make_inputs()
    rr = ref_runner.load()['rend_util']
    ro, rd, _ = rr.get_rays(c2w[None].to(device), K[None].to(device), H, W, -1)
    sel = np.linspace(0, H * W - 1, n_rays_sample).astype(np.int64) if stride_sel is None else stride_sel
    st = torch.as_tensor(sel, device=device)
    if device != 'cpu':                                   # warm-up on a small slice (cuBLAS / allocator start-up is not the reference's cost)
        ref_runner.volume_render('volsdf', model, ro[:, st[:256]], rd[:, st[:256]], **REF_KW)
    rgb, _, _, dt = ref_runner.volume_render('volsdf', model, ro[:, st], rd[:, st], **REF_KW)
    return len(sel) * P / dt, dt, rgb[0].detach().cpu().numpy(), sel


def cpu_baseline(n_rays_sample, repeats=1):
    """The oracle port on the host cores, on a strided ray sample of the same frame.  Returns (samples/s, seconds)."""
    import nerfart_oracle as orc
    from helpers import make_volsdf, oracle_net
    torch.set_num_threads(os.cpu_count())
    net = oracle_net(make_volsdf(0.1, 0.0), 'volsdf')
    c2w, K = make_inputs()
    ro, rd = orc.get_rays(c2w.numpy(), K.numpy(), H, W)
    sel = np.linspace(0, H * W - 1, n_rays_sample).astype(np.int64)
    best = None
    for _ in range(repeats):
        t0 = time.time()
        orc.volsdf_render(net, ro[sel], rd[sel], N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, rayschunk=2048)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    return n_rays_sample * P / best, best


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_sample = 256
    times = []
    import ref_runner
    kind = 'reference' if ref_runner.reference_root() is not None else 'port'
    for i in range(args.warmup + args.steps):
        dt = reference_baseline(n_sample)[1] if kind == 'reference' else cpu_baseline(n_sample)[1]
        if i >= args.warmup:
            times.append(dt)
    t = float(np.sum(times))
    value = n_sample * P * args.steps / t
    what = ("the unmodified reference's volume_render (PyTorch fp32, CPU, rayschunk 2048)" if kind == 'reference'
            else 'numpy+BLAS fp32 oracle port of the reference (no staged reference tree)')
    line = {'impl': 'reference', 'metric': f'MLP samples/sec (VolSDF {H}x{W}x{N_SAMPLES})', 'value': value, 'unit': 'samples/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, extra={'sample': f'{n_sample} of {H*W} rays (strided), same 128+64 samples/ray'}),
            'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': kind,
                             'sample': f'{n_sample} strided rays of the 480x270 frame per step, {what}'},
            'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def train_probe(dev, precision, n_patches=8):
    """Short probe of the fine-tune step's pass 2 (BASELINE configs 3 / 5): `n_patches` patches of 1200 rays, forward render with
    detailed outputs + backward kernels, for the VolSDF frame of this bench and for a NeuS model (configs/neus_fangzhou_vangogh.yaml
    geometry, 64 + 64 samples).  Event-timed after one warm-up patch; the full step is `bench.py --workload train`."""
    import nerfart_b200  # noqa: F401
    from helpers import make_volsdf, make_neus
    import fixtures as fx
    from nerfart_b200.models.frameworks import volsdf as pv, neus as pn
    from nerfart_b200.utils import rend_util
    F_BWD = 2 * (2 * 265216 + 2 * 524544 + 2 * 459008)
    out = {}
    pk, _ = peaks()
    split = precision in ('tc', 'tc_mixed') and os.environ.get('NA_BW_SPLIT', '1') != '0'      # split training program, as Trainer.forward runs it
    for fw in ('volsdf', 'neus'):
        if fw == 'volsdf':
            m = make_volsdf(0.1, 0.0, device=dev).train()
            c2w, K = fx.closed_form_camera(H, W)
            kw = dict(near=0.0, far=6.0, perturb=True, max_upsample_steps=6, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE,
                      train_stash=split)
            patch = lambda ro, rd: pv.render_patch(m, ro, rd, **kw)
            pts = P
        else:
            m = make_neus(0.05, 0.0, device=dev).train()
            c2w, K = fx.closed_form_camera(H, W)
            c2w = c2w.clone(); c2w[2, 3] = -0.9                                  # inside NeuS' unit bounding sphere
            kw = dict(upsample_algo='official_solution', N_upsample_iters=4, N_outside=0, obj_bounding_radius=1.0, perturb=True,
                      N_samples=64, N_importance=64, train_stash=split)
            patch = lambda ro, rd: pn.render_patch(m, ro, rd, **kw)
            pts = 128
        m.engine().precision = precision
        with torch.no_grad():
            ro, rd, _ = rend_util.get_rays(c2w[None].to(dev), K[None].to(dev), H, W)
        eng = m.engine(); eng.grad_zero()
        G = torch.full((1200, 3), 1e-3, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_f = t_b = 0.0
        for k in range(n_patches + 1):
            i = 1200 * (40 + k)
            rop, rdp = ro[0, i:i + 1200].contiguous(), rd[0, i:i + 1200].contiguous()
            ev[0].record()
            fwd, scal = patch(rop, rdp)
            ev[1].record()
            eng.render_bwd(rop, rdp, scal, fwd, G, w_eikonal=0.1, eikonal_count=1200 * pts, white_bkgd=False, speed_factor=m.speed_factor,
                           train_radiance=(fw == 'volsdf'))
            ev[2].record(); torch.cuda.synchronize()
            if k > 0:
                t_f += ev[0].elapsed_time(ev[1]); t_b += ev[1].elapsed_time(ev[2])
        eng.unpack_grads(True, fw == 'volsdf')
        torch.cuda.synchronize()
        out[fw] = {'patches': n_patches, 'rays_per_patch': 1200, 'points_per_ray': pts, 'patch_forward_ms': t_f / n_patches,
                   'patch_backward_ms': t_b / n_patches, 'samples_per_s': 1200 * pts * n_patches / ((t_f + t_b) * 1e-3)}
        if fw == 'volsdf':
            ach = 1200 * pts * F_BWD / (t_b / n_patches * 1e-3) / 1e12
            out[fw]['backward_roofline'] = {'bound': 'tensor', 'achieved': ach, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': ach / pk['bf16_tflops'],
                                            'flop_per_sample': F_BWD}
        del m, eng
        torch.cuda.empty_cache()
    return out


def small_beta_probe(dev, precision, step=8):
    """Informational line beside the headline: the same frame with beta = 0.01, where the error-bound sampler runs
    its upsampling iterations on most rays (trained checkpoints live there; BASELINE.md section 3 quotes ~1385 evaluations per ray).
    Reports ms / frame, SDF evaluations per ray, and -- on every `step`-th ray -- the share of rays whose rgb differs by more than 1e-3
    from the UNMODIFIED reference rendered on this GPU, for the default mode and for the fp32 CUDA-core mode (the sampler is
    discontinuous: two fp32-correct evaluations already disagree on ~12 % of the rays of the 400-ray beta = 0.01 fixture)."""
    import nerfart_b200  # noqa: F401
    import ref_runner
    import fixtures as fx
    from helpers import make_volsdf
    from nerfart_b200.models.frameworks.volsdf import volume_render
    from nerfart_b200.utils import rend_util
    beta = 0.01
    m = make_volsdf(beta, 0.0, device=dev)
    m.engine().precision = precision
    c2w, K = make_inputs()
    out = {'beta': beta}
    with torch.no_grad():
        ro, rd, _ = rend_util.get_rays(c2w[None].to(dev), K[None].to(dev), H, W)
        volume_render(ro, rd, m, **RENDER_KW)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            volume_render(ro, rd, m, **RENDER_KW)
        e1.record(); torch.cuda.synchronize()
        out['ms_per_frame'] = e0.elapsed_time(e1) / 2
        sel = torch.arange(0, H * W, step, device=dev)
        kw = dict(RENDER_KW, detailed_output=True)
        rgb, _, ex = volume_render(ro[:, sel], rd[:, sel], m, **kw)
        usage = ex['iter_usage'].reshape(-1)
        iters = torch.where(usage < 0, torch.full_like(usage, float(RENDER_KW.get('max_upsample_steps', 6))), usage)
        out['sdf_evals_per_ray'] = float((4 * N_SAMPLES * (1 + iters)).mean()) + P
        out['rays_not_converged_frac'] = float((usage < 0).float().mean())
        out['samples_per_s'] = H * W * P / (out['ms_per_frame'] * 1e-3)
        if ref_runner.reference_root() is not None:
            rm = ref_runner.build_model('volsdf', make_volsdf(beta, 0.0).state_dict(), fx.volsdf_kwargs(beta), device=str(dev))
            rr = ref_runner.load()['rend_util']
            rro, rrd, _ = rr.get_rays(c2w[None].to(dev), K[None].to(dev), H, W, -1)
            ref_rgb, _, _, dt = ref_runner.volume_render('volsdf', rm, rro[:, sel], rrd[:, sel], **REF_KW)
            out['reference_gpu_s'] = dt
            d_def = (rgb[0] - ref_rgb[0]).abs().amax(-1)
            m.engine().precision = 'fp32'
            rgb32, _, _ = volume_render(ro[:, sel], rd[:, sel], m, **RENDER_KW)
            d_32 = (rgb32[0] - ref_rgb[0]).abs().amax(-1)
            out['vs_reference_gpu'] = {'rays': int(sel.numel()),
                                       precision: {'rgb_diff_gt_1e-3_frac': float((d_def > 1e-3).float().mean()), 'rgb_diff_median': float(d_def.median())},
                                       'fp32': {'rgb_diff_gt_1e-3_frac': float((d_32 > 1e-3).float().mean()), 'rgb_diff_median': float(d_32.median())}}
    del m
    torch.cuda.empty_cache()
    return out


def workload_config(args, extra=None):
    c = {'workload': f'VolSDF fangzhou_nature-shaped synthetic render {H}x{W} ({H*W} rays), N_samples={N_SAMPLES}, N_importance={N_IMPORTANCE}, '
                     f'd_init={4 * N_SAMPLES}, seed-0 sphere init beta=0.1, radiance gains x3, closed-form camera',
         'evals_per_ray': {'sdf_only': 4 * N_SAMPLES, 'full': P},
         'parallelism': f'ray-partition x{args.gpus}' + (' + NCCL all-gather of RGB tiles' if args.gpus > 1 else ''),
         'precision_mode': args.precision,
         'l2': 'per-step working set (3.7 GB per-ray depth/sdf arrays + 0.9 GB per-sample outputs) >> 126 MB L2; no explicit flush'}
    if extra:
        c.update(extra)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('NA_PRECISION', 'auto'))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='render', choices=['render', 'train'],
                    help="'render' (default): BASELINE configs[1]; 'train': the fine-tune step, see bench_train.py")
    ap.add_argument('--style', default='clip', choices=['clip', 'mse'], help='--workload train: style loss (see bench_train.py)')
    ap.add_argument('--framework', default='volsdf', choices=['volsdf', 'neus'],
                    help="--workload train: 'neus' = BASELINE configs[2] (NeuS fine-tune step: radiance net frozen, 64+64 samples)")
    ap.add_argument('--config', type=int, default=2, choices=[2, 4],
                    help='BASELINE.json configs[] entry: 2 = VolSDF 480x270, 128 samples (the metric, default); 4 = 960x540, 256 samples (8-GPU config)')
    args = ap.parse_args()
    if args.config == 4:
        global H, W, N_SAMPLES, P
        H, W, N_SAMPLES = 960, 540, 256
        P = N_SAMPLES + N_IMPORTANCE
        RENDER_KW.update(N_samples=N_SAMPLES); REF_KW.update(N_samples=N_SAMPLES)
    if args.workload == 'train':
        import bench_train
        return bench_train.main(args)
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    assert args.warmup >= 3, 'timing rules: at least 3 warm-up steps'
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (nerfart_b200 has no CPU path); use --impl reference for the CPU arm')
    import nerfart_b200
    from helpers import make_volsdf
    from nerfart_b200.models.frameworks.volsdf import volume_render
    from nerfart_b200.utils import rend_util
    import torch.distributed as dist
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)
        dist.barrier()
    if args.precision == 'auto':
        args.precision = nerfart_b200.default_precision() if hasattr(nerfart_b200, 'default_precision') else 'fp32'
    model = make_volsdf(0.1, 0.0, device=dev)
    model.engine().precision = args.precision
    c2w_h, K_h = make_inputs()
    c2w_pin, K_pin = c2w_h[None].pin_memory(), K_h[None].pin_memory()
    n_rays = H * W
    from nerfart_b200 import parallel
    lo, hi, per = parallel.ray_block(n_rays, rank, world)
    emu = int(os.environ.get('NA_BENCH_EMULATE_WORLD', '0'))     # diagnostics: time rank 0's share of an `emu`-GPU job on ONE GPU (what of a
    if emu > 1 and world == 1:                                   # frame does not shrink with N, apart from chip-to-chip clock spread)
        lo, hi, per = parallel.ray_block(n_rays, 0, emu)
    rgb_host = torch.empty(n_rays, 3, dtype=torch.float32).pin_memory()

    def step(e2e):
        with torch.no_grad():
            if e2e:
                c2w, K = c2w_pin.to(dev, non_blocking=True), K_pin.to(dev, non_blocking=True)
                ro, rd, _ = rend_util.get_rays(c2w, K, H, W)
                step.rays = (ro, rd)
            ro, rd = step.rays
            rgb, depth, ex = volume_render(ro[:, lo:hi], rd[:, lo:hi], model, **RENDER_KW)
            img = parallel.gather_tiles(rgb[0], n_rays)        # NCCL all-gather of the RGB tiles (no-op at world 1)
            if e2e:
                rgb_host[:img.shape[0]].copy_(img, non_blocking=True)
        return img

    def timed(e2e, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            a.record(); step(e2e); b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e-3

    step(True)                                     # builds rays once, sizes the workspace
    if saved_stdout is not None:
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    for _ in range(args.warmup):
        step(False)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = nerfart_b200.launch_count()
    t_dev = timed(False, args.steps)
    launches = nerfart_b200.launch_count() - l0
    t_e2e = timed(True, args.steps) if os.environ.get('NA_BENCH_LIGHT') != '1' else float('nan')
    clk = clocks.stop() if rank == 0 else None
    img = step(False)
    torch.cuda.synchronize()

    # ---- roofline of the dominant kernel (the fused MLP kernel, SDF-only mode: 512 of the 704 evaluations per ray) --------
    roof = None
    light = os.environ.get('NA_BENCH_LIGHT') == '1'          # profiling runs: skip the isolated-kernel and e2e sections
    if rank == 0 and not light:
        pk, pk_kind = peaks()
        m = 8 * 1024 * 1024                                            # 8 Mi points per launch (6.3 % of one frame's sampler work)
        g = torch.Generator(device=dev); g.manual_seed(1)
        x = (torch.rand(m, 3, device=dev, generator=g) * 4 - 2)
        eng = model.engine()
        for _ in range(2):
            eng.sdf_eval(x, apply_bg=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            eng.sdf_eval(x, apply_bg=True)
        e1.record(); torch.cuda.synchronize()
        t_k = e0.elapsed_time(e1) * 1e-3 / reps
        v = torch.nn.functional.normalize(torch.randn(m // 4, 3, device=dev, generator=g), dim=-1)
        eng.full_eval(x[:m // 4], v, want_feat=False)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            eng.full_eval(x[:m // 4], v, want_feat=False)
        e1.record(); torch.cuda.synchronize()
        t_f = e0.elapsed_time(e1) * 1e-3 / reps
        ach = m * F_SDF / t_k / 1e12
        peak = pk['bf16_tflops']
        # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel (profiles/r1n_tmem_v2.md:
        # dram__bytes_read.sum + dram__bytes_write.sum = 13.9 B/sample SDF-only, 5.5 KB/sample full), scaled to this launch size.
        cap = traffic_capture()
        traffic = None if cap is None or cap.get('precision') != args.precision else (cap['sdf_only_bytes_per_sample'], cap['full_bytes_per_sample'])
        roof = {'bound': 'tensor', 'kernel': 'mlp kernel, SDF-only mode', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s',
                'frac': ach / peak, 'traffic': None if traffic is None else traffic[0] * m,
                'traffic_src': None if traffic is None else f"profiles/traffic_capture.json: {cap['source']} ({cap['date']}); bytes/sample x samples_per_launch",
                'peak_kind': f'{pk_kind} bf16 burst (MEASURED_PEAKS.json)',
                'flop_per_sample': F_SDF, 'samples_per_launch': m, 'launch_ms': t_k * 1e3,
                'full_mode': {'traffic': None if traffic is None else traffic[1] * (m // 4), 'achieved': (m // 4) * F_FULL / t_f / 1e12, 'flop_per_sample': F_FULL, 'launch_ms': t_f * 1e3,
                              'frac': (m // 4) * F_FULL / t_f / 1e12 / peak},
                'frame_flop': n_rays * (4 * N_SAMPLES * F_SDF + P * F_FULL),
                'frame_frac_of_peak': n_rays * (4 * N_SAMPLES * F_SDF + P * F_FULL) / (t_dev / args.steps) / 1e12 / peak / world}
        del x, v
    probe = None
    beta_probe = None
    if rank == 0 and not light and world == 1:
        probe = train_probe(dev, args.precision)
        if args.config == 2:
            beta_probe = small_beta_probe(dev, args.precision)
    cpu = None
    ref_gpu = None
    if rank == 0 and not args.no_cpu_baseline:
        import nerfart_oracle as orc
        from helpers import oracle_net
        n_cpu = 512
        rb = reference_baseline(n_cpu)
        if rb is not None:
            cpu = {'value': rb[0], 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'reference',
                   'sample': f"{n_cpu} strided rays of the same frame ({rb[1]:.1f} s), the unmodified reference's volume_render (PyTorch fp32, CPU)",
                   'rgb_linf_vs_reference': float(np.abs(img[torch.as_tensor(rb[3], device=dev)].cpu().numpy() - rb[2]).max())}
            # the same unmodified reference code on this B200 (plain PyTorch CUDA): every 8th ray of the frame
            sel8 = np.arange(0, n_rays, 8, dtype=np.int64)
            rg = reference_baseline(len(sel8), device=str(dev), stride_sel=sel8)
            ref_gpu = {'value': rg[0], 'unit': 'samples/s', 'kind': 'reference', 'device': torch.cuda.get_device_name(dev),
                       'sample': f'every 8th ray of the frame ({len(sel8)} rays, {rg[1]:.2f} s), unmodified reference PyTorch CUDA path, rayschunk 2048',
                       'ms_per_frame_extrapolated': 1e3 * rg[1] * n_rays / len(sel8),
                       'rgb_linf_ours_vs_reference_gpu': float(np.abs(img[torch.as_tensor(sel8, device=dev)].cpu().numpy() - rg[2]).max())}
            torch.cuda.empty_cache()
        else:
            v, dt = cpu_baseline(n_cpu)
            cpu = {'value': v, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
                   'sample': f'{n_cpu} strided rays of the same frame ({dt:.1f} s), numpy+BLAS fp32 oracle port of the reference'}
        # parity spot check of the rendered frame on the same sample (reported, not timed)
        sel = np.linspace(0, n_rays - 1, n_cpu).astype(np.int64)
        ro, rd = orc.get_rays(c2w_h.numpy(), K_h.numpy(), H, W)
        ref = orc.volsdf_render(oracle_net(model, 'volsdf'), ro[sel], rd[sel], N_samples=N_SAMPLES, N_importance=N_IMPORTANCE)
        cpu['rgb_linf_vs_oracle'] = float(np.abs(img[torch.as_tensor(sel, device=dev)].cpu().numpy() - ref['rgb']).max())
    if rank == 0:
        samples = n_rays * P * args.steps
        line = {'metric': f'MLP samples/sec (VolSDF {H}x{W}x{N_SAMPLES})', 'value': samples / t_dev, 'unit': 'samples/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev / args.steps, 'higher_is_better': True,
                'scaling': 'strong', 'vs_baseline': None, 'dtype': DTYPES.get(args.precision, args.precision),
                'data': 'synthetic', 'config': workload_config(args),
                'all_evals_per_s': n_rays * (4 * N_SAMPLES + P) * args.steps / t_dev,
                'e2e': {'value': samples / t_e2e, 'unit': 'samples/s', 'ms_per_step': 1e3 * t_e2e / args.steps,
                        'h2d_bytes_per_step': 2 * 16 * 4, 'd2h_bytes_per_step': n_rays * 3 * 4},
                'gpu_launches': int(launches), 'clocks': clk, 'roofline': roof, 'cpu_baseline': cpu, 'reference_gpu': ref_gpu, 'train_probe': probe, 'small_beta_probe': beta_probe}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
