#!/usr/bin/env python
"""bench_train.py -- the fine-tune step (BASELINE.json configs[2]/[4] shape) on the synthetic VolSDF 480x270x128 scene.
Reached through `python bench.py --workload train [...]`; `bench.py` without flags stays the render bench (configs[1]).

One "step" = one `Trainer.forward` + `optimizer.step()`:  pass 1 (no-grad render of all 129 600 rays), style loss, pass 2
(108 patches of 1200 rays: forward render with detailed outputs + backward kernels), weight-norm unpack, Adam.
Style loss (`--style`): 'clip' (default) = the three CLIP losses of nerfart_b200.criteria (directional + global contrastive +
PatchNCE over 12 crops: 17 image-tower passes forward and 15 backward per step on csrc/clip_vit.cu) with SEEDED RANDOM tower
weights and seeded stand-in text features -- openai/CLIP weights are not available offline, the arithmetic and its cost are the
same; the VGG perceptual term is omitted (torchvision weights unavailable).  'mse' = weighted MSE to a
constant target (render + backward only).
value = n_rays * 192 / t  (full-MLP samples per second of wall-clock step time, same unit as the render bench).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
H, W = 480, 270
N_SAMPLES, N_IMPORTANCE = 128, 64
P = N_SAMPLES + N_IMPORTANCE
# algorithmic matmul FLOPs per sample (SURVEY.md 8d MAC counts x 2)
F_FULL = 2 * (524544 + 459008 + 265216)                      # forward: sdf + reverse sweep (nabla) + radiance
F_BWD = 2 * (2 * 265216 + 2 * 524544 + 2 * 459008)           # backward: data + weight gradients of radiance, trunk, second-order sweep


class _A(dict):
    __getattr__ = dict.__getitem__


def cpu_baseline_train(n_rays_sample=48):
    """The numpy port of the reference's pass 2 (oracle: forward re-evaluation + closed-form backward, float32) on the host cores."""
    import nerfart_oracle as orc
    import nerfart_oracle_train as ot
    from helpers import make_volsdf
    import fixtures as fx
    m = make_volsdf(0.1, 0.0)
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    net = ot.TrainNet(sd, 'volsdf', dtype=np.float32)
    c2w, K = fx.closed_form_camera(H, W)
    ro, rd = orc.get_rays(c2w.numpy(), K.numpy(), H, W)
    sel = np.linspace(0, H * W - 1, n_rays_sample).astype(np.int64)
    t0 = time.time()
    fwd = orc.volsdf_render(orc.Net(sd, 'volsdf'), ro[sel], rd[sel], N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, detailed_output=True)
    G = np.full((n_rays_sample, 3), 1e-3, np.float32)
    ot.volsdf_backward(net, ro[sel], rd[sel], fwd['d_vals'], G, 0.1, False)
    dt = time.time() - t0
    return n_rays_sample * P / dt, dt


def main(args):
    # a step that stops making progress is reported (Python stacks on stderr) instead of blocking the caller forever
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('NA_BENCH_WATCHDOG_S', '900')), exit=True)
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        if rank == 0:
            v, dt = cpu_baseline_train()
            print(json.dumps({'impl': 'reference', 'metric': 'MLP samples/sec (VolSDF 480x270x128 fine-tune step)', 'value': v,
                              'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': 1, 'warmup': 0, 'ms_per_step': dt * 1e3,
                              'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                              'config': {'workload': 'pass 2 (render + backward) of the fine-tune step on 48 strided rays of the frame'},
                              'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
                                               'sample': '48 strided rays x 192 samples, numpy float32 port (forward + closed-form backward)'},
                              'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}), flush=True)
        return
    import nerfart_b200
    import bench as B
    from helpers import make_volsdf, make_neus
    import fixtures as fx
    from nerfart_b200.models.frameworks import volsdf as pv, neus as pn
    import torch.distributed as dist
    neus = getattr(args, 'framework', 'volsdf') == 'neus'
    P = 128 if neus else N_SAMPLES + N_IMPORTANCE                  # NeuS: configs/neus_fangzhou_vangogh.yaml geometry, 64 + 64 samples
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    saved_stdout = None
    if world > 1:
        sys.stdout.flush(); saved_stdout = os.dup(1); os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev); dist.barrier()
    if args.precision == 'auto':
        args.precision = nerfart_b200.default_precision()
    model = (make_neus(0.05, 0.0, device=dev) if neus else make_volsdf(0.1, 0.0, device=dev)).train()
    model.engine().precision = args.precision
    n_rays = H * W
    target = torch.full((1, n_rays, 3), 0.25)
    wts = torch.linspace(0.5, 1.5, n_rays * 3, device=dev).reshape(1, 3, H, W)
    zero = lambda *a, **k: torch.zeros((), device=dev)
    style = getattr(args, 'style', 'clip')
    if style == 'clip':
        from nerfart_b200.criteria import make_loss_dict, TextFeatures
        from nerfart_b200.criteria.clip_vit import ClipVisionB32

        def fake_text(strings):                       # stand-in text tower: seeded features per prompt (cached by TextFeatures)
            out = []
            for s_ in strings:
                gg = torch.Generator(device='cpu'); gg.manual_seed(sum(ord(c) * (i + 1) for i, c in enumerate(s_)) % (2 ** 31))
                out.append(torch.randn(512, generator=gg))
            return torch.stack(out).to(dev)
        text = TextFeatures(fake_text, templates=['a photo of a {}.'] * 79)
        loss_dict = make_loss_dict(ClipVisionB32.random(0, dev), text, [H, W])
        # steady state of the text-feature cache (criteria/text.py): after the first iterations of a run every prompt of the fixed
        # negative-prompt list has been encoded once; the stand-in text tower's host-side cost must not be billed to the timed steps
        for s_ in [f'negative prompt {i}' for i in range(40)] + ['photo', 'painting']:
            text(s_, True)
    else:
        loss_dict = {'clip': lambda gt, s, pred, t: ((pred - gt) ** 2 * wts).mean(), 'perceptual': None, 'contrastive': zero, 'patchnce': zero}
    trainer = (pn if neus else pv).Trainer(model, is_finetune=True, target_hw=[H, W], loss_dict=loss_dict)
    trainer.neg_texts = [f'negative prompt {i}' for i in range(40)]
    targs = _A(training=_A(is_finetune=True), data=_A(downscale=2), model=_A(radiance=_A(use_view_dirs=True)),
               finetune=_A(use_eikonal=True, w_eikonal=0.1, w_clip=1.0, w_perceptual=2.0, w_contrastive=0.2, w_patchnce=0.1,
                           src_text='photo', target_text='painting'))
    c2w, K = fx.closed_form_camera(H, W)
    if neus:
        c2w = c2w.clone(); c2w[2, 3] = -0.9                                      # inside NeuS' unit bounding sphere
        kw = dict(batched=True, perturb=True, white_bkgd=False, upsample_algo='official_solution', N_upsample_iters=4, N_outside=0,
                  obj_bounding_radius=1.0, H=H, W=W, N_samples=64, N_importance=64)
    else:
        kw = dict(near=0.0, far=6.0, batched=True, perturb=True, white_bkgd=False, max_upsample_steps=6, use_nerfplusplus=False,
                  obj_bounding_radius=3.0, H=H, W=W, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE)
    c2w_pin, K_pin, tgt_pin = c2w[None].pin_memory(), K[None].pin_memory(), target.pin_memory()
    opt = torch.optim.Adam(model.parameters(), lr=1e-6)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    # phase timers around the engine calls of pass 2
    eng = model.engine()
    phases = {'render_bwd': [], 'patch_fwd': []}
    orig_bwd, orig_fwd = eng.render_bwd, (eng.neus_render if neus else eng.volsdf_render)

    def timed_call(name, fn):
        def wrap(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(*a, **k); e1.record()
            phases[name].append((e0, e1, bool(k.get('detailed_output', True))))
            return r
        return wrap
    eng.render_bwd = timed_call('render_bwd', orig_bwd)
    if neus:
        eng.neus_render = timed_call('patch_fwd', orig_fwd)
    else:
        eng.volsdf_render = timed_call('patch_fwd', orig_fwd)
    from nerfart_b200.models.frameworks import _finetune
    phases['style'] = []
    style_host = []
    orig_style = _finetune.calc_style_loss

    def style_timed(*a, **k):                          # forward of the style losses; their backward runs inside losses.backward()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        e0.record(); r = orig_style(*a, **k); e1.record()
        phases['style'].append((e0, e1, True))
        style_host.append(1e3 * (time.perf_counter() - h0))
        return r
    _finetune.calc_style_loss = style_timed

    def step():
        ret = trainer(targs, None, {'intrinsics': K_pin.to(dev, non_blocking=True), 'c2w': c2w_pin.to(dev, non_blocking=True)},
                      {'rgb': tgt_pin}, kw, 0, optimizer=opt)
        opt.step()
        loss_host.copy_(ret['losses'].detach(), non_blocking=True)
        return ret

    import contextlib, io
    def quiet_step():
        with contextlib.redirect_stdout(io.StringIO()):
            return step()
    quiet_step()
    if saved_stdout is not None:
        torch.cuda.synchronize(); sys.stdout.flush(); os.dup2(saved_stdout, 1); os.close(saved_stdout)
    for _ in range(max(args.warmup - 1, 0)):
        quiet_step()
    for v in phases.values():
        v.clear()
    clocks = B.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = nerfart_b200.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        quiet_step()
    e1.record(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt.item())
    launches = nerfart_b200.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    t_bwd = sum(a.elapsed_time(b) for a, b, _ in phases['render_bwd']) * 1e-3 / args.steps
    t_pfwd = sum(a.elapsed_time(b) for a, b, det in phases['patch_fwd'] if det) * 1e-3 / args.steps
    t_p1 = sum(a.elapsed_time(b) for a, b, det in phases['patch_fwd'] if not det) * 1e-3 / args.steps
    t_style = sum(a.elapsed_time(b) for a, b, _ in phases['style']) * 1e-3 / args.steps
    if rank == 0:
        pk, kind = B.peaks()
        n_local = n_rays if world == 1 else None
        cpu = None
        if not args.no_cpu_baseline and not neus:
            v, dt = cpu_baseline_train()
            cpu = {'value': v, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
                   'sample': f'pass 2 only on 48 strided rays x 192 samples ({dt:.1f} s), numpy float32 port (render + closed-form backward)'}
        # NeuS (radiance net frozen): P points carry trunk + second-order sweep (data + weight gradients), P - 1 midpoints the same
        # plus the radiance net's backward-data GEMMs
        flop_ray = (P * 2 * (2 * 524544 + 2 * 459008) + (P - 1) * 2 * (265216 + 2 * 524544 + 2 * 459008)) if neus else P * F_BWD
        ach = n_rays * flop_ray / max(t_bwd, 1e-9) / 1e12 / world
        line = {'metric': 'MLP samples/sec (NeuS 480x270x64+64 fine-tune step)' if neus else 'MLP samples/sec (VolSDF 480x270x128 fine-tune step)', 'value': n_rays * P * args.steps / t, 'unit': 'samples/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps,
                'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': ('backward: the patch forward render leaves the activations (16-bit stash), the 20 backward-data GEMMs per tile in one tcgen05 launch (f16 operands, f32 accumulate), weight-gradient GEMMs tcgen05 bf16 operands from the stash, f32 accumulate; ' if args.precision != 'fp32' else 'backward: f32 recompute + tf32 mma.sync; ') + 'CLIP tower linear layers tcgen05 tf32; forward: ' + args.precision, 'data': 'synthetic',
                'config': {'workload': (f'NeuS fine-tune step {H}x{W} ({n_rays} rays), 64+64 samples/ray (sdf / nabla at 128 points + radiance at 127 '
                                        'midpoints), radiance net frozen, ' if neus else f'VolSDF fine-tune step {H}x{W} ({n_rays} rays), 128+64 samples/ray, ') +
                                       'pass 1 + pass 2 in 108 patches of 1200 rays, eikonal on, perturb on, style=' + style + (' (3 CLIP losses, seeded random ViT-B/32 weights + stand-in text features, no VGG)' if style == 'clip' else ' (weighted MSE)') + ', Adam step',
                           'parallelism': f'patch round-robin x{world}' + (' + NCCL all-reduce of the packed gradient' if world > 1 else ''),
                           'l2': 'per-patch stash (5.0 GB) >> 126 MB L2; no explicit flush'},
                'phases_ms': {'pass1_render': 1e3 * t_p1, 'pass2_patch_forward': 1e3 * t_pfwd, 'pass2_backward': 1e3 * t_bwd,
                              'style_loss_forward': 1e3 * t_style,
                              'other (style backward, unpack, Adam, host)': 1e3 * (t / args.steps - t_p1 - t_pfwd - t_bwd - t_style)},
                'style_ms_per_step': [round(a.elapsed_time(b), 2) for a, b, _ in phases['style']],
                'style_host_ms_per_step': [round(v, 2) for v in style_host[-args.steps:]],
                'e2e': {'value': n_rays * P * args.steps / t, 'unit': 'samples/s', 'ms_per_step': 1e3 * t / args.steps,
                        'h2d_bytes_per_step': 2 * 64 + n_rays * 12, 'd2h_bytes_per_step': 4,
                        'note': 'the step is timed through Trainer.forward with pinned-host camera / target image in and the loss out'},
                'gpu_launches': int(launches), 'clocks': clk,
                'roofline': {'bound': 'tensor', 'kernel': 'tm::mlp_tmem_kernel<1,1,1,*> (backward half) + wf::wgrad_f16_kernel (pass-2 backward)',
                             'achieved': ach, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': ach / pk['bf16_tflops'], 'traffic': None,
                             'flop_per_sample': flop_ray / P, 'note': 'algorithmic backward FLOPs / time inside render_bwd (compositing backward + backward half of the split training program '
                             '+ weight gradients); NA_BW_SPLIT=0: the launch also re-evaluates the forward pass (41 tile GEMMs); all weight-gradient GEMMs on '
                             'tcgen05 kind::f16 fed by TMA from the 16-bit stash', 'split_program': os.environ.get('NA_BW_SPLIT', '1') != '0' and args.precision != 'fp32', 'peak_kind': f'{kind} bf16 burst'},
                'cpu_baseline': cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
