"""The CLIP fine-tune step shared by the VolSDF and NeuS trainers (reference: Trainer.forward, fine-tune branch,
models/frameworks/volsdf.py:719-783 and models/frameworks/neus.py:520-576).

Two passes over the whole image, exactly like the reference:
  pass 1  no-grad render of all H*W rays -> style loss on the image -> d loss / d rgb            (volsdf.py:724-749)
  pass 2  re-render in patches of `batch_size` rays and back-propagate the image gradient slice plus the eikonal loss of
          the patch                                                                               (volsdf.py:754-783)
Pass 2 is where autograd spends its time in the reference; here it is `NetEngine.render_bwd` (csrc/train.cu): the forward
render of the patch with its detailed outputs, then hand-written backward kernels that accumulate into a packed gradient
buffer, mapped to the parameters' `.grad` once per step (`NetEngine.unpack_grads`).
"""
import os
import random
from collections import OrderedDict

import torch
from einops import rearrange

from ...utils import rend_util
from ... import parallel

BATCH_SIZE = 1200            # volsdf.py:754 / neus.py:541 ("hardcoded for 3090Ti")


def create_fine_neg_texts(args, path=None):
    """Trainer.create_fine_neg_texts (volsdf.py:649-683 / neus.py:458-491): negative prompts grouped under '#key' lines of
    criteria/neg_text.txt; the group matching the target style is dropped."""
    if path is None:
        path = "criteria/neg_text.txt"
        if not os.path.exists(path):
            path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'criteria', 'neg_text.txt')
    results = {}
    curr_key = 0
    with open(path, 'r') as fr:
        for item in fr.readlines():
            item = item.strip()
            if item.startswith("#"):
                curr_key = item[1:]
                results[curr_key] = []
            else:
                results[curr_key].append(item.split(".")[1])
    remove_ids = []
    ttext = args.finetune.target_text.lower()
    if 'botero' in ttext or 'monalisa' in ttext or 'portrait' in ttext or 'painting' in ttext:
        remove_ids = ['portrait']
    elif 'zombie' in ttext:
        remove_ids = ['zombie']
    elif 'wolf' in ttext:
        remove_ids = ['wolf']
    elif 'pixlar' in ttext or 'disney' in ttext:
        remove_ids = ['disney']
    elif 'sketch' in ttext:
        remove_ids = ['sketch']
    all_texts = []
    for key in results:
        if key not in remove_ids:
            all_texts += results[key]
    return all_texts


def calc_style_loss(trainer, rgb, rgb_gt, args, H):
    """Trainer.calc_style_loss (volsdf.py:878-915 / neus.py:629-665): directional CLIP + perceptual + global contrastive +
    PatchNCE, with the reference's weights and its `random` draws (one negative prompt, then 8 for the patches)."""
    loss = 0.0
    rgb_pred = rearrange(rgb, "B (H W) C -> B C H W", H=H)
    rgb_gt = rearrange(rgb_gt, "B (H W) C -> B C H W", H=H)
    s_text = args.finetune.src_text
    t_text = args.finetune.target_text
    ld = trainer.loss_dict
    dir_clip_loss = ld["clip"](rgb_gt, s_text, rgb_pred, t_text)
    loss = loss + dir_clip_loss * args.finetune.w_clip
    if ld.get("perceptual") is not None and args.finetune.w_perceptual:
        perp_loss = ld["perceptual"](rgb_pred, rgb_gt)
        loss = loss + perp_loss * args.finetune.w_perceptual
    s_text = random.choice(trainer.neg_texts)
    loss_contrastive = ld["contrastive"](rgb_gt, s_text, rgb_pred, t_text)
    loss = loss + loss_contrastive * args.finetune.w_contrastive
    neg_counts = 8
    s_text_list = random.sample(trainer.neg_texts, neg_counts)
    is_full_res = args.data.downscale == 1
    loss_patchnce = ld["patchnce"](s_text_list, rgb_pred, t_text, is_full_res)
    loss = loss + loss_patchnce * args.finetune.w_patchnce
    return loss


def _assign_grads(model, engine, scalar_param, train_surface, train_radiance):
    pairs, scal = engine.unpack_grads(train_surface, train_radiance)
    for p, g in pairs:
        if p.requires_grad:
            p.grad = g if p.grad is None else p.grad + g
    if scalar_param.requires_grad:
        g = scal[0:1].to(torch.float32).reshape(scalar_param.shape)
        scalar_param.grad = g if scalar_param.grad is None else scalar_param.grad + g
    return scal


def patch_groups(starts, n, batch_size, group=None):
    """Patches of one rank -> launch groups.  The reference back-propagates patch by patch (volsdf.py:754-783); the parameter gradient is
    the SUM over patches and every per-sample term (the eikonal mean's 1/N included, N = rays of ONE patch x samples) is independent of
    the other patches, so full-size patches can share a launch: 1200 rays x 192 samples are 1800 tiles = 12.2 waves of 148 CTAs (13 run:
    7 % idle), six patches are 72.97 waves.  NA_PATCH_GROUP (default 6; 1 = the reference's launch structure) bounds the group; groups are
    sized evenly; a short last patch (its eikonal normaliser differs) stays alone."""
    import os
    group = int(os.environ.get('NA_PATCH_GROUP', '6')) if group is None else group
    full = [i for i in starts if i + batch_size <= n]
    part = [i for i in starts if i + batch_size > n]
    out = []
    if full:
        n_groups = -(-len(full) // max(group, 1))
        base, extra = divmod(len(full), n_groups)
        k = 0
        for gi in range(n_groups):
            sz = base + (1 if gi < extra else 0)
            out.append(full[k:k + sz]); k += sz
    out += [[i] for i in part]
    return out


def backward_patches(model, framework, rays_o, rays_d, gradient, render_patch, *, w_eikonal, white_bkgd, batch_size=None):
    """Pass 2.  rays [1,N,3], gradient [1,N,3]; `render_patch(ro, rd)` = the flat detailed forward outputs of one patch
    (NetEngine.volsdf_render / neus_render).  Returns the mean eikonal loss over patches (what the reference prints)."""
    batch_size = BATCH_SIZE if batch_size is None else batch_size
    eng = model.engine()
    train_surface = any(p.requires_grad for p in model.implicit_surface.parameters())
    train_radiance = any(p.requires_grad for p in model.radiance_net.parameters())
    eng.grad_zero()
    ro = rays_o.reshape(-1, 3).float().contiguous()
    rd = rays_d.reshape(-1, 3).float().contiguous()
    g = gradient.reshape(-1, 3).float().contiguous()
    n = ro.shape[0]
    n_patches = 0
    # one process per GPU: patches are dealt round-robin to the ranks, the packed gradient is summed once per step
    # (replaces DDP's bucketed all-reduce, train.py:155)
    world, rank = parallel.world_rank()
    # (NetEngine.pack() re-folds the weights only when a parameter changed: once per step, not once per patch)
    mine = [i for pi, i in enumerate(range(0, n, batch_size)) if pi % world == rank]
    for starts in patch_groups(mine, n, batch_size):
        if len(starts) == 1:
            i = starts[0]
            rop, rdp, gp = ro[i:i + batch_size].contiguous(), rd[i:i + batch_size].contiguous(), g[i:i + batch_size]
        elif all(b - a == batch_size for a, b in zip(starts, starts[1:])):
            i, j = starts[0], starts[-1] + batch_size
            rop, rdp, gp = ro[i:j].contiguous(), rd[i:j].contiguous(), g[i:j]
        else:
            rop = torch.cat([ro[i:i + batch_size] for i in starts]); rdp = torch.cat([rd[i:i + batch_size] for i in starts])
            gp = torch.cat([g[i:i + batch_size] for i in starts])
        fwd, scal = render_patch(rop, rdp)
        P = (fwd['d_vals'] if framework == 'volsdf' else fwd['d_all']).shape[-1]
        # the eikonal term is a mean over ONE patch's samples (volsdf.py:775-777): the normaliser stays batch_size * P for a group
        eng.render_bwd(rop, rdp, scal, fwd, gp, w_eikonal=w_eikonal, eikonal_count=min(batch_size, rop.shape[0]) * P,
                       white_bkgd=white_bkgd, speed_factor=model.speed_factor, train_surface=train_surface,
                       train_radiance=train_radiance)
        n_patches += len(starts)
    parallel.allreduce_sum(eng._gpack, eng._gscal)
    scalar_param = model.ln_beta if framework == 'volsdf' else model.ln_s
    scal = _assign_grads(model, eng, scalar_param, train_surface, train_radiance)
    return scal, n_patches


def finetune_forward(trainer, framework, args, model_input, ground_truth, render_kwargs_train, optimizer, render_patch):
    """Body of Trainer.forward's fine-tune branch.  `render_patch(ro, rd, **render_kwargs)` renders one flat ray patch with
    detailed outputs."""
    if trainer.neg_texts is None:
        trainer.neg_texts = create_fine_neg_texts(args)
    model = trainer.model
    device = next(model.parameters()).device
    intrinsics = model_input["intrinsics"].to(device)
    c2w = model_input['c2w'].to(device)
    H = render_kwargs_train['H']
    W = render_kwargs_train['W']
    world, rank = parallel.world_rank()
    gt_rgb = ground_truth['rgb'].to(device)
    if world > 1:
        # one image per step for the whole job: rank 0's camera / target and one RNG stream for the style-loss draws
        c2w, intrinsics, gt_rgb = [t.contiguous() for t in (c2w.float(), intrinsics.float(), gt_rgb.float())]
        parallel.sync_step_inputs([c2w, intrinsics, gt_rgb])
    rays_o, rays_d, select_inds = rend_util.get_rays(c2w, intrinsics, H, W, -1)     # fine-tune: all rays, not shuffled
    target_rgb = torch.gather(gt_rgb, 1, torch.stack(3 * [select_inds], -1))
    use_eik = bool(args.finetune.use_eikonal)
    n_rays = rays_o.shape[1]
    idx = parallel.rank_rays(n_rays, rank, world, H, W, device=device)               # None: contiguous block; else interleaved tiles
    if idx is None:
        lo, hi, _ = parallel.ray_block(n_rays, rank, world)
        ro1, rd1 = rays_o[:, lo:hi], rays_d[:, lo:hi]
    else:
        ro1, rd1 = rays_o[:, idx], rays_d[:, idx]
    with torch.no_grad():                                                            # pass 1 (this rank's share of the rays)
        rgb, depth_v, _ = trainer.renderer(ro1, rd1, detailed_output=False,
                                           use_view_dirs=args.model.radiance.use_view_dirs,
                                           require_nablas=use_eik or args.model.radiance.use_view_dirs, **render_kwargs_train)
        rgb = parallel.gather_rays(rgb[0], n_rays, idx, H, W)[None]                  # every rank scores the whole image
    rgb = rgb.detach().requires_grad_(True)
    losses = calc_style_loss(trainer, rgb, target_rgb, args, H)
    losses.backward()
    gradient = rgb.grad.clone().detach()
    optimizer.zero_grad()
    kw = {k: v for k, v in render_kwargs_train.items() if k not in ('H', 'W', 'batched')}
    # tensor-core modes: the patch's forward render is the forward half of the training program, the backward launch runs the backward
    # half only (one network evaluation per sample point, as in the reference's autograd).  NA_BW_SPLIT=0: the backward launch
    # re-evaluates the forward pass (the only form in fp32 mode).
    if model.engine().precision in ('tc', 'tc_mixed') and os.environ.get('NA_BW_SPLIT', '1') != '0':
        kw['train_stash'] = True
    scal, n_patches = backward_patches(model, framework, rays_o, rays_d, gradient, lambda ro, rd: render_patch(ro, rd, **kw),
                                       w_eikonal=args.finetune.w_eikonal if use_eik else 0.0,
                                       white_bkgd=render_kwargs_train.get('white_bkgd', False))
    avg_eikonal_loss = float(scal[1].item()) / max(gradient.shape[1] // BATCH_SIZE, 1)        # volsdf.py:784
    print("\tEikonal loss: ", avg_eikonal_loss * args.finetune.w_perceptual)                   # volsdf.py:785 (sic)
    return losses, select_inds
