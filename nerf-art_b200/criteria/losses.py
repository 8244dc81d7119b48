"""The reference's three CLIP style losses on one shared image tower + the text-feature cache.

Same class names, call signatures and arithmetic as criteria/clip_loss.py (CLIPLoss, 155-304), criteria/contrastive_loss.py
(ContrastiveLoss, 91-186) and criteria/patchnce_loss.py (PatchNCELoss, 91-220).  What differs is where the work happens:
`encode_image` is `ClipVisionB32.encode_image` (csrc/clip_vit.cu) and text features come from `TextFeatures`.
Image pre-processing (bicubic / bilinear resize, crop, pad) stays on torch's interpolate: the reference pins
torchvision 0.9.1 (setup_env.sh:4), whose tensor `Resize` is `F.interpolate(..., align_corners=False)` without antialiasing.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .clip_vit import CLIP_MEAN, CLIP_STD


def _clip_normalize(x):
    mean = x.new_tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = x.new_tensor(CLIP_STD).view(1, 3, 1, 1)
    return (x - mean) / std


def _resize_short_side(x, size, mode):
    """transforms.Resize(int): the shorter side becomes `size`, aspect kept (torchvision 0.9 functional_tensor.resize)."""
    h, w = x.shape[-2:]
    if w <= h:
        nw, nh = size, int(size * h / w)
    else:
        nh, nw = size, int(size * w / h)
    if (nh, nw) == (h, w):
        return x
    return F.interpolate(x, size=(nh, nw), mode=mode, align_corners=False)


def _center_crop(x, size):
    h, w = x.shape[-2:]
    top, left = int(round((h - size) / 2.)), int(round((w - size) / 2.))
    return x[..., top:top + size, left:left + size]


class DirectionLoss(nn.Module):
    """clip_loss.py:137-153"""

    def __init__(self, loss_type='mse'):
        super().__init__()
        self.loss_type = loss_type
        self.loss_func = {'mse': nn.MSELoss, 'cosine': nn.CosineSimilarity, 'mae': nn.L1Loss}[loss_type]()

    def forward(self, x, y):
        if self.loss_type == "cosine":
            return 1. - self.loss_func(x, y)
        return self.loss_func(x, y)


class _ClipLossBase(nn.Module):
    def __init__(self, tower, text):
        super().__init__()
        self.tower, self.text = tower, text

    def preprocess(self, images):
        raise NotImplementedError

    def encode_images(self, images):
        return self.tower.encode_image(self.preprocess(images))

    def get_image_features(self, img, norm=True):
        f = self.encode_images(img)
        if norm:
            f = f / f.clone().norm(dim=-1, keepdim=True)
        return f

    def get_text_features(self, class_str, norm=True):
        return self.text(class_str, norm)


class CLIPLoss(_ClipLossBase):
    """Directional CLIP loss: mean(1 - cos(E(target_img) - E(src_img), T(target) - T(source)))  (clip_loss.py:219-254)."""

    def __init__(self, tower, text, direction_loss_type='cosine'):
        super().__init__(tower, text)
        self.text_direction = None
        self.direction_loss = DirectionLoss(direction_loss_type)

    def preprocess(self, images):                                           # clip_loss.py:166-168
        return _clip_normalize(F.interpolate(images, size=(224, 224), mode='bicubic', align_corners=False))

    def compute_text_direction(self, source_class, target_class, norm=True):
        d = (self.get_text_features(target_class, norm) - self.get_text_features(source_class, norm)).mean(axis=0, keepdim=True)
        if norm:
            d = d / d.norm(dim=-1, keepdim=True)
        return d

    def clip_directional_loss(self, src_img, source_class, target_img, target_class):
        if self.text_direction is None:                                     # cached on first use, like clip_loss.py:245-246
            self.text_direction = self.compute_text_direction(source_class, target_class)
        src_encoding = self.get_image_features(src_img)
        target_encoding = self.get_image_features(target_img)
        edit_direction = target_encoding - src_encoding
        edit_direction = edit_direction / edit_direction.clone().norm(dim=-1, keepdim=True)
        return self.direction_loss(edit_direction, self.text_direction).mean()

    def forward(self, src_img, source_class, target_img, target_class):
        return self.clip_directional_loss(src_img, source_class, target_img, target_class)


class ContrastiveLoss(_ClipLossBase):
    """Global contrastive loss, 'euclidean' variant (contrastive_loss.py:139-153)."""

    def __init__(self, tower, text, margin=2.0, distance_type='euclidean'):
        super().__init__(tower, text)
        if distance_type != 'euclidean':
            raise NotImplementedError("only distance_type='euclidean' (the reference's default, the one Trainer builds)")
        self.margin = margin

    def preprocess(self, images):                                           # contrastive_loss.py:98-101
        x = (images + 1.0) / 2.0                                            # Normalize(mean=-1, std=2)
        x = _center_crop(_resize_short_side(x, 224, 'bicubic'), 224)
        return _clip_normalize(x)

    def clip_contrastive_loss(self, src_img, source_class, target_img, target_class):
        source_features = self.get_text_features(source_class, norm=True)
        target_features = self.get_text_features(target_class, norm=True)
        src_encoding = self.get_image_features(src_img)
        target_encoding = self.get_image_features(target_img)
        near = F.pairwise_distance(target_encoding, target_features.detach(), keepdim=True)
        far_text = F.pairwise_distance(target_encoding, source_features.detach(), keepdim=True)
        far_img = F.pairwise_distance(target_encoding, src_encoding.detach(), keepdim=True)
        return torch.mean(torch.pow(near, 2) + torch.pow(torch.clamp(self.margin - far_text, min=0.0), 2) +
                          torch.pow(torch.clamp(self.margin - far_img, min=0.0), 2))

    def forward(self, src_img, source_class, target_img, target_class):
        return self.clip_contrastive_loss(src_img, source_class, target_img, target_class)


class PatchNCELoss(_ClipLossBase):
    """Local (patch) contrastive loss over 12 random crops (patchnce_loss.py:146-220)."""

    def __init__(self, tower, text, target_hw):
        super().__init__(tower, text)
        self.cos = nn.CosineSimilarity()
        self.temperature = 0.07
        self.ZeroPad = nn.ZeroPad2d(padding=(270, 270, 480, 480))            # patchnce_loss.py:114 (the last assignment wins)
        self.target_hw = [int(target_hw[0]), int(target_hw[1])]
        self.crops_per_call = 12

    def preprocess(self, images):                                           # patchnce_loss.py:98-102
        x = (images + 1.0) / 2.0
        x = F.interpolate(x, size=(224, 224), mode='bilinear', align_corners=False)
        return _clip_normalize(x)

    def clip_contrastive_loss(self, source_classes, target_img, target_class):
        source_feature_list = [self.get_text_features(s, norm=True) for s in source_classes]
        target_feature = self.get_text_features(target_class, norm=True)
        target_encoding = self.get_image_features(target_img)
        near = self.cos(target_encoding, target_feature.detach())
        neg_texts_sum = 0
        for sf in source_feature_list:
            neg_texts_sum = neg_texts_sum + torch.exp(self.cos(target_encoding, sf.detach()) / self.temperature)
        pos = torch.exp(near / self.temperature)
        return torch.mean(-torch.log(pos / (pos + neg_texts_sum)))

    def sample_crops(self, H, W, th, tw, is_full_res):
        """the reference's torch.randint draws, in its order (patchnce_loss.py:196-212)"""
        out = []
        for _ in range(self.crops_per_call):
            i = torch.randint(0, H - th + 1, size=(1,)).item()
            if H != W:
                m = 200 if is_full_res else 100
            else:
                m = 80 if is_full_res else 40
            i = torch.randint(m, H - th + 1 - m, size=(1,)).item()
            j = torch.randint(0, W - tw + 1, size=(1,)).item()
            out.append((i, j))
        return out

    def forward(self, source_classes, target_img, target_class, is_full_res):
        target_img = self.ZeroPad(target_img)
        target_img = F.interpolate(target_img, size=tuple(self.target_hw), mode='bicubic', align_corners=False)   # 117,184
        B, C_, H, W = target_img.shape
        th, tw = (224, 224) if is_full_res else (112, 112)
        crops = []
        for i, j in self.sample_crops(H, W, th, tw, is_full_res):
            img = target_img[..., i:i + th, j:j + tw]
            if not is_full_res:
                img = F.interpolate(img, size=(224, 224), mode='bicubic', align_corners=False)
            crops.append(img)
        # the 12 crops go through the image tower as ONE batch; the loss is the sum of the per-crop means (B = 1 in every
        # shipped config, so mean over the batch of one crop == that crop's value) -- identical to the reference's loop
        batch = torch.cat(crops, 0)
        source_feature_list = [self.get_text_features(s, norm=True) for s in source_classes]
        target_feature = self.get_text_features(target_class, norm=True)
        enc = self.get_image_features(batch)                                 # [12*B, 512]
        return self.crop_losses(enc, source_feature_list, target_feature, B)

    def crop_losses(self, enc, source_feature_list, target_feature, B):
        """sum over the crops of mean(-log(pos / (pos + sum_s exp(cos(e, T_s) / tau))))  (patchnce_loss.py:146-160 per crop)"""
        if B == 1:
            # the same arithmetic for all crops, classes and templates at once (the reference's loop is 12 x 9 cosine / exp / add
            # launches forward and twice that backward; the step is host-bound there): cos [12, 8, T], near [12, T]
            sf = torch.stack([f.detach() for f in source_feature_list])      # [8, T, 512]
            neg = torch.exp(F.cosine_similarity(enc[:, None, None, :], sf[None], dim=-1) / self.temperature).sum(1)
            pos = torch.exp(F.cosine_similarity(enc[:, None, :], target_feature.detach()[None], dim=-1) / self.temperature)
            return torch.mean(-torch.log(pos / (pos + neg)), dim=-1).sum()
        total = 0
        for c in range(enc.shape[0] // B):
            e = enc[c * B:(c + 1) * B]
            near = self.cos(e, target_feature.detach())
            neg = 0
            for sf in source_feature_list:
                neg = neg + torch.exp(self.cos(e, sf.detach()) / self.temperature)
            pos = torch.exp(near / self.temperature)
            total = total + torch.mean(-torch.log(pos / (pos + neg)))
        return total
