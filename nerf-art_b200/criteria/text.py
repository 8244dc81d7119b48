"""Prompt-template text features, computed once and shared.

The reference re-encodes the 79-template prompt set of every class string on every call: 2 sets per ContrastiveLoss call
(contrastive_loss.py:140-141) and 9 sets for each of PatchNCE's 12 crops (patchnce_loss.py:153-160) -- about 8.7 k text-tower
passes per training iteration -- although the features are pure functions of the class string.  `TextFeatures` maps a class
string to its L2-normalised [n_templates, 512] feature matrix, calling the text tower at most once per string.
The arithmetic is unchanged: `encode_text(tokenize(templates.format(s)))` then `/= norm` (clip_loss.py:222-233).
"""
import torch


def reference_templates():
    """`imagenet_templates` of the reference tree (criteria/clip_loss.py:11-91).  Not duplicated here: when this package is
    used inside the reference repository the list is imported from it."""
    try:
        from criteria.clip_loss import imagenet_templates        # reference tree on sys.path (train.py's working directory)
        return list(imagenet_templates)
    except Exception as e:                                        # pragma: no cover
        raise RuntimeError('prompt templates: pass `templates=[...]` or run inside the NeRF-Art tree so that '
                           'criteria.clip_loss.imagenet_templates is importable') from e


class TextFeatures:
    def __init__(self, encode_fn, templates=None):
        """encode_fn(list[str]) -> [len, 512] un-normalised text features (e.g. lambda t: model.encode_text(clip.tokenize(t)))."""
        self.encode_fn = encode_fn
        self.templates = templates
        self.cache = {}
        self.encodes = 0

    def __call__(self, class_str, norm=True):
        key = (class_str, bool(norm))
        f = self.cache.get(key)
        if f is None:
            if self.templates is None:
                self.templates = reference_templates()
            with torch.no_grad():
                f = self.encode_fn([t.format(class_str) for t in self.templates]).detach().float()
            self.encodes += 1
            if norm:
                f = f / f.norm(dim=-1, keepdim=True)
            self.cache[key] = f
        return f
