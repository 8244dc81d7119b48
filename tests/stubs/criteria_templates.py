"""A short prompt-template list for tests that must not depend on the reference tree's criteria/clip_loss.py."""
TEMPLATES = ['a photo of a {}.', 'a rendering of a {}.', 'a painting of the {}.', 'art of a {}.']
