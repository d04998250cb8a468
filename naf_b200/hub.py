"""Hub-style factory and the test-harness wrapper of the reference, for the NAF path only.

`naf(pretrained, device)` mirrors hubconf.py:8-24; `ModelWrapper` mirrors utils/wrapper.py:8-52
restricted to name="NAF" (the other names are the reference's competitor baselines, out of scope).
"""
from __future__ import annotations

import torch
from torch import nn

from .model import NAF

RELEASE_URL = "https://github.com/valeoai/NAF/releases/download/model/naf_release.pth"


def naf(pretrained: bool = True, device="cpu"):
    model = NAF().to(device)
    if pretrained:
        state = torch.hub.load_state_dict_from_url(RELEASE_URL, progress=True, map_location=device)
        model.load_state_dict(state)
    return model


class ModelWrapper(nn.Module):
    def __init__(self, name="NAF", embed_dim=384, ratio=16, ckpt_path: str = None):
        super().__init__()
        if name != "NAF":
            raise ValueError(f"Unknown upsampler: {name} (naf_b200 implements 'NAF' only)")
        self.name, self.embed_dim, self.ratio = name, embed_dim, ratio
        self.model = NAF()
        if ckpt_path is not None:
            self.model.load_state_dict(torch.load(ckpt_path, map_location="cpu"), strict=False)

    def forward(self, image, features, output_size):
        return self.model(image, features, output_size)
