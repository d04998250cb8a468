"""NAF upsampler: forward of the reference `src/model/naf.py` rebuilt around the B200 kernels.

Module tree, constructor keywords and `state_dict` keys are those of the reference (`NAF`,
`ImageEncoder`, `QueryEncoder`, `KeyEncoder`; src/model/naf.py:11-116) so released checkpoints
load with `strict=True`.  The data flow differs:

    reference                                          here
    ---------                                          ----
    conv encoder -> adaptive pool            (ATen)    same library ops, channels_last storage
    RoPE: ~12 elementwise passes over Q      (ATen)    folded into the two kernels below
    K = adaptive_avg_pool2d(RoPE(x))         (ATen)    naf_rope_kpool_f32   (reads x once)
    nearest-exact K,V to target res + NATTEN           naf_xattn_fwd_f32    (reads x once, rotates
      QK / softmax / AV on the dilated grid                                  on the fly, low-res K,V
                                                                             windows in smem)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import encoder_fast, ops
from ..layers import CrossAttention, RoPE, encoder


class ImageEncoder(nn.Module):
    def __init__(self, in_channels=3, out_channels=256, heads_rope=1, use_encoder=True,
                 rope_base=None, rope_rescale=None, img_layers=2):
        super().__init__()
        self.use_encoder = use_encoder
        self.out_channels = out_channels
        self.fast_encoder = True   # False: run the conv stacks as plain torch modules
        self.tc_encoder = True     # False: cuDNN convolutions between our GroupNorm kernels
        half = out_channels // 2
        self.encoder = encoder(in_channels, half, kernel_size=1, ks_res=1, num_layers=img_layers)
        self.sem_encoder = encoder(in_channels, half, kernel_size=3, ks_res=3, num_layers=img_layers)
        self.rope = RoPE(embed_dim=out_channels, num_heads=heads_rope, base=rope_base,
                         rescale_coords=rope_rescale)

    @staticmethod
    def encoder_resolution(in_hw, out_hw):
        """Resolution the conv stack runs at: the guidance image is capped at 4x the target
        (src/model/naf.py:39-48, including its mixed H/W `min`)."""
        (Hi, Wi), (Ho, Wo) = in_hw, out_hw
        if Hi > 4 * Ho or Wi > 4 * Wo:
            return min(Hi, 4 * Ho, 4 * Wo), min(Wi, 4 * Wo, 4 * Ho)
        return Hi, Wi

    def forward_encoder(self, x, output_size):
        if self.use_encoder:
            x = torch.cat([self.encoder(x), self.sem_encoder(x)], dim=1)
        if tuple(x.shape[-2:]) != tuple(int(s) for s in output_size):
            x = F.adaptive_avg_pool2d(x, output_size=output_size)  # identity when sizes match
        return x

    def _capped(self, image, Ho, Wo):
        size = self.encoder_resolution(image.shape[-2:], (Ho, Wo))
        if size != tuple(image.shape[-2:]):
            image = F.interpolate(image, size=size, mode="bilinear", align_corners=False)
        return image

    def guidance(self, image, output_size):
        """Pooled, UN-rotated guidance map (B, D, Ho, Wo) -- the reference's tensor, materialised."""
        Ho, Wo = int(output_size[0]), int(output_size[1])
        return self.forward_encoder(self._capped(image, Ho, Wo), (Ho, Wo))

    def guidance_source(self, image, output_size):
        """(x_src, (ry, rx)): the pooled guidance map WITHOUT materialising replication.

        When the target size is an integer multiple of the encoder resolution (equal sizes
        included) the reference's `adaptive_avg_pool2d` (src/model/naf.py:34) only replicates
        pixels, so the kernels read the encoder-resolution map through `rep` instead; the two
        branch outputs are concatenated and packed pixel-major by our packing kernel in the same
        pass.  Otherwise this falls back to the materialised `guidance()` with rep (1, 1)."""
        Ho, Wo = int(output_size[0]), int(output_size[1])
        image = self._capped(image, Ho, Wo)
        Hs, Ws = image.shape[-2:]
        if not (image.is_cuda and self.use_encoder):
            return self.forward_encoder(image, (Ho, Wo)), (1, 1)
        replicate = Ho % Hs == 0 and Wo % Ws == 0
        if (self.fast_encoder and self.tc_encoder and encoder_fast.tc_supported(self.encoder)
                and encoder_fast.tc_supported(self.sem_encoder) and Hs >= 2 and Ws >= 2):
            # whole conv stack on our tcgen05 kernels; both branches write their 128-channel slab
            # of the concatenated guidance map directly (no torch.cat pass)
            B = image.shape[0]
            x = torch.empty((B, Hs, Ws, 2 * 128), device=image.device, dtype=torch.float32)
            encoder_fast.forward_tc(self.encoder, image, out=x, ch_off=0)
            encoder_fast.forward_tc(self.sem_encoder, image, out=x, ch_off=128)
            x = x.permute(0, 3, 1, 2)
            if replicate:
                return x, (Ho // Hs, Wo // Ws)
            # true pooling (target below the encoder resolution, or a non-integer ratio): ATen's
            # adaptive_avg_pool2d on the pixel-major map, exactly the reference's op (naf.py:34)
            return F.adaptive_avg_pool2d(x, output_size=(Ho, Wo)), (1, 1)
        if not replicate:
            return self.forward_encoder(image, (Ho, Wo)), (1, 1)
        if (self.fast_encoder and image.dtype == torch.float32 and encoder_fast.supported(self.encoder)
                and encoder_fast.supported(self.sem_encoder)):
            # our GroupNorm/SiLU/pad kernels between cuDNN convs, pixel-major end to end
            ya, ba = encoder_fast.forward(self.encoder, image)
            yb, bb = encoder_fast.forward(self.sem_encoder, image)
            return ops.concat_bias_nhwc(ya, ba, yb, bb), (Ho // Hs, Wo // Ws)
        parts = [self.encoder(image), self.sem_encoder(image)]
        return ops.pack_concat_nhwc(parts), (Ho // Hs, Wo // Ws)

    def forward(self, x, output_size):
        return self.rope(self.guidance(x, output_size))


class QueryEncoder(nn.Module):
    def forward(self, x):
        return x


class KeyEncoder(nn.Module):
    """K = block mean of the rotated guidance map at the feature resolution."""

    def forward(self, x, features):
        if torch.is_grad_enabled() and x.requires_grad:
            from ..autograd import KeyPoolFn
            return KeyPoolFn.apply(x, tuple(features.shape[-2:]))
        k, _ = ops.rope_kpool(x, None, 1, pooled_hw=features.shape[-2:], want_q=False)
        return k


class NAF(nn.Module):
    def __init__(self, dim=256, heads_attn=4, heads_rope=4, kernel_size=9, use_encoder=True,
                 rope_base=100.0, rope_rescale=2.0, img_layers=2, **kwargs):
        super().__init__()
        self.image_encoder = ImageEncoder(in_channels=3, out_channels=dim, heads_rope=heads_rope,
                                          use_encoder=use_encoder, rope_base=rope_base,
                                          img_layers=img_layers, rope_rescale=rope_rescale)
        self.query_encoder = QueryEncoder()
        self.key_encoder = KeyEncoder()
        self.upsampler = CrossAttention(dim=dim, num_heads=heads_attn,
                                        kernel_size=(kernel_size, kernel_size))

    def upsample_from_guidance(self, x, features, return_weights=False, rep=(1, 1), out_dtype=None):
        """The hot path proper: pooled un-rotated guidance x (B,D,Ho/ry,Wo/rx) + features
        (B,C,h,w) -> (B,C,Ho,Wo).  Two kernel launches (+ one tiny packing launch for V).
        Differentiable w.r.t. x and features when grad mode is on (naf_b200/autograd.py)."""
        rope = self.image_encoder.rope
        Ho, Wo = x.shape[-2] * rep[0], x.shape[-1] * rep[1]
        tables = rope.axis_tables(Ho, Wo)
        D = x.shape[1]
        if out_dtype is None:
            from ..layers.attentions import autocast_out_dtype
            out_dtype = autocast_out_dtype()
        fused = rope.D_head == D // self.upsampler.num_heads
        h, w = features.shape[-2:]
        ry, rx = int(rep[0]), int(rep[1])
        # Key pooling over replicated guidance: every source pixel stands for an ry x rx block
        # whose rotations differ only through the angles, and the rotation is linear, so the
        # block mean of RoPE(x) is x rotated by the block-MEAN cos/sin.  Pool the source map
        # with mean tables: 1/(ry*rx) of the work, identical result up to summation order.
        pool_tables = None
        if fused and (ry > 1 or rx > 1) and Ho % h == 0 and Wo % w == 0 and (Ho // h) % ry == 0 \
                and (Wo // w) % rx == 0:
            pool_tables = rope.mean_axis_tables(Ho, Wo, ry, rx)
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or features.requires_grad)
        if needs_grad:
            if return_weights:
                raise NotImplementedError("naf_b200: return_weights=True is an inference feature (no gradient "
                                          "flows through the returned scores); call it under torch.no_grad()")
            if fused:
                from ..autograd import NAFUpsampleFn
                return NAFUpsampleFn.apply(x.float(), features, tables, pool_tables, rope.num_heads,
                                           self.upsampler.num_heads, self.upsampler._window(),
                                           self.upsampler.scale, self.upsampler.algo, (ry, rx), out_dtype)
            # rope heads != attention heads: compose the differentiable operators on the materialised map
            if (ry, rx) != (1, 1):
                x = x.repeat_interleave(ry, 2).repeat_interleave(rx, 3)
            q = rope(x.float())
            k = self.key_encoder(q, features)
            return self.upsampler(q, k, features, out_dtype=out_dtype)
        x = ops.as_pixel_major(x)  # once, shared by both kernels
        if pool_tables is not None:
            k, q = ops.rope_kpool(x, pool_tables, rope.num_heads, pooled_hw=(h, w), want_q=False)
        else:
            k, q = ops.rope_kpool(x, tables, rope.num_heads, pooled_hw=(h, w), want_q=not fused, rep=rep)
        if fused:
            return self.upsampler(x, k, features, return_weights=return_weights, rope_tables=tables,
                                  rep=rep, out_dtype=out_dtype)
        return self.upsampler(q, k, features, return_weights=return_weights, out_dtype=out_dtype)

    def forward(self, image, features, output_size, return_weights=False, *args, out_dtype=None, **kwargs):
        """`out_dtype`: None = what the reference returns in this context (bf16 under bf16 autocast,
        else fp32); torch.float32 / torch.bfloat16 to choose.  All arithmetic is fp32 either way.

        With grad mode on the call is differentiable like the reference's when there is something to
        differentiate: the module is in train() mode with trainable parameters (a training loop,
        train.py:127-136), or the features / the image require grad (a backbone or a probe trained through
        a frozen NAF).  The conv encoder then runs as plain torch modules under autograd (cuDNN), the
        attention path through `naf_b200.autograd.NAFUpsampleFn` (our forward AND backward kernels).
        Otherwise -- including eval() mode with grad mode left on and only the parameters requiring
        grad, which every reference eval caller avoids with no_grad -- the inference path runs, all on
        our kernels, and the result is detached (the parameters are treated as frozen)."""
        from ..layers.attentions import autocast_out_dtype
        if out_dtype is None:
            out_dtype = autocast_out_dtype()
        needs_grad = torch.is_grad_enabled() and (
            features.requires_grad or image.requires_grad or
            (self.training and any(p.requires_grad for p in self.parameters())))
        if needs_grad:
            enc = self.image_encoder
            Ho, Wo = int(output_size[0]), int(output_size[1])
            x = enc.forward_encoder(enc._capped(image, Ho, Wo), (Ho, Wo))     # torch modules, autograd
            with torch.autocast("cuda", enabled=False):
                return self.upsample_from_guidance(x.float(), features, return_weights=return_weights,
                                                   out_dtype=out_dtype)
        with torch.no_grad():
            x, rep = self.image_encoder.guidance_source(image, output_size)
            with torch.autocast("cuda", enabled=False):
                return self.upsample_from_guidance(x.float(), features, return_weights=return_weights, rep=rep,
                                                   out_dtype=out_dtype)
