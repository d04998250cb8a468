"""Deviation of the TF32-class (1-pass fp16 operands) encoder from the fp32-class (3-pass) one,
end to end, on a C2-shaped image: max-abs over the upsampled features and over the guidance map."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, C, gi, to, lo, K = 2, 768, 448, 896, 32, 7
m = naf_b200.NAF(kernel_size=K).eval().to(dev)
img = torch.randn(B, 3, gi, gi, device=dev)
ft = torch.randn(B, C, lo, lo, device=dev)
res = {}
with torch.no_grad():
    for name, flag in (("tf32_class_1pass", True), ("fp32_class_3pass", False)):
        torch.backends.cudnn.allow_tf32 = flag
        x, rep = m.image_encoder.guidance_source(img, (to, to))
        out = m(img, ft, (to, to))
        res[name] = (x.float().clone(), out.float().clone())
    # cuDNN strict fp32 encoder as a third opinion on the guidance map
    torch.backends.cudnn.allow_tf32 = False
    m.image_encoder.tc_encoder = False
    x_cudnn, _ = m.image_encoder.guidance_source(img, (to, to))
    out_cudnn = m(img, ft, (to, to))
x1, o1 = res["tf32_class_1pass"]
x3, o3 = res["fp32_class_3pass"]
print(f"guidance map: |x| max {x3.abs().max():.3f} rms {x3.pow(2).mean().sqrt():.3f}; 1-pass vs 3-pass max-abs {(x1 - x3).abs().max():.3e} rms {(x1 - x3).pow(2).mean().sqrt():.3e}")
print(f"guidance map: 3-pass vs cuDNN strict fp32 max-abs {(x3 - x_cudnn).abs().max():.3e}")
print(f"output: |out| max {o3.abs().max():.3f}; 1-pass vs 3-pass max-abs {(o1 - o3).abs().max():.3e} rms {(o1 - o3).pow(2).mean().sqrt():.3e}")
print(f"output: 3-pass vs cuDNN-strict-encoder path max-abs {(o3 - out_cudnn).abs().max():.3e}")
torch.backends.cudnn.allow_tf32 = True
m.image_encoder.tc_encoder = False
with torch.no_grad():
    out_cudnn_tf32 = m(img, ft, (to, to))
print(f"output: cuDNN TF32 encoder (what the reference runs by default on GPU) vs strict max-abs {(out_cudnn_tf32 - out_cudnn).abs().max():.3e} rms {(out_cudnn_tf32 - out_cudnn).pow(2).mean().sqrt():.3e}")
