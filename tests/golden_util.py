"""Load the committed golden fixtures (tests/golden/*.npz, made by oracle/gen_golden.py from the
unmodified reference modules) and regenerate their seeded inputs."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def seeded_normal(seed, shape):
    # numpy's legacy MT19937 stream is frozen across numpy versions
    return torch.from_numpy(np.random.RandomState(int(seed)).standard_normal(tuple(int(s) for s in shape)).astype(np.float32))


def names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz")
                  and not f.endswith("_state.npz"))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def attention_case(name):
    d = load(name)
    seed = int(d["seed"])
    q = seeded_normal(seed, d["q_shape"]) * float(d["gain"])
    k = seeded_normal(seed + 1, d["k_shape"])
    v = seeded_normal(seed + 2, d["v_shape"])
    scores = torch.from_numpy(d["scores"]) if "scores" in d else None
    return dict(q=q, k=k, v=v, out=torch.from_numpy(d["out"]), scores=scores, heads=int(d["heads"]),
                K=int(d["kernel_size"]), dilation=tuple(int(x) for x in d["dilation"]))


def rope_case(name):
    d = load(name)
    return dict(x=seeded_normal(int(d["seed"]), d["x_shape"]), out=torch.from_numpy(d["out"]),
                periods=torch.from_numpy(d["periods"]), heads=int(d["heads"]))


def module_case(name):
    d = load(name)
    seed = int(d["seed"])
    return dict(image=seeded_normal(seed, d["image_shape"]), features=seeded_normal(seed + 1, d["features_shape"]),
                out=torch.from_numpy(d["out"]), queries=torch.from_numpy(d["queries"]),
                output_size=tuple(int(x) for x in d["output_size"]))


def module_state():
    d = np.load(os.path.join(GOLDEN, "naf_dim128_k7_state.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


# ---- default-width model (dim=256: the shape the tensor-core encoder implements) --------------
def default_case(name):
    d = load(name)
    seed = int(d["seed"])
    return dict(image=seeded_normal(seed, d["image_shape"]), features=seeded_normal(seed + 1, d["features_shape"]),
                out=torch.from_numpy(d["out"]), queries_sub=torch.from_numpy(d["queries_sub"]),
                qstep=int(d["qstep"]), output_size=tuple(int(x) for x in d["output_size"]))


def default_state():
    d = np.load(os.path.join(GOLDEN, "naf256_k7_state.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


# ---- gradient fixtures (oracle/gen_golden_bwd.py: autograd through the unmodified reference) -------
def bwd_attention_case(name):
    d = load(name)
    seed = int(d["seed"])
    return dict(q=seeded_normal(seed, d["q_shape"]) * float(d["gain"]), k=seeded_normal(seed + 1, d["k_shape"]),
                v=seeded_normal(seed + 2, d["v_shape"]),
                dout=seeded_normal(seed + 3, (d["q_shape"][0], d["v_shape"][1], d["q_shape"][2], d["q_shape"][3])),
                out=torch.from_numpy(d["out"]), dq=torch.from_numpy(d["dq"]), dk=torch.from_numpy(d["dk"]),
                dv=torch.from_numpy(d["dv"]), heads=int(d["heads"]), K=int(d["kernel_size"]))


def bwd_module_case(name):
    d = load(name)
    seed = int(d["seed"])
    Ho, Wo = (int(x) for x in d["output_size"])
    feats_shape = d["features_shape"]
    grads = {k[len("grad__"):]: torch.from_numpy(d[k]) for k in d if k.startswith("grad__")}
    return dict(image=seeded_normal(seed, d["image_shape"]), features=seeded_normal(seed + 1, feats_shape),
                dout=seeded_normal(seed + 2, (1, int(feats_shape[1]), Ho, Wo)), out=torch.from_numpy(d["out"]),
                dimage=torch.from_numpy(d["dimage"]), dfeatures=torch.from_numpy(d["dfeatures"]),
                output_size=(Ho, Wo), grads=grads)


# ---- rectangular windows (oracle/gen_golden_rect.py: forward, scores and gradients from the unmodified reference)
def rect_attention_case(name):
    d = load(name)
    seed = int(d["seed"])
    q_shape, v_shape = d["q_shape"], d["v_shape"]
    return dict(q=seeded_normal(seed, q_shape) * float(d["gain"]), k=seeded_normal(seed + 1, d["k_shape"]),
                v=seeded_normal(seed + 2, v_shape), dout=seeded_normal(seed + 3, (q_shape[0], v_shape[1], q_shape[2], q_shape[3])),
                out=torch.from_numpy(d["out"]), scores=torch.from_numpy(d["scores"]), dq=torch.from_numpy(d["dq"]),
                dk=torch.from_numpy(d["dk"]), dv=torch.from_numpy(d["dv"]), heads=int(d["heads"]),
                K=tuple(int(x) for x in d["kernel_size"]), dilation=tuple(int(x) for x in d["dilation"]))
