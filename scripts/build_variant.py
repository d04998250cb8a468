"""Build a profiling variant of libnaf_b200.so: ONE source recompiled with extra -D flags, linked with the
regular objects, into scripts/exp/libnaf_<name>.so (run it here; the .so travels to the GPU box; select it
with NAF_B200_LIB=scripts/exp/libnaf_<name>.so).
    python scripts/build_variant.py nostore naf_xattn_tma.cu -DNAF_TMA_EXP=1"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from naf_b200.csrc import build as B  # noqa: E402

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
objdir = os.path.join(B.HERE, "_obj")
out_dir = os.path.join(ROOT, "scripts", "exp")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, f"{name}_{src.replace('.cu', '.o')}")
subprocess.check_call([B.nvcc_path()] + B._flags() + flags + ["-c", os.path.join(B.HERE, src), "-o", obj])
objs = [obj if s == src else os.path.join(objdir, s.replace(".cu", ".o")) for s in B.SOURCES]
lib = os.path.join(out_dir, f"libnaf_{name}.so")
subprocess.check_call([B.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
print(lib)
