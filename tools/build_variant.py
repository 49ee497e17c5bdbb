"""Build an A/B variant of the library with extra -D flags:  python tools/build_variant.py NAME -DFOO [-DBAR ...]
-> paillier_halo2_b200/variants/lib_NAME.so ; select it at run time with PB200_LIB=<path>."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paillier_halo2_b200 import build as b

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(b.HERE, "variants")
obj_dir = os.path.join(out_dir, "_obj_" + name)
os.makedirs(obj_dir, exist_ok=True)
objs = []
procs = []
for src in b.SOURCES:
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    objs.append(obj)
    procs.append(subprocess.Popen([b.NVCC] + b.ARCH + b.FLAGS + flags + ["-c", os.path.join(b.CSRC, src), "-o", obj]))
assert all(p.wait() == 0 for p in procs)
lib = os.path.join(out_dir, f"lib_{name}.so")
subprocess.check_call([b.NVCC] + b.ARCH + ["-shared", "-o", lib] + objs + ["-lcudart"])
print(lib)
