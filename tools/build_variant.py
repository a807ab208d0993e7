"""Developer A/B builds: recompile ONE translation unit with extra nvcc flags and link it with the
other objects of the in-tree build into zksnark-rs_b200/_var/libzkb200_<name>.so (git-ignored; it
travels to the GPU box).  Select it at run time with ZKB200_LIB=<path>.
Usage: python tools/build_variant.py <name> <source.cu>[,<source2.cu>...] [nvcc flags ...]"""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
    b = importlib.import_module("zksnark-rs_b200.build")
    b.build()
    var = os.path.join(b.HERE, "_var")
    os.makedirs(var, exist_ok=True)
    swapped = {}
    for one in src.split(","):
        obj = os.path.join(var, f"{name}_{one.replace('.cu', '.o')}")
        cmd = [b._nvcc()] + b.NVCC_FLAGS + flags + ["-c", os.path.join(b.CSRC, one), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stdout + r.stderr)
        for line in (r.stdout + r.stderr).splitlines():
            if "registers" in line or "spill" in line and " 0 bytes spill stores" not in line:
                print("  ", line.strip()[:160])
        swapped[one] = obj
    objs = [swapped.get(s, os.path.join(b.OBJ, s.replace(".cu", ".o"))) for s in b.SOURCES]
    lib = os.path.join(var, f"libzkb200_{name}.so")
    subprocess.check_call([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    print(lib)


if __name__ == "__main__":
    main()
